// The per-read side of the pileup parser interface, for hosts that ask for it (readFile(in, true, true), call.cpp:11-20):
//   PileupLine::bases / strands     parseReadBases pileup.cpp:84-123 (one entry per counted base: the letter in upper
//                                   case, strand 1 for an upper-case character, 0 for a lower-case one)
//   base / mapping qualities        parseQualities pileup.cpp:155-167 (character - 33, at least 1)
// and the per-profile likelihoods / p-values behind the host mirrors of lynch.hpp and stats.hpp.
// One thread per line: this is an interface for inspection and tests, not the calling path (which never materialises
// per-read vectors: k_quality consumes the text directly).
#pragma once
#include "calls.cuh"
#include "k_quality.cuh"
#include "parse.cuh"

namespace sid {

#if defined(__CUDACC__)

// Lengths of the three per-read vectors of every line (line_off: first byte of the line in `text`).
// The five columns up to the bases must be there (pileup.cpp:20-40); the quality columns only when they are asked
// for: a missing base-quality column is SIDGPU_EMALFORMED then (the reference hands a null pointer to parseQualities
// there), a missing mapping-quality column SIDGPU_EMISSING_MAPQ (pileup.cpp:60-66).  A quality string shorter than
// the counted bases is fine, as it is for the reference's parser (its own test line has six bases and five qualities).
__global__ void k_read_counts(const uint8_t* text, uint64_t text_len, const uint64_t* line_off, uint64_t n, int want_bq, int want_mq,
                              uint32_t* n_bases, uint32_t* n_bq, uint32_t* n_mq, unsigned long long* error) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FlatSrc src {text, text_len};
    ParsedLine pl;
    parse_line(src, line_off[i], false, pl);
    n_bases[i] = n_bq[i] = n_mq[i] = 0;
    int bad = pl.status != LINE_OK ? LINE_MALFORMED : LINE_OK;
    if (bad == LINE_OK) {
        n_bases[i] = pl.n_bases;
        if (want_bq || want_mq) {
            parse_line(src, line_off[i], true, pl);
            if (pl.status == LINE_MALFORMED) {                      // no sixth column
                if (want_bq) bad = LINE_MALFORMED;
                else if (want_mq) bad = LINE_MISSING_MAPQ;
            } else {
                n_bq[i] = pl.bq_len;
                if (pl.status == LINE_MISSING_MAPQ) { if (want_mq) bad = LINE_MISSING_MAPQ; }
                else n_mq[i] = pl.mq_len;
            }
        }
    }
    if (bad != LINE_OK) atomicMin(error, (unsigned long long)((line_off[i] << 3) | (uint64_t)bad));
}

__global__ void k_read_fill(const uint8_t* text, uint64_t text_len, const uint64_t* line_off, uint64_t n, const uint64_t* base_off,
                            const uint64_t* bq_off, const uint64_t* mq_off, char* bases, uint8_t* strands, uint8_t* bq, uint8_t* mq) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FlatSrc src {text, text_len};
    ParsedLine pl;
    const uint64_t p = line_off[i];
    parse_line(src, p, false, pl);
    if (pl.status != LINE_OK) return;
    if (bases || strands) {
        BasesState st;
        st.init((uint8_t)pl.ref);
        uint64_t o = base_off[i];
        for (uint32_t k = 0; k < pl.bases_len; ++k) {
            const uint8_t c = src.at(p + pl.bases_off + k);
            const int idx = st.feed(c);
            if (idx < 0) continue;
            // the character the switch saw (pileup.cpp:78-83): its case is the strand
            const uint8_t seen = c == '.' ? st.dot_as : c == ',' ? st.comma_as : c;
            if (bases) bases[o] = "ACGT"[idx];
            if (strands) strands[o] = (seen & 0x20u) ? 0 : 1;
            ++o;
        }
    }
    if (!bq && !mq) return;
    parse_line(src, p, true, pl);
    if (pl.status == LINE_MALFORMED) return;
    if (bq) for (uint32_t k = 0; k < pl.bq_len; ++k) bq[bq_off[i] + k] = (uint8_t)phred_of(src.at(p + pl.bq_off + k));
    if (mq && pl.status != LINE_MISSING_MAPQ) for (uint32_t k = 0; k < pl.mq_len; ++k) mq[mq_off[i] + k] = (uint8_t)phred_of(src.at(p + pl.mq_off + k));
}

// Per-strand profiles of every line (the strands parseReadBases derives and nobody reads, pileup.hpp:15,
// pileup.cpp:87-123): the counted bases of the upper-case characters ('.' included) and of the lower-case ones (',' included),
// packed like the profile (A | C<<16 | G<<32 | T<<48, each count mod 65536); fwd + rev is the line's profile, lane by lane.
__global__ void k_strand_counts(const uint8_t* text, uint64_t text_len, const uint64_t* line_off, uint64_t n, unsigned long long* fwd,
                                unsigned long long* rev, unsigned long long* error) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FlatSrc src {text, text_len};
    ParsedLine pl;
    const uint64_t p = line_off[i];
    parse_line(src, p, false, pl);
    uint32_t c[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    if (pl.status != LINE_OK) atomicMin(error, (unsigned long long)((p << 3) | (uint64_t)LINE_MALFORMED));
    else {
        BasesState st;
        st.init((uint8_t)pl.ref);
        for (uint32_t k = 0; k < pl.bases_len; ++k) {
            const uint8_t ch = src.at(p + pl.bases_off + k);
            const int idx = st.feed(ch);
            if (idx < 0) continue;
            const uint8_t seen = ch == '.' ? st.dot_as : ch == ',' ? st.comma_as : ch;      // pileup.cpp:78-83
            ++c[(seen & 0x20u) ? 1 : 0][idx];
        }
    }
    if (fwd) fwd[i] = pack_profile(c[0][0], c[0][1], c[0][2], c[0][3]);
    if (rev) rev[i] = pack_profile(c[1][0], c[1][1], c[1][2], c[1][3]);
}

// parseQualities of one string: stops at NUL, tab or line end (pileup.cpp:158).  One block.
__global__ void __launch_bounds__(256) k_qualities(const uint8_t* q, uint64_t n, uint8_t* out, unsigned long long* n_out) {
    __shared__ unsigned long long s_end;
    if (threadIdx.x == 0) s_end = n;
    __syncthreads();
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint8_t c = q[i];
        if (c == 0 || c == '\t' || c == '\n') { atomicMin(&s_end, (unsigned long long)i); break; }
    }
    __syncthreads();
    const uint64_t end = s_end;
    for (uint64_t i = threadIdx.x; i < end; i += blockDim.x) out[i] = (uint8_t)phred_of(q[i]);
    if (threadIdx.x == 0) *n_out = end;
}

// log homozygousLikelihood / heterozygousLikelihood (lynch.hpp:57-74,82-90) of packed profiles, multinomial coefficient included.
__global__ void k_profile_loglik(const unsigned long long* profiles, uint64_t n, LynchConsts k, double* log_hom, double* log_het) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lh, lt;
    lynch_loglik(profiles[i], k, lh, lt);
    const double m = log_multinomial(profiles[i]);
    log_hom[i] = lh + m;
    log_het[i] = lt + m;
}

// likelihoodRatioTest (stats.cpp:29-37) on log-likelihoods (-inf: l == 0).
__global__ void k_lr_test(const double* log_h0, const double* log_h1, uint64_t n, double* p) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = lrt_log(log_h0[i], log_h1[i]);
}

#endif  // __CUDACC__

}  // namespace sid
