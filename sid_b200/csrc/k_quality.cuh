// `-m quality`: callQualityBasedSimple body (call.cpp:309-370), one thread per site.
// The j-th counted base is paired with the j-th base-quality and the j-th mapping-quality
// character (call.cpp:330-331), exactly as the reference does, including the drift that
// characters without a base ('*', 'N', '<', '>') cause.  The per-read terms come from four
// 256-entry tables built on the host with the reference's own expressions
// (pow(10., q / -10.), log(1 - e), ...), so each term is bit-identical; the running sums use a
// compensated accumulator in place of x87 long double.
#pragma once
#include "calls.cuh"
#include "common.cuh"
#include "parse.cuh"
#include "parse_win.cuh"
#include "table.cuh"

namespace sid {

constexpr int QUAL_THREADS = 128;
// log of the smallest value expl() can return on x87 (2^-16446 rounds to zero): below it the
// reference's exp(log_probability) is exactly 0 and likelihoodRatioTest takes its l_H0 == 0 branch.
constexpr double LOG_LDBL_ZERO = -11399.498531488861;

struct QualityParams {
    const uint8_t* text;
    uint64_t text_len;
    const uint64_t* line_off;
    const uint64_t* profile;    // per site, as stored by the tokenizer
    const double* qual_l;       // optional, two per site: the log-likelihood sums the tokenizer already formed (+inf: it did not)
    const uint32_t* order;      // file index -> storage index; NULL: the launch covers the whole store, thread i takes storage index i
                                // (no indirection, neighbouring threads read and write neighbouring records)
    uint64_t site_begin, n_sites;
    const double* lut;          // [0,256) log(1-e)  [256,512) log(e)  [512,768) log(1-2e/3)  [768,1024) log(2e/3)  [1024,1280) lgamma(n+1)
    double prior, alpha;
    int het_only;               // rows of hom sites get length 0
    char* site_suffix;          // SUFFIX_BYTES per site
    unsigned long long* error;
    // optional per-site records (OutputRecord, call.hpp:14-27) in file order, index i - site_begin
    uint8_t* rec_label;
    char* rec_gt;
    double* rec_hom;
    double* rec_het;
};

SID_HD uint32_t phred_of(uint8_t c) {                    // parseQualities pileup.cpp:158-163
    const uint32_t q = (uint8_t)(c - 33);
    return q < 1 ? 1u : q;
}

// The bases field walked once more for the order of the counted bases (parseReadBases pileup.cpp:70-153
// without the counters): 0..3 = A C G T, -1 = nothing counted.  '.' / ',' are replaced by toupper / tolower of
// the reference character before they are looked at (pileup.cpp:78-83).
struct BasesWalk {
    uint64_t skip, num;
    int mode;                // 0 normal, 1 just saw '+'/'-', 2 reading the indel length
    uint8_t dot_as, comma_as;
    SID_HD void init(uint8_t ref) { skip = 0; num = 0; mode = 0; dot_as = ascii_upper(ref); comma_as = ascii_lower(ref); }
    SID_HD int feed(uint8_t c) {
        if (mode) {
            const uint32_t d = (uint32_t)c - (uint32_t)'0';
            if (d <= 9) {
                if (mode == 1) { mode = 2; num = d; } else if (num < (1ull << 40)) num = num * 10 + d;
                return -1;
            }
            if (mode == 2) skip = num;          // pileup.cpp:144
            mode = 0;                           // pileup.cpp:131-133: a sign without digits is ignored
        }
        if (skip) { --skip; return -1; }
        c = c == '.' ? dot_as : c;
        c = c == ',' ? comma_as : c;
        // no switch: the lanes of a warp look at different characters, selects keep them together
        const uint32_t u = (uint32_t)c | 0x20u;                       // letters to lower case; '+' '-' unchanged, '^' -> '~'
        int r = -1;
        r = u == 'a' ? 0 : r;
        r = u == 'c' ? 1 : r;
        r = u == 'g' ? 2 : r;
        r = u == 't' ? 3 : r;
        skip = c == '^' ? 1 : 0;                                      // pileup.cpp:125-127 (skip was 0 here)
        mode = (c == '+' || c == '-') ? 1 : 0;                        // (mode was 0 here)
        return r;
    }
};

constexpr int LOG_FACT_N = 256;   // lut[1024 + n] = lgamma(n + 1) for n < LOG_FACT_N
SID_HD double log_factorial(const double* lut, uint32_t n) { return n < (uint32_t)LOG_FACT_N ? lut[1024 + n] : lgamma((double)n + 1.0); }

// The end of the per-read sums: the last partial block, then the binomial coefficient and 2^-n of the heterozygous
// model (call.cpp:347-350).  l1 / l2: log-likelihood of the homozygous / heterozygous model.
SID_HD void quality_sums_finish(uint64_t profile, int ref0, int ref1, const double* lut, CompSum lh, CompSum lt, double bh, double bt,
                                double& l1, double& l2) {
    lh.add(bh);
    lt.add(bt);
    const uint32_t n = profile_count(profile, ref0) + profile_count(profile, ref1);   // call.cpp:347-349
    const uint32_t k = profile_count(profile, ref1);
    lt.add(log_factorial(lut, n) - log_factorial(lut, n - k) - log_factorial(lut, k));
    lt.add(-(double)n * 0.69314718055994530942);
    l1 = lh.value();
    l2 = lt.value();
}

// From the two log-likelihoods to the call (call.cpp:352-367).
SID_HD CallResult quality_result(uint64_t profile, double l1, double l2, double prior, double alpha) {
    int ref0, ref1;
    major_alleles(profile, ref0, ref1);
    if (l1 < LOG_LDBL_ZERO) l1 = neg_inf();               // call.cpp:352-353 exp() underflow
    if (l2 < LOG_LDBL_ZERO) l2 = neg_inf();
    if (prior > 0) { l1 += log1p(-prior); l2 += log(prior); }   // call.cpp:354-357
    CallResult res;
    res.hom = lrt_log(l2, l1);                            // call.cpp:359
    res.het = lrt_log(l1, l2);                            // call.cpp:360
    res.label = 0;
    res.gt0 = res.gt1 = base_char(ref0);
    if (res.het < alpha) { res.label = 1; res.gt1 = base_char(ref1); }    // call.cpp:364-367
    return res;
}

// ---- the per-read sums inside the tokenizer -------------------------------------------------------------------------
// k_quality walks every bases field a second time, one byte per step, from global memory: 60 % of its instructions.
// In a quality session the tokenizer (k_tok2<..., QUAL>) has the class windows of the line at hand: the counted bases
// are the set bits of (base | dot) & live, their order is the bit order, their letter two bit planes.  WinQuality is
// the visitor of win_bases_walk that forms the same sums in the same order as call_quality (bit-identical doubles).
struct WinQuality {
    const uint8_t* s;           // staged text
    uint32_t bq, mq;            // offsets in `s` of the base-quality and the mapping-quality string
    const double* lut;
    int ref0, ref1;             // major alleles of the line's profile (call.cpp:311-319)
    int ref_idx;                // index of the reference base that '.' / ',' stand for; -1: not one of ACGT
    uint32_t j;
    CompSum lh, lt;
    double bh, bt;
    SID_HD void init() { j = 0; lh.init(); lt.init(); bh = bt = 0; }
    SID_HD void operator()(const Win64& w, uint64_t live) {
        // which counted bytes are the major allele, which one of the two major alleles: as masks, once per window
        const uint64_t b = w.base & live;
        uint64_t is[4] = {b & ~w.p1 & ~w.p2, b & w.p1 & ~w.p2, b & w.p1 & w.p2, b & ~w.p1 & w.p2};      // A C G T
        const uint64_t d = ref_idx >= 0 ? (w.dot & live) : 0ull;
#pragma unroll
        for (int i = 0; i < 4; ++i) is[i] |= (i == ref_idx) ? d : 0ull;
        uint64_t m0 = 0, m1 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { m0 |= (i == ref0) ? is[i] : 0ull; m1 |= (i == ref1) ? is[i] : 0ull; }
        const uint64_t all = b | d, m01 = m0 | m1;
        // the counted bytes in text order, 32 bits at a time (64-bit shifts and scans cost twice as much here)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t counted = (uint32_t)(all >> (32 * h));
            const uint32_t h0 = (uint32_t)(m0 >> (32 * h)), h01 = (uint32_t)(m01 >> (32 * h));
            while (counted) {
                const uint32_t bit = first_bit(counted);
                counted &= counted - 1;
                const uint32_t q1 = phred_of(s[bq + j]), q2 = phred_of(s[mq + j]);
                const uint32_t q = q1 < q2 ? q1 : q2;                                 // call.cpp:330
                ++j;
                bh += lut[q + (((h0 >> bit) & 1u) ? 0u : 256u)];                      // call.cpp:331-335
                bt += lut[q + (((h01 >> bit) & 1u) ? 512u : 768u)];                   // call.cpp:336-340
                if ((j & 7u) == 0) { lh.add(bh); lt.add(bt); bh = bt = 0; }
            }
        }
    }
};

// The log-likelihood sums of the line at line_off, given what parse_line_win found (profile, header, end of the bases
// field).  Returns false -- and the line goes to k_quality's own walk -- unless the two quality strings follow the
// bases field separated by single delimiters, lie within the classified bytes and are long enough (anything else is
// an error or an oddity that the byte-wise code reports exactly as the reference does).
SID_HD bool quality_sums_win(const uint8_t* s, uint32_t region_off, const uint32_t* cw, const uint32_t* nlw, uint32_t n_bits, const WinHeader& hd,
                             uint32_t bases_end, uint64_t profile, const double* lut, double& l1, double& l2) {
    const uint32_t b = bases_end;
    if (b + 2 >= n_bits || ((nlw[b >> 5] >> (b & 31)) & 1u)) return false;            // no sixth column on this line
    const uint32_t t6 = win_field_end(cw, n_bits, b + 1);
    if (t6 == 0xFFFFFFFFu || t6 == b + 1 || t6 + 2 >= n_bits || ((nlw[t6 >> 5] >> (t6 & 31)) & 1u)) return false;
    const uint32_t t7 = win_field_end(cw, n_bits, t6 + 1);
    if (t7 == 0xFFFFFFFFu || t7 == t6 + 1) return false;
    const uint32_t n_bases = profile_count(profile, 0) + profile_count(profile, 1) + profile_count(profile, 2) + profile_count(profile, 3);
    if (n_bases > t6 - b - 1 || n_bases > t7 - t6 - 1) return false;                  // LINE_QUAL_SHORT: reported by the byte-wise code
    WinQuality v;
    v.init();
    v.s = s;
    v.bq = region_off + b + 1;
    v.mq = region_off + t6 + 1;
    v.lut = lut;
    major_alleles(profile, v.ref0, v.ref1);
    const uint32_t rc = (hd.ref_p2 ? 2u : 0u) + (hd.ref_p1 ? 1u : 0u);
    v.ref_idx = hd.ref_base ? (int)(rc ^ (rc >> 1)) : -1;
    const Win64 w = load_window(cw, hd.l0);
    if (!win_bases_walk(s, region_off, cw, n_bits, hd, w, true, v, nullptr)) return false;
    quality_sums_finish(profile, v.ref0, v.ref1, lut, v.lh, v.lt, v.bh, v.bt, l1, l2);
    return true;
}

// `text` holds the whole line: every index below was validated by quality_fields (bq_len, mq_len >= counted bases).
SID_HD CallResult call_quality(const uint8_t* text, uint64_t line_abs, const ParsedLine& pl, const double* lut, double prior,
                               double alpha) {
    int ref0, ref1;
    major_alleles(pl.profile, ref0, ref1);                // call.cpp:311-319
    // sums of per-read terms: eight plain additions at a time, the blocks added with compensation (in place of
    // the reference's x87 long double accumulation)
    CompSum lh, lt;
    lh.init();
    lt.init();
    double bh = 0, bt = 0;
    BasesWalk b;
    b.init((uint8_t)pl.ref);
    uint32_t j = 0;
    const uint8_t* bases = text + line_abs + pl.bases_off;
    const uint8_t* bq = text + line_abs + pl.bq_off;
    const uint8_t* mq = text + line_abs + pl.mq_off;
    for (uint32_t i = 0; i < pl.bases_len; ++i) {
        const int r = b.feed(bases[i]);                  // '.' / ',' stand for the reference character
        if (r < 0) continue;
        const uint32_t q1 = phred_of(bq[j]), q2 = phred_of(mq[j]);
        const uint32_t q = q1 < q2 ? q1 : q2;             // call.cpp:330
        ++j;
        bh += r == ref0 ? lut[q] : lut[256 + q];          // call.cpp:331-335
        bt += (r == ref0 || r == ref1) ? lut[512 + q] : lut[768 + q];   // call.cpp:336-340
        if ((j & 7u) == 0) { lh.add(bh); lt.add(bt); bh = bt = 0; }
    }
    double l1, l2;
    quality_sums_finish(pl.profile, ref0, ref1, lut, lh, lt, bh, bt, l1, l2);
    return quality_result(pl.profile, l1, l2, prior, alpha);
}

#if !defined(__CUDACC__)
// Host check: the quality call of the line at p the way the QUAL tokenizer + k_quality form it (window walk, sums,
// quality_result).  Returns false when the line leaves that path (k_quality's own walk takes it then).
inline bool quality_line_win_host(const uint8_t* text, uint64_t len, uint64_t p, const double* lut, double prior, double alpha, CallResult& r) {
    const HostLineClasses h = classify_line_host(text, len, p);
    if (!h.usable) return false;
    WinLine wl;
    WinHeader hd;
    uint32_t end_bit = 0;
    if (!parse_line_win<true>(h.scratch, 0, h.cw, h.nlw, h.units * 32, h.line_off, wl, &hd, &end_bit)) return false;
    double l1, l2;
    if (!quality_sums_win(h.scratch, 0, h.cw, h.nlw, h.units * 32, hd, end_bit, wl.profile, lut, l1, l2)) return false;
    r = quality_result(wl.profile, l1, l2, prior, alpha);
    return true;
}
#endif

#if defined(__CUDACC__)
__global__ void __launch_bounds__(QUAL_THREADS) k_quality(const QualityParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * QUAL_THREADS + threadIdx.x;
    if (i >= p.n_sites) return;
    const uint64_t site = p.order ? p.order[p.site_begin + i] : i;
    const uint64_t line_abs = p.line_off[site];
    char* dst = p.site_suffix + site * SUFFIX_BYTES;
    CallResult r;
    const double l1_tok = p.qual_l ? p.qual_l[2 * site] : bits_double(0x7FF0000000000000ull);
    if (l1_tok != bits_double(0x7FF0000000000000ull)) {
        // the tokenizer formed the sums while it had the line in shared memory: only the call is left
        r = quality_result(p.profile[site], l1_tok, p.qual_l[2 * site + 1], p.prior, p.alpha);
    } else {
        ParsedLine pl;
        quality_fields(WordSrc {p.text, p.text_len}, line_abs, p.profile[site], pl);      // parse_line's offsets, lengths, status
        if (pl.status != LINE_OK) {
            atomicMin(p.error, (unsigned long long)((line_abs << 3) | (uint64_t)pl.status));
            dst[SUFFIX_BYTES - 1] = 0;
            return;
        }
        r = call_quality(p.text, line_abs, pl, p.lut, p.prior, p.alpha);
    }
    if (p.rec_label) p.rec_label[i] = r.label;
    if (p.rec_gt) { p.rec_gt[2 * i] = r.gt0; p.rec_gt[2 * i + 1] = r.gt1; }
    if (p.rec_hom) p.rec_hom[i] = r.hom;
    if (p.rec_het) p.rec_het[i] = r.het;
    // the record leaves as three 16-byte stores (byte stores of 30 characters per thread, 48 bytes apart from the
    // neighbouring thread's, were a third of this kernel's time)
    __align__(16) char buf[SUFFIX_BYTES];
    const int n = (p.het_only && r.label != 1) ? 0 : format_suffix(r, false, buf);
    uint32_t w[12];
    {
        const uint4* bv = reinterpret_cast<const uint4*>(buf);
        const uint4 v0 = bv[0], v1 = bv[1], v2 = bv[2];
        w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        w[8] = v2.x; w[9] = v2.y; w[10] = v2.z; w[11] = v2.w;
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {              // bytes past the text are whatever the stack held
        const int rem = n - 4 * k;
        w[k] = rem >= 4 ? w[k] : (rem > 0 ? (w[k] & ((1u << (8 * rem)) - 1u)) : 0u);
    }
    w[11] = (w[11] & 0x00FFFFFFu) | ((uint32_t)n << 24);
    uint4* dv = reinterpret_cast<uint4*>(dst);
    dv[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dv[1] = make_uint4(w[4], w[5], w[6], w[7]);
    dv[2] = make_uint4(w[8], w[9], w[10], w[11]);
}
#endif

}  // namespace sid
