// K2: per-unique-profile classification (+ the text every site with that profile will print).
//   callSiteMLError body call.cpp:238-273, callBayes body call.cpp:176-194,
//   callLikelihoodRatio body call.cpp:93-127, likelihoodRatioTest stats.cpp:29-37.
// K6: CSV rows, operator<< call.hpp:29-38 + sid.cpp:103-105.
#pragma once
#include "calls.cuh"
#include "common.cuh"
#include "k_tokenize.cuh"
#include "table.cuh"

namespace sid {

struct ClassifyParams {
    TableView table;
    uint32_t first, last;        // range of table.entry_list to classify
    int method;                  // 0 local, 1 bayes, 2 likelihood_ratio (after BH: p-values given)
    double prior, error_threshold, alpha;
    LynchConsts lynch;
    double pi;
    int use_prior;
    const double* adj_hom;       // likelihood_ratio: BH-adjusted p-values per entry index (or NULL)
    const double* adj_het;
    const uint32_t* entry_to_unique;   // likelihood_ratio: entry index -> row of adj_* (0xFFFFFFFF: dropped)
    int het_only;                // rows of hom profiles get length 0: the CSV writer skips them
};

#if defined(__CUDACC__)

__device__ __forceinline__ void store_class(const TableView& t, uint32_t slot, const CallResult& r, bool probability, bool het_only) {
    t.label[slot] = r.label;
    t.gt[2 * slot] = r.gt0;
    t.gt[2 * slot + 1] = r.gt1;
    t.hom[slot] = r.hom;
    t.het[slot] = r.het;
    char buf[SUFFIX_BYTES];
    const int n = (het_only && r.label != 1) ? 0 : format_suffix(r, probability, buf);
    char* dst = t.suffix + (size_t)slot * SUFFIX_BYTES;
    for (int i = 0; i < n; ++i) dst[i] = buf[i];
    dst[SUFFIX_BYTES - 1] = (char)(n | SUFFIX_READY);      // length + "record complete" (the fused row writer waits for it)
}

__global__ void __launch_bounds__(128) k_classify(const ClassifyParams p) {
    const uint32_t e = p.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.last) return;
    const uint32_t slot = p.table.entry_list[e];
    const uint64_t profile = p.table.keys[slot];
    CallResult r;
    if (p.method == 0) {
        r = call_local(profile, p.prior, p.error_threshold, p.alpha);
    } else if (p.method == 1) {
        r = call_bayes(profile, p.lynch, p.pi);
    } else {
        int f, s;
        major_alleles(profile, f, s);
        r.gt0 = r.gt1 = base_char(f);
        r.label = 0;
        const uint32_t u = p.entry_to_unique[e];
        if (u == 0xFFFFFFFFu) { r.label = 255; r.hom = r.het = 0; }
        else {
            r.hom = p.adj_hom[u];
            r.het = p.adj_het[u];
            if (r.het < p.alpha) { r.label = 1; r.gt1 = base_char(s); }     // call.cpp:120-123
        }
    }
    store_class(p.table, slot, r, p.method == 1, p.het_only != 0);
}

// ------------------------------------------------------------------------------------------- K6
// CSV writer (operator<< call.hpp:29-38 + the print loop sid.cpp:102-106).  Every warp works on its own:
// it draws a tile of 128 consecutive sites (file order), computes the row lengths, publishes their sum
// and finds its byte offset with a warp-wide decoupled look-back, assembles the rows in its private
// piece of shared memory and copies them out with aligned 16-byte stores.  No CTA-wide barrier: a warp
// waiting for memory or for its predecessors does not hold up the others.
#ifndef SID_CSV_THREADS
#define SID_CSV_THREADS 128     // 4 warps per CTA: at 72 registers 7 CTAs = 28 warps fit an SM (256 threads: 3 CTAs = 24 warps)
#endif
constexpr int CSV_THREADS = SID_CSV_THREADS;
constexpr int CSV_WARPS = CSV_THREADS / 32;
#ifndef SID_CSV_PER_THREAD
#define SID_CSV_PER_THREAD 4
#endif
constexpr int CSV_PER_THREAD = SID_CSV_PER_THREAD;
constexpr int CSV_TILE = 32 * CSV_PER_THREAD;     // sites per warp tile
constexpr int CSV_WSTAGE = 1728 * CSV_PER_THREAD; // bytes of rows a warp may stage (else: direct global writes)
constexpr int CSV_STAGE = CSV_WARPS * CSV_WSTAGE;

struct CsvParams {
    uint64_t site_begin, n_sites;    // file-order range of the store
    const uint32_t* order;           // file index -> storage index
    const int32_t* pos;
    const uint32_t* slot;
    const uint32_t* name_ref;
    const char* site_suffix;     // per-site suffixes (quality) or NULL -> table.suffix[slot]
    TableView table;
    const char* pool;
    char* out;
    uint64_t out_cap;
    unsigned int* ticket;
    unsigned long long* status;  // one look-back word per tile, zeroed
    unsigned long long* bytes_out;
    unsigned long long* rows_out;
    unsigned long long* error;   // atomicMin target; a look-back that never resolves reports LINE_MALFORMED + 4
    uint32_t n_tiles;
};

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

// Bytes before tile `tile`: sums the published aggregates of its predecessors, 32 at a time, back to the
// first one whose inclusive prefix is known.
__device__ __forceinline__ unsigned long long csv_look_back(const CsvParams& p, uint32_t tile, int lane, bool& ok) {
    unsigned long long base = 0;
    int64_t idx = (int64_t)tile - 1;
    uint32_t idle = 0;
    ok = true;
    while (idx >= 0) {
        const int64_t i = idx - lane;
        unsigned long long w = LB_FLAG_PREFIX;                                   // before tile 0: prefix 0
        if (i >= 0) w = *((volatile unsigned long long*)&p.status[i]);
        const uint32_t flag = (uint32_t)(w >> 62);
        const uint32_t not_ready = __ballot_sync(0xFFFFFFFFu, flag == 0);
        const uint32_t usable = not_ready ? ((1u << (__ffs((int)not_ready) - 1)) - 1u) : 0xFFFFFFFFu;   // lanes before the first gap
        const uint32_t pref = __ballot_sync(0xFFFFFFFFu, flag == 2) & usable;
        if (pref) {
            const int first = __ffs((int)pref) - 1;
            base += warp_sum_u64(lane <= first ? (w & LB_VALUE_MASK) : 0ull);
            return base;
        }
        if (usable) {
            base += warp_sum_u64(((usable >> lane) & 1u) ? (w & LB_VALUE_MASK) : 0ull);
            idx -= __popc(usable);
            idle = 0;
        } else {
            if (++idle > (1u << 22)) { ok = false; return base; }
            __nanosleep(40);
        }
    }
    return base;
}

#ifndef SID_CSV_CTAS
#define SID_CSV_CTAS 3      // minimum CTAs per SM for the register allocator: loose on purpose.  A 64-register cap (32 warps per
                            // SM) costs more than the extra warps bring: 0.69 ms against 0.61 ms at 72 registers and 28 warps
#endif
__global__ void __launch_bounds__(CSV_THREADS, SID_CSV_CTAS) k_csv(const CsvParams p) {
    extern __shared__ __align__(16) char s_stage_all[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    char* s_stage = s_stage_all + warp * CSV_WSTAGE;
    unsigned long long my_rows = 0;
    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(p.ticket, 1u);      // in order: the look-back needs every earlier tile started
        tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
        if (tile >= p.n_tiles) break;
        const uint64_t first = (uint64_t)tile * CSV_TILE + (uint64_t)lane * CSV_PER_THREAD;
        // storage indices of this lane's four sites (0xFFFFFFFF past the end)
        uint32_t ord[CSV_PER_THREAD];
#pragma unroll
        for (int k = 0; k < CSV_PER_THREAD; ++k) ord[k] = 0xFFFFFFFFu;
        if (first < p.n_sites) {
            const uint32_t* src = p.order + p.site_begin + first;
            if (CSV_PER_THREAD % 4 == 0 && first + CSV_PER_THREAD <= p.n_sites && ((uintptr_t)src & 15u) == 0) {
#pragma unroll
                for (int k = 0; k < CSV_PER_THREAD / 4; ++k) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + k);
                    ord[4 * k] = v.x; ord[4 * k + 1] = v.y; ord[4 * k + 2] = v.z; ord[4 * k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < CSV_PER_THREAD; ++k) if (first + k < p.n_sites) ord[k] = __ldg(src + k);
            }
        }
        uint32_t len[CSV_PER_THREAD], nlen[CSV_PER_THREAD], slen[CSV_PER_THREAD], nref[CSV_PER_THREAD];
        int32_t pos[CSV_PER_THREAD];
        const char* sfxp[CSV_PER_THREAD];
        uint32_t mine = 0, rows = 0;
#pragma unroll
        for (int k = 0; k < CSV_PER_THREAD; ++k) {
            len[k] = 0;
            if (ord[k] != 0xFFFFFFFFu) {
                const uint64_t site = ord[k];
                const char* sfx = p.site_suffix ? p.site_suffix + site * SUFFIX_BYTES
                                                : p.table.suffix + (size_t)p.slot[site] * SUFFIX_BYTES;
                sfxp[k] = sfx;
                const uint32_t sl = (uint8_t)sfx[SUFFIX_BYTES - 1] & 0x7Fu;
                slen[k] = sl;
                if (sl) {
                    nref[k] = p.name_ref[site];
                    pos[k] = p.pos[site];
                    const uint32_t hdr = *reinterpret_cast<const uint32_t*>(p.pool + nref[k]);   // pool records are 4-byte aligned
                    nlen[k] = hdr & 0xFFFFu;
                    len[k] = nlen[k] + 1 + (uint32_t)digits_i32(pos[k]) + sl;
                    ++rows;
                }
            }
            mine += len[k];
        }
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        my_rows += rows;
        // ---- byte offset of the tile
        if (lane == 0) atomicExch(&p.status[tile], (tile == 0 ? LB_FLAG_PREFIX : LB_FLAG_AGG) | (unsigned long long)total);
        unsigned long long tile_base = 0;
        if (tile) {
            bool ok;
            tile_base = csv_look_back(p, tile, lane, ok);
            if (!ok && lane == 0) atomicMin(p.error, (unsigned long long)(LINE_MALFORMED + 4));
            if (lane == 0) atomicExch(&p.status[tile], LB_FLAG_PREFIX | (tile_base + total));
        }
        if (tile == p.n_tiles - 1 && lane == 0) *p.bytes_out = tile_base + total;
        if (tile_base + total > p.out_cap) continue;                   // the host reports SIDGPU_ECAPACITY from bytes_out
        // ---- rows: assembled in shared memory at (global offset mod 16), so that 16-byte chunks line up
        const uint32_t mis = (uint32_t)((uintptr_t)(p.out + tile_base) & 15u);
        const bool staged = total + mis + 16 <= CSV_WSTAGE;
        uint32_t o = incl - mine;
#pragma unroll
        for (int k = 0; k < CSV_PER_THREAD; ++k) {
            if (!len[k]) continue;
            char* w = staged ? s_stage + mis + o : p.out + tile_base + o;
            const uint8_t* nm = (const uint8_t*)p.pool + nref[k];
            for (uint32_t i = 0; i < nlen[k]; ++i) w[i] = (char)nm[2 + i];
            w += nlen[k];
            *w++ = ',';
            w += fmt_i32_fast(pos[k], w);
            // suffix: three aligned 16-byte loads, copied byte-wise into the (unaligned) row
            const uint4* sv = reinterpret_cast<const uint4*>(sfxp[k]);
            const uint32_t sl = slen[k];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if ((uint32_t)(16 * c) < sl) {
                    const uint4 v = sv[c];
                    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int b = 0; b < 16; ++b) {
                        if ((uint32_t)(16 * c + b) < sl) w[16 * c + b] = (char)((wd[b >> 2] >> (8 * (b & 3))) & 0xFF);
                    }
                }
            }
            o += len[k];
        }
        if (staged) {
            __syncwarp();
            char* g = p.out + tile_base;               // first byte of the tile in global memory
            // head: bytes up to the first 16-byte boundary; body: aligned vectors; tail: the rest
            const uint32_t head = mis ? min(16u - mis, total) : 0u;
            if ((uint32_t)lane < head) g[lane] = s_stage[mis + lane];
            const uint32_t body = (total - head) >> 4;
            const uint4* sv = reinterpret_cast<const uint4*>(s_stage + mis + head);    // 16-byte aligned by construction
            uint4* gv = reinterpret_cast<uint4*>(g + head);
            for (uint32_t i = lane; i < body; i += 32) gv[i] = sv[i];
            const uint32_t done = head + (body << 4);
            if (done + (uint32_t)lane < total) g[done + lane] = s_stage[mis + done + lane];
            __syncwarp();
        }
    }
    my_rows = warp_sum_u64(my_rows);
    if (lane == 0 && my_rows) atomicAdd(p.rows_out, my_rows);
}

// Per-site records instead of text (OutputRecord, call.hpp:14-27).
struct RecordParams {
    uint64_t site_begin, n_sites;
    const uint32_t* order;
    const uint32_t* slot;
    TableView table;
    uint8_t* label;
    char* gt;
    double* hom;
    double* het;
    const int32_t* pos_in;          // columns of the site store, copied out in file order when asked for
    const uint32_t* name_ref_in;
    int32_t* pos;
    uint32_t* name_ref;
    const unsigned long long* fwd_in;       // forward-strand profile per site of the store (sessions with want_strands)
    unsigned long long* profile;
    unsigned long long* fwd;
};

__global__ void k_records(const RecordParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_sites) return;
    const uint32_t site = p.order[p.site_begin + i];
    const uint32_t s = p.slot[site];
    if (p.pos) p.pos[i] = p.pos_in[site];
    if (p.name_ref) p.name_ref[i] = p.name_ref_in[site];
    if (p.label) p.label[i] = p.table.label[s];
    if (p.gt) { p.gt[2 * i] = p.table.gt[2 * s]; p.gt[2 * i + 1] = p.table.gt[2 * s + 1]; }
    if (p.hom) p.hom[i] = p.table.hom[s];
    if (p.het) p.het[i] = p.table.het[s];
    if (p.profile) p.profile[i] = p.table.keys[s];
    if (p.fwd) p.fwd[i] = p.fwd_in[site];
}

#endif  // __CUDACC__

}  // namespace sid
