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
};

#if defined(__CUDACC__)

__device__ __forceinline__ void store_class(const TableView& t, uint32_t slot, const CallResult& r, bool probability) {
    t.label[slot] = r.label;
    t.gt[2 * slot] = r.gt0;
    t.gt[2 * slot + 1] = r.gt1;
    t.hom[slot] = r.hom;
    t.het[slot] = r.het;
    char buf[SUFFIX_BYTES];
    const int n = format_suffix(r, probability, buf);
    buf[SUFFIX_BYTES - 1] = (char)n;
    char* dst = t.suffix + (size_t)slot * SUFFIX_BYTES;
    for (int i = 0; i < n; ++i) dst[i] = buf[i];
    dst[SUFFIX_BYTES - 1] = (char)n;
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
    store_class(p.table, slot, r, p.method == 1);
}

// ------------------------------------------------------------------------------------------- K6
constexpr int CSV_THREADS = 256;
constexpr int CSV_PER_THREAD = 4;
constexpr int CSV_TILE = CSV_THREADS * CSV_PER_THREAD;

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
    unsigned long long* status;
    unsigned long long* bytes_out;
    unsigned long long* rows_out;
    uint32_t n_tiles;
};

// Rows of one tile are assembled in shared memory and copied out with aligned 16-byte stores.
// The staging area starts at (global offset of the tile's first byte) mod 16, so 16-byte chunks of
// shared memory line up with 16-byte chunks of the output buffer.
constexpr int CSV_STAGE = 52 * 1024;          // bytes of rows one tile may stage (else: direct global writes)

__global__ void __launch_bounds__(CSV_THREADS) k_csv(const CsvParams p) {
    extern __shared__ __align__(16) char s_stage[];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    __shared__ uint32_t s_warp_sums[CSV_THREADS / 32];
    __shared__ uint32_t s_warp_rows[CSV_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // storage indices of this thread's four sites of a tile (0xFFFFFFFF past the end)
    auto load_order = [&](uint32_t tile, uint32_t (&ord)[CSV_PER_THREAD]) {
        const uint64_t first = (uint64_t)tile * CSV_TILE + (uint64_t)tid * CSV_PER_THREAD;
#pragma unroll
        for (int k = 0; k < CSV_PER_THREAD; ++k) ord[k] = 0xFFFFFFFFu;
        if (tile >= p.n_tiles || first >= p.n_sites) return;
        const uint32_t* src = p.order + p.site_begin + first;
        if (first + CSV_PER_THREAD <= p.n_sites && ((uintptr_t)src & 15u) == 0) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
            ord[0] = v.x; ord[1] = v.y; ord[2] = v.z; ord[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < CSV_PER_THREAD; ++k) if (first + k < p.n_sites) ord[k] = __ldg(src + k);
        }
    };
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);     // in order: the look-back below needs every earlier tile started
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= p.n_tiles) break;
        uint32_t ord[CSV_PER_THREAD];
        load_order(tile, ord);
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
                const uint32_t sl = (uint8_t)sfx[SUFFIX_BYTES - 1];
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
        uint32_t incl = mine, rincl = rows;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            const uint32_t r = __shfl_up_sync(0xFFFFFFFFu, rincl, d);
            if (lane >= d) { incl += o; rincl += r; }
        }
        if (lane == 31) { s_warp_sums[warp] = incl; s_warp_rows[warp] = rincl; }
        __syncthreads();
        uint32_t warp_off = 0, total = 0, total_rows = 0;
#pragma unroll
        for (int w = 0; w < CSV_THREADS / 32; ++w) {
            if (w < warp) warp_off += s_warp_sums[w];
            total += s_warp_sums[w];
            total_rows += s_warp_rows[w];
        }
        if (tid == 0) {
            uint64_t base = 0;
            if (tile == 0) {
                atomicExch(&p.status[0], LB_FLAG_PREFIX | (unsigned long long)total);
            } else {
                atomicExch(&p.status[tile], LB_FLAG_AGG | (unsigned long long)total);
                uint32_t i = tile - 1;
                for (;;) {
                    unsigned long long w;
                    unsigned int spins = 0;
                    do {
                        w = *((volatile unsigned long long*)&p.status[i]);
                        if ((w >> 62) == 0 && ++spins > (1u << 24)) w = LB_FLAG_PREFIX;
                    } while ((w >> 62) == 0);
                    base += w & LB_VALUE_MASK;
                    if (w & LB_FLAG_PREFIX) break;
                    --i;
                }
                atomicExch(&p.status[tile], LB_FLAG_PREFIX | (unsigned long long)(base + total));
            }
            s_base = base;
            if (tile == p.n_tiles - 1) *p.bytes_out = base + total;
            if (total_rows) atomicAdd(p.rows_out, (unsigned long long)total_rows);
        }
        __syncthreads();
        const uint64_t tile_base = s_base;
        const uint32_t local = warp_off + incl - mine;                 // byte offset of this thread's rows inside the tile
        if (tile_base + total > p.out_cap) continue;                   // the host reports SIDGPU_ECAPACITY from bytes_out
        const uint32_t mis = (uint32_t)((uintptr_t)(p.out + tile_base) & 15u);
        const bool staged = total + mis + 16 <= CSV_STAGE;
        uint32_t o = local;
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
            __syncthreads();
            char* g = p.out + tile_base;               // first byte of the tile in global memory
            // head: bytes up to the first 16-byte boundary; body: aligned vectors; tail: the rest
            const uint32_t head = mis ? min(16u - mis, total) : 0u;
            if ((uint32_t)tid < head) g[tid] = s_stage[mis + tid];
            const uint32_t body = (total - head) >> 4;
            const uint4* sv = reinterpret_cast<const uint4*>(s_stage + mis + head);    // 16-byte aligned by construction
            uint4* gv = reinterpret_cast<uint4*>(g + head);
            for (uint32_t i = tid; i < body; i += CSV_THREADS) gv[i] = sv[i];
            const uint32_t done = head + (body << 4);
            if (done + (uint32_t)tid < total) g[done + tid] = s_stage[mis + done + tid];
        }
    }
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
};

__global__ void k_records(const RecordParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_sites) return;
    const uint32_t s = p.slot[p.order[p.site_begin + i]];
    if (p.label) p.label[i] = p.table.label[s];
    if (p.gt) { p.gt[2 * i] = p.table.gt[2 * s]; p.gt[2 * i + 1] = p.table.gt[2 * s + 1]; }
    if (p.hom) p.hom[i] = p.table.hom[s];
    if (p.het) p.het[i] = p.table.het[s];
}

#endif  // __CUDACC__

}  // namespace sid
