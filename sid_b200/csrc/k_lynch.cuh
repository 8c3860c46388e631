// K3: unique-profile histogram   countUniqueProfiles pileup.cpp:169-196 (compaction + lexicographic
//     order), computeNucleotideDistribution pileup.cpp:198-217.
// K4: Lynch objective            compoundLikelihood lynch.cpp:37-61 with lynch.hpp:57-74,82-90.
// K5: Benjamini-Hochberg         adjustBenjaminiHochberg stats.cpp:58-80.
#pragma once
#include "calls.cuh"
#include "common.cuh"
#include "nelder_mead.hpp"
#include "table.cuh"

namespace sid {

#if defined(__CUDACC__)

// ---- generic helpers ---------------------------------------------------------------------------

// Order on (key, value) pairs: the value breaks key ties, so padding entries (all-ones key AND
// all-ones value) sort strictly after every real entry -- also after a real all-ones key (the
// profile {65535,65535,65535,65535}; a p-value of exactly 0 in the BH keys).
__device__ __forceinline__ bool pair_greater(unsigned long long ka, uint32_t va, unsigned long long kb, uint32_t vb) {
    return ka > kb || (ka == kb && va > vb);
}

// One compare-exchange step of a bitonic sorting network over (key, value) pairs, ascending.
__global__ void k_bitonic_step(unsigned long long* keys, uint32_t* vals, uint32_t n, uint32_t j, uint32_t k) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t ixj = i ^ j;
    if (ixj <= i || ixj >= n) return;
    const unsigned long long a = keys[i], b = keys[ixj];
    const uint32_t va = vals[i], vb = vals[ixj];
    const bool up = (i & k) == 0;
    if (pair_greater(a, va, b, vb) == up) {
        keys[i] = b;
        keys[ixj] = a;
        vals[i] = vb;
        vals[ixj] = va;
    }
}

// All steps with j < BITONIC_BLOCK for one k, done in shared memory (one launch instead of log2 j).
constexpr int BITONIC_BLOCK = 2048;
__global__ void __launch_bounds__(BITONIC_BLOCK / 2) k_bitonic_local(unsigned long long* keys, uint32_t* vals, uint32_t n,
                                                                      uint32_t j_start, uint32_t k, int all_k) {
    __shared__ unsigned long long sk[BITONIC_BLOCK];
    __shared__ uint32_t sv[BITONIC_BLOCK];
    const uint32_t base = blockIdx.x * BITONIC_BLOCK;
    for (uint32_t t = threadIdx.x; t < BITONIC_BLOCK; t += blockDim.x) {
        const uint32_t i = base + t;
        sk[t] = i < n ? keys[i] : 0xFFFFFFFFFFFFFFFFull;
        sv[t] = i < n ? vals[i] : 0xFFFFFFFFu;
    }
    __syncthreads();
    // all_k: run the whole network for k = 2 .. BITONIC_BLOCK (first phase); else only the tail of one k
    for (uint32_t kk = all_k ? 2 : k; kk <= (all_k ? (uint32_t)BITONIC_BLOCK : k); kk <<= 1) {
        for (uint32_t j = all_k ? kk >> 1 : j_start; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < BITONIC_BLOCK; t += blockDim.x) {
                const uint32_t x = t ^ j;
                if (x > t) {
                    const bool up = ((base + t) & kk) == 0;
                    const unsigned long long a = sk[t], b = sk[x];
                    const uint32_t va = sv[t], vb = sv[x];
                    if (pair_greater(a, va, b, vb) == up) {
                        sk[t] = b; sk[x] = a;
                        sv[t] = vb; sv[x] = va;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t t = threadIdx.x; t < BITONIC_BLOCK; t += blockDim.x) {
        const uint32_t i = base + t;
        if (i < n) { keys[i] = sk[t]; vals[i] = sv[t]; }
    }
}

// ---- K3 ----------------------------------------------------------------------------------------

// Select the table entries with coverage >= min_cov and a non-zero count; emit sort keys.
__global__ void k_hist_select(TableView t, uint32_t n_entries, uint32_t min_cov, unsigned long long* keys, uint32_t* vals,
                              unsigned int* n_out) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const uint32_t slot = t.entry_list[e];
    const uint64_t prof = t.keys[slot];
    if (profile_coverage(prof) < min_cov || t.counts[slot] == 0) return;
    const uint32_t o = atomicAdd(n_out, 1u);
    keys[o] = profile_sort_key(prof);
    vals[o] = e;
}

__global__ void k_fill_pad(unsigned long long* keys, uint32_t* vals, uint32_t from, uint32_t to) {
    const uint32_t i = from + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < to) { keys[i] = 0xFFFFFFFFFFFFFFFFull; vals[i] = 0xFFFFFFFFu; }
}

// After sorting: materialise the unique table in lexicographic order + per-entry back map.
__global__ void k_hist_gather(TableView t, const uint32_t* sorted_entry, uint32_t n_unique, unsigned long long* u_profile,
                              unsigned long long* u_count, double* u_logM, uint32_t* entry_to_unique,
                              unsigned long long* nd_acc /* [5]: A C G T total */) {
    __shared__ unsigned long long s_acc[5];
    if (threadIdx.x < 5) s_acc[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_unique) {
        const uint32_t e = sorted_entry[u];
        const uint32_t slot = t.entry_list[e];
        const uint64_t prof = t.keys[slot];
        const unsigned long long cnt = t.counts[slot];
        u_profile[u] = prof;
        u_count[u] = cnt;
        u_logM[u] = log_multinomial(prof);
        entry_to_unique[e] = u;
        // computeNucleotideDistribution (pileup.cpp:198-217) in 64-bit (the reference's uint32
        // product count*coverage wraps above 2^32)
        for (int i = 0; i < 4; ++i) atomicAdd(&s_acc[i], cnt * (unsigned long long)profile_count(prof, i));
        atomicAdd(&s_acc[4], cnt * (unsigned long long)profile_coverage(prof));
    }
    __syncthreads();
    if (threadIdx.x < 5 && s_acc[threadIdx.x]) atomicAdd(&nd_acc[threadIdx.x], s_acc[threadIdx.x]);
}

__global__ void k_fill_u32(uint32_t* p, uint32_t n, uint32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- K4 ----------------------------------------------------------------------------------------
constexpr int OBJ_THREADS = 256;

struct ObjParams {
    const unsigned long long* u_profile;
    const unsigned long long* u_count;
    const double* u_logM;
    uint32_t n_unique;
    LynchConsts k;
    double log1m_pi, log_pi;
    double* partials;          // 2 per block (sum, compensation)
    unsigned int* done_blocks;
    double* out;               // -log likelihood (local sum)
};

__global__ void __launch_bounds__(OBJ_THREADS) k_lynch_objective(const ObjParams p) {
    __shared__ double s_s[OBJ_THREADS], s_c[OBJ_THREADS];
    __shared__ bool s_last;
    CompSum acc;
    acc.init();
    for (uint32_t u = blockIdx.x * OBJ_THREADS + threadIdx.x; u < p.n_unique; u += gridDim.x * OBJ_THREADS) {
        double t;
        if (lynch_term(p.u_profile[u], p.u_logM[u], p.k, p.log1m_pi, p.log_pi, t)) acc.add(t * (double)p.u_count[u]);
    }
    s_s[threadIdx.x] = acc.s;
    s_c[threadIdx.x] = acc.c;
    __syncthreads();
    // fixed-order tree: the result does not depend on scheduling
    for (int d = OBJ_THREADS / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) {
            CompSum a;
            a.s = s_s[threadIdx.x]; a.c = s_c[threadIdx.x];
            a.add(s_s[threadIdx.x + d]);
            a.c += s_c[threadIdx.x + d];
            s_s[threadIdx.x] = a.s; s_c[threadIdx.x] = a.c;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.partials[2 * blockIdx.x] = s_s[0];
        p.partials[2 * blockIdx.x + 1] = s_c[0];
        __threadfence();
        s_last = atomicAdd(p.done_blocks, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        CompSum a;
        a.init();
        for (uint32_t b = 0; b < gridDim.x; ++b) {
            a.add(__ldcg(&p.partials[2 * b]));
            a.c += __ldcg(&p.partials[2 * b + 1]);
        }
        *p.out = -a.value();
        *p.done_blocks = 0;
    }
}

// ---- K4, whole fit in one launch -----------------------------------------------------------------
// estimateProfileGenotypeLikelihoods (lynch.cpp:17-35) + FunctionMinimizer<2>::run (optimization.hpp:51-89) as ONE
// cooperative kernel: every thread walks the same Nelder-Mead trajectory (nelder_mead.hpp); an objective evaluation
// (compoundLikelihood, lynch.cpp:37-61) is a grid-wide reduction: per-thread compensated sums over the histogram,
// a fixed-order tree per block, one grid barrier, then every block adds the block partials in the same order.
// No host round trip per evaluation; the result depends only on the histogram (its order included), so every rank
// that holds the same merged histogram gets bit-identical (pi, eps).
struct FitParams {
    const unsigned long long* u_profile;
    const unsigned long long* u_count;
    const double* u_logM;
    uint32_t n_unique;
    double nd[4];
    double x0[2], step[2];
    double* partials;          // 2 buffers x 2 doubles per block
    unsigned int* barrier;     // monotonic arrival counter, zero at launch
    double* out;               // pi, eps, fval, iterations, evaluations, converged (as doubles)
    unsigned long long* error;
};

struct FitEval {
    const FitParams& p;
    double* s_s;
    double* s_c;
    double* s_val;
    uint32_t epoch;
    bool dead;
    __device__ double operator()(double pi, double eps) {
        if (pi < 0 || pi > 1 || eps < 0 || eps > 1) return 1.7976931348623157e308;      // lynch.cpp:41-43
        const LynchConsts k = lynch_consts(p.nd, eps);
        const double log1m_pi = log1p(-pi), log_pi = log(pi);
        CompSum acc;
        acc.init();
        for (uint32_t u = blockIdx.x * OBJ_THREADS + threadIdx.x; u < p.n_unique; u += gridDim.x * OBJ_THREADS) {
            double t;
            if (lynch_term(p.u_profile[u], p.u_logM[u], k, log1m_pi, log_pi, t)) acc.add(t * (double)p.u_count[u]);
        }
        s_s[threadIdx.x] = acc.s;
        s_c[threadIdx.x] = acc.c;
        __syncthreads();
        for (int d = OBJ_THREADS / 2; d > 0; d >>= 1) {
            if ((int)threadIdx.x < d) {
                CompSum a;
                a.s = s_s[threadIdx.x]; a.c = s_c[threadIdx.x];
                a.add(s_s[threadIdx.x + d]);
                a.c += s_c[threadIdx.x + d];
                s_s[threadIdx.x] = a.s; s_c[threadIdx.x] = a.c;
            }
            __syncthreads();
        }
        double* buf = p.partials + (size_t)(epoch & 1u) * 2 * gridDim.x;
        if (threadIdx.x == 0) {
            buf[2 * blockIdx.x] = s_s[0];
            buf[2 * blockIdx.x + 1] = s_c[0];
            __threadfence();
            atomicAdd(p.barrier, 1u);
            // grid barrier: all blocks are resident (cooperative launch); bounded like every wait of this library
            const unsigned int target = (epoch + 1u) * gridDim.x;
            uint32_t spins = 0;
            while (*((volatile unsigned int*)p.barrier) < target) {
                if (++spins > (1u << 26)) { dead = true; break; }
            }
            __threadfence();
            *s_val = dead ? 1.0 : 0.0;
        }
        __syncthreads();
        const bool lost = *s_val != 0.0;
        __syncthreads();
        if (threadIdx.x < 32) {
            // the block partials, read by the 32 lanes side by side (one thread walking them pays an L2 round trip
            // each: measured 13 us per evaluation at 28 blocks); lane i adds partials i, i + 32, ... in that order,
            // then a fixed tree over the lanes: the same sum in every block, on every rank
            CompSum a;
            a.init();
            for (uint32_t b = threadIdx.x; b < gridDim.x; b += 32) {
                a.add(__ldcg(&buf[2 * b]));
                a.c += __ldcg(&buf[2 * b + 1]);
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const double os = __shfl_down_sync(0xFFFFFFFFu, a.s, d), oc = __shfl_down_sync(0xFFFFFFFFu, a.c, d);
                a.add(os);
                a.c += oc;
            }
            if (threadIdx.x == 0) *s_val = lost ? 1.7976931348623157e308 : -a.value();
        }
        __syncthreads();
        const double v = *s_val;
        __syncthreads();
        ++epoch;
        return v;
    }
};

__global__ void __launch_bounds__(OBJ_THREADS) k_lynch_fit(const FitParams p) {
    __shared__ double s_s[OBJ_THREADS], s_c[OBJ_THREADS];
    __shared__ double s_val;
    FitEval f {p, s_s, s_c, &s_val, 0u, false};
    const NelderMeadResult r = nelder_mead_2d(f, p.x0, p.step);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.out[0] = r.x[0];
        p.out[1] = r.x[1];
        p.out[2] = r.fval;
        p.out[3] = (double)r.iterations;
        p.out[4] = (double)r.evaluations;
        p.out[5] = r.converged ? 1.0 : 0.0;
        if (f.dead) atomicMin(p.error, (unsigned long long)(LINE_MALFORMED + 4));
    }
}

// Per-unique-profile p-values for likelihood_ratio before BH (call.cpp:93-103).
__global__ void k_lr_pvalues(const unsigned long long* u_profile, uint32_t n, LynchConsts k, int use_prior, double pi,
                             double* p_hom, double* p_het) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    lr_pvalues(u_profile[u], k, use_prior != 0, pi, p_hom[u], p_het[u]);
}

// ---- K5 ----------------------------------------------------------------------------------------
// adjusted[s_i] = min(1, min_{j<=i} p[s_j] * m / (m - j)) with s the descending order of p.
__global__ void k_bh_keys(const double* p, uint32_t n, unsigned long long* keys, uint32_t* vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // descending order of non-negative doubles == ascending order of the complemented bit pattern
    keys[i] = ~double_bits(p[i]);
    vals[i] = i;
}

constexpr int BH_THREADS = 256;
// pass 1: block-local inclusive running minimum of c_i, block minima out
__global__ void __launch_bounds__(BH_THREADS) k_bh_scan1(const double* p, const uint32_t* order, uint32_t n, double* c,
                                                         double* block_min) {
    __shared__ double s[BH_THREADS];
    const uint32_t i = blockIdx.x * BH_THREADS + threadIdx.x;
    double v = bits_double(0x7FF0000000000000ull);
    if (i < n) v = i == 0 ? p[order[0]]                              // stats.cpp:71
                          : p[order[i]] * (double)n / (double)(n - i);   // stats.cpp:73
    s[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < BH_THREADS; d <<= 1) {
        double o = (int)threadIdx.x >= d ? s[threadIdx.x - d] : bits_double(0x7FF0000000000000ull);
        __syncthreads();
        if (o < s[threadIdx.x]) s[threadIdx.x] = o;
        __syncthreads();
    }
    if (i < n) c[i] = s[threadIdx.x];
    if (threadIdx.x == BH_THREADS - 1) block_min[blockIdx.x] = s[threadIdx.x];
}
// pass 2: exclusive running minimum over the block minima (single thread: n / 256 values)
__global__ void k_bh_scan2(double* block_min, uint32_t n_blocks) {
    if (blockIdx.x || threadIdx.x) return;
    double run = bits_double(0x7FF0000000000000ull);
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const double v = block_min[b];
        block_min[b] = run;
        if (v < run) run = v;
    }
}
// pass 3: combine, clamp to 1 (stats.cpp:76-78), scatter back to profile order
__global__ void __launch_bounds__(BH_THREADS) k_bh_scatter(const double* c, const double* block_min, const uint32_t* order,
                                                           uint32_t n, double* adjusted) {
    const uint32_t i = blockIdx.x * BH_THREADS + threadIdx.x;
    if (i >= n) return;
    double v = c[i];
    const double b = block_min[blockIdx.x];
    if (b < v) v = b;
    if (v > 1.0) v = 1.0;
    adjusted[order[i]] = v;
}

__global__ void k_format_g(const double* v, uint64_t n, char* out16) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    char buf[16];
    const int len = fmt_g6(v[i], buf);
    for (int k = 0; k < 16; ++k) out16[16 * i + k] = k < len ? buf[k] : 0;
}

#endif  // __CUDACC__

}  // namespace sid
