// Small bit helpers shared by the tokenizers (funnel shift, population count, first set bit) and the
// result record of the fast grammar.
#pragma once
#include "common.cuh"
#include "parse.cuh"

namespace sid {
struct FastLine {
    int status;
    int32_t pos;
    uint64_t profile;
    uint32_t chrom_off, chrom_len;
};

SID_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t shift_bits) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift_bits);
#else
    return shift_bits ? (lo >> shift_bits) | (hi << (32 - shift_bits)) : lo;
#endif
}

SID_HD uint32_t pop_count(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
SID_HD uint32_t first_bit(uint32_t x) {        // index of the lowest set bit; 32 when x == 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)__brev(x));                 // 32 for x == 0
#else
    return x ? (uint32_t)__builtin_ctz(x) : 32u;
#endif
}

#if defined(__CUDA_ARCH__)
#define SID_SYNCWARP() __syncwarp()
#else
#define SID_SYNCWARP() ((void)0)
#endif


}  // namespace sid
