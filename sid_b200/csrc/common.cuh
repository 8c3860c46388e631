// Shared definitions for the sm_100a kernels.  Everything marked SID_HD also compiles as plain
// host C++ so that tests/hostcheck can exercise the exact same arithmetic on the CPU build box
// (test infrastructure only: the product library exports none of it as a CPU path).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SID_HD __host__ __device__ __forceinline__
#define SID_D __device__ __forceinline__
#else
#define SID_HD inline
#define SID_D inline
#endif

namespace sid {

// Status codes a parsed line can carry (mapped to SIDGPU_E* by the host side).
enum LineStatus : int {
    LINE_OK = 0,
    LINE_MALFORMED = 1,      // pileup.cpp:9
    LINE_MISSING_MAPQ = 2,   // pileup.cpp:10
    LINE_QUAL_SHORT = 3      // call.cpp:330-331 would read past the quality vectors
};

// profile_t (pileup.hpp:7) packed little endian: A | C<<16 | G<<32 | T<<48.
SID_HD uint64_t pack_profile(uint32_t a, uint32_t c, uint32_t g, uint32_t t) {
    return (uint64_t)(a & 0xFFFFu) | ((uint64_t)(c & 0xFFFFu) << 16) | ((uint64_t)(g & 0xFFFFu) << 32) |
           ((uint64_t)(t & 0xFFFFu) << 48);
}
SID_HD uint32_t profile_count(uint64_t p, int i) { return (uint32_t)(p >> (16 * i)) & 0xFFFFu; }
// UniqueProfile::coverage (pileup.hpp:37-39): plain sum of the four (already wrapped) counts.
SID_HD uint32_t profile_coverage(uint64_t p) {
    return profile_count(p, 0) + profile_count(p, 1) + profile_count(p, 2) + profile_count(p, 3);
}
// Lexicographic order of std::array<uint16_t,4> (pileup.cpp:179-182) as a single 64-bit compare key.
SID_HD uint64_t profile_sort_key(uint64_t p) {
    return ((p & 0xFFFFull) << 48) | ((p & 0xFFFF0000ull) << 16) | ((p >> 16) & 0xFFFF0000ull) | (p >> 48);
}

SID_HD uint64_t double_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return b;
#endif
}
SID_HD double bits_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

SID_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}

SID_HD int ctz64(uint64_t x) {              // x != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

SID_HD uint64_t mix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xFF51AFD7ED558CCDull;
    k ^= k >> 33;
    k *= 0xC4CEB9FE1A85EC53ull;
    k ^= k >> 33;
    return k;
}

}  // namespace sid
