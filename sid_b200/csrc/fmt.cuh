// Text formatting for the CSV rows of operator<< (call.hpp:29-38): `int` and `double` through
// iostream defaults, i.e. printf("%d") and printf("%g") (6 significant digits, glibc: correctly
// rounded, round-half-even on the exact binary value).  fmt_g6 is exact: the decimal digits come
// from big-integer arithmetic on the double's exact value, never from floating point.
#pragma once
#include "common.cuh"

namespace sid {

SID_HD int fmt_i32(int32_t v, char* out) {
    char tmp[12];
    int n = 0, len = 0;
    uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) out[len++] = '-';
    while (n) out[len++] = tmp[--n];
    return len;
}

SID_HD int digits_i32(int32_t v) {
    const uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    int n = v < 0 ? 2 : 1;
    n += u >= 10u;
    n += u >= 100u;
    n += u >= 1000u;
    n += u >= 10000u;
    n += u >= 100000u;
    n += u >= 1000000u;
    n += u >= 10000000u;
    n += u >= 100000000u;
    n += u >= 1000000000u;
    return n;
}

// Four decimal digits of x < 10000 as four bytes, most significant digit in the lowest byte
// (multiply-shift divisions: x/100 == x*5243>>19 for x < 43699, a/10 == a*205>>11 for a < 1029).
SID_HD uint32_t digits4(uint32_t x) {
    const uint32_t a = (x * 5243u) >> 19, b = x - a * 100u;
    const uint32_t d0 = (a * 205u) >> 11, d1 = a - d0 * 10u;
    const uint32_t d2 = (b * 205u) >> 11, d3 = b - d2 * 10u;
    return d0 | (d1 << 8) | (d2 << 16) | (d3 << 24);
}

// fmt_i32 without a division loop: the digits of |v| in three groups (2 + 4 + 4), then the
// significant ones are written out.  Same text as fmt_i32.
SID_HD int fmt_i32_fast(int32_t v, char* out) {
    const uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    const uint32_t top = u / 100000000u;                   // 0..42
    const uint32_t rest = u - top * 100000000u;
    const uint32_t mid = rest / 10000u, low = rest - mid * 10000u;
    const uint32_t g0 = digits4(top), g1 = digits4(mid), g2 = digits4(low);
    // twelve digit bytes d[0..11]: g0 g1 g2, text order
    const int nd = digits_i32(v < 0 ? (int32_t)(0u - u) : v) - (v < 0 ? 1 : 0);
    int len = 0;
    if (v < 0) out[len++] = '-';
    const int skip = 12 - nd;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const uint32_t g = k < 4 ? g0 : (k < 8 ? g1 : g2);
        if (k >= skip) out[len + k - skip] = (char)('0' + ((g >> (8 * (k & 3))) & 0xFFu));
    }
    return len + nd;
}

// round_half_even(m * 2^e2 * 10^j) for j >= 0, result known to be < 2^40.
SID_HD uint64_t scaled_round(uint64_t m, int e2, int j) {
    const int NL = 30;               // 53 + 2.33*345 bits < 30*32
    uint32_t w[NL];
    int len = 2;
    w[0] = (uint32_t)m;
    w[1] = (uint32_t)(m >> 32);
    for (int i = 2; i < NL; ++i) w[i] = 0;
    int rem = j;
    while (rem > 0) {                // multiply by 5^rem in chunks of 5^13 < 2^32
        int step = rem > 13 ? 13 : rem;
        uint32_t mul = 1;
        for (int i = 0; i < step; ++i) mul *= 5u;
        uint64_t carry = 0;
        for (int i = 0; i < len; ++i) {
            uint64_t t = (uint64_t)w[i] * mul + carry;
            w[i] = (uint32_t)t;
            carry = t >> 32;
        }
        if (carry && len < NL) w[len++] = (uint32_t)carry;
        rem -= step;
    }
    const int s = e2 + j;            // value = W * 2^s
    if (s >= 0) {                    // exact integer; small by contract
        uint64_t v = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
        return v << s;
    }
    const int k = -s;                // drop k bits with round-half-even
    const int limb = k >> 5, bit = k & 31;
    uint64_t lo = 0;
    for (int i = 0; i < 3; ++i) {
        int idx = limb + i;
        uint64_t part = idx < len ? (uint64_t)w[idx] : 0;
        if (i == 0) lo = part >> bit;
        else if (32 * i - bit < 64) lo |= part << (32 * i - bit);
    }
    // half bit is bit (k-1); sticky = any bit below it
    const int hk = k - 1;
    const int hl = hk >> 5, hb = hk & 31;
    const bool half = hl < len ? ((w[hl] >> hb) & 1u) != 0 : false;
    bool sticky = false;
    if (hl < len && (w[hl] & ((1u << hb) - 1u))) sticky = true;
    for (int i = 0; i < hl && i < len && !sticky; ++i) if (w[i]) sticky = true;
    if (half && (sticky || (lo & 1))) ++lo;
    return lo;
}

// 10^k, k = 0..300, as double-double (tools/gen_pow10_dd.py)
#if defined(__CUDA_ARCH__)
__device__ const double POW10_DD[301][2] = {
#include "pow10_dd.inc"
};
#else
static const double POW10_DD[301][2] = {
#include "pow10_dd.inc"
};
#endif

// The fast way to round_half_even(x * 10^k): the product in double-double arithmetic (x * hi error-free through an
// FMA, plus x * lo): about 104 correct bits of a value below 2^20, so the integer part and which side of one half the
// fraction lies on are certain unless the fraction is within 1e-9 of one half.  Then -- and for powers or values
// outside the table's comfortable range -- the caller takes the exact big-integer path.  Returns false for "not sure".
SID_HD bool scaled_round_fast(double x, int k, uint64_t& d) {
    if (k < 0 || k > 300 || !(x >= 1e-290) || !(x < 1e300)) return false;
    const double hi = POW10_DD[k][0], lo = POW10_DD[k][1];
    const double p = x * hi;
    if (!(p < 4.0e15)) return false;
#if defined(__CUDA_ARCH__)
    const double e = __fma_rn(x, hi, -p);
#else
    const double e = __builtin_fma(x, hi, -p);
#endif
    const double t = e + x * lo;
    const double yh = p + t;
    const double yl = t - (yh - p);                  // fast two-sum: |p| >= |t|
    double f = floor(yh);
    double s = (yh - f) + yl;                        // the fraction, in (-tiny, 1 + tiny)
    if (s < 0) { f -= 1.0; s += 1.0; }
    else if (s >= 1.0) { f += 1.0; s -= 1.0; }
    const double off = s - 0.5;
    if (off < 1e-9 && off > -1e-9) return false;     // too close to a tie to call
    d = (uint64_t)f + (off > 0 ? 1u : 0u);
    return true;
}

// printf("%g", x).  out needs 16 bytes; returns the length (no terminator written).
SID_HD int fmt_g6(double x, char* out) {
    const uint64_t bits = double_bits(x);
    const bool neg = (bits >> 63) != 0;
    const int be = (int)((bits >> 52) & 0x7FF);
    const uint64_t frac = bits & ((1ull << 52) - 1);
    int n = 0;
    if (neg) out[n++] = '-';
    if (be == 0x7FF) {
        if (frac) { out[n++] = 'n'; out[n++] = 'a'; out[n++] = 'n'; }
        else { out[n++] = 'i'; out[n++] = 'n'; out[n++] = 'f'; }
        return n;
    }
    if (be == 0 && frac == 0) { out[n++] = '0'; return n; }
    uint64_t m;
    int e2;
    if (be == 0) { m = frac; e2 = -1074; } else { m = frac | (1ull << 52); e2 = be - 1075; }
    const int l2 = 63 - clz64(m) + e2;                 // floor(log2 x)
    int e10 = (l2 * 78913) >> 18;                      // ~ floor(l2 * log10 2), at most one off
    uint64_t d = 0;
    if (e10 > 5) {
        // Values >= 1e6 never occur on this path (p-values and probabilities); keep the output
        // well-formed with ordinary floating point instead of big-integer division.
        double y = x < 0 ? -x : x;
        int e = 0;
        while (y >= 10.0) { y /= 10.0; ++e; }
        d = (uint64_t)(y * 100000.0 + 0.5);
        if (d >= 1000000) { d /= 10; ++e; }
        e10 = e;
    } else {
        const double ax = neg ? -x : x;
        for (int it = 0; it < 4; ++it) {
            if (!scaled_round_fast(ax, 5 - e10, d)) d = scaled_round(m, e2, 5 - e10);
            if (d >= 1000000) { ++e10; if (e10 > 5) { d = 100000; break; } continue; }
            if (d < 100000) { --e10; continue; }
            break;
        }
    }
    char dig[6];
    for (int i = 5; i >= 0; --i) { dig[i] = (char)('0' + d % 10); d /= 10; }
    int nd = 6;
    while (nd > 1 && dig[nd - 1] == '0') --nd;        // %g strips trailing zeros
    if (e10 < -4 || e10 >= 6) {
        out[n++] = dig[0];
        if (nd > 1) { out[n++] = '.'; for (int i = 1; i < nd; ++i) out[n++] = dig[i]; }
        out[n++] = 'e';
        int e = e10;
        if (e < 0) { out[n++] = '-'; e = -e; } else out[n++] = '+';
        if (e >= 100) { out[n++] = (char)('0' + e / 100); e %= 100; }
        out[n++] = (char)('0' + e / 10);
        out[n++] = (char)('0' + e % 10);
    } else if (e10 >= 0) {
        const int ip = e10 + 1;                        // digits before the decimal point
        for (int i = 0; i < ip; ++i) out[n++] = i < nd ? dig[i] : '0';
        if (nd > ip) { out[n++] = '.'; for (int i = ip; i < nd; ++i) out[n++] = dig[i]; }
    } else {
        out[n++] = '0';
        out[n++] = '.';
        for (int i = 0; i < -e10 - 1; ++i) out[n++] = '0';
        for (int i = 0; i < nd; ++i) out[n++] = dig[i];
    }
    return n;
}

}  // namespace sid
