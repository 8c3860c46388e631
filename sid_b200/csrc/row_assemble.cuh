// CSV rows assembled by the tokenizer itself (operator<< call.hpp:29-38 for `-m local`, fused into K1).
//
// A row is  name ',' position  +  ",hom,AA,0.000196638,1,p_value\n".  For a line of the fast grammar whose position
// is in canonical decimal form the first part is byte for byte the beginning of the pileup line with its first
// separator replaced by ',', so nothing is converted or formatted per site: the lane copies `hdr_len` bytes of
// staged text and the (<= 46 byte) suffix its profile's table slot carries.
//
// The 32 rows of a group are laid end to end in the warp's staging buffer.  A lane moves its row as 32-bit words
// (funnel shifts align source and destination), in two phases so that lanes never need each other's data:
//   phase A  every word that holds bytes of the row, except the first one when that word is shared with the row
//            before (a != 0).  The last word is written in full: its upper bytes belong to the next row and are
//            garbage for now.
//   phase B  (after __syncwarp) the lane's own bytes of the shared first word, byte-wise, and the ','.
// Rows of lines that took the byte-wise tokenizer are written byte-wise in phase B.
#pragma once
#include "common.cuh"
#include "fmt.cuh"
#include "parse_fast.cuh"

namespace sid {

SID_HD uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t shift_bits) {      // shift 0..32 (32 gives hi)
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(lo, hi, shift_bits);
#else
    return shift_bits >= 32 ? hi : (shift_bits ? (lo >> shift_bits) | (hi << (32 - shift_bits)) : lo);
#endif
}

constexpr int ROW_HDR_WORDS = 9;      // dest words a header of <= 30 bytes can touch (3 + 30 bytes)
constexpr int ROW_SFX_WORDS = 13;     // dest words a suffix of <= 46 bytes can touch (3 + 46 bytes)

struct RowSrc {
    uint32_t line_off;      // offset of the line's first byte in the staged text (>= 12)
    uint32_t hdr_len;       // bytes of "name<sep>position" (3..30)
    uint32_t name_len;
    uint32_t sfx[12];       // the slot's suffix record (48 bytes, text first)
    uint32_t sfx_len;       // 1..46
};

// Phase A.  `text`: staged text; `stage`: the warp's staging buffer (both 4-byte aligned); d: offset of the row in
// it.  hdr_words / sfx_words: warp-uniform upper bounds of the words any lane needs (the loops leave early together).
// Returns the shared first word for phase B.
SID_HD uint32_t row_phase_a(const uint8_t* text, uint8_t* stage, uint32_t d, const RowSrc& r, uint32_t hdr_words, uint32_t sfx_words) {
    const uint32_t* tw = reinterpret_cast<const uint32_t*>(text);
    uint32_t* sw = reinterpret_cast<uint32_t*>(stage);
    const uint32_t a = d & 3u, dw = d >> 2;
    // ---- header: dest word q holds the text bytes [line_off - a + 4q, +4)
    const uint32_t base = r.line_off - a;
    const uint32_t* tp = tw + (base >> 2);
    const uint32_t sh = (base & 3u) * 8u;
    const uint32_t nh = (a + r.hdr_len + 3u) >> 2;          // words that hold header bytes
    uint32_t prev = tp[0], first = 0;
#pragma unroll
    for (int q = 0; q < ROW_HDR_WORDS; ++q) {
        if ((uint32_t)q >= hdr_words) break;
        const uint32_t next = tp[q + 1];
        const uint32_t v = funnel_r(prev, next, sh);
        prev = next;
        if (q == 0) first = v;
        if ((uint32_t)q < nh && (q > 0 || a == 0)) sw[dw + q] = v;
    }
    // ---- suffix, from byte d + hdr_len on; the word at the junction takes its low bytes from the end of the header
    const uint32_t ds = d + r.hdr_len, a2 = ds & 3u, dw2 = ds >> 2;
    const uint32_t xb = r.line_off + r.hdr_len - 4u;
    const uint32_t x = funnel_r(tw[xb >> 2], tw[(xb >> 2) + 1], (xb & 3u) * 8u);          // header bytes [hdr_len - 4, hdr_len)
    const uint32_t sh2 = 8u * (4u - a2);
    const uint32_t ns = (a2 + r.sfx_len + 3u) >> 2;
    uint32_t lo = x;
#pragma unroll
    for (int q = 0; q < ROW_SFX_WORDS; ++q) {
        if ((uint32_t)q >= sfx_words) break;
        const uint32_t hi = q < 12 ? r.sfx[q < 12 ? q : 0] : 0u;       // q is a constant after unrolling
        const uint32_t v = funnel_rc(lo, hi, sh2);
        lo = hi;
        if ((uint32_t)q < ns) sw[dw2 + q] = v;
    }
    return first;
}

// Phase B for a row moved by row_phase_a.
SID_HD void row_phase_b(uint8_t* stage, uint32_t d, const RowSrc& r, uint32_t first) {
    const uint32_t a = d & 3u;
    uint8_t* w0 = stage + (d & ~3u);
    if (a) {
#pragma unroll
        for (int k = 1; k < 4; ++k)
            if ((uint32_t)k >= a) w0[k] = (uint8_t)(first >> (8 * k));
    }
    stage[d + r.name_len] = (uint8_t)',';
}

// A whole row byte by byte (lines of the byte-wise tokenizer; any name length): name, ',', printf("%d") of the
// position, suffix (`sfx`: the slot's record in the table).
template <class Src>
SID_HD void row_bytewise(uint8_t* out, const Src& src, uint64_t name_abs, uint32_t name_len, int32_t pos, const uint8_t* sfx, uint32_t sfx_len) {
    for (uint32_t i = 0; i < name_len; ++i) out[i] = src.at(name_abs + i);
    out += name_len;
    *out++ = (uint8_t)',';
    char digits[12];
    const int nd = fmt_i32(pos, digits);
    for (int i = 0; i < nd; ++i) out[i] = (uint8_t)digits[i];
    out += nd;
    for (uint32_t i = 0; i < sfx_len; ++i) out[i] = sfx[i];
}

}  // namespace sid
