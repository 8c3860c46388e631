// Bit-parallel form of the tokenizer.
//
// Stage 1 (flat, no per-line work): 32 bytes of text are transposed into their 8 bit planes (byte
// permutes, then an 8x8 bit-matrix transpose across eight words: three rounds of block swaps), and every
// character class the grammar of parseReadBases (pileup.cpp:70-153) distinguishes becomes a boolean
// function of the planes: one 32-bit word per class and 32 bytes of text.  The words go to bit
// arrays in shared memory (bit i <-> byte i of the warp's slice).
//
// Stage 2 (one line per lane): the header separators come from the TERM bit array, the bases field
// is walked 32 bytes per step: '^' masks its successor by a shift of the CARET word, the counts of
// A/C/G/T/./, are population counts of class words under the step's mask, an indel reads its length
// from the text and restarts the walk after the number with that many bytes to ignore.
//
// Same contract as parse_fast.cuh: a line outside the fast grammar is REFUSED (returns false) and
// the caller re-parses it with the byte-wise state machine of parse.cuh.
#pragma once
#include "common.cuh"
#include "parse.cuh"
#include "parse_fast.cuh"

namespace sid {

SID_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, s);
#else
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

// Three-input logic with the truth table spelled out (a = 0xF0, b = 0xCC, c = 0xAA): on the device one
// LOP3 per call.  Stage 1 is bound by the integer ALU pipe, so the number of LOP3s is counted by hand.
constexpr int TA = 0xF0, TB = 0xCC, TC = 0xAA;
template <int LUT>
SID_HD uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT & 0xFF));
    return d;
#else
    uint32_t r = 0;
    for (int i = 0; i < 8; ++i)
        if ((LUT >> i) & 1) r |= ((i & 4) ? a : ~a) & ((i & 2) ? b : ~b) & ((i & 1) ? c : ~c);
    return r;
#endif
}

// x >> k for a constant k.  (The high half of x * 2^(32-k) would run on the FMA pipe instead of the
// saturated ALU pipe, but IMAD.HI is slower: 1.256 vs 1.236 ms per 1.63 GB.  -DSID_SHR_FMA to try again.)
template <int K>
SID_HD uint32_t shr_fma(uint32_t x) {
#if defined(__CUDA_ARCH__) && defined(SID_SHR_FMA)
    return __umulhi(x, 1u << (32 - K));
#else
    return x >> K;
#endif
}

// One step of an 8x8 bit-matrix transpose whose rows are eight words: rows lo and hi exchange the bit
// blocks selected by M / S (two LOP3 selects, one shift each way).
template <uint32_t M, int S>
SID_HD void swap_blocks(uint32_t& lo, uint32_t& hi) {
    constexpr int SELECT = (TA & TC) | (TB & ~TC);                     // c ? a : b, bit by bit
    const uint32_t nl = lop3<SELECT>(lo, hi << S, M);
    const uint32_t nh = lop3<SELECT>(shr_fma<S>(lo), hi, M);
    lo = nl;
    hi = nh;
}

struct ClassWords {     // bit i of each word <-> byte i of the 32-byte unit
    uint32_t term;      // byte <= 0x20: field separators, line end, NUL and the other control bytes
    uint32_t nl;        // '\n'
    uint32_t a, c, g, t;    // A/a C/c G/g T/t
    uint32_t dot;       // '.' or ','
    uint32_t caret;     // '^'
    uint32_t pm;        // '+' or '-'
    uint32_t high;      // byte >= 0x80
};

// w[0..7]: the unit's 32 bytes as little-endian words.
SID_HD ClassWords classify32(const uint32_t w[8]) {
    // bytes: n[i] = text bytes (i, 8 + i, 16 + i, 24 + i), so that after the bit transpose across the eight
    // words byte lane k of plane j holds bit j of bytes 8k .. 8k+7 in order
    uint32_t n[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t ab01 = byte_perm(w[h], w[2 + h], 0x5140u), ab23 = byte_perm(w[h], w[2 + h], 0x7362u);
        const uint32_t cd01 = byte_perm(w[4 + h], w[6 + h], 0x5140u), cd23 = byte_perm(w[4 + h], w[6 + h], 0x7362u);
        n[4 * h + 0] = byte_perm(ab01, cd01, 0x5410u);
        n[4 * h + 1] = byte_perm(ab01, cd01, 0x7632u);
        n[4 * h + 2] = byte_perm(ab23, cd23, 0x5410u);
        n[4 * h + 3] = byte_perm(ab23, cd23, 0x7632u);
    }
    // bits: 8x8 transpose (rows = words, columns = bit in byte), all four byte lanes at once
#pragma unroll
    for (int i = 0; i < 4; ++i) swap_blocks<0x0F0F0F0Fu, 4>(n[i], n[i + 4]);
    swap_blocks<0x33333333u, 2>(n[0], n[2]);
    swap_blocks<0x33333333u, 2>(n[1], n[3]);
    swap_blocks<0x33333333u, 2>(n[4], n[6]);
    swap_blocks<0x33333333u, 2>(n[5], n[7]);
#pragma unroll
    for (int i = 0; i < 8; i += 2) swap_blocks<0x55555555u, 1>(n[i], n[i + 1]);
    const uint32_t p0 = n[0], p1 = n[1], p2 = n[2], p3 = n[3], p4 = n[4], p5 = n[5], p6 = n[6], p7 = n[7];
    ClassWords k;
    k.high = p7;
    // high nibbles (p7 p6 p5 p4)
    const uint32_t h00 = lop3<~TA & ~TB & ~TC>(p7, p6, p5);            // 000x: 0x00..0x1f
    const uint32_t h2 = lop3<~TA & ~TB & TC>(p7, p6, p5) & ~p4;        // 0010: 0x2_
    const uint32_t h46 = lop3<~TA & TB & ~TC>(p7, p6, p4);             // 01x0: 0x4_ 0x6_
    const uint32_t h57 = lop3<~TA & TB & TC>(p7, p6, p4);              // 01x1: 0x5_ 0x7_
    // low nibbles (p3 p2 p1 p0), three planes first, then p0 together with the high-nibble term
    const uint32_t l000 = lop3<~TA & ~TB & ~TC>(p3, p2, p1);
    const uint32_t l001 = lop3<~TA & ~TB & TC>(p3, p2, p1);
    const uint32_t l011 = lop3<~TA & TB & TC>(p3, p2, p1);
    const uint32_t l010 = lop3<~TA & TB & ~TC>(p3, p2, p1);
    const uint32_t l111 = lop3<TA & TB & TC>(p3, p2, p1);
    const uint32_t l101 = lop3<TA & ~TB & TC>(p3, p2, p1);
    const uint32_t l1x = lop3<TA & (TB ^ TC)>(p3, p2, p1);             // 101_ or 110_
    k.a = lop3<TA & TB & TC>(l000, p0, h46);                           // 0x41 0x61
    k.c = lop3<TA & TB & TC>(l001, p0, h46);                           // 0x43 0x63
    k.g = lop3<TA & TB & TC>(l011, p0, h46);                           // 0x47 0x67
    k.t = lop3<TA & ~TB & TC>(l010, p0, h57);                          // 0x54 0x74
    k.dot = lop3<TA & ~TB & TC>(p3 & p2, p0, h2);                      // 0x2c 0x2e
    k.pm = lop3<TA & TB & TC>(l1x, p0, h2);                            // 0x2b 0x2d
    k.caret = lop3<TA & ~TB & TC>(l111, p0, h57) & ~p5;                // 0x5e
    k.nl = lop3<TA & ~TB & TC>(l101, p0, h00) & ~p4;                   // 0x0a
    k.term = h00 | lop3<TA & ~TB & TC>(l000, p0, h2);                  // 0x00..0x1f, 0x20
    return k;
}

// The class bit arrays of one region of staged text (bit i <-> byte region_off + i).  Each array
// holds n_words valid words followed by at least one padding word.
struct BitArrays {
    const uint32_t* term;
    const uint32_t* a;
    const uint32_t* c;
    const uint32_t* g;
    const uint32_t* t;
    const uint32_t* dot;
    const uint32_t* caret;
    const uint32_t* pm;
    uint32_t n_bits;        // classified bytes
};

SID_HD uint32_t bits32(const uint32_t* arr, uint32_t pos) {        // 32 bits starting at bit `pos`
    const uint32_t k = pos >> 5;
    return funnel_r(arr[k], arr[k + 1], pos & 31);
}

// `s`: staged text (4-byte aligned, `avail` bytes, byte 0 at absolute offset abs0); the bit arrays
// cover the bytes from offset region_off of s.  Same result contract as parse_line_fast_smem.
SID_HD bool parse_line_bits(const uint8_t* s, uint32_t avail, uint32_t region_off, const BitArrays& B, uint32_t line_off,
                            FastLine& o) {
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s);
    const uint32_t ls = line_off - region_off;                      // bit index of the line's first byte
    bool ok = line_off >= region_off && ls + 96 <= B.n_bits && line_off + 64 <= avail && line_off >= 12;
    const uint32_t l0 = ok ? ls : 0, h0 = ok ? line_off : 16;
    // ---- header: the bytes <= 0x20 among the first 32 locate the four separators
    const uint32_t sepmask = bits32(B.term, l0);
    ok = ok && pop_count(sepmask) >= 4;
    uint32_t m = sepmask;
    const uint32_t p1 = first_bit(m); m &= m - 1;
    const uint32_t p2 = first_bit(m); m &= m - 1;
    const uint32_t p3 = first_bit(m); m &= m - 1;
    const uint32_t p4 = first_bit(m);
    const uint32_t nd = p2 - p1 - 1;
    // chrom non-empty, 1..9 digits, one reference character, depth non-empty, bases non-empty within reach
    ok = ok && p1 >= 1 && nd >= 1 && nd <= 9 && p3 == p2 + 2 && p4 > p3 + 1 && p4 <= 30 && ((sepmask >> (p4 + 1)) & 1u) == 0;
    const uint32_t q1 = ok ? p1 : 1, q2 = ok ? p2 : 3, q3 = ok ? p3 : 5, q4 = ok ? p4 : 7;
    {
        const uint32_t c1 = s[h0 + q1], c2 = s[h0 + q2], c3 = s[h0 + q3], c4 = s[h0 + q4];
        ok = ok && (c1 == '\t' || c1 == ' ') && (c2 == '\t' || c2 == ' ') && (c3 == '\t' || c3 == ' ') && (c4 == '\t' || c4 == ' ');
    }
    o.chrom_off = 0;
    o.chrom_len = q1;
    const uint32_t ref = s[h0 + q2 + 1];
    ok = ok && !ref_is_control((uint8_t)ref);               // '.' / ',' would become '^', '+' or '-' (pileup.cpp:78-83)
    // ---- position: the (up to) eight characters before the second separator, leading ones forced to '0'
    uint32_t acc;
    {
        const uint32_t e = h0 + q2;
        const uint32_t ndd = ok ? nd : 1;
        const uint32_t* pw = sw + ((e - 8) >> 2);
        const uint32_t ps = ((e - 8) & 3) * 8;
        const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2];
        uint32_t lo = funnel_r(w0, w1, ps), hi = funnel_r(w1, w2, ps);
        const uint32_t zero = ndd >= 8 ? 0u : 8u - ndd;
        if (zero >= 4) {
            lo = 0x30303030u;
            const uint32_t mz = zero == 4 ? 0u : ((1u << (8 * (zero - 4))) - 1u);
            hi = (hi & ~mz) | (0x30303030u & mz);
        } else if (zero) {
            const uint32_t mz = (1u << (8 * zero)) - 1u;
            lo = (lo & ~mz) | (0x30303030u & mz);
        }
        const bool dig = ((lo & 0xF0F0F0F0u) == 0x30303030u) && ((hi & 0xF0F0F0F0u) == 0x30303030u) &&
                         ((((lo & 0x0F0F0F0Fu) + 0x06060606u) | ((hi & 0x0F0F0F0Fu) + 0x06060606u)) & 0x10101010u) == 0;
        ok = ok && dig;
        const uint32_t xl = lo & 0x0F0F0F0Fu, xh = hi & 0x0F0F0F0Fu;
        const uint32_t tl = xl * 10u + (xl >> 8), th = xh * 10u + (xh >> 8);
        const uint32_t vl = (tl & 0xFFu) * 100u + ((tl >> 16) & 0xFFu), vh = (th & 0xFFu) * 100u + ((th >> 16) & 0xFFu);
        acc = vl * 10000u + vh;
        if (ndd == 9) {
            const uint32_t d9 = (uint32_t)s[e - 9] - (uint32_t)'0';
            ok = ok && d9 <= 9;
            acc += d9 * 100000000u;
        }
    }
    SID_SYNCWARP();
    // ---- bases field: 32 bytes per step
    uint32_t cur = l0 + q4 + 1;             // bit index of the next byte to look at
    uint32_t na = 0, nc = 0, ng = 0, nt = 0, ndot = 0;
    uint32_t skip = 0;                      // bytes at `cur` still covered by a '^' or an indel
    // The loop has one exit, at its bottom: a refusal clears `ok` and lets the step finish on harmless values
    // (no `break`: the lanes of a warp leave together more often and the compiler emits no break blocks).
    bool running = ok;
    while (running) {
        if (cur + 64 > B.n_bits) { ok = false; cur = 0; running = false; }   // ran out of classified bytes
        const uint32_t term = bits32(B.term, cur);
        uint32_t vmask = 0xFFFFFFFFu;       // bytes of this step that belong to the field
        bool last = false;
        if (term) {
            const uint32_t n = first_bit(term);
            const uint32_t tb = s[region_off + cur + n];
            if (tb != '\t' && tb != ' ' && tb != '\n' && tb != 0) ok = false;   // a control byte inside the field
            vmask = n ? (0xFFFFFFFFu >> (32 - n)) : 0u;
            last = true;
        }
        if (skip) {
            const uint32_t sk = skip < 32 ? skip : 32;
            vmask &= sk == 32 ? 0u : (0xFFFFFFFFu << sk);
            skip -= sk;
        }
        const uint32_t car = bits32(B.caret, cur) & vmask;
        if (car & (car << 1)) ok = false;                           // "^^": leave the parity to the byte-wise path
        const uint32_t live = vmask & ~(car << 1);                  // '^' hides the byte after it
        const uint32_t pm = bits32(B.pm, cur) & live;
        uint32_t cm = live;                 // the bytes to count in this step
        uint32_t next = cur + 32;
        if (pm) {
            // everything before the sign counts; the indel length is read from the text
            // (pileup.cpp:131-136) and the walk restarts right after the number
            const uint32_t p = first_bit(pm);
            cm = live & (p ? (0xFFFFFFFFu >> (32 - p)) : 0u);
            uint32_t q = region_off + cur + p + 1;                  // first byte after the sign
            uint32_t n = 0;
            bool any = false;
            while (q + 8 < avail) {
                const uint32_t d = (uint32_t)s[q] - (uint32_t)'0';
                if (d > 9) break;
                if (n < (1u << 26)) n = n * 10 + d;
                any = true;
                ++q;
            }
            skip = any ? n : 0;             // a sign without digits is ignored (pileup.cpp:131-133)
            next = q - region_off;
            last = false;
            if (q + 8 >= avail) { ok = false; last = true; }
        } else if (car >> 31) {
            skip = 1;                       // the hidden byte is the first of the next step
        }
        na += pop_count(bits32(B.a, cur) & cm);
        nc += pop_count(bits32(B.c, cur) & cm);
        ng += pop_count(bits32(B.g, cur) & cm);
        nt += pop_count(bits32(B.t, cur) & cm);
        ndot += pop_count(bits32(B.dot, cur) & cm);
        cur = next;
        if (last || !ok) running = false;
    }
    SID_SYNCWARP();
    // '.' and ',' stand for the reference base (pileup.cpp:78-83); other reference characters drop them
    const uint32_t rf = ref & 0xDFu;
    o.profile = pack_profile(na + (rf == 'A' ? ndot : 0u), nc + (rf == 'C' ? ndot : 0u), ng + (rf == 'G' ? ndot : 0u),
                             nt + (rf == 'T' ? ndot : 0u));
    o.pos = (int32_t)acc;
    o.status = LINE_OK;
    return ok;
}

#if !defined(__CUDACC__)
// Host check: classifies the whole line (plus slack) like the kernel's stage 1, then runs stage 2.
inline bool parse_line_bits_host(const uint8_t* text, uint64_t len, uint64_t p, FastLine& o) {
    const int64_t first = (int64_t)(p & ~(uint64_t)31) - 32;       // region and staging start (32-byte aligned, with lead-in)
    uint64_t end = p;
    while (end < len && text[end] != '\n') ++end;
    const uint64_t avail64 = (((int64_t)end - first) + 256 + 31) & ~(uint64_t)31;
    if (avail64 > (1u << 20)) return false;
    static thread_local uint8_t scratch[(1u << 20) + 64] __attribute__((aligned(16)));
    static thread_local uint32_t arr[8][(1u << 15) + 8];
    for (uint64_t k = 0; k < avail64; ++k) {
        const int64_t q = first + (int64_t)k;
        scratch[k] = (q >= 0 && (uint64_t)q < len) ? text[q] : (uint8_t)'\n';
    }
    const uint32_t units = (uint32_t)(avail64 / 32);
    for (uint32_t u = 0; u < units; ++u) {
        uint32_t w[8];
        memcpy(w, scratch + 32 * u, 32);
        const ClassWords k = classify32(w);
        arr[0][u] = k.term; arr[1][u] = k.a; arr[2][u] = k.c; arr[3][u] = k.g; arr[4][u] = k.t;
        arr[5][u] = k.dot; arr[6][u] = k.caret; arr[7][u] = k.pm;
    }
    for (int i = 0; i < 8; ++i) arr[i][units] = arr[i][units + 1] = 0;
    BitArrays B {arr[0], arr[1], arr[2], arr[3], arr[4], arr[5], arr[6], arr[7], units * 32};
    return parse_line_bits(scratch, (uint32_t)avail64, 0, B, (uint32_t)((int64_t)p - first), o);
}
#endif

}  // namespace sid
