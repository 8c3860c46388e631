// Word-at-a-time (SWAR) form of the line tokenizer: same results as parse.cuh on every line it
// accepts, and it REFUSES (returns false) anything outside its fast grammar so the caller falls back
// to the byte-wise state machine of parse.cuh.  tests/hostcheck compares the two on every test text
// and on adversarial random lines.
//
//   parsePileupLine  pileup.cpp:13-68   header: chrom \t pos \t ref \t depth \t  (single delimiters,
//                                       1..9 digit unsigned position)
//   parseReadBases   pileup.cpp:70-153  bases field, 4 bytes per step:
//       - bytes outside [0x21,0x7f] end the field (tab/space/newline/NUL) or refuse the line
//       - '^' masks the following byte without a branch; "^^" refuses
//       - '+' / '-' (outside a masked byte): the length is read byte-wise, the word loop restarts
//         right after the number with that many bytes to neutralise
//       - A/C/G/T (case folded) and '.'/',' are counted with per-byte equality flags summed by dp4a
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

// Adds 128 * (number of bytes of f7 whose bit 7 is set) to acc; f7 has only bits 7,15,23,31.
SID_HD uint32_t add_flags(uint32_t f7, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp4a(f7, 0x01010101u, acc);
#else
    return acc + 128u * (uint32_t)__builtin_popcount(f7);
#endif
}

SID_HD int first_flag_byte(uint32_t f7) {   // index of the lowest byte whose bit 7 is set; f7 != 0
#if defined(__CUDA_ARCH__)
    return (__ffs((int)f7) - 1) >> 3;
#else
    return __builtin_ctz(f7) >> 3;
#endif
}

constexpr uint32_t M80 = 0x80808080u, M7F = 0x7F7F7F7Fu, NEUTRAL = 0x21212121u;   // '!' is ignored by the grammar

// Bit 7 of each byte set iff the byte (all bytes must be < 0x80) equals the pattern byte, and the
// same byte of `excl` has bit 7 clear.
SID_HD uint32_t eq7(uint32_t x, uint32_t pat, uint32_t excl) {
    const uint32_t t = (x ^ pat) + M7F;          // bit 7 set iff the byte differs
    return ~t & M80 & ~excl;
}

#if defined(__CUDA_ARCH__)
#define SID_SYNCWARP() __syncwarp()
#else
#define SID_SYNCWARP() ((void)0)
#endif

// `s` is a 4-byte aligned staging buffer whose byte 0 is absolute offset abs0 (abs0 % 4 == 0) with
// `avail` valid bytes (multiple of 4).  Returns false when the line must take the byte-wise path.
// On the device ALL 32 lanes of a warp must call this together (lanes without a line of their own
// pass any valid line): the header and the word loop end in a warp-wide reconvergence point, so
// the code after them runs with full warps again.  A refusal therefore never returns early; it
// clears `ok` and lets the lane idle to the next reconvergence point.
SID_HD bool parse_line_fast_smem(const uint8_t* s, uint64_t abs0, uint32_t avail, uint64_t line_abs, FastLine& o) {
    const uint32_t start = (uint32_t)(line_abs - abs0);
    bool ok = start + 64 <= avail;
    const uint32_t safe_end = avail - 8;
    uint32_t i = ok ? start : 0;
    uint32_t c;
    // ---- chromosome name
    c = s[i];
    ok = ok && c > 0x20;
    do { c = s[++i]; } while (c > 0x20 && i < safe_end);
    ok = ok && (c == '\t' || c == ' ');
    o.chrom_off = 0;
    o.chrom_len = i - start;
    if (i >= safe_end) i = safe_end - 16;     // refused already; keeps the remaining header reads in bounds
    ++i;
    // ---- position: 1..9 digits
    uint32_t acc = 0, nd = 0;
    for (;;) {
        const uint32_t d = (uint32_t)s[i] - (uint32_t)'0';
        if (d > 9 || nd > 9) break;
        acc = acc * 10 + d;
        ++i;
        ++nd;
    }
    ok = ok && nd >= 1 && nd <= 9;
    c = s[i];
    ok = ok && (c == '\t' || c == ' ');
    ++i;
    // ---- reference base: exactly one character
    const uint32_t ref = s[i];
    ok = ok && ref > 0x20;
    c = s[++i];
    ok = ok && (c == '\t' || c == ' ');
    ++i;
    // ---- depth column: skipped
    c = s[i];
    ok = ok && c > 0x20 && i < safe_end;
    if (i >= safe_end) i = safe_end;
    do { c = s[++i]; } while (c > 0x20 && i < safe_end);
    ok = ok && (c == '\t' || c == ' ');
    ++i;
    ok = ok && i < safe_end && s[i] > 0x20;      // empty bases field or doubled delimiter
    if (i >= safe_end) i = safe_end;
    SID_SYNCWARP();

    // ---- bases field, one 32-bit word per step
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s);
    const uint32_t n_words = avail >> 2;
    uint32_t idx = i >> 2;
    uint32_t sh = (i & 3) * 8;
    uint32_t cur = sw[idx];
    uint32_t a7 = 0, c7 = 0, g7 = 0, t7 = 0, d7 = 0;   // 128 * count
    uint32_t skip = 0;
    bool running = ok;
    while (running) {
        if (idx + 1 >= n_words) { ok = false; break; }  // ran out of staged bytes
        const uint32_t nxt = sw[idx + 1];
        uint32_t w = funnel_r(cur, nxt, sh);
        cur = nxt;
        ++idx;
        // bytes outside [0x21, 0x7f]
        const uint32_t ok7 = ((w | M80) - NEUTRAL) & ~w & M80;
        bool last = false;
        if (ok7 != M80) {
            const uint32_t nvalid = (uint32_t)first_flag_byte(ok7 ^ M80);
            const uint32_t b = (w >> (8 * nvalid)) & 0xFFu;
            if (b != '\t' && b != ' ' && b != '\n' && b != 0) { ok = false; break; }   // a control or 8-bit byte inside the field
            const uint32_t keep = nvalid ? (0xFFFFFFFFu >> (32 - 8 * nvalid)) : 0u;
            w = (w & keep) | (NEUTRAL & ~keep);
            last = true;
        }
        // bytes still covered by a skip that started in an earlier word
        if (skip) {
            const uint32_t sk = skip < 4 ? skip : 4;
            const uint32_t m = sk == 4 ? 0xFFFFFFFFu : ((1u << (8 * sk)) - 1u);
            w = (w & ~m) | (NEUTRAL & m);
            skip -= sk;
        }
        // '^' masks the byte after it
        const uint32_t caret7 = eq7(w, 0x5E5E5E5Eu, 0);
        if (caret7 & (caret7 << 8)) { ok = false; break; }   // "^^": leave the parity to the byte-wise path
        const uint32_t masked7 = caret7 << 8;
        // '+' / '-' outside masked bytes
        const uint32_t pm7 = (eq7(w, 0x2B2B2B2Bu, 0) | eq7(w, 0x2D2D2D2Du, 0)) & ~masked7;
        uint32_t wc = w;                                 // the bytes to count in this step
        bool restart = false;
        uint32_t q = 0;
        if (pm7) {
            // everything before the sign is plain; the indel length is read byte-wise
            // (pileup.cpp:131-136); the skipped bases are then neutralised by the word loop itself,
            // restarted right after the number
            const uint32_t k0 = (uint32_t)first_flag_byte(pm7);
            const uint32_t before = k0 ? (0xFFFFFFFFu >> (32 - 8 * k0)) : 0u;
            wc = (w & before) | (NEUTRAL & ~before);
            q = (idx - 1) * 4 + (sh >> 3) + k0 + 1;      // first byte after the sign
            uint32_t n = 0;
            bool any = false;
            while (q < safe_end) {
                const uint32_t d = (uint32_t)s[q] - (uint32_t)'0';
                if (d > 9) break;
                if (n < (1u << 26)) n = n * 10 + d;
                any = true;
                ++q;
            }
            if (q >= safe_end) { ok = false; break; }
            skip = any ? n : 0;                          // a sign without digits is ignored (pileup.cpp:131-133)
            restart = true;
            last = false;
        } else if (caret7 >> 31) {
            skip = 1;                                    // the masked byte is the first of the next word
        }
        const uint32_t f = wc & 0xDFDFDFDFu;
        a7 = add_flags(eq7(f, 0x41414141u, masked7), a7);
        c7 = add_flags(eq7(f, 0x43434343u, masked7), c7);
        g7 = add_flags(eq7(f, 0x47474747u, masked7), g7);
        t7 = add_flags(eq7(f, 0x54545454u, masked7), t7);
        d7 = add_flags(eq7(wc & 0xFDFDFDFDu, 0x2C2C2C2Cu, masked7), d7);
        if (restart) {
            idx = q >> 2;
            sh = (q & 3) * 8;
            cur = sw[idx];
        }
        if (last) running = false;
    }
    SID_SYNCWARP();
    // '.' and ',' stand for the reference base (pileup.cpp:78-83); other reference characters drop them
    const uint32_t rf = ref & 0xDFu, dots = d7 >> 7;
    const uint32_t na = (a7 >> 7) + (rf == 'A' ? dots : 0u);
    const uint32_t nc = (c7 >> 7) + (rf == 'C' ? dots : 0u);
    const uint32_t ng = (g7 >> 7) + (rf == 'G' ? dots : 0u);
    const uint32_t nt = (t7 >> 7) + (rf == 'T' ? dots : 0u);
    o.profile = pack_profile(na, nc, ng, nt);
    o.pos = (int32_t)acc;
    o.status = LINE_OK;
    return ok;
}

#if !defined(__CUDACC__)
// Flat-buffer entry used by the host checks only: stages the line into an aligned scratch copy.
inline bool parse_line_fast(const uint8_t* text, uint64_t len, uint64_t p, FastLine& o) {
    // the caller guarantees text is readable up to a multiple of 16 past len (padding reads as '\n')
    const uint64_t abs0 = p & ~(uint64_t)15;
    uint64_t end = p;
    while (end < len && text[end] != '\n') ++end;
    const uint64_t avail64 = ((end - abs0) + 64 + 15) & ~(uint64_t)15;
    if (avail64 > (1u << 20)) return false;
    static thread_local uint8_t scratch[(1u << 20) + 64] __attribute__((aligned(16)));
    for (uint64_t k = 0; k < avail64; ++k) scratch[k] = abs0 + k < len ? text[abs0 + k] : (uint8_t)'\n';
    return parse_line_fast_smem(scratch, abs0, (uint32_t)avail64, p, o);
}
#endif

}  // namespace sid
