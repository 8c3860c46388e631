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
//       - '+' / '-' (outside a masked byte) switch to the byte-wise BasesState for the indel, then
//         the word loop resumes
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

// `s` is a 4-byte aligned staging buffer whose byte 0 is absolute offset abs0 (abs0 % 4 == 0) with
// `avail` valid bytes (multiple of 4).  Returns false when the line must take the byte-wise path.
SID_HD bool parse_line_fast_smem(const uint8_t* s, uint64_t abs0, uint32_t avail, uint64_t line_abs, FastLine& o) {
    const uint32_t start = (uint32_t)(line_abs - abs0);
    if (start + 64 > avail) return false;
    const uint32_t safe_end = avail - 8;
    uint32_t i = start;
    uint32_t c;
    // ---- chromosome name
    c = s[i];
    if (c <= 0x20) return false;
    do { c = s[++i]; } while (c > 0x20 && i < safe_end);
    if (c != '\t' && c != ' ') return false;
    o.chrom_off = 0;
    o.chrom_len = i - start;
    ++i;
    // ---- position: 1..9 digits
    uint32_t acc = 0, nd = 0;
    for (;;) {
        const uint32_t d = (uint32_t)s[i] - (uint32_t)'0';
        if (d > 9) break;
        acc = acc * 10 + d;
        ++i;
        if (++nd > 9) return false;
    }
    if (nd == 0) return false;
    c = s[i];
    if (c != '\t' && c != ' ') return false;
    ++i;
    // ---- reference base: exactly one character
    const uint32_t ref = s[i];
    if (ref <= 0x20) return false;
    c = s[++i];
    if (c != '\t' && c != ' ') return false;
    ++i;
    // ---- depth column: skipped
    c = s[i];
    if (c <= 0x20) return false;
    do { c = s[++i]; } while (c > 0x20 && i < safe_end);
    if (c != '\t' && c != ' ') return false;
    ++i;
    if (s[i] <= 0x20) return false;          // empty bases field or doubled delimiter
    if (i >= safe_end) return false;

    // ---- bases field, one 32-bit word per step
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s);
    const uint32_t n_words = avail >> 2;
    uint32_t idx = i >> 2;
    const uint32_t sh = (i & 3) * 8;
    uint32_t cur = sw[idx];
    uint32_t a7 = 0, c7 = 0, g7 = 0, t7 = 0, d7 = 0;   // 128 * count
    uint32_t skip = 0;
    for (;;) {
        if (idx + 1 >= n_words) return false;          // ran out of staged bytes
        const uint32_t nxt = sw[idx + 1];
        uint32_t w = funnel_r(cur, nxt, sh);
        cur = nxt;
        ++idx;
        // bytes outside [0x21, 0x7f]
        const uint32_t ok7 = ((w | M80) - NEUTRAL) & ~w & M80;
        bool last = false;
        uint32_t nvalid = 4;
        if (ok7 != M80) {
            nvalid = (uint32_t)first_flag_byte(ok7 ^ M80);
            const uint32_t b = (w >> (8 * nvalid)) & 0xFFu;
            if (b != '\t' && b != ' ' && b != '\n' && b != 0) return false;     // a control or 8-bit byte inside the field
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
        if (caret7 & (caret7 << 8)) return false;       // "^^": leave the parity to the byte-wise path
        const uint32_t masked7 = caret7 << 8;
        // '+' / '-' outside masked bytes: byte-wise for the rest of this word and the indel that follows
        const uint32_t pm7 = (eq7(w, 0x2B2B2B2Bu, 0) | eq7(w, 0x2D2D2D2Du, 0)) & ~masked7;
        if (pm7) {
            BasesState b;
            b.init();
            // everything before the sign is plain: count it with the flags, then feed from the sign on
            const uint32_t k0 = (uint32_t)first_flag_byte(pm7);
            const uint32_t before = k0 ? (0xFFFFFFFFu >> (32 - 8 * k0)) : 0u;
            const uint32_t wb = (w & before) | (NEUTRAL & ~before);
            const uint32_t ex = masked7;
            const uint32_t f = wb & 0xDFDFDFDFu;
            a7 = add_flags(eq7(f, 0x41414141u, ex), a7);
            c7 = add_flags(eq7(f, 0x43434343u, ex), c7);
            g7 = add_flags(eq7(f, 0x47474747u, ex), g7);
            t7 = add_flags(eq7(f, 0x54545454u, ex), t7);
            d7 = add_flags(eq7(wb & 0xFDFDFDFDu, 0x2C2C2C2Cu, ex), d7);
            // byte-wise from the sign: bytes come from the staged text directly
            uint32_t q = (idx - 1) * 4 + (sh >> 3) + k0;      // byte offset of the sign in s
            // (idx was advanced; the word started at ((idx-1)*4 + sh/8))
            bool ended = false;
            for (;;) {
                if (q >= safe_end) return false;
                const uint32_t ch = s[q];
                if (ch <= 0x20 || ch >= 0x80) {
                    if (ch == '\t' || ch == ' ' || ch == '\n' || ch == 0) { ended = true; break; }
                    return false;
                }
                b.feed((uint8_t)ch);
                ++q;
                // resume the word loop once the state machine is idle again and we are word aligned
                // with respect to the field's word grid
                if (b.mode == 0 && b.skip == 0 && ((q - (sh >> 3)) & 3) == 0) break;
            }
            a7 += 128u * b.cnt[0];
            c7 += 128u * b.cnt[1];
            g7 += 128u * b.cnt[2];
            t7 += 128u * b.cnt[3];
            d7 += 128u * (b.dots + b.commas);
            if (ended) break;
            // a trailing '^' inside the byte-wise stretch leaves skip == 1 only if we stopped right after it,
            // which the resume condition (skip == 0) excludes
            idx = (q - (sh >> 3)) >> 2;
            if (idx >= n_words) return false;
            cur = sw[idx];
            continue;
        }
        if (caret7 >> 31) skip = 1;                     // the masked byte is the first of the next word
        const uint32_t f = w & 0xDFDFDFDFu;
        a7 = add_flags(eq7(f, 0x41414141u, masked7), a7);
        c7 = add_flags(eq7(f, 0x43434343u, masked7), c7);
        g7 = add_flags(eq7(f, 0x47474747u, masked7), g7);
        t7 = add_flags(eq7(f, 0x54545454u, masked7), t7);
        d7 = add_flags(eq7(w & 0xFDFDFDFDu, 0x2C2C2C2Cu, masked7), d7);
        if (last) break;
    }
    uint32_t cnt[4] = {a7 >> 7, c7 >> 7, g7 >> 7, t7 >> 7};
    const int ri = ref_index((uint8_t)ref);
    if (ri >= 0) cnt[ri] += d7 >> 7;
    o.profile = pack_profile(cnt[0], cnt[1], cnt[2], cnt[3]);
    o.pos = (int32_t)acc;
    o.status = LINE_OK;
    return true;
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
