// TEST INFRASTRUCTURE ONLY: the word-at-a-time (SWAR) tokenizer of round 1, superseded in the kernels by the
// bit-parallel form (sid_b200/csrc/parse_bits.cuh).  Kept as a second, independently written cross-check
// of the byte-wise state machine: same results as parse.cuh on every line it accepts, refusal otherwise.
#pragma once
#include "../../sid_b200/csrc/parse_fast.cuh"

namespace sid {

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
// On the device ALL 32 lanes of a warp must call this together (lanes without a line of their own
// pass any valid line): the header and the word loop end in a warp-wide reconvergence point, so
// the code after them runs with full warps again.  A refusal therefore never returns early; it
// clears `ok` and lets the lane idle to the next reconvergence point.
SID_HD bool parse_line_fast_smem(const uint8_t* s, uint64_t abs0, uint32_t avail, uint64_t line_abs, FastLine& o) {
    const uint32_t start = (uint32_t)(line_abs - abs0);
    bool ok = start + 64 <= avail && start >= 12;
    const uint32_t safe_end = avail - 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s);
    // ---- header: the first 32 bytes of the line as eight words; a bit mask of the bytes <= 0x20
    //      locates the four separators after chrom, pos, ref and depth
    const uint32_t h0 = ok ? start : 16;
    uint32_t sepmask = 0;
    {
        const uint32_t* hw = sw + (h0 >> 2);
        const uint32_t hs = (h0 & 3) * 8;
        uint32_t prev = hw[0];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t next = hw[k + 1];
            const uint32_t w = funnel_r(prev, next, hs);
            prev = next;
            const uint32_t sep7 = ~((w | M80) - NEUTRAL) & ~w & M80;     // bit 7 set iff the byte is <= 0x20
            sepmask |= ((sep7 * 0x00204081u) >> 28) << (4 * k);
        }
    }
    ok = ok && pop_count(sepmask) >= 4;
    uint32_t m = sepmask;
    const uint32_t p1 = first_bit(m); m &= m - 1;
    const uint32_t p2 = first_bit(m); m &= m - 1;
    const uint32_t p3 = first_bit(m); m &= m - 1;
    const uint32_t p4 = first_bit(m);
    const uint32_t nd = p2 - p1 - 1;
    // chrom non-empty, 1..9 digits, one reference character, depth non-empty, bases non-empty within reach
    ok = ok && p1 >= 1 && nd >= 1 && nd <= 9 && p3 == p2 + 2 && p4 > p3 + 1 && p4 <= 30 && ((sepmask >> (p4 + 1)) & 1u) == 0;
    if (!ok) { /* keep every read below in bounds */ }
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
        const uint32_t e = h0 + q2;                           // offset of the separator after the digits
        const uint32_t ndd = ok ? nd : 1;
        const uint32_t* pw = sw + ((e - 8) >> 2);
        const uint32_t ps = ((e - 8) & 3) * 8;
        const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2];
        uint32_t lo = funnel_r(w0, w1, ps), hi = funnel_r(w1, w2, ps);
        const uint32_t zero = ndd >= 8 ? 0u : 8u - ndd;       // leading bytes that are not digits of this number
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
    uint32_t i = h0 + q4 + 1;                                  // first byte of the bases field
    if (i >= safe_end) { ok = false; i = 16; }
    SID_SYNCWARP();

    // ---- bases field, one 32-bit word per step
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
    // like a staged tile: 16 bytes of lead-in before the aligned start, '\n' outside the text
    const int64_t first = (int64_t)(p & ~(uint64_t)15) - 16;
    uint64_t end = p;
    while (end < len && text[end] != '\n') ++end;
    const uint64_t avail64 = (((int64_t)end - first) + 64 + 15) & ~(uint64_t)15;
    if (avail64 > (1u << 20)) return false;
    static thread_local uint8_t scratch[(1u << 20) + 64] __attribute__((aligned(16)));
    for (uint64_t k = 0; k < avail64; ++k) {
        const int64_t q = first + (int64_t)k;
        scratch[k] = (q >= 0 && (uint64_t)q < len) ? text[q] : (uint8_t)'\n';
    }
    return parse_line_fast_smem(scratch, (uint64_t)first, (uint32_t)avail64, p, o);
}
#endif

}  // namespace sid
