// Stage 2 of the tokenizer on the classes of stage 1 as they lie: one line per lane, the bases field walked UNIT BY UNIT
// (32 bytes = one 32-bit word per class, aligned as stage 1 stored them) instead of in 64-bit windows aligned to the line
// (parse_win.cuh).  No funnel shifts outside the header, no 64-bit mask arithmetic on a 32-bit ALU; a depth-30 line is a
// header plus two or three trips.  Same grammar and the same refusals as parse_line_win (the caller re-parses a refused
// line byte by byte); tests/hostcheck checks it against the byte-wise parser on every test text and on adversarial lines.
#pragma once
#include "parse_win.cuh"

namespace sid {

struct UnitRec { uint32_t w[CW_WORDS]; };

SID_HD UnitRec load_unit(const uint32_t* cw, uint32_t u) {
    UnitRec r;
#if defined(__CUDA_ARCH__)
    const uint4 lo = *reinterpret_cast<const uint4*>(cw + (size_t)u * CW_WORDS);
    const uint4 hi = *reinterpret_cast<const uint4*>(cw + (size_t)u * CW_WORDS + 4);
    r.w[0] = lo.x; r.w[1] = lo.y; r.w[2] = lo.z; r.w[3] = lo.w;
    r.w[4] = hi.x; r.w[5] = hi.y; r.w[6] = hi.z; r.w[7] = hi.w;
#else
    for (int c = 0; c < CW_WORDS; ++c) r.w[c] = cw[(size_t)u * CW_WORDS + c];
#endif
    return r;
}

SID_HD uint32_t low_bits32(uint32_t n) {                   // bits [0, n), n = 0 .. 32 (and beyond: all)
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(0xFFFFFFFFu, 0u, n);            // the upper word of (0 : ~0) << min(n, 32): one instruction
#else
    return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
#endif
}

// parsePileupLine + parseReadBases (pileup.cpp:13-153) of the line at line_off; arguments as parse_line_win.
// STRANDS: p5w holds the raw bit plane 5 of every unit (lower case); *fwd receives the profile of the bases read on the forward
// strand (upper-case letters and '.', pileup.cpp:78-124: ReadStack::strands summed per letter; reverse = profile - fwd).
template <bool WANT_POS, bool STRANDS = false>
SID_HD bool parse_line_units(const uint8_t* s, uint32_t region_off, const uint32_t* cw, const uint32_t* nlw, uint32_t n_bits,
                             uint32_t line_off, WinLine& o, const uint32_t* p5w = nullptr, uint64_t* fwd = nullptr) {
    const uint32_t ls = line_off - region_off;                      // bit index of the line's first byte
    bool ok = line_off >= region_off && ls + 64 <= n_bits && line_off >= 12;
    const uint32_t l0 = ok ? ls : 0, h0 = ok ? line_off : region_off + 16;
    const uint32_t u0 = l0 >> 5, sh = l0 & 31;
    const UnitRec r0 = load_unit(cw, u0), r1 = load_unit(cw, u0 + 1);
    // ---- header on the 32 bits from the line start (pileup.cpp:17-40): the bytes <= 0x20 locate the four separators
    const uint32_t sepmask = funnel_r(r0.w[CW_TERM], r1.w[CW_TERM], sh);
    const uint32_t digit = funnel_r(r0.w[CW_DIGIT], r1.w[CW_DIGIT], sh);
    const uint32_t ctl = funnel_r(r0.w[CW_CARET] | r0.w[CW_PM], r1.w[CW_CARET] | r1.w[CW_PM], sh);
    ok = ok && pop_count(sepmask) >= 4;
    uint32_t m = sepmask;
    const uint32_t p1 = first_bit(m); m &= m - 1;
    const uint32_t p2 = first_bit(m); m &= m - 1;
    const uint32_t p3 = first_bit(m); m &= m - 1;
    const uint32_t p4 = first_bit(m);
    const uint32_t nd = p2 - p1 - 1;
    // name non-empty, 1..9 digits, one reference character, depth non-empty, bases non-empty, all of it within reach
    ok = ok && p1 >= 1 && nd >= 1 && nd <= 9 && p3 == p2 + 2 && p4 > p3 + 1 && p4 <= 30 && ((sepmask >> (p4 + 1)) & 1u) == 0;
    const uint32_t q1 = ok ? p1 : 1, q2 = ok ? p2 : 3, q4 = ok ? p4 : 7;
    {
        // none of the four is a line end (fewer than five columns: the reference throws, pileup.cpp:22-40)
        const uint32_t nl32 = funnel_r(nlw[u0], nlw[u0 + 1], sh);
        ok = ok && (nl32 & (0xFFFFFFFFu >> (31 - q4))) == 0;
        const uint32_t dm = (0xFFFFFFFFu >> (32 - q2)) & ~(0xFFFFFFFFu >> (31 - q1));       // bits (q1, q2): the position is all digits
        ok = ok && (digit & dm) == dm;
    }
    // the reference character: '.' / ',' count as it (pileup.cpp:78-83); '^', '+', '-' would turn them into control
    // characters of the bases grammar -> byte-wise path
    const uint32_t rabs = l0 + q2 + 1;                              // its bit; in unit u0 or u0 + 1
    const bool rin0 = (rabs >> 5) == u0;
    const uint32_t rb = rabs & 31;
    ok = ok && ((ctl >> (q2 + 1)) & 1u) == 0;
    const bool ref_base = (((rin0 ? r0.w[CW_BASE] : r1.w[CW_BASE]) >> rb) & 1u) != 0;
    const bool ref_p1 = (((rin0 ? r0.w[CW_P1] : r1.w[CW_P1]) >> rb) & 1u) != 0;
    const bool ref_p2 = (((rin0 ? r0.w[CW_P2] : r1.w[CW_P2]) >> rb) & 1u) != 0;
    o.name_len = q1;
    o.hdr_len = q2;
    o.pos = 0;
    o.pos_canonical = s[h0 + q1 + 1] != (uint8_t)'0' || nd == 1;
    if (WANT_POS) {
        // the (up to) eight characters before the second separator, leading ones forced to '0' (digits checked above)
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s);
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
        const uint32_t xl = lo & 0x0F0F0F0Fu, xh = hi & 0x0F0F0F0Fu;
        const uint32_t tl = xl * 10u + (xl >> 8), th = xh * 10u + (xh >> 8);
        const uint32_t vl = (tl & 0xFFu) * 100u + ((tl >> 16) & 0xFFu), vh = (th & 0xFFu) * 100u + ((th >> 16) & 0xFFu);
        uint32_t acc = vl * 10000u + vh;
        if (ndd == 9) acc += ((uint32_t)s[e - 9] - (uint32_t)'0') * 100000000u;
        o.pos = (int32_t)acc;
    }
    SID_SYNCWARP();
    // ---- the bases field, unit by unit from the byte after the fourth separator
    const uint32_t a = l0 + q4 + 1;
    uint32_t uu = a >> 5;                                           // u0 or u0 + 1
    uint32_t valid = 0xFFFFFFFFu << (a & 31);                       // bits of the unit that belong to the field (so far as it starts here)
    uint32_t skip = 0;                                              // bytes at the start of the next unit still covered by a '^' or an indel
    uint32_t cn = 0, c1 = 0, c2 = 0, c12 = 0, cd = 0;
    uint32_t fn = 0, f1 = 0, f2 = 0, f12 = 0, fd = 0;              // STRANDS: the same counts over the forward strand only
    UnitRec r = uu == u0 ? r0 : r1;
    bool running = ok;
    while (running) {
        const uint32_t t = r.w[CW_TERM] & valid;
        const bool last = t != 0;
        const uint32_t e = first_bit(t);                            // 32 when the field goes on
        uint32_t live_all = valid & low_bits32(e);
        if (skip) {
            const uint32_t sk = skip < 32 ? skip : 32;
            live_all &= ~low_bits32(sk);
            skip -= sk;
        }
        uint32_t pmw = r.w[CW_PM], live;
        for (;;) {
            // '^' hides the byte after it (pileup.cpp:125-127); "^^": leave the parity to the byte-wise path
            const uint32_t car = r.w[CW_CARET] & live_all;
            if (car & (car << 1)) ok = false;
            live = live_all & ~(car << 1);
            const uint32_t pv = pmw & live;
            if (!pv || !ok) break;
            // '+' / '-' (pileup.cpp:128-147): its number is read from the raw text
            const uint32_t p = first_bit(pv);
            pmw &= ~(1u << p);
            const uint8_t* q = s + region_off + uu * 32 + p + 1;
            uint32_t ndig = 0, n = 0;
            while (ndig < 11 && (uint32_t)q[ndig] - (uint32_t)'0' <= 9u) {
                if (n < (1u << 26)) n = n * 10 + ((uint32_t)q[ndig] - (uint32_t)'0');
                ++ndig;
            }
            if (ndig == 0) continue;                                // a sign without a digit is ignored (pileup.cpp:131-133)
            if (ndig > 10) { ok = false; break; }                   // strtol's overflow rules: byte-wise path
            const uint64_t to = (uint64_t)p + 1 + ndig + n;         // first byte after the skipped ones (pileup.cpp:144), from the unit's start
            live_all &= ~(low_bits32(to >= 32 ? 32u : (uint32_t)to) & ~low_bits32(p + 1));
            if (to > 32 && !last) {
                const uint64_t over = to - 32;
                const uint32_t ov = over > (1u << 27) ? (1u << 27) : (uint32_t)over;
                if (ov > skip) skip = ov;
            }
        }
        if (!last && ((r.w[CW_CARET] & live_all) >> 31) && skip == 0) skip = 1;   // the hidden byte opens the next unit
        const uint32_t b = r.w[CW_BASE] & live;
        cn += pop_count(b);
        c1 += pop_count(b & r.w[CW_P1]);
        c2 += pop_count(b & r.w[CW_P2]);
        c12 += pop_count(b & r.w[CW_P1] & r.w[CW_P2]);
        cd += pop_count(r.w[CW_DOT] & live);
        if (STRANDS) {
            const uint32_t fb = b & ~p5w[uu];                       // upper-case letters
            fn += pop_count(fb);
            f1 += pop_count(fb & r.w[CW_P1]);
            f2 += pop_count(fb & r.w[CW_P2]);
            f12 += pop_count(fb & r.w[CW_P1] & r.w[CW_P2]);
            fd += pop_count(r.w[CW_DOT] & r.w[CW_P1] & live);       // '.' (0x2e) has bit 1, ',' (0x2c) has not
        }
        if (last || !ok) running = false;
        else {
            ++uu;
            valid = 0xFFFFFFFFu;
            if ((uu + 1) * 32 > n_bits) { ok = false; running = false; }        // ran out of classified bytes
            else r = load_unit(cw, uu);
        }
    }
    SID_SYNCWARP();
    WinHeader hd;
    hd.l0 = l0; hd.q4 = q4; hd.ref_base = ref_base; hd.ref_p1 = ref_p1; hd.ref_p2 = ref_p2;
    o.profile = win_profile(cn, c1, c2, c12, cd, hd);
    if (STRANDS) *fwd = win_profile(fn, f1, f2, f12, fd, hd);
    o.status = LINE_OK;
    return ok;
}

#if !defined(__CUDACC__)
// Host check: stage 2 by units on the line at p, classified like the kernel's stage 1.
template <bool WANT_POS>
inline bool parse_line_units_host(const uint8_t* text, uint64_t len, uint64_t p, WinLine& o, uint64_t* fwd = nullptr) {
    const HostLineClasses h = classify_line_host(text, len, p);
    if (!h.usable) return false;
    if (!fwd) return parse_line_units<WANT_POS>(h.scratch, 0, h.cw, h.nlw, h.units * 32, h.line_off, o);
    // plane 5 of every classified unit, as the STRANDS tokenizer keeps it
    static thread_local uint32_t p5w[(1u << 15) + 8];
    for (uint32_t u = 0; u < h.units + CW_PAD_UNITS; ++u) {
        uint32_t v = 0;
        if (u < h.units) for (int i = 0; i < 32; ++i) v |= (uint32_t)((h.scratch[32 * u + i] >> 5) & 1u) << i;
        p5w[u] = v;
    }
    return parse_line_units<WANT_POS, true>(h.scratch, 0, h.cw, h.nlw, h.units * 32, h.line_off, o, p5w, fwd);
}
#endif

}  // namespace sid
