// Bit-parallel tokenizer, round 2 form: one 64-bit window per line.
//
// Stage 1 (flat, per 32-byte unit): the bytes are transposed into their 8 bit planes and every character class
// the grammar of parsePileupLine / parseReadBases (pileup.cpp:13-153) distinguishes becomes a boolean function of
// the planes (classify_unit).  The class words of a unit are stored together (array of structures, eight words per
// unit, two 16-byte stores), the '\n' word separately (it only feeds the line starts and one header check).
//
// Stage 2 (one line per lane, parse_line_win): six 16-byte loads fetch all classes for the 96 bits that hold the
// line's first 64 bytes; two funnel shifts per class align them to the line start.  Header (four separators, digits
// of the position, reference character) and bases field (counts as population counts) are then pure register
// arithmetic on that window: no byte loads, no per-class shared-memory traffic, and for an ordinary depth-30 line
// (header + bases <= 64 bytes) no loop.  Longer fields continue in further 64-bit windows.  A/C/G/T are not four
// classes but one (BASE) plus the raw bit planes 1 and 2 of the byte, which tell the four letters apart:
//   A 0x41 -> (p2,p1) = 00   C 0x43 -> 01   G 0x47 -> 11   T 0x54 -> 10      (same for lower case)
//
// Same contract as before: a line outside the fast grammar is REFUSED (returns false) and the caller re-parses it
// with the byte-wise state machine of parse.cuh.  tests/hostcheck checks classifier == per-byte definition and
// window parser == byte-wise parser on every test text and on adversarial random lines.
#pragma once
#include "common.cuh"
#include "parse.cuh"
#include "parse_bits.cuh"

namespace sid {

enum : int { CW_TERM = 0, CW_BASE = 1, CW_P1 = 2, CW_P2 = 3, CW_DOT = 4, CW_CARET = 5, CW_PM = 6, CW_DIGIT = 7, CW_WORDS = 8 };
constexpr uint32_t CW_PAD_UNITS = 3;       // zero units after the classified ones: a window may start in the last unit

struct UnitClasses {
    uint32_t w[CW_WORDS];   // bit i of each word <-> byte i of the unit
    uint32_t nl;            // '\n'
    uint32_t bad;           // control bytes other than '\t' and '\n' (NUL included): their lines leave the fast path
    uint32_t p5;            // raw bit plane 5: on a letter, lower case = the reverse strand (pileup.cpp:85-124)
};

SID_HD uint32_t popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popcll(x);
#else
    return (uint32_t)__builtin_popcountll(x);
#endif
}

// b[0..7]: the unit's 32 bytes as little-endian words.
SID_HD UnitClasses classify_unit(const uint32_t b[8]) {
    uint32_t n[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t ab01 = byte_perm(b[h], b[2 + h], 0x5140u), ab23 = byte_perm(b[h], b[2 + h], 0x7362u);
        const uint32_t cd01 = byte_perm(b[4 + h], b[6 + h], 0x5140u), cd23 = byte_perm(b[4 + h], b[6 + h], 0x7362u);
        n[4 * h + 0] = byte_perm(ab01, cd01, 0x5410u);
        n[4 * h + 1] = byte_perm(ab01, cd01, 0x7632u);
        n[4 * h + 2] = byte_perm(ab23, cd23, 0x5410u);
        n[4 * h + 3] = byte_perm(ab23, cd23, 0x7632u);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) swap_blocks<0x0F0F0F0Fu, 4>(n[i], n[i + 4]);
    swap_blocks<0x33333333u, 2>(n[0], n[2]);
    swap_blocks<0x33333333u, 2>(n[1], n[3]);
    swap_blocks<0x33333333u, 2>(n[4], n[6]);
    swap_blocks<0x33333333u, 2>(n[5], n[7]);
#pragma unroll
    for (int i = 0; i < 8; i += 2) swap_blocks<0x55555555u, 1>(n[i], n[i + 1]);
    const uint32_t p0 = n[0], p1 = n[1], p2 = n[2], p3 = n[3], p4 = n[4], p5 = n[5], p6 = n[6], p7 = n[7];
    UnitClasses k;
    // high nibbles (p7 p6 p5 p4)
    const uint32_t h00 = lop3<~TA & ~TB & ~TC>(p7, p6, p5);            // 000x: 0x00..0x1f
    const uint32_t x25 = lop3<~TA & ~TB & TC>(p7, p6, p5);             // 001x: 0x20..0x3f
    const uint32_t h2 = x25 & ~p4;                                     // 0x2_
    const uint32_t h46 = lop3<~TA & TB & ~TC>(p7, p6, p4);             // 0x4_ 0x6_
    const uint32_t h57 = lop3<~TA & TB & TC>(p7, p6, p4);              // 0x5_ 0x7_
    // low nibbles (p3 p2 p1 p0)
    const uint32_t lacg = lop3<~TA & ~(TB & ~TC)>(p3, p2, p1);         // 0 (p2 p1) in {00, 01, 11}, p0 below
    const uint32_t l010 = lop3<~TA & TB & ~TC>(p3, p2, p1);
    const uint32_t bacg = lop3<TA & TB & TC>(lacg, p0, h46);           // 0x41 0x43 0x47 (+0x20)
    const uint32_t bt = lop3<TA & ~TB & TC>(l010, p0, h57);            // 0x54 0x74
    k.w[CW_BASE] = bacg | bt;
    k.w[CW_P1] = p1;
    k.w[CW_P2] = p2;
    k.w[CW_DOT] = lop3<TA & ~TB & TC>(p3 & p2, p0, h2);                // 0x2c 0x2e
    k.w[CW_PM] = lop3<TA & TB & TC>(lop3<TA & (TB ^ TC)>(p3, p2, p1), p0, h2);          // 0x2b 0x2d
    k.w[CW_CARET] = lop3<TA & ~TB & TC>(lop3<TA & TB & TC>(p3, p2, p1), p0, h57) & ~p5; // 0x5e
    k.w[CW_DIGIT] = lop3<TA & TB & TC>(lop3<(~TA | (~TB & ~TC)) & 0xFF>(p3, p2, p1), x25, p4);   // 0x30..0x39
    k.nl = lop3<TA & ~TB & TC>(lop3<TA & ~TB & TC>(p3, p2, p1), p0, h00) & ~p4;         // 0x0a
    const uint32_t tab = lop3<TA & TB & TC>(lop3<TA & ~TB & ~TC>(p3, p2, p1), p0, h00) & ~p4;   // 0x09
    const uint32_t space = lop3<TA & ~TB & TC>(lop3<~TA & ~TB & ~TC>(p3, p2, p1), p0, h2);      // 0x20
    k.w[CW_TERM] = h00 | space;                                        // byte <= 0x20
    k.bad = lop3<TA & ~TB & ~TC>(h00, tab, k.nl);
    k.p5 = p5;
    return k;
}

struct WinLine {
    int status;
    int32_t pos;            // only with WANT_POS
    uint64_t profile;
    uint32_t name_len;      // the name starts at the first byte of the line
    uint32_t hdr_len;       // bytes of "name<sep>position": with a canonical position the CSV row starts with exactly these
    bool pos_canonical;     // the digits are what printf("%d") prints for the position (no leading zero); 1..9 digits always
};

struct Win64 {              // one 64-bit window of every class
    uint64_t term, base, p1, p2, dot, caret, pm, digit;
};

// The classes of the 64 bytes starting at bit `pos` of the class arrays (pos + 64 must lie within the padded arrays).
SID_HD Win64 load_window(const uint32_t* cw, uint32_t pos) {
    const uint32_t u = pos >> 5, sh = pos & 31;
    uint32_t a[3][CW_WORDS];
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint4 lo = *reinterpret_cast<const uint4*>(cw + (size_t)(u + k) * CW_WORDS);
        const uint4 hi = *reinterpret_cast<const uint4*>(cw + (size_t)(u + k) * CW_WORDS + 4);
        a[k][0] = lo.x; a[k][1] = lo.y; a[k][2] = lo.z; a[k][3] = lo.w;
        a[k][4] = hi.x; a[k][5] = hi.y; a[k][6] = hi.z; a[k][7] = hi.w;
    }
#else
    for (int k = 0; k < 3; ++k) for (int c = 0; c < CW_WORDS; ++c) a[k][c] = cw[(size_t)(u + k) * CW_WORDS + c];
#endif
    uint64_t v[CW_WORDS];
#pragma unroll
    for (int c = 0; c < CW_WORDS; ++c)
        v[c] = (uint64_t)funnel_r(a[0][c], a[1][c], sh) | ((uint64_t)funnel_r(a[1][c], a[2][c], sh) << 32);
    Win64 w;
    w.term = v[CW_TERM]; w.base = v[CW_BASE]; w.p1 = v[CW_P1]; w.p2 = v[CW_P2];
    w.dot = v[CW_DOT]; w.caret = v[CW_CARET]; w.pm = v[CW_PM]; w.digit = v[CW_DIGIT];
    return w;
}

SID_HD uint64_t low_bits64(uint32_t n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }      // bits [0, n)

// What the header of a line leaves for the bases field.
struct WinHeader {
    uint32_t l0;            // bit index (byte offset from region_off) of the line's first byte
    uint32_t q4;            // offset of the fourth separator from there: the bases field starts at l0 + q4 + 1
    bool ref_base, ref_p1, ref_p2;     // the reference character is A/C/G/T (any case) and its two telling bit planes
};

// `s`: staged text (4-byte aligned, byte 0 of the class arrays is s[region_off]); `cw`: class words (array of
// structures) of n_bits classified bytes followed by CW_PAD_UNITS zero units; `nlw`: the '\n' words (same padding).
// WANT_POS: also convert the position (the row writer copies its digits from the text instead).
// Header of the line at line_off (parsePileupLine, pileup.cpp:13-40): separators, position, reference character.
// `w` receives the class window of the line's first 64 bytes.  Returns false when the line leaves the fast grammar.
template <bool WANT_POS>
SID_HD bool win_header(const uint8_t* s, uint32_t region_off, const uint32_t* cw, const uint32_t* nlw, uint32_t n_bits,
                       uint32_t line_off, WinLine& o, WinHeader& hd, Win64& w) {
    const uint32_t ls = line_off - region_off;                      // bit index of the line's first byte
    bool ok = line_off >= region_off && ls + 64 <= n_bits && line_off >= 12;
    const uint32_t l0 = ok ? ls : 0, h0 = ok ? line_off : region_off + 16;
    w = load_window(cw, l0);
    // ---- header: the bytes <= 0x20 among the first 32 locate the four separators (pileup.cpp:17-36)
    const uint32_t sepmask = (uint32_t)w.term;
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
        // none of the four is a line end (fewer than five columns: the reference throws, pileup.cpp:22-40); other
        // control bytes are the caller's business (UnitClasses::bad)
        const uint32_t nl32 = funnel_r(nlw[l0 >> 5], nlw[(l0 >> 5) + 1], l0 & 31);
        ok = ok && (nl32 & (0xFFFFFFFFu >> (31 - q4))) == 0;
        // the position is all digits
        const uint32_t dm = (0xFFFFFFFFu >> (32 - q2)) & ~(0xFFFFFFFFu >> (31 - q1));       // bits (q1, q2)
        ok = ok && ((uint32_t)w.digit & dm) == dm;
    }
    // the reference character: '.' / ',' count as it (pileup.cpp:78-83); '^', '+', '-' would turn them into control
    // characters of the bases grammar -> byte-wise path
    const uint32_t rbit = q2 + 1;
    ok = ok && (((uint32_t)(w.caret | w.pm) >> rbit) & 1u) == 0;
    hd.l0 = l0;
    hd.q4 = q4;
    hd.ref_base = (((uint32_t)w.base >> rbit) & 1u) != 0;
    hd.ref_p1 = (((uint32_t)w.p1 >> rbit) & 1u) != 0;
    hd.ref_p2 = (((uint32_t)w.p2 >> rbit) & 1u) != 0;
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
    return ok;
}

// The counts of a bases field as the packed profile; '.' and ',' stand for the reference base (pileup.cpp:78-83).
SID_HD uint64_t win_profile(uint32_t cn, uint32_t c1, uint32_t c2, uint32_t c12, uint32_t cd, const WinHeader& hd) {
    if (hd.ref_base) {
        cn += cd;
        c1 += hd.ref_p1 ? cd : 0u;
        c2 += hd.ref_p2 ? cd : 0u;
        c12 += (hd.ref_p1 && hd.ref_p2) ? cd : 0u;
    }
    return pack_profile(cn - c1 - c2 + c12, c1 - c12, c12, c2 - c12);
}

// The bases field of a line window by window (a loop of ceil(field / 64) steps; the first window is the one that holds
// the header).  visit(w, live) sees every window once, in text order, with the bytes of the field that the grammar
// looks at ('^'-hidden and indel-skipped bytes removed).  *end_bit receives the bit index of the byte that ends the field.
template <class Visit>
SID_HD bool win_bases_walk(const uint8_t* s, uint32_t region_off, const uint32_t* cw, uint32_t n_bits, const WinHeader& hd, Win64 w, bool ok,
                           Visit& visit, uint32_t* end_bit) {
    uint32_t cur = hd.l0;                       // bit index of the window
    uint64_t below = (2ull << hd.q4) - 1ull;    // bits of the window that precede the field
    uint32_t skip = 0;                          // bytes at the start of the next window still covered by a '^' or an indel
    bool running = ok;
    while (running) {
        const uint64_t t = w.term & ~below;
        const bool last = t != 0;
        const uint32_t e = last ? (uint32_t)ctz64(t) : 64u;
        uint64_t live_all = low_bits64(e) & ~below;             // the bytes of this window that belong to the field
        if (skip) {
            const uint32_t sk = skip < 64 ? skip : 64;
            live_all &= ~low_bits64(sk);
            skip -= sk;
        }
        uint32_t next = cur + 64;
        uint64_t pmw = w.pm, live;
        for (;;) {
            // '^' hides the byte after it (pileup.cpp:125-127); "^^": leave the parity to the byte-wise path
            const uint64_t car = w.caret & live_all;
            if (car & (car << 1)) ok = false;
            live = live_all & ~(car << 1);
            const uint64_t pv = pmw & live;
            if (!pv || !ok) break;
            // '+' / '-' (pileup.cpp:128-147): the digits after it are read from the raw text
            const uint32_t p = (uint32_t)ctz64(pv);
            pmw &= ~(1ull << p);
            const uint64_t dg = p == 63 ? 0ull : (w.digit >> (p + 1));
            const uint32_t ndig = dg == ~0ull ? 64u : (uint32_t)ctz64(~dg);
            if (p + 1 + ndig >= 64) {
                // the sign or its number touches the end of the window: count what precedes it, restart there
                if (p == 0) { ok = false; break; }
                live_all &= low_bits64(p);
                next = cur + p;
                skip = 0;
                continue;
            }
            if (ndig == 0) continue;                            // a sign without a digit is ignored (pileup.cpp:131-133)
            uint32_t n = 0;
            const uint8_t* q = s + region_off + cur + p + 1;
            for (uint32_t i = 0; i < ndig; ++i)
                if (n < (1u << 26)) n = n * 10 + ((uint32_t)q[i] - (uint32_t)'0');
            const uint64_t to = (uint64_t)p + 1 + ndig + n;     // first byte after the skipped ones (pileup.cpp:144)
            live_all &= ~(low_bits64(to >= 64 ? 64u : (uint32_t)to) & ~low_bits64(p + 1));
            if (to > 64 && !last) skip = (uint32_t)(to - 64 > (1u << 27) ? (1u << 27) : to - 64);
        }
        if (!last && next == cur + 64 && ((w.caret & live_all) >> 63)) skip = 1;   // the hidden byte opens the next window
        visit(w, live);
        if (last || !ok) {
            running = false;
            if (end_bit) *end_bit = cur + e;
        } else {
            cur = next;
            below = 0;
            if (cur + 64 > n_bits) { ok = false; running = false; }        // ran out of classified bytes
            else w = load_window(cw, cur);
        }
    }
    return ok;
}

// The tokenizer's visitor: counts by the bit planes that tell A/C/G/T apart, and '.'/','.
struct WinCounts {
    uint32_t cn = 0, c1 = 0, c2 = 0, c12 = 0, cd = 0;
    SID_HD void operator()(const Win64& w, uint64_t live) {
        const uint64_t b = w.base & live;
        cn += popc64(b);
        c1 += popc64(b & w.p1);
        c2 += popc64(b & w.p2);
        c12 += popc64(b & w.p1 & w.p2);
        cd += popc64(w.dot & live);
    }
};

// One line per lane: header, then the bases field.
template <bool WANT_POS>
SID_HD bool parse_line_win(const uint8_t* s, uint32_t region_off, const uint32_t* cw, const uint32_t* nlw, uint32_t n_bits,
                           uint32_t line_off, WinLine& o, WinHeader* hd_out = nullptr, uint32_t* end_bit = nullptr) {
    WinHeader hd;
    Win64 w;
    bool ok = win_header<WANT_POS>(s, region_off, cw, nlw, n_bits, line_off, o, hd, w);
    SID_SYNCWARP();
    WinCounts c;
    ok = win_bases_walk(s, region_off, cw, n_bits, hd, w, ok, c, end_bit);
    SID_SYNCWARP();
    o.profile = win_profile(c.cn, c.c1, c.c2, c.c12, c.cd, hd);
    o.status = LINE_OK;
    if (hd_out) *hd_out = hd;
    return ok;
}

// ---- long lines: one WINDOW per lane ----------------------------------------------------------------------------
// A deep pileup (depth 500: 1.2 KB per line) leaves the loop above on three or four lanes for nineteen steps.  The
// kernel then turns the field into windows of 64 bytes on a fixed grid from the field's first byte and gives every
// window its own lane (win_window), all lines of the slice at once.  What a window needs from its predecessor is one
// number: how many of its first bytes a '^' or an indel that began earlier still covers (skip_in).  Lanes start with
// skip_in = 0, compare with what the lane before them reports (skip_out) and redo their window when it differs; a
// lane is final one round after its predecessor, in practice after one or two rounds.

// First byte <= 0x20 at or after bit `from` (the end of the bases field that starts there); 0xFFFFFFFF when the
// classified bytes end first.
// term_groups (optional): bit i of word g says that unit 32 g + i holds such a byte (stage 1 has them as ballots);
// with it the search skips 32 units at a time.
SID_HD uint32_t win_field_end(const uint32_t* cw, uint32_t n_bits, uint32_t from, const uint32_t* term_groups = nullptr) {
    uint32_t u = from >> 5;
    uint32_t t = cw[(size_t)u * CW_WORDS + CW_TERM] & (0xFFFFFFFFu << (from & 31));
    const uint32_t n_units = n_bits >> 5;
    if (t == 0 && term_groups) {
        ++u;
        if (u >= n_units) return 0xFFFFFFFFu;
        uint32_t g = u >> 5;
        uint32_t m = term_groups[g] & (0xFFFFFFFFu << (u & 31));
        const uint32_t n_groups = (n_units + 31) >> 5;
        while (m == 0) {
            if (++g >= n_groups) return 0xFFFFFFFFu;
            m = term_groups[g];
        }
        u = g * 32 + first_bit(m);
        if (u >= n_units) return 0xFFFFFFFFu;
        t = cw[(size_t)u * CW_WORDS + CW_TERM];
    }
    while (t == 0) {
        if (++u >= n_units) return 0xFFFFFFFFu;
        t = cw[(size_t)u * CW_WORDS + CW_TERM];
    }
    return u * 32 + first_bit(t);
}

struct WinPart {
    uint32_t cn, c1, c2, c12, cd;   // counted bases of the window, by the bit planes that tell them apart, and '.'/','
    uint32_t skip_out;              // bytes at the start of the next window that this one's '^' / indels cover
    bool ok;
};

// The window [pos, pos + 64) of a bases field that ends at bit `end` (> pos); skip_in: bytes at its start covered
// from before.  Numbers that run past the window's end are read from the text.
SID_HD WinPart win_window_of(const Win64& w, bool in_reach, const uint8_t* s, uint32_t region_off, uint32_t pos, uint32_t end, uint32_t skip_in) {
    WinPart r;
    r.cn = r.c1 = r.c2 = r.c12 = r.cd = 0;
    r.skip_out = 0;
    r.ok = in_reach;
    const uint32_t len = end - pos < 64 ? end - pos : 64;
    uint64_t live_all = low_bits64(len);
    if (skip_in) {
        live_all &= ~low_bits64(skip_in < 64 ? skip_in : 64);
        if (skip_in > 64) r.skip_out = skip_in - 64;
    }
    uint64_t pmw = w.pm, live;
    for (;;) {
        const uint64_t car = w.caret & live_all;                // pileup.cpp:125-127
        if (car & (car << 1)) r.ok = false;                     // "^^": the byte-wise path keeps the parity
        live = live_all & ~(car << 1);
        const uint64_t pv = pmw & live;
        if (!pv || !r.ok) break;
        const uint32_t p = (uint32_t)ctz64(pv);                 // '+' / '-' (pileup.cpp:128-147)
        pmw &= ~(1ull << p);
        const uint64_t dg = p == 63 ? 0ull : (w.digit >> (p + 1));
        uint32_t ndig = dg == ~0ull ? 64u : (uint32_t)ctz64(~dg);
        if (p + 1 + ndig >= 64) {
            // the digits reach the end of the window: the number continues in the text (if the field does)
            ndig = 63 - p;
            uint32_t extra = 0;
            while (pos + 64 + extra < end && extra < 12 && (uint32_t)s[region_off + pos + 64 + extra] - (uint32_t)'0' <= 9u) ++extra;
            ndig += extra;
        }
        if (ndig == 0) continue;                                // a sign without a digit is ignored (pileup.cpp:131-133)
        if (ndig > 10) { r.ok = false; break; }
        uint32_t n = 0;
        const uint8_t* q = s + region_off + pos + p + 1;
        for (uint32_t i = 0; i < ndig; ++i)
            if (n < (1u << 26)) n = n * 10 + ((uint32_t)q[i] - (uint32_t)'0');
        const uint64_t to = (uint64_t)p + 1 + ndig + n;         // first byte after the skipped ones (pileup.cpp:144)
        live_all &= ~(low_bits64(to >= 64 ? 64u : (uint32_t)to) & ~low_bits64(p + 1));
        if (to > 64) {
            const uint32_t over = (uint32_t)(to - 64 > (1u << 27) ? (1u << 27) : to - 64);
            if (over > r.skip_out) r.skip_out = over;
        }
    }
    if (((w.caret & live_all) >> 63) && r.skip_out == 0) r.skip_out = 1;       // the hidden byte opens the next window
    const uint64_t b = w.base & live;
    r.cn = popc64(b);
    r.c1 = popc64(b & w.p1);
    r.c2 = popc64(b & w.p2);
    r.c12 = popc64(b & w.p1 & w.p2);
    r.cd = popc64(w.dot & live);
    return r;
}

SID_HD WinPart win_window(const uint8_t* s, uint32_t region_off, const uint32_t* cw, uint32_t n_bits, uint32_t pos, uint32_t end,
                          uint32_t skip_in) {
    const bool in_reach = pos + 64 <= n_bits;
    const Win64 w = load_window(cw, in_reach ? pos : 0);
    return win_window_of(w, in_reach, s, region_off, pos, end, skip_in);
}

// The result for skip_in = s from the result r0 for skip_in = 0 of the same window, when the first s bytes hold neither
// a '^' nor a sign (then they only ever contributed counts, and nothing after them changes).  Returns false when the
// window has to be evaluated again.
SID_HD bool win_window_patch(const Win64& w, uint32_t pos, uint32_t end, const WinPart& r0, uint32_t s, WinPart& r) {
    const uint32_t len = end - pos < 64 ? end - pos : 64;
    if (s >= len) return false;
    const uint64_t low = low_bits64(s);
    if ((w.caret | w.pm) & low) return false;
    const uint64_t b = w.base & low;
    r = r0;
    r.cn -= popc64(b);
    r.c1 -= popc64(b & w.p1);
    r.c2 -= popc64(b & w.p2);
    r.c12 -= popc64(b & w.p1 & w.p2);
    r.cd -= popc64(w.dot & low);
    return true;
}

#if !defined(__CUDACC__)
// Host check: the line at p (plus slack) classified like the kernel's stage 1.
struct HostLineClasses {
    const uint8_t* scratch;     // the bytes that were classified; scratch[0] is 32..63 bytes before the line
    const uint32_t* cw;
    const uint32_t* nlw;
    uint32_t units;
    uint32_t line_off;          // offset of the line in scratch
    bool usable;                // false: the line is too long for the scratch arrays or holds a control byte
};
inline HostLineClasses classify_line_host(const uint8_t* text, uint64_t len, uint64_t p) {
    HostLineClasses h {nullptr, nullptr, nullptr, 0, 0, false};
    const int64_t first = (int64_t)(p & ~(uint64_t)31) - 32;
    uint64_t end = p;
    while (end < len && text[end] != '\n') ++end;
    const uint64_t avail64 = (((int64_t)end - first) + 256 + 31) & ~(uint64_t)31;
    if (avail64 > (1u << 20)) return h;
    static thread_local uint8_t scratch[(1u << 20) + 64] __attribute__((aligned(16)));
    static thread_local uint32_t cw[((1u << 15) + 8) * CW_WORDS], nlw[(1u << 15) + 8];
    for (uint64_t k = 0; k < avail64; ++k) {
        const int64_t q = first + (int64_t)k;
        scratch[k] = (q >= 0 && (uint64_t)q < len) ? text[q] : (uint8_t)'\n';
    }
    const uint32_t units = (uint32_t)(avail64 / 32);
    uint32_t bad = 0;
    for (uint32_t u = 0; u < units; ++u) {
        uint32_t w[8];
        memcpy(w, scratch + 32 * u, 32);
        const UnitClasses k = classify_unit(w);
        for (int c = 0; c < CW_WORDS; ++c) cw[(size_t)u * CW_WORDS + c] = k.w[c];
        nlw[u] = k.nl;
        // only the bytes of this line matter for the refusal
        for (int i = 0; i < 32; ++i) {
            const int64_t q = first + 32 * (int64_t)u + i;
            if (q >= (int64_t)p && q < (int64_t)end && ((k.bad >> i) & 1u)) bad = 1;
        }
    }
    for (uint32_t u = units; u < units + CW_PAD_UNITS; ++u) {
        for (int c = 0; c < CW_WORDS; ++c) cw[(size_t)u * CW_WORDS + c] = 0;
        nlw[u] = 0;
    }
    h.scratch = scratch; h.cw = cw; h.nlw = nlw; h.units = units;
    h.line_off = (uint32_t)((int64_t)p - first);
    h.usable = bad == 0;
    return h;
}

// Runs stage 2 on the classified line.  A line with a control byte (UnitClasses::bad) is refused here as the kernel
// refuses its whole slice.
template <bool WANT_POS, bool COOP = false>
inline bool parse_line_win_host(const uint8_t* text, uint64_t len, uint64_t p, WinLine& o) {
    const HostLineClasses h = classify_line_host(text, len, p);
    if (!h.usable) return false;
    const uint8_t* scratch = h.scratch;
    const uint32_t* cw = h.cw;
    const uint32_t* nlw = h.nlw;
    const uint32_t units = h.units;
    const int64_t first = (int64_t)p - (int64_t)h.line_off;
    if (!COOP) return parse_line_win<WANT_POS>(scratch, 0, cw, nlw, units * 32, (uint32_t)((int64_t)p - first), o);
    // the window-per-lane form, windows in order: skip_in of a window is the final skip_out of the one before it,
    // which is the fixed point the lanes of the kernel converge to
    WinHeader hd;
    Win64 w0;
    if (!win_header<WANT_POS>(scratch, 0, cw, nlw, units * 32, (uint32_t)((int64_t)p - first), o, hd, w0)) return false;
    const uint32_t a = hd.l0 + hd.q4 + 1;
    static thread_local uint32_t groups[((1u << 15) + 8) / 32 + 2];
    for (uint32_t g = 0; g < (units + 31) / 32; ++g) {
        groups[g] = 0;
        for (uint32_t i = 0; i < 32 && g * 32 + i < units; ++i)
            if (cw[(size_t)(g * 32 + i) * CW_WORDS + CW_TERM]) groups[g] |= 1u << i;
    }
    const uint32_t b = win_field_end(cw, units * 32, a, groups);
    if (b != win_field_end(cw, units * 32, a)) return false;            // the two searches agree (the caller counts this as a failure)
    if (b == 0xFFFFFFFFu || b <= a) return false;
    uint32_t cn = 0, c1 = 0, c2 = 0, c12 = 0, cd = 0, skip = 0;
    for (uint32_t pos = a; pos < b; pos += 64) {
        const WinPart r = win_window(scratch, 0, cw, units * 32, pos, b, skip);
        if (!r.ok) return false;
        cn += r.cn; c1 += r.c1; c2 += r.c2; c12 += r.c12; cd += r.cd;
        skip = r.skip_out;
    }
    o.profile = win_profile(cn, c1, c2, c12, cd, hd);
    o.status = LINE_OK;
    return true;
}
#endif

}  // namespace sid
