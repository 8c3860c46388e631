// DEFLATE (RFC 1951) on the device, for BGZF input (device half): the pileups of the reference's
// pipeline are stored gzipped (scripts/prepare-data.sh:14) and inflated by `zcat` on one core before sid sees a byte
// (scripts/sid-pipeline/run-sid.sh:15).  A BGZF file (`bgzip`, samtools) is a series of independent gzip members of at most
// 64 KiB of text each; the host only walks their headers (bgzf_scan), the compressed bytes cross the link (a third to a
// quarter of the text) and every member is inflated here by ONE WARP:
//   * lane 0 owns the bit reader and the Huffman decoding (canonical codes; a 10-bit / 8-bit first-level table per
//     alphabet in shared memory, the rare longer codes by the count/offset walk of the canonical order);
//   * runs of literals are stored by lane 0 as it decodes them; a match is broadcast and copied by all 32 lanes
//     (out[p + k] = out[p - dist + k mod dist]: every byte comes from text that existed before the match began);
//   * stored and fixed-Huffman blocks included; every loop is bounded by the member's compressed and inflated sizes
//     (a damaged member is an error code, never a hang); ISIZE is checked, the CRC-32 is not (documented in DESIGN.md).
// The table builder and the symbol decoder are SID_HD: tests/hostcheck inflates whole files with them on the CPU and
// compares with zlib byte for byte.
#pragma once
#include "common.cuh"

namespace sid {

struct BgzfBlock {          // one gzip member of a BGZF file (filled by bgzf_scan on the host)
    uint64_t c_off;         // offset of its deflate stream in the compressed buffer
    uint64_t out_off;       // offset of its text in the output buffer
    uint32_t c_len;         // bytes of the deflate stream
    uint32_t isize;         // bytes of text (ISIZE of the trailer)
};

enum : int { INF_OK = 0, INF_BAD_BLOCK_TYPE = 1, INF_BAD_STORED = 2, INF_BAD_LENGTHS = 3, INF_BAD_SYMBOL = 4, INF_BAD_DISTANCE = 5,
             INF_OUTPUT_OVERRUN = 6, INF_INPUT_OVERRUN = 7, INF_SIZE_MISMATCH = 8 };

constexpr int INF_LIT_BITS = 10, INF_DIST_BITS = 8;

struct InflateTables {
    uint16_t lit_fast[1 << INF_LIT_BITS];   // (symbol << 4) | length for codes of at most INF_LIT_BITS bits, else 0
    uint16_t dist_fast[1 << INF_DIST_BITS];
    uint16_t lit_sym[288], dist_sym[32];    // symbols in canonical order (by length, then by value)
    uint16_t lit_cnt[16], dist_cnt[16];     // codes per length
};

// The input as a little-endian bit stream: two consecutive aligned words and a bit offset; the next 32 bits of the stream are
// one funnel shift away at any time.  The buffer must be readable up to 8 bytes past its end (the callers pad).
struct BitReader {
    const uint32_t* wp;     // the word after `hi`
    const uint32_t* end;    // first word that holds no byte of the stream
    uint32_t lo, hi;        // the words that hold the next bits
    uint32_t p;             // offset of the next bit in `lo` (0..31)
    bool over;              // the stream was read past its end

    SID_HD void init(const uint8_t* in, uint32_t len) {
        const uintptr_t a = (uintptr_t)in;
        wp = (const uint32_t*)(a & ~(uintptr_t)3);
        end = (const uint32_t*)(((a + len) + 3) & ~(uintptr_t)3);
        p = (uint32_t)(a & 3) * 8;
        lo = *wp++;
        hi = *wp++;
        over = false;
    }
    SID_HD uint32_t window() const {        // the next 32 bits
#if defined(__CUDA_ARCH__)
        return __funnelshift_r(lo, hi, p);
#else
        return (uint32_t)((((uint64_t)hi << 32) | lo) >> p);
#endif
    }
    SID_HD uint32_t peek(uint32_t n) const { return window() & ((1u << n) - 1u); }
    SID_HD void drop(uint32_t n) {          // n <= 32
        p += n;
        if (p >= 32) {
            p -= 32;
            lo = hi;
            if (wp > end) {                             // `hi` may be the word after the stream, never a later one
                hi = 0;
                if (wp > end + 1) over = true;          // `lo` now lies wholly beyond the stream
                ++wp;
            } else hi = *wp++;
        }
    }
    SID_HD uint32_t take(uint32_t n) { const uint32_t v = peek(n); drop(n); return v; }     // n <= 16
    SID_HD void to_byte_boundary() { drop((8u - (p & 7u)) & 7u); }
    // bits consumed beyond the stream's last byte?  (exact: position of the next bit against the stream's length)
    SID_HD bool overrun(const uint8_t* in, uint32_t len) const {
        const uintptr_t first = (uintptr_t)in & ~(uintptr_t)3;
        const uint64_t pos = ((uint64_t)((uintptr_t)wp - first) - 8) * 8 + p;       // bit offset of the next bit from `first`
        return over || pos > ((uint64_t)((uintptr_t)in - first) + len) * 8;
    }
};

SID_HD uint32_t bit_reverse(uint32_t v, uint32_t n) {      // the low n bits of v, reversed
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// Canonical Huffman code of `n` symbols with the given lengths (0 = unused): fast table of `bits` index bits, symbols in
// canonical order, codes per length.  Returns false for an over-subscribed set of lengths (an incomplete one is accepted
// like zlib accepts a single distance code; its unused codes decode to "bad symbol").
SID_HD bool build_huffman(const uint8_t* lengths, uint32_t n, uint16_t* fast, uint32_t bits, uint16_t* sym, uint16_t* cnt) {
    for (int l = 0; l < 16; ++l) cnt[l] = 0;
    for (uint32_t s = 0; s < n; ++s) ++cnt[lengths[s]];
    cnt[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left = (left << 1) - (int)cnt[l];
        if (left < 0) return false;
    }
    uint16_t offs[16], next[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + cnt[l]);
    uint32_t code = 0;
    for (int l = 1; l < 16; ++l) {
        code = (code + cnt[l - 1]) << 1;
        next[l] = (uint16_t)code;
    }
    for (uint32_t i = 0; i < (1u << bits); ++i) fast[i] = 0;
    for (uint32_t s = 0; s < n; ++s) {
        const uint32_t l = lengths[s];
        if (l == 0) continue;
        sym[offs[l]++] = (uint16_t)s;
        const uint32_t c = next[l]++;
        if (l <= bits) {
            const uint16_t e = (uint16_t)((s << 4) | l);
            for (uint32_t i = bit_reverse(c, l); i < (1u << bits); i += 1u << l) fast[i] = e;
        }
    }
    return true;
}

// Next symbol of the alphabet (fast, bits, cnt, sym); -1 for a code that is not in it.
SID_HD int decode_symbol(BitReader& br, const uint16_t* fast, uint32_t bits, const uint16_t* cnt, const uint16_t* sym) {
    const uint32_t w = br.window();
    const uint32_t e = fast[w & ((1u << bits) - 1u)];
    if (e) {
        br.drop(e & 15u);
        return (int)(e >> 4);
    }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        code |= (int)((w >> (l - 1)) & 1u);
        const int count = cnt[l];
        if (code - count < first) {
            br.drop((uint32_t)l);
            return sym[index + (code - first)];
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

SID_HD uint32_t length_base(uint32_t i) {       // length symbols 257..285 -> i = 0..28
    return i < 8 ? 3 + i : i == 28 ? 258 : ((4 + (i & 3)) << ((i >> 2) - 1)) + 3;
}
SID_HD uint32_t length_extra(uint32_t i) { return i < 8 || i == 28 ? 0 : (i >> 2) - 1; }
SID_HD uint32_t dist_base(uint32_t i) {         // distance symbols 0..29
    return i < 4 ? 1 + i : ((2 + (i & 1)) << ((i >> 1) - 1)) + 1;
}
SID_HD uint32_t dist_extra(uint32_t i) { return i < 4 ? 0 : (i >> 1) - 1; }

// Header of a dynamic block (RFC 1951 3.2.7) or the fixed code (3.2.6) -> the two decoding tables.
SID_HD int read_block_tables(BitReader& br, bool fixed, InflateTables& t) {
    uint8_t lengths[320];
    uint32_t nlen = 288, ndist = 30;
    if (fixed) {
        for (int s = 0; s < 144; ++s) lengths[s] = 8;
        for (int s = 144; s < 256; ++s) lengths[s] = 9;
        for (int s = 256; s < 280; ++s) lengths[s] = 7;
        for (int s = 280; s < 288; ++s) lengths[s] = 8;
        for (int s = 0; s < 30; ++s) lengths[288 + s] = 5;
    } else {
        nlen = br.take(5) + 257;
        ndist = br.take(5) + 1;
        const uint32_t ncode = br.take(4) + 4;
        if (nlen > 286 || ndist > 30) return INF_BAD_LENGTHS;
        uint8_t cl[19];
        for (int i = 0; i < 19; ++i) cl[i] = 0;
        for (uint32_t i = 0; i < ncode; ++i) {
                // order of the code length code lengths: 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
            const uint32_t pos = i < 3 ? 16 + i : i == 3 ? 0 : (i & 1) ? (19 - i) / 2 : 6 + i / 2;
            cl[pos] = (uint8_t)br.take(3);
        }
        // the code length alphabet reuses the distance arrays (they are rebuilt right after)
        if (!build_huffman(cl, 19, t.dist_fast, 7, t.dist_sym, t.dist_cnt)) return INF_BAD_LENGTHS;
        uint32_t i = 0;
        while (i < nlen + ndist) {
            const int s = decode_symbol(br, t.dist_fast, 7, t.dist_cnt, t.dist_sym);
            if (s < 0) return INF_BAD_LENGTHS;
            if (br.over) return INF_INPUT_OVERRUN;
            if (s < 16) { lengths[i++] = (uint8_t)s; continue; }
                uint32_t rep, val = 0;
            if (s == 16) {
                if (i == 0) return INF_BAD_LENGTHS;
                val = lengths[i - 1];
                rep = 3 + br.take(2);
            } else if (s == 17) rep = 3 + br.take(3);
            else rep = 11 + br.take(7);
            if (i + rep > nlen + ndist) return INF_BAD_LENGTHS;
            while (rep--) lengths[i++] = (uint8_t)val;
        }
        if (lengths[256] == 0) return INF_BAD_LENGTHS;          // no end-of-block code
    }
    if (!build_huffman(lengths, nlen, t.lit_fast, INF_LIT_BITS, t.lit_sym, t.lit_cnt)) return INF_BAD_LENGTHS;
    if (!build_huffman(lengths + nlen, ndist, t.dist_fast, INF_DIST_BITS, t.dist_sym, t.dist_cnt)) return INF_BAD_LENGTHS;
    return INF_OK;
}

// What lane 0 hands to the warp: literals it has already stored, then one of match / end of block / error.
enum : uint32_t { EV_MATCH = 0, EV_END = 1, EV_ERROR = 2 };

// Decodes symbols until something other than a literal comes; literals go straight to out[pos...].  Returns the event;
// *n_lit literals were stored, for EV_MATCH *len and *dist are set, for EV_ERROR *len is the code.
SID_HD uint32_t decode_run(BitReader& br, const InflateTables& t, uint8_t* out, uint32_t pos, uint32_t out_len, uint32_t* n_lit,
                           uint32_t* len, uint32_t* dist) {
    uint32_t n = 0;
    for (;;) {
        const int s = decode_symbol(br, t.lit_fast, INF_LIT_BITS, t.lit_cnt, t.lit_sym);
        if (s < 0) { *n_lit = n; *len = INF_BAD_SYMBOL; return EV_ERROR; }
        if (br.over) { *n_lit = n; *len = INF_INPUT_OVERRUN; return EV_ERROR; }
        if (s < 256) {
            if (pos + n >= out_len) { *n_lit = n; *len = INF_OUTPUT_OVERRUN; return EV_ERROR; }
            out[pos + n] = (uint8_t)s;
            ++n;
            continue;
        }
        *n_lit = n;
        if (s == 256) return EV_END;
        const uint32_t li = (uint32_t)s - 257;
        if (li > 28) { *len = INF_BAD_SYMBOL; return EV_ERROR; }
        const uint32_t l = length_base(li) + br.take(length_extra(li));
        const int d = decode_symbol(br, t.dist_fast, INF_DIST_BITS, t.dist_cnt, t.dist_sym);
        if (d < 0 || d > 29) { *len = INF_BAD_DISTANCE; return EV_ERROR; }
        const uint32_t dd = dist_base((uint32_t)d) + br.take(dist_extra((uint32_t)d));
        if (dd > pos + n) { *len = INF_BAD_DISTANCE; return EV_ERROR; }
        if (pos + n + l > out_len) { *len = INF_OUTPUT_OVERRUN; return EV_ERROR; }
        *len = l;
        *dist = dd;
        return EV_MATCH;
    }
}

// One symbol.  Returns STEP_LITERAL (stored at out[pos]), STEP_MATCH (*len, *dist set, nothing copied yet), STEP_END (end of
// block) or STEP_ERROR (*len = code).
enum : uint32_t { STEP_LITERAL = 0, STEP_MATCH = 1, STEP_END = 2, STEP_ERROR = 3 };
SID_HD uint32_t decode_step(BitReader& br, const InflateTables& t, uint8_t* out, uint32_t pos, uint32_t out_len, uint32_t* len, uint32_t* dist) {
    const int s = decode_symbol(br, t.lit_fast, INF_LIT_BITS, t.lit_cnt, t.lit_sym);
    if (s < 0 || br.over) { *len = s < 0 ? INF_BAD_SYMBOL : INF_INPUT_OVERRUN; return STEP_ERROR; }
    if (s < 256) {
        if (pos >= out_len) { *len = INF_OUTPUT_OVERRUN; return STEP_ERROR; }
        out[pos] = (uint8_t)s;
        return STEP_LITERAL;
    }
    if (s == 256) return STEP_END;
    const uint32_t li = (uint32_t)s - 257;
    if (li > 28) { *len = INF_BAD_SYMBOL; return STEP_ERROR; }
    const uint32_t l = length_base(li) + br.take(length_extra(li));
    const int d = decode_symbol(br, t.dist_fast, INF_DIST_BITS, t.dist_cnt, t.dist_sym);
    if (d < 0 || d > 29) { *len = INF_BAD_DISTANCE; return STEP_ERROR; }
    const uint32_t dd = dist_base((uint32_t)d) + br.take(dist_extra((uint32_t)d));
    if (dd > pos) { *len = INF_BAD_DISTANCE; return STEP_ERROR; }
    if (pos + l > out_len) { *len = INF_OUTPUT_OVERRUN; return STEP_ERROR; }
    *len = l;
    *dist = dd;
    return STEP_MATCH;
}

// Block header: BFINAL, BTYPE; a stored block is copied right here.  *final_block, *kind (0 stored: done, 1/2: tables built).
SID_HD int begin_block(BitReader& br, InflateTables& t, uint8_t* out, uint32_t* pos, uint32_t out_len, bool* final_block, uint32_t* kind) {
    *final_block = br.take(1) != 0;
    *kind = br.take(2);
    if (*kind == 3) return INF_BAD_BLOCK_TYPE;
    if (*kind == 0) {
        br.to_byte_boundary();
        const uint32_t n = br.take(16);
        const uint32_t nn = br.take(16);
        if ((n ^ nn) != 0xFFFFu) return INF_BAD_STORED;
        if (*pos + n > out_len) return INF_OUTPUT_OVERRUN;
        for (uint32_t i = 0; i < n; ++i) out[*pos + i] = (uint8_t)br.take(8);
        *pos += n;
        return br.over ? INF_INPUT_OVERRUN : INF_OK;
    }
    return read_block_tables(br, *kind == 1, t);
}

// One member on one thread (tests/hostcheck; the kernel below is the same walk with the matches copied by the warp).
SID_HD int inflate_member_steps(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, InflateTables& t) {
    BitReader br;
    br.init(in, in_len);
    uint32_t pos = 0;
    for (uint32_t guard = 0; guard < (1u << 20); ++guard) {
        bool final_block;
        uint32_t kind;
        const int rc = begin_block(br, t, out, &pos, out_len, &final_block, &kind);
        if (rc != INF_OK) return rc;
        if (kind != 0) {
            for (;;) {
                uint32_t len = 0, dist = 0;
                const uint32_t ev = decode_step(br, t, out, pos, out_len, &len, &dist);
                if (ev == STEP_ERROR) return (int)len;
                if (ev == STEP_END) break;
                if (ev == STEP_LITERAL) { ++pos; continue; }
                for (uint32_t k = 0; k < len; ++k) out[pos + k] = (out + pos - dist)[dist >= len ? k : k % dist];
                pos += len;
            }
        }
        if (final_block) return pos == out_len ? (br.overrun(in, in_len) ? INF_INPUT_OVERRUN : INF_OK) : INF_SIZE_MISMATCH;
    }
    return INF_BAD_BLOCK_TYPE;
}

SID_HD int inflate_member_serial(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, InflateTables& t) {
    BitReader br;
    br.init(in, in_len);
    uint32_t pos = 0;
    for (uint32_t guard = 0; guard < (1u << 20); ++guard) {
        bool final_block;
        uint32_t kind;
        const int rc = begin_block(br, t, out, &pos, out_len, &final_block, &kind);
        if (rc != INF_OK) return rc;
        if (kind != 0) {
            for (;;) {
                uint32_t n_lit, len = 0, dist = 0;
                const uint32_t ev = decode_run(br, t, out, pos, out_len, &n_lit, &len, &dist);
                pos += n_lit;
                if (ev == EV_ERROR) return (int)len;
                if (ev == EV_END) break;
                for (uint32_t k = 0; k < len; ++k) out[pos + k] = out[pos + k - dist];
                pos += len;
            }
        }
        if (final_block) return pos == out_len ? (br.overrun(in, in_len) ? INF_INPUT_OVERRUN : INF_OK) : INF_SIZE_MISMATCH;
    }
    return INF_BAD_BLOCK_TYPE;
}

#if defined(__CUDACC__)

constexpr int INF_WARPS = 8;        // members per CTA (one warp each; 3.3 KB of tables per warp)

// One warp per member, members dealt round robin.  error: (member index << 4 | code), the smallest wins.
__global__ void __launch_bounds__(INF_WARPS * 32) k_inflate_bgzf(const uint8_t* comp, const BgzfBlock* blocks, uint32_t n_blocks, uint8_t* text,
                                                                   unsigned long long* error) {
    __shared__ InflateTables s_tables[INF_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    InflateTables& t = s_tables[warp];
    for (uint32_t m = blockIdx.x * INF_WARPS + warp; m < n_blocks; m += gridDim.x * INF_WARPS) {
        const BgzfBlock b = blocks[m];
        uint8_t* out = text + b.out_off;
        const uint32_t out_len = b.isize;
        BitReader br;
        if (lane == 0) br.init(comp + b.c_off, b.c_len);
        uint32_t pos = 0;
        int rc = INF_OK;
        bool done = out_len == 0 && b.c_len == 0;
        for (uint32_t guard = 0; !done && guard < (1u << 20); ++guard) {
            bool final_block = false;
            uint32_t kind = 0;
            if (lane == 0) rc = begin_block(br, t, out, &pos, out_len, &final_block, &kind);
            rc = __shfl_sync(0xFFFFFFFFu, rc, 0);
            if (rc != INF_OK) break;
            pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
            kind = __shfl_sync(0xFFFFFFFFu, kind, 0);
            final_block = __shfl_sync(0xFFFFFFFFu, (int)final_block, 0) != 0;
            if (kind != 0) {
                for (;;) {
                    uint32_t n_lit = 0, len = 0, dist = 0, ev = EV_END;
                    if (lane == 0) ev = decode_run(br, t, out, pos, out_len, &n_lit, &len, &dist);
                    // one packet: event (2 bits), literals (17 bits: at most 65536)
                    const uint32_t head = __shfl_sync(0xFFFFFFFFu, ev | (n_lit << 2), 0);
                    const uint32_t ld = __shfl_sync(0xFFFFFFFFu, len | (dist << 16), 0);     // dist <= 32768, len <= 258
                    ev = head & 3u;
                    pos += head >> 2;
                    if (ev == EV_ERROR) { rc = (int)(ld & 0xFFFFu); break; }
                    if (ev == EV_END) break;
                    len = ld & 0xFFFFu;
                    dist = ld >> 16;
                    __syncwarp();                               // lane 0's literals are visible to the lanes that copy
                    const uint8_t* src = out + pos - dist;
                    if (dist >= len) {
                        for (uint32_t k = lane; k < len; k += 32) out[pos + k] = src[k];
                    } else {
                        for (uint32_t k = lane; k < len; k += 32) out[pos + k] = src[k % dist];
                    }
                    __syncwarp();
                    pos += len;
                }
                if (rc != INF_OK) break;
            }
            if (final_block) {
                int over = 0;
                if (lane == 0) over = br.overrun(comp + b.c_off, b.c_len) ? 1 : 0;
                over = __shfl_sync(0xFFFFFFFFu, over, 0);
                rc = pos != out_len ? INF_SIZE_MISMATCH : over ? INF_INPUT_OVERRUN : INF_OK;
                done = true;
            }
        }
        if (!done && rc == INF_OK) rc = INF_BAD_BLOCK_TYPE;
        if (rc != INF_OK && lane == 0) atomicMin(error, ((unsigned long long)m << 4) | (unsigned long long)rc);
        __syncwarp();
    }
}

// Lockstep form: a warp inflates 32 / SW members at once.  Lane 0 of every group of SW lanes (the leader) owns one member's
// bit reader; every trip of the warp's loop each leader decodes ONE symbol -- the same instructions for all of them --
// and the groups that met a match copy it with their SW lanes.  Block headers (table building) run on the leaders that
// need them while the others wait; members are dealt group by group, a group that is done with its member takes the
// next one of its warp's share.
constexpr int INF2_WARPS = 4;
template <int SW>
__global__ void __launch_bounds__(INF2_WARPS * 32) k_inflate_bgzf_lockstep(const uint8_t* comp, const BgzfBlock* blocks, uint32_t n_blocks, uint8_t* text,
                                                                            unsigned long long* error) {
    constexpr int G = 32 / SW;
    extern __shared__ __align__(16) uint8_t s_inf[];
    InflateTables* const tables = reinterpret_cast<InflateTables*>(s_inf);
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t sub = lane / SW, sl = lane % SW;
    const bool leader = sl == 0;
    InflateTables& t = tables[warp * G + sub];
    enum : int { ST_NEXT = 0, ST_HEADER = 1, ST_SYMBOLS = 2, ST_DONE = 3 };
    // members: decoder d of the grid takes d, d + D, d + 2 D, ...
    const uint32_t n_dec = gridDim.x * INF2_WARPS * G;
    uint32_t m = (blockIdx.x * INF2_WARPS + warp) * G + sub;
    int st = ST_NEXT;
    BitReader br;
    br.wp = br.end = nullptr; br.lo = br.hi = br.p = 0; br.over = false;
    uint8_t* out = text;
    uint32_t out_len = 0, pos = 0, c_len = 0;
    const uint8_t* in = comp;
    bool final_block = false;
    bool first = true;
    for (uint32_t guard = 0; guard < 0x7FFFFFF0u; ++guard) {
        uint32_t ev = STEP_LITERAL, len = 0, dist = 0;
        int rc = INF_OK;
        if (leader) {
            if (st == ST_NEXT) {
                if (!first) m += n_dec;
                first = false;
                if (m >= n_blocks) st = ST_DONE;
                else {
                    const BgzfBlock b = blocks[m];
                    in = comp + b.c_off;
                    c_len = b.c_len;
                    out = text + b.out_off;
                    out_len = b.isize;
                    pos = 0;
                    br.init(in, c_len);
                    st = ST_HEADER;
                }
            }
            if (st == ST_HEADER) {
                uint32_t kind = 0;
                rc = begin_block(br, t, out, &pos, out_len, &final_block, &kind);
                if (rc == INF_OK) {
                    if (kind != 0) st = ST_SYMBOLS;
                    else if (final_block) {
                        rc = pos != out_len ? INF_SIZE_MISMATCH : br.overrun(in, c_len) ? INF_INPUT_OVERRUN : INF_OK;
                        st = ST_NEXT;
                    }
                }
            } else if (st == ST_SYMBOLS) {
                ev = decode_step(br, t, out, pos, out_len, &len, &dist);
                if (ev == STEP_LITERAL) ++pos;
                else if (ev == STEP_END) {
                    if (final_block) {
                        rc = pos != out_len ? INF_SIZE_MISMATCH : br.overrun(in, c_len) ? INF_INPUT_OVERRUN : INF_OK;
                        st = ST_NEXT;
                    } else st = ST_HEADER;
                } else if (ev == STEP_ERROR) rc = (int)len;
            }
            if (rc != INF_OK) {
                atomicMin(error, ((unsigned long long)m << 4) | (unsigned long long)rc);
                st = ST_NEXT;                                   // give the member up, go on with the next one
                ev = STEP_LITERAL;
            }
        }
        // ---- matches: copied by the lanes of the group
        const bool match = leader && ev == STEP_MATCH;
        if (__any_sync(FULL, match)) {
            const uint32_t ld = __shfl_sync(FULL, match ? (len | (dist << 16)) : 0u, 0, SW);      // dist <= 32768, len <= 258
            const unsigned long long dst = __shfl_sync(FULL, (unsigned long long)(uintptr_t)(out + pos), 0, SW);
            __syncwarp();                                       // the leaders' literals are visible to the lanes that copy
            const uint32_t l = ld & 0xFFFFu, d = ld >> 16;
            if (l) {
                uint8_t* q = (uint8_t*)(uintptr_t)dst;
                const uint8_t* src = q - d;
                if (d >= l) for (uint32_t k = sl; k < l; k += SW) q[k] = src[k];
                else for (uint32_t k = sl; k < l; k += SW) q[k] = src[k % d];
            }
            __syncwarp();
            if (match) pos += len;
        }
        if (__all_sync(FULL, !leader || st == ST_DONE)) break;
    }
}

// Offset of the byte after the last '\n' of text[0, len) (0: none).  One CTA, scanning backwards in strides.
__global__ void __launch_bounds__(256) k_last_line_end(const uint8_t* text, uint64_t len, unsigned long long* out) {
    __shared__ unsigned long long s_best;
    if (threadIdx.x == 0) s_best = 0;
    __syncthreads();
    for (uint64_t hi = len; hi > 0;) {
        const uint64_t lo = hi > 4096 ? hi - 4096 : 0;
        unsigned long long best = 0;
        for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
            if (text[i] == (uint8_t)'\n') best = i + 1;
        if (best) atomicMax(&s_best, best);
        __syncthreads();
        const unsigned long long found = s_best;
        __syncthreads();
        if (found) break;
        hi = lo;
    }
    if (threadIdx.x == 0) *out = s_best;
}

#endif  // __CUDACC__

}  // namespace sid
