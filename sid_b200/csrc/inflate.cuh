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
//     (a damaged member is an error code, never a hang); ISIZE is checked by the decoder, the CRC-32 by k_crc32_members.
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
    uint32_t crc;           // CRC-32 of the text (trailer)
    uint32_t reserved;
};

enum : int { INF_OK = 0, INF_BAD_BLOCK_TYPE = 1, INF_BAD_STORED = 2, INF_BAD_LENGTHS = 3, INF_BAD_SYMBOL = 4, INF_BAD_DISTANCE = 5,
             INF_OUTPUT_OVERRUN = 6, INF_INPUT_OVERRUN = 7, INF_SIZE_MISMATCH = 8, INF_CRC_MISMATCH = 9 };

constexpr int INF_LIT_BITS = 10, INF_DIST_BITS = 8;

// A table entry says everything the walk needs about a symbol: bits 0-3 the length of its code (0: not in the fast table),
// bits 4-7 the number of extra bits that follow it, bits 8-9 what it is, bits 16-31 its value (the literal, the base of
// the match length or of the distance).
enum : uint32_t { K_LITERAL = 0, K_LENGTH = 1, K_END = 2, K_INVALID = 3 };
SID_HD uint32_t make_entry(uint32_t kind, uint32_t extra, uint32_t value) { return (extra << 4) | (kind << 8) | (value << 16); }
SID_HD uint32_t entry_kind(uint32_t e) { return (e >> 8) & 3u; }
SID_HD uint32_t entry_value(uint32_t e) { return e >> 16; }

struct InflateTables {
    uint32_t lit_fast[1 << INF_LIT_BITS];   // entries of the codes of at most INF_LIT_BITS bits, by the next bits of the stream
    uint32_t dist_fast[1 << INF_DIST_BITS];
    uint16_t lit_sym[288], dist_sym[32];    // symbols in canonical order (by length, then by value): the longer codes
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

SID_HD uint32_t length_base(uint32_t i) {       // length symbols 257..285 -> i = 0..28
    return i < 8 ? 3 + i : i == 28 ? 258 : ((4 + (i & 3)) << ((i >> 2) - 1)) + 3;
}
SID_HD uint32_t length_extra(uint32_t i) { return i < 8 || i == 28 ? 0 : (i >> 2) - 1; }
SID_HD uint32_t dist_base(uint32_t i) {         // distance symbols 0..29
    return i < 4 ? 1 + i : ((2 + (i & 1)) << ((i >> 1) - 1)) + 1;
}
SID_HD uint32_t dist_extra(uint32_t i) { return i < 4 ? 0 : (i >> 1) - 1; }

// The three alphabets of a block: literals / lengths, distances, and the code lengths of its header.
enum : int { ALPHA_LITLEN = 0, ALPHA_DIST = 1, ALPHA_PLAIN = 2 };
SID_HD uint32_t symbol_entry(int alphabet, uint32_t s) {
    if (alphabet == ALPHA_LITLEN) {
        if (s < 256) return make_entry(K_LITERAL, 0, s);
        if (s == 256) return make_entry(K_END, 0, 0);
        if (s > 285) return make_entry(K_INVALID, 0, 0);
        return make_entry(K_LENGTH, length_extra(s - 257), length_base(s - 257));
    }
    if (alphabet == ALPHA_DIST) return s > 29 ? make_entry(K_INVALID, 0, 0) : make_entry(K_LITERAL, dist_extra(s), dist_base(s));
    return make_entry(K_LITERAL, 0, s);
}

// Canonical Huffman code of `n` symbols with the given lengths (0 = unused): fast table of `bits` index bits, symbols in
// canonical order, codes per length.  Returns false for an over-subscribed set of lengths (an incomplete one is accepted
// like zlib accepts a single distance code; its unused codes decode to K_INVALID).
SID_HD bool build_huffman(int alphabet, const uint8_t* lengths, uint32_t n, uint32_t* fast, uint32_t bits, uint16_t* sym, uint16_t* cnt) {
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
            const uint32_t e = symbol_entry(alphabet, s) | l;
            for (uint32_t i = bit_reverse(c, l); i < (1u << bits); i += 1u << l) fast[i] = e;
        }
    }
    return true;
}

// The entry (with its code length in bits 0-3) of a code longer than the fast table's index: the count/offset walk of the
// canonical order over the bits of w.  K_INVALID | 15 for a code that is not in the alphabet.
SID_HD uint32_t slow_entry(uint32_t w, int alphabet, const uint16_t* cnt, const uint16_t* sym) {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        code |= (int)((w >> (l - 1)) & 1u);
        const int count = cnt[l];
        if (code - count < first) return symbol_entry(alphabet, sym[index + (code - first)]) | (uint32_t)l;
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return make_entry(K_INVALID, 0, 0) | 15u;
}

// Next symbol with its extra bits: its entry with the value completed (base + extra bits) in bits 16-31; code and extra
// bits are consumed.  K_INVALID for a code that is not in the alphabet.
SID_HD uint32_t decode_entry(BitReader& br, int alphabet, const uint32_t* fast, uint32_t bits, const uint16_t* cnt, const uint16_t* sym) {
    const uint32_t w = br.window();
    uint32_t e = fast[w & ((1u << bits) - 1u)];
    if ((e & 15u) == 0) e = slow_entry(w, alphabet, cnt, sym);
    const uint32_t cl = e & 15u, xb = (e >> 4) & 15u;
    const uint32_t extra = (w >> cl) & ((1u << xb) - 1u);          // cl + xb <= 28: all of it is in this window
    br.drop(cl + xb);
    return e + (extra << 16);
}

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
        // the code length alphabet borrows the distance arrays (they are rebuilt right after)
        if (!build_huffman(ALPHA_PLAIN, cl, 19, t.dist_fast, 7, t.dist_sym, t.dist_cnt)) return INF_BAD_LENGTHS;
        uint32_t i = 0;
        while (i < nlen + ndist) {
            const uint32_t e = decode_entry(br, ALPHA_PLAIN, t.dist_fast, 7, t.dist_cnt, t.dist_sym);
            if (entry_kind(e) == K_INVALID) return INF_BAD_LENGTHS;
            if (br.over) return INF_INPUT_OVERRUN;
            const uint32_t s = entry_value(e);
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
    if (!build_huffman(ALPHA_LITLEN, lengths, nlen, t.lit_fast, INF_LIT_BITS, t.lit_sym, t.lit_cnt)) return INF_BAD_LENGTHS;
    if (!build_huffman(ALPHA_DIST, lengths + nlen, ndist, t.dist_fast, INF_DIST_BITS, t.dist_sym, t.dist_cnt)) return INF_BAD_LENGTHS;
    return INF_OK;
}

// What lane 0 hands to the warp: literals it has already stored, then one of match / end of block / error.
enum : uint32_t { EV_MATCH = 0, EV_END = 1, EV_ERROR = 2 };

// Decodes symbols until something other than a literal comes; literals go straight to out[pos...].  Returns the event;
// *n_lit literals were stored, for EV_MATCH *len and *dist are set, for EV_ERROR *len is the code.
SID_HD uint32_t decode_run(BitReader& br, const InflateTables& t, uint8_t* out, uint32_t pos, uint32_t out_len, uint32_t* n_lit,
                           uint32_t* len, uint32_t* dist) {
    uint32_t q = pos;
    for (;;) {
        const uint32_t e = decode_entry(br, ALPHA_LITLEN, t.lit_fast, INF_LIT_BITS, t.lit_cnt, t.lit_sym);
        const uint32_t kind = entry_kind(e);
        if (kind == K_LITERAL) {
            if (q >= out_len || br.over) { *n_lit = q - pos; *len = br.over ? INF_INPUT_OVERRUN : INF_OUTPUT_OVERRUN; return EV_ERROR; }
            out[q++] = (uint8_t)entry_value(e);
            continue;
        }
        *n_lit = q - pos;
        if (kind == K_END) return EV_END;
        if (kind == K_INVALID) { *len = INF_BAD_SYMBOL; return EV_ERROR; }
        const uint32_t l = entry_value(e);
        const uint32_t d = decode_entry(br, ALPHA_DIST, t.dist_fast, INF_DIST_BITS, t.dist_cnt, t.dist_sym);
        const uint32_t dd = entry_value(d);
        if (entry_kind(d) == K_INVALID || dd > q) { *len = INF_BAD_DISTANCE; return EV_ERROR; }
        if (q + l > out_len) { *len = INF_OUTPUT_OVERRUN; return EV_ERROR; }
        if (br.over) { *len = INF_INPUT_OVERRUN; return EV_ERROR; }
        *len = l;
        *dist = dd;
        return EV_MATCH;
    }
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

constexpr int INF_WARPS = 8;        // members per CTA (one warp each; 5.8 KB of tables per warp)

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

// ---- CRC-32 (the gzip trailer's, polynomial 0xEDB88320 reflected) of every member's text, one warp per member.
// The register after a message is linear in (message, start value): lane i runs the byte-wise table walk over its 2 KiB
// piece COUNTED FROM THE END of the text (the lane that holds byte 0 starts from 0xFFFFFFFF, the others from 0), so lane i's
// register has to be advanced by exactly i * 2048 zero bytes; Horner over the lanes with the one operator "advance by 2048
// zero bytes" (a 32 x 32 bit matrix, column `lane` in lane `lane`; a matrix-vector product is one select and an
// exclusive-or reduction over the warp) gives the register of the whole text.
struct CrcTables {
    uint32_t byte_table[256];       // the usual table of the byte-wise walk
    uint32_t advance_2k[32];        // column j: register 1 << j advanced by 2048 zero bytes
};
constexpr uint32_t CRC_PIECE = 2048;
constexpr int CRC_WARPS = 8;

__global__ void __launch_bounds__(CRC_WARPS * 32) k_crc32_members(const uint8_t* text, const BgzfBlock* blocks, uint32_t n_blocks, const CrcTables* tables,
                                                                    unsigned long long* error) {
    __shared__ uint32_t s_table[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = tables->byte_table[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t column = tables->advance_2k[lane];
    for (uint32_t m = blockIdx.x * CRC_WARPS + warp; m < n_blocks; m += gridDim.x * CRC_WARPS) {
        const BgzfBlock b = blocks[m];
        const uint8_t* p = text + b.out_off;
        const uint32_t n = b.isize;
        // lane i: bytes [n - (i + 1) * 2048, n - i * 2048) cut at 0
        const uint32_t hi = n > lane * CRC_PIECE ? n - lane * CRC_PIECE : 0u;
        const uint32_t lo = hi > CRC_PIECE ? hi - CRC_PIECE : 0u;
        uint32_t reg = (hi > 0 && lo == 0) ? 0xFFFFFFFFu : 0u;
        for (uint32_t k = lo; k < hi; ++k) reg = s_table[(reg ^ p[k]) & 0xFFu] ^ (reg >> 8);
        // Horner from the lane farthest from the end: acc = advance(acc) ^ reg_i, i = 31 .. 0
        uint32_t acc = 0;
        for (int i = 31; i >= 0; --i) {
            uint32_t v = ((acc >> lane) & 1u) ? column : 0u;        // advance: exclusive-or of the columns of the set bits
#pragma unroll
            for (int d = 16; d; d >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, d);
            acc = v ^ __shfl_sync(0xFFFFFFFFu, reg, i);
        }
        if (lane == 0 && (acc ^ 0xFFFFFFFFu) != b.crc) atomicMin(error, ((unsigned long long)m << 4) | (unsigned long long)INF_CRC_MISMATCH);
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
