// K1: pileup tokenizer + profile builder + join against the unique-profile table.
//   readFile call.cpp:11-20, parsePileupLine pileup.cpp:13-68, parseReadBases pileup.cpp:70-153,
//   countUniqueProfiles pileup.cpp:169-196 (the counting half), call.cpp:217-221 (profile index).
//
// Layout: the text is cut into tiles of eight slices.  Each persistent CTA is warp specialised:
//   * warp 0 (service) draws tiles in order from an atomic ticket and stages them (+ a tail for
//     the line that straddles the end) in shared memory with one bulk asynchronous copy
//     (cp.async.bulk -> mbarrier), one tile ahead of the parse warps.
//   * warps 1..8 (parse) each own one slice of the staged tile: the warp finds the line starts of
//     its slice, then parses them one line per lane.  The slice length is chosen by the host so
//     that a slice holds about 30 lines.
// Site storage is dense but NOT in file order: a warp reserves room for the lines of its slice with
// one atomicAdd and records (first storage index, count) in the block table entry of (tile, slice).
// The block table is in file order; k_blk_sums/k_blk_order turn it into order[file index] = storage
// index, and the consumers (CSV writer, records, quality, the ordered view) go through that.  A file-ordered store would
// need every tile to wait for the line counts of all tiles before it (a look-back chain over some
// 450 resident CTAs): measured 18 % of the kernel, see profiles/README.md.
// Hand-over in both directions goes through mbarriers; there is no CTA-wide barrier in the loop.
#pragma once
#include "common.cuh"
#include "parse.cuh"
#include "parse_bits.cuh"
#include "parse_fast.cuh"
#include "ptx.cuh"
#include "table.cuh"

namespace sid {

constexpr int TOK_PARSE_WARPS = 8;
constexpr int TOK_THREADS = 32 * (1 + TOK_PARSE_WARPS);
constexpr int TOK_STAGES_MAX = 2;                 // tiles staged per CTA: 2 (copy of tile n+1 under the parsing of n) or 1 (long lines)
#ifndef SID_SVC_SLEEP
#define SID_SVC_SLEEP 500
#endif
#ifndef SID_PARSE_SLEEP
#define SID_PARSE_SLEEP 200
#endif
constexpr unsigned SVC_SLEEP = SID_SVC_SLEEP, PARSE_SLEEP = SID_PARSE_SLEEP;   // ns between polls of a waiting warp
constexpr int SLICE_MAX = 4096;                   // bytes; slices are multiples of 16
constexpr int SLICE_MIN = 256;                    // slices are multiples of 32
constexpr int TILE_TAIL = 2048;                   // staged past the tile end for straddling lines
constexpr int TILE_PAD = 16;                      // staged before the tile begin (previous byte)
constexpr int TILE_SMEM_MAX = TILE_PAD + TOK_PARSE_WARPS * SLICE_MAX + TILE_TAIL;
constexpr int SLICE_MAX_LINES = SLICE_MAX / 8;    // a valid line has >= 10 bytes

// decoupled look-back status words (k_csv)
constexpr unsigned long long LB_FLAG_AGG = 1ull << 62;
constexpr unsigned long long LB_FLAG_PREFIX = 2ull << 62;
constexpr unsigned long long LB_VALUE_MASK = (1ull << 62) - 1;

struct TokParams {
    const uint8_t* text;
    uint64_t text_len, range_begin, range_end;
    uint64_t tile0;                 // absolute offset of tile 0 (range_begin rounded down to 16)
    uint32_t n_tiles;
    // outputs, indexed by storage index (site_base + the room a warp reserved)
    uint64_t site_base;             // first storage index of this call
    uint64_t site_cap;              // capacity of the site arrays (storage indices, < 2^32)
    uint64_t* profile;              // optional
    int32_t* pos;
    uint32_t* slot;                 // optional (needs table)
    uint32_t* name_ref;
    uint64_t* line_off;             // optional (quality path)
    // scheduling / look-back
    unsigned int* tile_ticket;
    unsigned long long* site_alloc; // storage indices handed out so far, relative to site_base (atomic)
    unsigned long long* blk;        // block table of this call, entry (tile * 8 + slice): lines << 32 | first storage index
    unsigned long long* error;      // out: min over (line offset << 3 | LineStatus)
    TableView table;
    NameDict names;
    int use_table, count_profiles, want_qual;
    uint32_t slice_bytes;           // multiple of 16 in [SLICE_MIN, SLICE_MAX]; a tile is 8 slices
    uint32_t text_stride;           // bytes of shared memory per staged tile (tile_smem rounded up to 128)
    uint32_t tail_bytes;            // bytes staged past the tile end for the lines that straddle it (<= TILE_TAIL)
    uint32_t lines_cap;             // line-start slots per slice (slice_bytes / 8)
    uint32_t ext_bytes;             // bytes past its slice a parse warp also classifies (multiple of 32, <= 2016)
    uint32_t words_cap;             // 32-bit words per class bit array: (slice_bytes + ext_bytes) / 32 + 2
};

// Dynamic shared memory a launch with this slice length needs.
inline uint32_t tok_tail_bytes(uint32_t ext) { return ext + 128u > 2048u ? 2048u : ext + 128u; }   // multiple of 32
inline uint32_t tok_text_stride(uint32_t slice, uint32_t ext) { return (16u + 8u * slice + tok_tail_bytes(ext) + 127u) & ~127u; }
inline uint32_t tok_words_cap(uint32_t slice, uint32_t ext) { return (slice + ext) / 32u + 2u; }
// staged text (1 or 2 stages) + line starts (8 warps) + class bit arrays (8 warps x 8 classes)
inline uint32_t tok_dyn_smem(uint32_t slice, uint32_t ext, uint32_t stages = 2) {
    return stages * tok_text_stride(slice, ext) + 8u * (slice / 8u) * 2u + 8u * 8u * tok_words_cap(slice, ext) * 4u;
}

#if defined(__CUDACC__)

struct SmemSrc {
    const uint8_t* s;       // s[0] is absolute offset abs0
    uint64_t abs0;
    uint32_t avail;
    mutable bool overrun;
    __device__ __forceinline__ uint8_t at(uint64_t off) const {
        const uint64_t i = off - abs0;
        if (i < avail) return s[i];
        overrun = true;
        return (uint8_t)'\n';
    }
};

// 16-bit mask of the bytes equal to '\n' in a 16-byte vector (exact for all byte values).
__device__ __forceinline__ uint32_t newline_mask16(uint4 v) {
    uint32_t m = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t x = w[i] ^ 0x0A0A0A0Au;
        const uint32_t t = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;      // bit 7 set iff the byte is not '\n'
        const uint32_t z7 = ~t & 0x80808080u;
        m |= ((z7 * 0x00204081u) >> 28) << (4 * i);                    // gather bits 7,15,23,31 into a nibble
    }
    return m;
}

// Eight bytes starting at byte offset o of a 4-byte aligned buffer.
__device__ __forceinline__ uint2 load8_unaligned(const uint8_t* s, uint32_t o) {
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s) + (o >> 2);
    const uint32_t sh = (o & 3) * 8;
    const uint32_t w0 = sw[0], w1 = sw[1], w2 = sw[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

__device__ __forceinline__ void report_error_at(unsigned long long* error, uint64_t line_abs, int status) {
    atomicMin(error, (unsigned long long)((line_abs << 3) | (uint64_t)status));
}
__device__ __forceinline__ void report_error(const TokParams& p, uint64_t line_abs, int status) {
    atomicMin(p.error, (unsigned long long)((line_abs << 3) | (uint64_t)status));
}

struct LineResult {
    int status;
    int32_t pos;
    uint64_t profile;
    uint32_t chrom_off, chrom_len;
};

// Does the name of this line equal the name of the tile's first line?  (8 bytes at a time for the
// usual short names; the first line's name is known to start at its first byte when l0_ok.)
__device__ __forceinline__ bool same_name_as_first(const uint8_t* s_text, uint32_t my_off, uint32_t len, uint32_t l0_off,
                                                   uint2 first8, uint32_t avail = TILE_SMEM_MAX) {
    if (len <= 7) {
        // the name AND the delimiter after it, as one masked compare of len + 1 bytes: equal names followed by different
        // delimiters (a space here, a tab there) only send this line through the dictionary, which finds the same name
        const uint2 mine = load8_unaligned(s_text, my_off);
        const uint32_t bits = 8 * (len + 1);                       // 8 .. 64
        const uint32_t mlo = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
        const uint32_t mhi = bits <= 32 ? 0u : (bits >= 64 ? 0xFFFFFFFFu : ((1u << (bits - 32)) - 1u));
        return (((mine.x ^ first8.x) & mlo) | ((mine.y ^ first8.y) & mhi)) == 0;
    }
    if (my_off + len + 1 > avail || l0_off + len + 1 > avail) return false;
    for (uint32_t i = 0; i < len; ++i) if (s_text[my_off + i] != s_text[l0_off + i]) return false;
    const uint8_t d = s_text[l0_off + len];
    return d == '\t' || d == ' ';
}

struct StageMeta {
    uint64_t tb;           // absolute offset of the tile
    int32_t tile;          // -1: no more tiles
    uint32_t smem_bytes;   // bytes staged for this tile
};

struct SmemBytes {         // byte source over a staged tile for name_intern
    const uint8_t* s;
    uint64_t abs0;
    __device__ __forceinline__ uint8_t at(uint64_t off) const { return s[(uint32_t)(off - abs0)]; }
};


template <bool FAST, int TOK_STAGES>
__global__ void __launch_bounds__(TOK_THREADS) k_tokenize(const TokParams p) {
    extern __shared__ __align__(128) uint8_t s_dyn[];      // staged text per stage, then line starts per stage and warp
    __shared__ StageMeta s_meta[TOK_STAGES];
    __shared__ __align__(8) uint64_t s_full[TOK_STAGES], s_done[TOK_STAGES];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t slice = p.slice_bytes;
    const uint32_t tile_bytes = slice * TOK_PARSE_WARPS;
    const uint32_t tile_smem = TILE_PAD + tile_bytes + p.tail_bytes;
    if (tid == 0) {
        for (int b = 0; b < TOK_STAGES; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&s_done[b], TOK_PARSE_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    FlatSrc gsrc {p.text, p.text_len};

    if (warp == 0) {
        // ================================================================= service warp
        auto issue_load = [&](int b, uint32_t tile) {
            const uint64_t tb = p.tile0 + (uint64_t)tile * tile_bytes;
            uint8_t* txt = s_dyn + (size_t)b * p.text_stride;
            if (lane == 0) {
                s_meta[b].tb = tb;
                s_meta[b].tile = (int32_t)tile;
            }
            // stage [tb - 16, tb + tile_bytes + tail_bytes); bytes outside the text read as '\n'
            const bool interior = tb >= TILE_PAD && tb - TILE_PAD + tile_smem <= p.text_len;
            if (interior) {
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(&s_full[b], tile_smem);
                    bulk_load(txt, p.text + (tb - TILE_PAD), tile_smem, &s_full[b]);
                }
            } else {
                for (uint32_t i = lane; i < tile_smem / 16; i += 32) {
                    const int64_t a = (int64_t)tb - TILE_PAD + 16 * (int64_t)i;
                    uint4 v;
                    if (a >= 0 && (uint64_t)a + 16 <= p.text_len) {
                        v = __ldg(reinterpret_cast<const uint4*>(p.text + a));
                    } else {
                        uint32_t w[4] = {0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au};
                        for (int k = 0; k < 16; ++k) {
                            const int64_t q = a + k;
                            if (q >= 0 && (uint64_t)q < p.text_len) {
                                w[k >> 2] = (w[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | ((uint32_t)p.text[q] << (8 * (k & 3)));
                            }
                        }
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    reinterpret_cast<uint4*>(txt)[i] = v;
                }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) mbar_arrive(&s_full[b]);
            }
        };
        uint32_t next_tile = 0;
        if (lane == 0) next_tile = atomicAdd(p.tile_ticket, 1u);
        next_tile = __shfl_sync(0xFFFFFFFFu, next_tile, 0);
        if (next_tile < p.n_tiles) issue_load(0, next_tile);
        for (uint32_t it = 0;; ++it) {
            const int b = it % TOK_STAGES;
            const uint32_t tile = next_tile;
            if (tile >= p.n_tiles) {
                if (lane == 0) {
                    s_meta[b].tile = -1;
                    __threadfence_block();
                    mbar_arrive(&s_full[b]);
                }
                break;
            }
            // ---- prefetch: free the next stage, draw the next ticket, start its copy
            {
                const int nb = (it + 1) % TOK_STAGES;
                if (it + 1 >= TOK_STAGES && !mbar_wait<SVC_SLEEP>(&s_done[nb], ((it + 1) / TOK_STAGES - 1) & 1)) {
                    if (lane == 0) report_error(p, 0, LINE_MALFORMED + 4);
                    break;
                }
                uint32_t t = 0;
                if (lane == 0) t = atomicAdd(p.tile_ticket, 1u);
                next_tile = __shfl_sync(0xFFFFFFFFu, t, 0);
                if (next_tile < p.n_tiles) issue_load(nb, next_tile);
            }
        }
        return;
    }

    // ===================================================================== parse warps
    const int pw = warp - 1;                               // slice index
    uint32_t cache_len = 0, cache_ref = 0;                 // name of the previous first line of this warp (lane 0)
    uint4 cache_name = make_uint4(0, 0, 0, 0);
    for (uint32_t it = 0;; ++it) {
        const int b = it % TOK_STAGES;
        const uint32_t use = it / TOK_STAGES;
        if (!mbar_wait<PARSE_SLEEP>(&s_full[b], use & 1)) {
            if (lane == 0) report_error(p, 0, LINE_MALFORMED + 4);
            break;
        }
        if (s_meta[b].tile < 0) break;
        const uint8_t* txt = s_dyn + (size_t)b * p.text_stride;
        uint16_t* starts = reinterpret_cast<uint16_t*>(s_dyn + (size_t)TOK_STAGES * p.text_stride) +
                           (size_t)pw * p.lines_cap;
        const uint64_t tb = s_meta[b].tb;
        const uint64_t abs0 = tb - TILE_PAD;              // wraps for a tile at offset 0; only differences are used
        // ---- stage 1, flat over the slice (+ ext_bytes so that the last lines can be finished): every
        //      32-byte unit is transposed into bit planes and classified (parse_bits.cuh); the class
        //      words go to this warp's bit arrays; the '\n' word gives the line starts of the slice:
        //      byte q starts a line iff text[q] != '\n' and text[q-1] == '\n'.
        //      Unit u is handled by lane u % 32 in round u / 32.
        uint32_t n_lines = 0;
        const uint32_t slice_off = (uint32_t)pw * slice;               // offset of the slice inside the tile
        const uint32_t region_off = TILE_PAD + slice_off;              // offset in txt of bit 0 of the bit arrays
        uint32_t* bits = reinterpret_cast<uint32_t*>(s_dyn + (size_t)TOK_STAGES * p.text_stride + (size_t)TOK_PARSE_WARPS * p.lines_cap * 2) +
                         (size_t)pw * 8 * p.words_cap;
        {
            const uint32_t own_units = slice / 32, units = (slice + p.ext_bytes) / 32;
            const bool inside = tb + slice_off >= p.range_begin && tb + slice_off + slice <= p.range_end;   // the usual case
            for (uint32_t u0 = 0; u0 < units; u0 += 32) {
                const uint32_t u = u0 + lane;
                uint32_t st = 0;
                if (u < units) {
                    const uint8_t* up = txt + region_off + u * 32;
                    const uint4 v0 = *reinterpret_cast<const uint4*>(up), v1 = *reinterpret_cast<const uint4*>(up + 16);
                    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    const ClassWords k = classify32(w);
                    bits[0 * p.words_cap + u] = k.term;
                    bits[1 * p.words_cap + u] = k.a;
                    bits[2 * p.words_cap + u] = k.c;
                    bits[3 * p.words_cap + u] = k.g;
                    bits[4 * p.words_cap + u] = k.t;
                    bits[5 * p.words_cap + u] = k.dot;
                    bits[6 * p.words_cap + u] = k.caret;
                    bits[7 * p.words_cap + u] = k.pm;
                    if (u < own_units) {
                        const uint32_t prev_nl = up[-1] == (uint8_t)'\n' ? 1u : 0u;
                        st = ((k.nl << 1) | prev_nl) & ~k.nl;
                        if (!inside) {                                 // restrict to the owned range [range_begin, range_end)
                            const uint64_t first = tb + slice_off + u * 32;
                            if (first + 32 <= p.range_begin || first >= p.range_end) st = 0;
                            else {
                                if (first < p.range_begin) st &= 0xFFFFFFFFu << (uint32_t)(p.range_begin - first);
                                if (first + 32 > p.range_end) st &= 0xFFFFFFFFu >> (32u - (uint32_t)(p.range_end - first));
                            }
                        }
                    }
                } else if (u < units + 2) {
                    for (int c = 0; c < 8; ++c) bits[c * p.words_cap + u] = 0;      // padding words read by bits32()
                }
                const uint32_t my_count = __popc(st);
                uint32_t incl = my_count;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += o;
                }
                uint32_t idx = n_lines + incl - my_count;
                while (st) {
                    const int bit = __ffs((int)st) - 1;
                    st &= st - 1;
                    if (idx < p.lines_cap) starts[idx] = (uint16_t)(slice_off + u * 32 + bit);
                    ++idx;
                }
                n_lines += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            if (units % 32 >= 30 || units % 32 == 0) {                 // the padding words did not fit into the last round
                if (lane < 2) for (int c = 0; c < 8; ++c) bits[c * p.words_cap + units + lane] = 0;
            }
            if (n_lines > p.lines_cap) {                   // only possible with lines shorter than 8 bytes
                if (lane == 0) report_error(p, tb, LINE_MALFORMED);
                n_lines = p.lines_cap;
            }
            __syncwarp();
        }
        // ---- room for the lines of this slice: one atomicAdd, recorded in the (file-ordered) block table
        uint64_t base = 0;
        {
            unsigned long long bb = 0;
            if (lane == 0) {
                if (n_lines) bb = p.site_base + atomicAdd(p.site_alloc, (unsigned long long)n_lines);
                p.blk[(uint64_t)s_meta[b].tile * TOK_PARSE_WARPS + pw] = ((unsigned long long)n_lines << 32) | (bb & 0xFFFFFFFFull);
            }
            base = __shfl_sync(0xFFFFFFFFu, bb, 0);
        }
        BitArrays B;
        B.term = bits; B.a = bits + p.words_cap; B.c = bits + 2 * p.words_cap; B.g = bits + 3 * p.words_cap;
        B.t = bits + 4 * p.words_cap; B.dot = bits + 5 * p.words_cap; B.caret = bits + 6 * p.words_cap; B.pm = bits + 7 * p.words_cap;
        B.n_bits = slice + p.ext_bytes;
        // ---- the slice's first line: its name is shared by (almost) all lines of the slice
        uint32_t l0_off = 0, name0_ref = 0;
        uint2 first8 = make_uint2(0, 0);
        if (n_lines) {
            l0_off = TILE_PAD + starts[0];
            first8 = load8_unaligned(txt, l0_off);
        }
        // ---- groups of 32 consecutive lines, one line per lane.  Every lane runs the tokenizer (lanes
        //      past the end re-parse the group's first line and drop the result) so that the warp
        //      reconverges inside it.
        for (uint32_t g = 0; g < n_lines; g += 32) {
            const uint32_t j = g + lane;
            const bool mine = j < n_lines;
            const uint32_t off = starts[mine ? j : g];
            const uint64_t line_abs = tb + off;
            LineResult r;
            r.status = LINE_MALFORMED;
            bool fast = false;
            if (FAST) {
                // (with want_qual the quality columns are validated by k_quality, which re-reads the line)
                FastLine fl;
                fast = parse_line_bits(txt, tile_smem, region_off, B, TILE_PAD + off, fl);
                r.status = fl.status; r.pos = fl.pos; r.profile = fl.profile; r.chrom_off = fl.chrom_off; r.chrom_len = fl.chrom_len;
            }
            if (!fast && mine) {
                SmemSrc ssrc {txt, abs0, tile_smem, false};
                ParsedLine pl;
                parse_line(ssrc, line_abs, p.want_qual != 0, pl);
                if (ssrc.overrun) parse_line(gsrc, line_abs, p.want_qual != 0, pl);
                r.status = pl.status; r.pos = pl.pos; r.profile = pl.profile; r.chrom_off = pl.chrom_off; r.chrom_len = pl.chrom_len;
            }
            __syncwarp();
            const bool good = mine && r.status == LINE_OK;
            if (g == 0) {
                // lane 0 holds the slice's first line: intern its name once (kept in registers from tile
                // to tile, so the dictionary is only consulted when the name changes); only a name that
                // starts at the first byte of the line can be shared (same_name_as_first)
                if (lane == 0 && good && r.chrom_off == 0) {
                    const uint32_t len = r.chrom_len;
                    uint4 nm = make_uint4(0, 0, 0, 0);
                    if (len <= 16) {
                        const uint2 hi = load8_unaligned(txt, l0_off + 8);
                        uint32_t wds[4] = {first8.x, first8.y, hi.x, hi.y};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int rem = (int)len - 4 * k;
                            if (rem <= 0) wds[k] = 0; else if (rem < 4) wds[k] &= (1u << (8 * rem)) - 1u;
                        }
                        nm = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                    }
                    if (len <= 16 && len == cache_len && cache_ref && nm.x == cache_name.x && nm.y == cache_name.y &&
                        nm.z == cache_name.z && nm.w == cache_name.w) {
                        name0_ref = cache_ref;
                    } else {
                        SmemBytes sb {txt, abs0};
                        name0_ref = name_intern(p.names, sb, line_abs, len);
                        if (len <= 16) { cache_len = len; cache_ref = name0_ref; cache_name = nm; }
                    }
                }
                name0_ref = __shfl_sync(0xFFFFFFFFu, name0_ref, 0);
            }
            bool same = false;
            uint32_t slot = 0;
            if (good) {
                same = name0_ref && r.chrom_off == 0 && same_name_as_first(txt, TILE_PAD + off, r.chrom_len, l0_off, first8, tile_smem);
                if (p.use_table) slot = table_find_or_insert(p.table, r.profile);
            }
            if (mine && r.status != LINE_OK) report_error(p, line_abs, r.status);
            if (good) {
                const uint64_t site = base + j;
                if (site >= p.site_cap) { report_error(p, line_abs, LINE_MALFORMED + 5); }
                else {
                    const uint32_t ref = same ? name0_ref : name_intern(p.names, gsrc, line_abs + r.chrom_off, r.chrom_len);
                    p.pos[site] = r.pos;
                    p.name_ref[site] = ref;
                    if (p.profile) p.profile[site] = r.profile;
                    if (p.line_off) p.line_off[site] = line_abs;
                    if (p.use_table) p.slot[site] = slot;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_done[b]);            // this warp is finished with the stage
    }
}

// ---- block table -> order[]: order[first_file_index + f] = storage index of the f-th line of the range.
constexpr int BLK_THREADS = 256;
constexpr int BLK_PER_THREAD = 2;
constexpr int BLK_CHUNK = BLK_THREADS * BLK_PER_THREAD;

__global__ void __launch_bounds__(BLK_THREADS) k_blk_sums(const unsigned long long* blk, uint32_t n_blocks, unsigned long long* part) {
    __shared__ uint32_t s_w[BLK_THREADS / 32];
    const uint32_t first = blockIdx.x * BLK_CHUNK + threadIdx.x;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < BLK_PER_THREAD; ++k) {
        const uint32_t b = first + k * BLK_THREADS;
        if (b < n_blocks) sum += (uint32_t)(blk[b] >> 32);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < BLK_THREADS / 32; ++w) t += s_w[w];
        part[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(BLK_THREADS) k_blk_order(const unsigned long long* blk, uint32_t n_blocks, const unsigned long long* part,
                                                           uint64_t file_base, uint32_t* order, uint64_t order_cap) {
    __shared__ unsigned long long s_red[BLK_THREADS / 32];
    __shared__ uint32_t s_first[BLK_CHUNK + 1];             // first file index (relative to the chunk) of each block
    __shared__ uint32_t s_base[BLK_CHUNK];                  // first storage index of each block
    __shared__ uint32_t s_wsum[BLK_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // file index of the chunk's first line: the lines of all chunks before it
    unsigned long long before = 0;
    for (uint32_t c = tid; c < blockIdx.x; c += BLK_THREADS) before += part[c];
#pragma unroll
    for (int d = 16; d; d >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, d);
    if (lane == 0) s_red[warp] = before;
    // exclusive scan of the chunk's counts, thread t owning blocks [8t, 8t+8)
    const uint32_t b0 = blockIdx.x * BLK_CHUNK + tid * BLK_PER_THREAD;
    uint32_t cnt[BLK_PER_THREAD], mine = 0;
#pragma unroll
    for (int k = 0; k < BLK_PER_THREAD; ++k) {
        const unsigned long long e = b0 + k < n_blocks ? blk[b0 + k] : 0ull;
        cnt[k] = (uint32_t)(e >> 32);
        s_base[tid * BLK_PER_THREAD + k] = (uint32_t)e;
        mine += cnt[k];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t off = incl - mine;
    for (int w = 0; w < warp; ++w) off += s_wsum[w];
    unsigned long long chunk_first = file_base;
    for (int w = 0; w < BLK_THREADS / 32; ++w) chunk_first += s_red[w];
#pragma unroll
    for (int k = 0; k < BLK_PER_THREAD; ++k) {
        s_first[tid * BLK_PER_THREAD + k] = off;
        off += cnt[k];
    }
    if (tid == BLK_THREADS - 1) s_first[BLK_CHUNK] = off;
    __syncthreads();
    // one warp per block: consecutive lanes write consecutive entries
    for (uint32_t j = warp; j < BLK_CHUNK; j += BLK_THREADS / 32) {
        const uint32_t f = s_first[j], n = s_first[j + 1] - f, base = s_base[j];
        const unsigned long long f0 = chunk_first + f;
        for (uint32_t i = lane; i < n; i += 32)
            if (f0 + i < order_cap) order[f0 + i] = base + i;
    }
}

#endif  // __CUDACC__

}  // namespace sid
