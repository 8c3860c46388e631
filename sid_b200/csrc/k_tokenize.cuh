// K1: pileup tokenizer + profile builder + join against the unique-profile table.
//   readFile call.cpp:11-20, parsePileupLine pileup.cpp:13-68, parseReadBases pileup.cpp:70-153,
//   countUniqueProfiles pileup.cpp:169-196 (the counting half), call.cpp:217-221 (profile index).
//
// Layout: the text is cut into fixed tiles of TILE_BYTES.  Persistent CTAs take tiles in order
// from an atomic ticket, stage the tile (+ a tail for the line that straddles its end) in shared
// memory, find the line starts with SWAR newline masks, obtain the global index of their first
// line by a decoupled look-back over per-tile line counts (single pass over the text, output
// dense and in file order), then parse one line per thread out of shared memory.
#pragma once
#include "common.cuh"
#include "parse.cuh"
#include "parse_fast.cuh"
#include "table.cuh"

namespace sid {

constexpr int TOK_PARSE_THREADS = 256;             // 8 parse warps
constexpr int TOK_THREADS = TOK_PARSE_THREADS + 32;   // + 1 service warp (look-back, name of the first line)
constexpr int TILE_BYTES = 32768;                 // 128 bytes per thread in the line-start scan
constexpr int TILE_TAIL = 2048;                   // staged past the tile end for straddling lines
constexpr int TILE_PAD = 16;                      // staged before the tile begin (previous byte)
constexpr int TILE_SMEM = TILE_PAD + TILE_BYTES + TILE_TAIL;
constexpr int TILE_MAX_LINES = TILE_BYTES / 8;    // a valid line has >= 10 bytes

constexpr unsigned long long LB_FLAG_AGG = 1ull << 62;
constexpr unsigned long long LB_FLAG_PREFIX = 2ull << 62;
constexpr unsigned long long LB_VALUE_MASK = (1ull << 62) - 1;

struct TokParams {
    const uint8_t* text;
    uint64_t text_len, range_begin, range_end;
    uint64_t tile0;                 // absolute offset of tile 0 (range_begin rounded down to 16)
    uint32_t n_tiles;
    // outputs, indexed by site (site_base + running index)
    uint64_t site_base;
    uint64_t site_cap;
    uint64_t* profile;              // optional
    int32_t* pos;
    uint32_t* slot;                 // optional (needs table)
    uint32_t* name_ref;
    uint64_t* line_off;             // optional (quality path)
    // scheduling / look-back
    unsigned int* tile_ticket;
    unsigned long long* tile_status;
    unsigned long long* n_sites;    // out: sites produced by this call
    unsigned long long* error;      // out: min over (line offset << 3 | LineStatus)
    TableView table;
    NameDict names;
    int use_table, count_profiles, want_qual;
};

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
                                                   uint2 first8) {
    if (len <= 7) {
        const uint2 mine = load8_unaligned(s_text, my_off);
        const uint32_t bits = 8 * len;
        uint32_t dlo, dhi, delim;
        if (len < 4) { dlo = (mine.x ^ first8.x) & ((1u << bits) - 1u); dhi = 0; delim = (first8.x >> bits) & 0xFFu; }
        else if (len == 4) { dlo = mine.x ^ first8.x; dhi = 0; delim = first8.y & 0xFFu; }
        else { dlo = mine.x ^ first8.x; dhi = (mine.y ^ first8.y) & ((1u << (bits - 32)) - 1u); delim = (first8.y >> (bits - 32)) & 0xFFu; }
        return (dlo | dhi) == 0 && (delim == '\t' || delim == ' ');
    }
    if (my_off + len + 1 > TILE_SMEM || l0_off + len + 1 > TILE_SMEM) return false;
    for (uint32_t i = 0; i < len; ++i) if (s_text[my_off + i] != s_text[l0_off + i]) return false;
    const uint8_t d = s_text[l0_off + len];
    return d == '\t' || d == ' ';
}

template <bool FAST>
__global__ void __launch_bounds__(TOK_THREADS) k_tokenize(const TokParams p) {
    __shared__ __align__(16) uint8_t s_text[TILE_SMEM];
    __shared__ uint16_t s_starts[TILE_MAX_LINES];
    __shared__ uint32_t s_warp_sums[TOK_PARSE_THREADS / 32];
    __shared__ uint32_t s_tile, s_name0_ref, s_ready;
    __shared__ uint64_t s_base;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const bool parser = tid < TOK_PARSE_THREADS;
    if (tid == 0) s_ready = 0;

    for (;;) {
        __syncthreads();                                   // protects s_* reuse across iterations
        if (tid == 0) s_tile = atomicAdd(p.tile_ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= p.n_tiles) break;
        const uint64_t tb = p.tile0 + (uint64_t)tile * TILE_BYTES;      // absolute offset of the tile
        const uint64_t abs0 = tb - TILE_PAD;                             // wraps for a tile at offset 0; only differences are used

        // ---- stage [tb - 16, tb + TILE_BYTES + TILE_TAIL) ; bytes outside the text read as '\n'
        for (int i = tid; i < TILE_SMEM / 16; i += TOK_THREADS) {
            const int64_t a = (int64_t)tb - TILE_PAD + 16 * (int64_t)i;
            uint4 v;
            if (a >= 0 && (uint64_t)a + 16 <= p.text_len) {
                v = __ldg(reinterpret_cast<const uint4*>(p.text + a));
            } else {
                uint32_t w[4] = {0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au};
                for (int b = 0; b < 16; ++b) {
                    const int64_t q = a + b;
                    if (q >= 0 && (uint64_t)q < p.text_len) {
                        w[b >> 2] = (w[b >> 2] & ~(0xFFu << (8 * (b & 3)))) | ((uint32_t)p.text[q] << (8 * (b & 3)));
                    }
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            reinterpret_cast<uint4*>(s_text)[i] = v;
        }
        __syncthreads();

        // ---- line starts: byte q starts a line iff text[q] != '\n' and text[q-1] == '\n'
        // parse thread t owns bytes [128 t, 128 t + 128) of the tile; 16-byte loads rotated across the
        // quarter-warp so that the eight lanes hit eight different bank groups.
        uint64_t st_lo = 0, st_hi = 0;
        uint32_t my_count = 0, incl = 0;
        if (parser) {
            uint64_t nl_lo = 0, nl_hi = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int g = (k + tid) & 7;
                const uint4 v = *reinterpret_cast<const uint4*>(s_text + TILE_PAD + tid * 128 + g * 16);
                const uint64_t m = newline_mask16(v);
                if (g < 4) nl_lo |= m << (16 * g); else nl_hi |= m << (16 * (g - 4));
            }
            const uint64_t prev_nl = s_text[TILE_PAD + tid * 128 - 1] == (uint8_t)'\n' ? 1ull : 0ull;
            st_lo = ((nl_lo << 1) | prev_nl) & ~nl_lo;
            st_hi = ((nl_hi << 1) | (nl_lo >> 63)) & ~nl_hi;
            // restrict to the owned range [range_begin, range_end)
            const uint64_t first = tb + (uint64_t)tid * 128;
            if (first + 128 <= p.range_begin || first >= p.range_end) { st_lo = 0; st_hi = 0; }
            else {
                if (first < p.range_begin) {
                    const int cut = (int)(p.range_begin - first);                 // 1..127 low bits dropped
                    if (cut >= 64) { st_lo = 0; st_hi &= ~0ull << (cut - 64); } else st_lo &= ~0ull << cut;
                }
                if (first + 128 > p.range_end) {
                    const int keep = (int)(p.range_end - first);                  // 1..127 low bits kept
                    if (keep <= 64) { st_hi = 0; st_lo &= keep == 64 ? ~0ull : ((1ull << keep) - 1); }
                    else st_hi &= (1ull << (keep - 64)) - 1;
                }
            }
            my_count = __popcll(st_lo) + __popcll(st_hi);
            incl = my_count;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) s_warp_sums[warp] = incl;
        }
        __syncthreads();
        uint32_t warp_off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TOK_PARSE_THREADS / 32; ++w) {
            const uint32_t v = s_warp_sums[w];
            if (w < warp) warp_off += v;
            total += v;
        }
        if (parser) {
            uint32_t idx = warp_off + incl - my_count;
            uint64_t m = st_lo;
            while (m) { const int b = __ffsll((long long)m) - 1; m &= m - 1; if (idx < TILE_MAX_LINES) s_starts[idx] = (uint16_t)(tid * 128 + b); ++idx; }
            m = st_hi;
            while (m) { const int b = __ffsll((long long)m) - 1; m &= m - 1; if (idx < TILE_MAX_LINES) s_starts[idx] = (uint16_t)(tid * 128 + 64 + b); ++idx; }
        }
        uint32_t n_lines = total;
        if (n_lines > TILE_MAX_LINES) {                    // only possible with lines shorter than 8 bytes
            if (tid == 0) report_error(p, tb, LINE_MALFORMED);
            n_lines = TILE_MAX_LINES;
        }
        __syncthreads();

        FlatSrc gsrc {p.text, p.text_len};
        if (!parser) {
            // ---- service warp: decoupled look-back over per-tile line counts (lane 0) and the
            //      interned name of the tile's first line (lane 1), while the parse warps work
            if (lane == 0) {
                uint64_t base = 0;
                if (tile == 0) {
                    atomicExch(&p.tile_status[0], LB_FLAG_PREFIX | (unsigned long long)n_lines);
                } else {
                    atomicExch(&p.tile_status[tile], LB_FLAG_AGG | (unsigned long long)n_lines);
                    uint32_t i = tile - 1;
                    for (;;) {
                        unsigned long long w;
                        unsigned int spins = 0;
                        do {
                            w = *((volatile unsigned long long*)&p.tile_status[i]);
                            if ((w >> 62) == 0 && ++spins > (1u << 24)) { report_error(p, tb, LINE_MALFORMED + 4); w = LB_FLAG_PREFIX; }
                        } while ((w >> 62) == 0);
                        base += w & LB_VALUE_MASK;
                        if (w & LB_FLAG_PREFIX) break;
                        --i;
                    }
                    atomicExch(&p.tile_status[tile], LB_FLAG_PREFIX | (unsigned long long)(base + n_lines));
                }
                s_base = base;
                if (tile == p.n_tiles - 1) *p.n_sites = base + n_lines;
            } else if (lane == 1) {
                uint32_t ref = 0;
                if (n_lines > 0) {
                    // only a name that starts at the first byte of the line can be shared (same_name_as_first)
                    const uint64_t nabs = tb + s_starts[0];
                    uint64_t q = nabs;
                    uint8_t c = gsrc.at(q);
                    while (!is_delim(c) && !is_eol(c)) c = gsrc.at(++q);
                    const uint32_t len = (uint32_t)(q - nabs);
                    if (len > 0 && is_delim(c)) ref = name_intern(p.names, gsrc, nabs, len);
                }
                s_name0_ref = ref;
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) *((volatile uint32_t*)&s_ready) = tile + 1;     // publishes s_base and s_name0_ref to the parse warps
            continue;
        }

        // ---- parse warps: groups of 32 consecutive lines, one line per lane.  Every lane runs the
        //      tokenizer (lanes past the end re-parse the group's first line and drop the result)
        //      so that the warp reconverges inside it.
        const uint32_t l0_off = n_lines ? TILE_PAD + s_starts[0] : 0;
        const uint2 first8 = n_lines ? load8_unaligned(s_text, l0_off) : make_uint2(0, 0);
        bool have_base = false;
        uint64_t base = 0;
        uint32_t name0_ref = 0;
        for (uint32_t g = warp * 32; g < n_lines; g += TOK_PARSE_THREADS) {
            const uint32_t j = g + lane;
            const bool mine = j < n_lines;
            const uint32_t off = s_starts[mine ? j : g];
            const uint64_t line_abs = tb + off;
            LineResult r;
            bool fast = false;
            if (FAST && !p.want_qual) {
                FastLine fl;
                fast = parse_line_fast_smem(s_text, abs0, TILE_SMEM, line_abs, fl);
                r.status = fl.status; r.pos = fl.pos; r.profile = fl.profile; r.chrom_off = fl.chrom_off; r.chrom_len = fl.chrom_len;
            }
            if (!fast && mine) {
                SmemSrc ssrc {s_text, abs0, (uint32_t)TILE_SMEM, false};
                ParsedLine pl;
                parse_line(ssrc, line_abs, p.want_qual != 0, pl);
                if (ssrc.overrun) parse_line(gsrc, line_abs, p.want_qual != 0, pl);
                r.status = pl.status; r.pos = pl.pos; r.profile = pl.profile; r.chrom_off = pl.chrom_off; r.chrom_len = pl.chrom_len;
            }
            __syncwarp();
            const bool good = mine && r.status == LINE_OK;
            bool same = false;
            uint32_t slot = 0;
            if (good) {
                same = r.chrom_off == 0 && same_name_as_first(s_text, TILE_PAD + off, r.chrom_len, l0_off, first8);
                if (p.use_table) slot = table_find_or_insert(p.table, r.profile);
            }
            if (!have_base) {                              // wait for the service warp (usually long done)
                while (*((volatile uint32_t*)&s_ready) != tile + 1) { }
                __threadfence_block();
                base = *((volatile uint64_t*)&s_base);
                name0_ref = *((volatile uint32_t*)&s_name0_ref);
                have_base = true;
            }
            if (mine && r.status != LINE_OK) report_error(p, line_abs, r.status);
            if (good) {
                const uint64_t site = p.site_base + base + j;
                if (site >= p.site_cap) { report_error(p, line_abs, LINE_MALFORMED + 5); }
                else {
                    const uint32_t ref = (same && name0_ref) ? name0_ref : name_intern(p.names, gsrc, line_abs + r.chrom_off, r.chrom_len);
                    p.pos[site] = r.pos;
                    p.name_ref[site] = ref;
                    if (p.profile) p.profile[site] = r.profile;
                    if (p.line_off) p.line_off[site] = line_abs;
                    if (p.use_table) p.slot[site] = slot;
                }
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace sid
