// K1, round 2: pileup tokenizer + profile builder + join against the unique-profile table and, for streaming
// `-m local`, the classification of new profiles and the CSV rows themselves (one kernel from text to rows).
//   readFile call.cpp:11-20, parsePileupLine pileup.cpp:13-68, parseReadBases pileup.cpp:70-153,
//   countUniqueProfiles pileup.cpp:169-196 (counting half), call.cpp:217-221 (profile index),
//   ROWS form: callSiteMLError body call.cpp:238-285 + operator<< call.hpp:29-38.
//
// Tile/slice geometry, the service warp with its bulk copies and the mbarrier hand-over are those of k_tokenize.cuh.
// What changed:
//   * stage 1 stores the class words of a 32-byte unit as one 32-byte record (two 16-byte stores instead of eight
//     4-byte ones), A/C/G/T as one class plus two raw bit planes, digits and control bytes as classes;
//   * stage 2 (parse_win.cuh) works on one 64-bit window per line held in registers: no byte loads;
//   * SITES form: as before (dense unordered site store + block table -> order[]);
//   * ROWS form: a profile is classified by the lane that inserts it (call_local + format_suffix, a few thousand
//     profiles per million sites), every lane then copies its row -- the line's own "name<sep>position" bytes and the
//     slot's suffix -- into the warp's staging buffer (row_assemble.cuh), and the warp copies the rows of its slice
//     to the slice's region of a scratch buffer with 16-byte stores.  The block table records bytes and rows per
//     region; k_rows_compact lays the regions end to end in file order.  No site store, no order[], no K6.
#pragma once
#include "calls.cuh"
#include "k_quality.cuh"
#include "k_tokenize.cuh"
#include "parse_win.cuh"
#include "parse_units.cuh"
#include "row_assemble.cuh"

namespace sid {

constexpr int ROW_STAGE = 2688;                   // bytes per warp: 32 rows of <= 30 + 46 bytes, a carried tail, slack
constexpr int LINE_ROWS_OVERFLOW = LINE_MALFORMED + 6;   // a slice's rows did not fit its region: the host retries with larger regions

struct Tok2Params {
    const uint8_t* text;
    uint64_t text_len, range_begin, range_end;
    uint64_t tile0;
    uint32_t n_tiles;
    // SITES form
    uint64_t site_base, site_cap;
    uint64_t* profile;
    int32_t* pos;
    uint32_t* slot;
    uint32_t* name_ref;
    uint64_t* line_off;
    uint64_t* fwd;                  // STRANDS: per site, the profile of the bases read on the forward strand
    double* qual_l;                 // QUAL: two doubles per site, the log-likelihood sums of `-m quality` (+inf: not formed here)
    const double* qual_lut;         // QUAL: the per-read term tables (k_quality.cuh)
    unsigned long long* site_alloc;
    // ROWS form
    uint8_t* rows;                  // scratch: region of (tile, slice) at rows + (tile * 8 + slice) * region_cap
    uint32_t region_cap;            // multiple of 16
    double prior, error_threshold, alpha;
    int het_only;
    // both
    unsigned int* tile_ticket;
    unsigned long long* blk;        // SITES: lines << 32 | first storage index;  ROWS: rows << 32 | bytes
    unsigned long long* error;
    TableView table;
    NameDict names;
    int use_table, want_qual, bytewise;
    uint32_t slice_bytes, text_stride, tail_bytes, lines_cap, ext_bytes, units_cap;
};

// units_cap: units (32 bytes) a parse warp classifies, slice + ext, plus the zero padding
SID_HD uint32_t tok2_units(uint32_t slice, uint32_t ext) { return (slice + ext) / 32u; }
inline uint32_t tok2_dyn_smem(uint32_t slice, uint32_t ext, uint32_t stages, bool rows, bool strands = false) {
    const uint32_t units = tok2_units(slice, ext) + CW_PAD_UNITS;
    return stages * tok_text_stride(slice, ext) + 8u * (slice / 8u) * 2u + 8u * units * (CW_WORDS + 1u) * 4u + (rows ? 8u * ROW_STAGE : 0u) +
           (strands ? 8u * units * 4u : 0u);
}

#if defined(__CUDACC__)

// table_find_or_insert with the first probe through L1 (keys never change once set; a stale EMPTY is settled by the
// CAS) and the information whether this lane created the entry.
__device__ __forceinline__ uint32_t table_join(const TableView& t, uint64_t key, bool& inserted) {
    inserted = false;
    if (key == TABLE_EMPTY) {
        if (atomicCAS(t.special_used, 0u, 1u) == 0u) {
            t.keys[t.cap] = key;
            t.entry_list[atomicAdd(t.n_entries, 1u)] = t.cap;
            inserted = true;
        }
        return t.cap;
    }
    uint32_t h = table_hash(key) & t.mask;
    for (uint32_t probes = 0; probes <= t.mask; ++probes) {
        unsigned long long k = t.keys[h];
        if (k == key) return h;
        if (k == TABLE_EMPTY) {
            const unsigned long long old = atomicCAS(&t.keys[h], (unsigned long long)TABLE_EMPTY, (unsigned long long)key);
            if (old == TABLE_EMPTY) {
                t.entry_list[atomicAdd(t.n_entries, 1u)] = h;
                inserted = true;
                return h;
            }
            if (old == key) return h;
        }
        h = (h + 1) & t.mask;
    }
    atomicExch(t.overflow, 1u);
    return 0;
}

// A load served by L2 that the compiler keeps in program order with the loads after it.  (Not ld.acquire: at gpu
// scope that is followed by CCTL.IVALL, an invalidation of the SM's whole L1, once per group of lines: measured, it
// doubled the kernel's time.  The readers below only issue their data loads after a branch on this value, and all of
// them go to L2, where the writer's release made the data visible before the flag.)
__device__ __forceinline__ uint32_t ld_cg_ordered_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Classifies a profile this lane has just inserted and publishes its record (ROWS form).
__device__ __noinline__ void classify_inserted(const TableView& t, uint32_t slot, uint64_t profile, double prior, double error_threshold,
                                                  double alpha, bool het_only) {
    const CallResult r = call_local(profile, prior, error_threshold, alpha);
    t.label[slot] = r.label;
    t.gt[2 * slot] = r.gt0;
    t.gt[2 * slot + 1] = r.gt1;
    t.hom[slot] = r.hom;
    t.het[slot] = r.het;
    uint32_t w[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) w[i] = 0;
    char buf[SUFFIX_BYTES];
    const int n = (het_only && r.label != 1) ? 0 : format_suffix(r, false, buf);
    uint32_t* dst = reinterpret_cast<uint32_t*>(t.suffix + (size_t)slot * SUFFIX_BYTES);
    for (int i = 0; i < n; ++i) w[i >> 2] |= (uint32_t)(uint8_t)buf[i] << (8 * (i & 3));
    for (int i = 0; i < 11; ++i) dst[i] = w[i];
    // the last word carries the length and the ready bit: released after the rest of the record
    st_release_u32(dst + 11, w[11] | ((uint32_t)n | SUFFIX_READY) << 24);
}

// The byte-wise tokenizer for the lines the window parser refuses: out of line, so that its registers and its
// divergent loops stay out of the hot path.
__device__ __noinline__ void parse_line_slow(const uint8_t* txt, uint64_t abs0, uint32_t tile_smem, const uint8_t* text, uint64_t text_len,
                                             uint64_t line_abs, bool want_qual, ParsedLine& pl) {
    SmemSrc ssrc {txt, abs0, tile_smem, false};
    parse_line(ssrc, line_abs, want_qual, pl);
    if (ssrc.overrun) {
        FlatSrc gsrc {text, text_len};
        parse_line(gsrc, line_abs, want_qual, pl);
    }
}

// The forward-strand profile of a line the byte-wise parser took (rare): the same walk over its bases field, from global memory.
__device__ __noinline__ uint64_t forward_profile_slow(const uint8_t* text, uint64_t text_len, uint64_t line_abs, const ParsedLine& pl) {
    FlatSrc src {text, text_len};
    BasesState st;
    st.init((uint8_t)pl.ref);
    uint32_t c[4] = {0, 0, 0, 0};
    for (uint32_t k = 0; k < pl.bases_len; ++k) {
        const uint8_t ch = src.at(line_abs + pl.bases_off + k);
        const int idx = st.feed(ch);
        if (idx < 0) continue;
        const uint8_t seen = ch == '.' ? st.dot_as : ch == ',' ? st.comma_as : ch;      // pileup.cpp:78-83
        if (!(seen & 0x20u)) ++c[idx];
    }
    return pack_profile(c[0], c[1], c[2], c[3]);
}

// Stage 2 of a slice of LONG lines, warp-collective: the lines of the slice one per lane (line_off; `mine` false on
// the lanes without one), their bases fields cut into 64-byte windows, one window per lane and pass (parse_win.cuh:
// win_window).  Returns per lane whether its line stayed within the fast grammar; wl as parse_line_win<true>.
__device__ __forceinline__ bool parse_lines_by_windows(const uint8_t* txt, uint32_t region_off, const uint32_t* cw, const uint32_t* nlw,
                                                    const uint32_t* term_groups, uint32_t n_bits, uint32_t line_off, bool mine, WinLine& wl) {
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31u;
    WinHeader hd;
    Win64 w0;
    bool ok = win_header<true>(txt, region_off, cw, nlw, n_bits, line_off, wl, hd, w0) && mine;
    const uint32_t a = hd.l0 + hd.q4 + 1;                        // first byte of the bases field
    uint32_t b = ok ? win_field_end(cw, n_bits, a, term_groups) : 0u;     // its end
    if (b == 0xFFFFFFFFu || b <= a) { ok = false; b = 0; }
    const uint32_t nw = ok ? (b - a + 63u) >> 6 : 0u;            // windows of this lane's line
    const uint32_t refbits = (hd.ref_base ? 1u : 0u) | (hd.ref_p1 ? 2u : 0u) | (hd.ref_p2 ? 4u : 0u);
    uint32_t incl = nw;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    const uint32_t excl = incl - nw;
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    uint64_t acc = 0;                                            // packed profile of this lane's line
    uint32_t pass_carry = 0;                                     // skip_out of the last window of the pass before
    for (uint32_t g0 = 0; g0 < total; g0 += 32) {
        const uint32_t g = g0 + lane;                            // this lane's window, counted over all lines
        const bool active = g < total;
        // its line: the first j with incl_j > g
        uint32_t j = 0;
#pragma unroll
        for (int step = 16; step; step >>= 1) {
            const uint32_t v = __shfl_sync(FULL, incl, (int)(j + step - 1));
            if (v <= g) j += step;
        }
        j = j < 31u ? j : 31u;
        const uint32_t aj = __shfl_sync(FULL, a, (int)j), bj = __shfl_sync(FULL, b, (int)j), ej = __shfl_sync(FULL, excl, (int)j);
        const uint32_t rj = __shfl_sync(FULL, refbits, (int)j);
        const uint32_t k = g - ej;                               // window k of line j
        const uint32_t pos = aj + 64u * k;
        const bool cont = active && k > 0;                       // the window before it belongs to the same line
        uint32_t skip_in = (cont && lane == 0) ? pass_carry : 0u;
        // every lane evaluates its window as if nothing reached into it (r0), then takes what its predecessor reports;
        // a skip that lands on plain letters only takes their counts off again (win_window_patch), anything else is a
        // second evaluation.  A lane is final one round after the lane before it: one or two rounds in practice.
        const bool in_reach = active && pos + 64u <= n_bits;
        const Win64 w = load_window(cw, in_reach ? pos : 0u);
        WinPart r0;
        r0.cn = r0.c1 = r0.c2 = r0.c12 = r0.cd = 0;
        r0.skip_out = 0;
        r0.ok = true;
        if (active) r0 = win_window_of(w, in_reach, txt, region_off, pos, bj, 0u);
        WinPart r = r0;
        if (skip_in) {                                           // lane 0 continuing a line from the pass before
            if (!win_window_patch(w, pos, bj, r0, skip_in, r)) r = win_window_of(w, in_reach, txt, region_off, pos, bj, skip_in);
        }
        for (int round = 0; round < 34; ++round) {
            const uint32_t up = __shfl_up_sync(FULL, r.skip_out, 1);
            const uint32_t want = (cont && lane > 0) ? up : skip_in;
            const bool redo = active && want != skip_in;
            skip_in = want;
            if (!__any_sync(FULL, redo)) break;
            bool full = false;
            if (redo) {
                if (skip_in == 0) r = r0;
                else full = !win_window_patch(w, pos, bj, r0, skip_in, r);
            }
            if (__any_sync(FULL, full)) {
                if (full) r = win_window_of(w, in_reach, txt, region_off, pos, bj, skip_in);
            }
        }
        pass_carry = __shfl_sync(FULL, r.skip_out, 31);
        // the window's share of its line's profile, summed over the lanes of the line (runs of consecutive lanes)
        WinHeader hj;
        hj.l0 = 0; hj.q4 = 0;
        hj.ref_base = (rj & 1u) != 0; hj.ref_p1 = (rj & 2u) != 0; hj.ref_p2 = (rj & 4u) != 0;
        uint64_t v = active ? win_profile(r.cn, r.c1, r.c2, r.c12, r.cd, hj) : 0ull;
        const uint32_t seg = active ? j : 0xFFFFFFFFu;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_down_sync(FULL, v, d);
            const uint32_t os = __shfl_down_sync(FULL, seg, d);
            if (lane + (uint32_t)d < 32u && active && os == seg) v += o;
        }
        // line `lane` collects from the first lane of its run in this pass
        const uint32_t s0 = excl > g0 ? excl : g0;
        const bool hit = nw != 0 && s0 < incl && s0 < g0 + 32u;
        const uint32_t src = hit ? s0 - g0 : 0u;
        const uint64_t part = __shfl_sync(FULL, v, (int)src);
        const uint32_t refused = __ballot_sync(FULL, active && !r.ok);
        if (hit) {
            acc += part;
            const uint32_t e0 = (incl < g0 + 32u ? incl : g0 + 32u) - g0;          // lanes [src, e0) hold this line's windows
            const uint32_t m = (e0 >= 32u ? FULL : ((1u << e0) - 1u)) & ~((1u << src) - 1u);
            if (refused & m) ok = false;
        }
    }
    wl.profile = acc;
    wl.status = LINE_OK;
    return ok;
}

#ifndef SID_STAGE2_UNITS
#define SID_STAGE2_UNITS 1      // stage 2 of ordinary lines: 1 = unit by unit (parse_units.cuh), 0 = 64-bit windows (parse_win.cuh)
#endif
#ifndef SID_TOK2_CTAS
#define SID_TOK2_CTAS 3         // CTAs per SM the register allocator must leave room for (288 threads each: 72 registers)
#endif
// DEEP: the instantiation for long lines (hundreds of bytes): slices of a few lines run stage 2 one window per lane
// (parse_lines_by_windows).  Its own kernel so that the code and the registers of that path stay out of the ordinary one
// (folded into one kernel behind a flag it cost the depth-30 path 30 %); shared memory holds two such CTAs per SM anyway.
#ifndef SID_TOK2_QUAL_CTAS
#define SID_TOK2_QUAL_CTAS 2      // CTAs per SM the register allocator leaves room for in the QUAL instantiation
#endif
// QUAL: the instantiation of quality sessions: the per-read sums of callQualityBasedSimple are formed right here, from the
// class windows of the line (k_quality.cuh: quality_sums_win), instead of a second byte-wise walk over the text in k_quality.
// STRANDS: the instantiation that also keeps, per site, the profile of the forward strand (the strands
// parseReadBases derives and nobody reads): stage 1 stores bit plane 5 of every unit, stage 2 counts the upper-case letters
// and the '.' beside the profile.
template <bool ROWS, int TOK_STAGES, bool DEEP = false, bool QUAL = false, bool STRANDS = false>
__global__ void __launch_bounds__(TOK_THREADS, DEEP ? 2 : (QUAL ? SID_TOK2_QUAL_CTAS : SID_TOK2_CTAS)) k_tok2(const Tok2Params p) {
    static_assert(!STRANDS || (!ROWS && !DEEP && !QUAL), "strand counts come from the ordinary sites form");
    extern __shared__ __align__(128) uint8_t s_dyn[];
    __shared__ StageMeta s_meta[TOK_STAGES];
    __shared__ __align__(8) uint64_t s_full[TOK_STAGES], s_done[TOK_STAGES];
    __shared__ uint32_t s_term_groups[DEEP ? TOK_PARSE_WARPS : 1][(SLICE_MAX + 2016) / 1024 + 2];     // DEEP, per warp: which units hold a byte <= 0x20

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
        // ================================================================= service warp (as in k_tokenize)
        auto issue_load = [&](int b, uint32_t tile) {
            const uint64_t tb = p.tile0 + (uint64_t)tile * tile_bytes;
            uint8_t* txt = s_dyn + (size_t)b * p.text_stride;
            if (lane == 0) {
                s_meta[b].tb = tb;
                s_meta[b].tile = (int32_t)tile;
            }
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
            {
                const int nb = (it + 1) % TOK_STAGES;
                if (it + 1 >= TOK_STAGES && !mbar_wait<SVC_SLEEP>(&s_done[nb], ((it + 1) / TOK_STAGES - 1) & 1)) {
                    if (lane == 0) report_error_at(p.error, 0, LINE_MALFORMED + 4);
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
    const int pw = warp - 1;
    const uint32_t units = tok2_units(slice, p.ext_bytes), own_units = slice / 32;
    uint8_t* const after_text = s_dyn + (size_t)TOK_STAGES * p.text_stride;
    uint16_t* const starts = reinterpret_cast<uint16_t*>(after_text) + (size_t)pw * p.lines_cap;
    uint32_t* const cw_all = reinterpret_cast<uint32_t*>(after_text + (size_t)TOK_PARSE_WARPS * p.lines_cap * 2);
    uint32_t* const cw = cw_all + (size_t)pw * p.units_cap * CW_WORDS;                    // 32-byte records: 16-byte aligned
    uint32_t* const nlw = cw_all + (size_t)TOK_PARSE_WARPS * p.units_cap * CW_WORDS + (size_t)pw * p.units_cap;
    uint8_t* const stage = reinterpret_cast<uint8_t*>(cw_all + (size_t)TOK_PARSE_WARPS * p.units_cap * (CW_WORDS + 1)) + (size_t)pw * ROW_STAGE;
    uint32_t* const p5w = cw_all + (size_t)TOK_PARSE_WARPS * p.units_cap * (CW_WORDS + 1) + (size_t)pw * p.units_cap;     // STRANDS (never with ROWS)
    uint32_t cache_len = 0, cache_ref = 0;
    uint4 cache_name = make_uint4(0, 0, 0, 0);
    for (uint32_t it = 0;; ++it) {
        const int b = it % TOK_STAGES;
        const uint32_t use = it / TOK_STAGES;
        if (!mbar_wait<PARSE_SLEEP>(&s_full[b], use & 1)) {
            if (lane == 0) report_error_at(p.error, 0, LINE_MALFORMED + 4);
            break;
        }
        if (s_meta[b].tile < 0) break;
        const uint8_t* txt = s_dyn + (size_t)b * p.text_stride;
        const uint64_t tb = s_meta[b].tb;
        const uint64_t abs0 = tb - TILE_PAD;
        const uint32_t slice_off = (uint32_t)pw * slice;
        const uint32_t region_off = TILE_PAD + slice_off;
        // ---- stage 1: classify the slice (+ ext) unit by unit; the '\n' words give the line starts
        uint32_t n_lines = 0, bad = 0;
        {
            const bool inside = tb + slice_off >= p.range_begin && tb + slice_off + slice <= p.range_end;
            uint32_t carry_nl = txt[region_off - 1] == (uint8_t)'\n' ? 1u : 0u;
            for (uint32_t u0 = 0; u0 < units; u0 += 32) {
                const uint32_t u = u0 + lane;
                uint32_t st = 0, nl = 0, tw = 0;
                if (u < units) {
                    const uint8_t* up = txt + region_off + u * 32;
                    const uint4 v0 = *reinterpret_cast<const uint4*>(up), v1 = *reinterpret_cast<const uint4*>(up + 16);
                    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    const UnitClasses k = classify_unit(w);
                    uint4* rec = reinterpret_cast<uint4*>(cw + (size_t)u * CW_WORDS);
                    rec[0] = make_uint4(k.w[0], k.w[1], k.w[2], k.w[3]);
                    rec[1] = make_uint4(k.w[4], k.w[5], k.w[6], k.w[7]);
                    nlw[u] = k.nl;
                    if (STRANDS) p5w[u] = k.p5;
                    nl = k.nl;
                    bad |= k.bad;
                    if (DEEP) tw = k.w[CW_TERM];
                }
                if (DEEP) {
                    const uint32_t tg = __ballot_sync(0xFFFFFFFFu, tw != 0);
                    if (lane == 0) s_term_groups[pw][u0 >> 5] = tg;
                }
                uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, nl >> 31, 1);
                if (lane == 0) prev = carry_nl;
                carry_nl = __shfl_sync(0xFFFFFFFFu, nl >> 31, 31);
                if (u < own_units) {
                    st = ((nl << 1) | prev) & ~nl;
                    if (!inside) {
                        const uint64_t first = tb + slice_off + u * 32;
                        if (first + 32 <= p.range_begin || first >= p.range_end) st = 0;
                        else {
                            if (first < p.range_begin) st &= 0xFFFFFFFFu << (uint32_t)(p.range_begin - first);
                            if (first + 32 > p.range_end) st &= 0xFFFFFFFFu >> (32u - (uint32_t)(p.range_end - first));
                        }
                    }
                }
                // lines of 32 bytes and more start at most once per unit: their index is a population count of the
                // ballot; a unit with two starts (short lines) sends the group through the general scan
                const uint32_t has = __ballot_sync(0xFFFFFFFFu, st != 0);
                if (!__any_sync(0xFFFFFFFFu, (st & (st - 1)) != 0)) {
                    const uint32_t idx = n_lines + __popc(has & ((1u << lane) - 1u));
                    if (st && idx < p.lines_cap) starts[idx] = (uint16_t)(slice_off + u * 32 + (__ffs((int)st) - 1));
                    n_lines += __popc(has);
                } else {
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
            }
            if (lane < (int)CW_PAD_UNITS) {
                uint4* rec = reinterpret_cast<uint4*>(cw + (size_t)(units + lane) * CW_WORDS);
                rec[0] = make_uint4(0, 0, 0, 0);
                rec[1] = make_uint4(0, 0, 0, 0);
                nlw[units + lane] = 0;
                if (STRANDS) p5w[units + lane] = 0;
            }
            if (n_lines > p.lines_cap) {
                if (lane == 0) report_error_at(p.error, tb, LINE_MALFORMED);
                n_lines = p.lines_cap;
            }
            bad = __any_sync(0xFFFFFFFFu, bad != 0) ? 1u : 0u;      // a control byte somewhere: the whole slice goes byte by byte
            __syncwarp();
        }
        const bool win_ok = !bad && !p.bytewise;
        const uint32_t n_bits = slice + p.ext_bytes;
        const uint32_t region = (uint32_t)s_meta[b].tile * TOK_PARSE_WARPS + pw;

        if (!ROWS) {
            // ================================================================= SITES form
            uint64_t base = 0;
            {
                unsigned long long bb = 0;
                if (lane == 0) {
                    if (n_lines) bb = p.site_base + atomicAdd(p.site_alloc, (unsigned long long)n_lines);
                    p.blk[region] = ((unsigned long long)n_lines << 32) | (bb & 0xFFFFFFFFull);
                }
                base = __shfl_sync(0xFFFFFFFFu, bb, 0);
            }
            uint32_t l0_off = 0, name0_ref = 0;
            uint2 first8 = make_uint2(0, 0);
            if (n_lines) {
                l0_off = TILE_PAD + starts[0];
                first8 = load8_unaligned(txt, l0_off);
            }
            const bool by_windows = DEEP && n_lines <= 10u;
            for (uint32_t g = 0; g < n_lines; g += 32) {
                const uint32_t j = g + lane;
                const bool mine = j < n_lines;
                const uint32_t off = starts[mine ? j : g];
                const uint64_t line_abs = tb + off;
                LineResult r;
                r.status = LINE_MALFORMED;
                bool fast = false;
                uint64_t fwd = 0;                                             // STRANDS
                double ql1 = bits_double(0x7FF0000000000000ull), ql2 = 0;     // QUAL: +inf = "k_quality walks this line itself"
                if (win_ok) {
                    WinLine wl;
                    // a handful of long lines: one 64-byte window of a bases field per lane instead of one line per lane
                    if (DEEP && by_windows) fast = parse_lines_by_windows(txt, region_off, cw, nlw, s_term_groups[pw], n_bits, TILE_PAD + off, mine, wl);
                    else if (QUAL) {
                        WinHeader hd;
                        uint32_t end_bit = 0;
                        fast = parse_line_win<true>(txt, region_off, cw, nlw, n_bits, TILE_PAD + off, wl, &hd, &end_bit);
                        if (fast && mine) {
                            double a, b;
                            if (quality_sums_win(txt, region_off, cw, nlw, n_bits, hd, end_bit, wl.profile, p.qual_lut, a, b)) { ql1 = a; ql2 = b; }
                        }
                        __syncwarp();
                    }
#if SID_STAGE2_UNITS
                    else if (STRANDS) fast = parse_line_units<true, true>(txt, region_off, cw, nlw, n_bits, TILE_PAD + off, wl, p5w, &fwd);
                    else fast = parse_line_units<true>(txt, region_off, cw, nlw, n_bits, TILE_PAD + off, wl);
#else
                    else fast = parse_line_win<true>(txt, region_off, cw, nlw, n_bits, TILE_PAD + off, wl);
#endif
                    r.status = wl.status; r.pos = wl.pos; r.profile = wl.profile; r.chrom_off = 0; r.chrom_len = wl.name_len;
                }
                if (!fast && mine) {
                    ParsedLine pl;
                    parse_line_slow(txt, abs0, tile_smem, p.text, p.text_len, line_abs, p.want_qual != 0, pl);
                    r.status = pl.status; r.pos = pl.pos; r.profile = pl.profile; r.chrom_off = pl.chrom_off; r.chrom_len = pl.chrom_len;
                    if (STRANDS && pl.status == LINE_OK) fwd = forward_profile_slow(p.text, p.text_len, line_abs, pl);
                }
                __syncwarp();
                const bool good = mine && r.status == LINE_OK;
                if (g == 0) {
                    if (lane == 0 && good && r.chrom_off == 0) {
                        const uint32_t len = r.chrom_len;
                        uint4 nm = make_uint4(0, 0, 0, 0);
                        if (len <= 16) {
                            const uint2 hi = load8_unaligned(txt, l0_off + 8);
                            // the name's bytes, zero beyond its length
                            const unsigned long long lo64 = ((unsigned long long)first8.y << 32) | first8.x, hi64 = ((unsigned long long)hi.y << 32) | hi.x;
                            const unsigned long long mlo = len >= 8 ? ~0ull : ((1ull << (8 * len)) - 1ull);
                            const unsigned long long mhi = len <= 8 ? 0ull : len >= 16 ? ~0ull : ((1ull << (8 * (len - 8))) - 1ull);
                            const unsigned long long a64 = lo64 & mlo, b64 = hi64 & mhi;
                            nm = make_uint4((uint32_t)a64, (uint32_t)(a64 >> 32), (uint32_t)b64, (uint32_t)(b64 >> 32));
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
                    bool inserted;
                    if (p.use_table) slot = table_join(p.table, r.profile, inserted);
                }
                if (mine && r.status != LINE_OK) report_error_at(p.error, line_abs, r.status);
                if (good) {
                    const uint64_t site = base + j;
                    if (site >= p.site_cap) { report_error_at(p.error, line_abs, LINE_MALFORMED + 5); }
                    else {
                        const uint32_t ref = same ? name0_ref : name_intern(p.names, gsrc, line_abs + r.chrom_off, r.chrom_len);
                        p.pos[site] = r.pos;
                        p.name_ref[site] = ref;
                        if (p.profile) p.profile[site] = r.profile;
                        if (p.line_off) p.line_off[site] = line_abs;
                        if (STRANDS) p.fwd[site] = fwd;
                        if (QUAL) { p.qual_l[2 * site] = fast ? ql1 : bits_double(0x7FF0000000000000ull); p.qual_l[2 * site + 1] = ql2; }
                        if (p.use_table) p.slot[site] = slot;
                    }
                }
            }
        } else {
            // ================================================================= ROWS form
            uint8_t* const reg = p.rows + (size_t)region * p.region_cap;
            if (lane == 0 && n_lines) atomicAdd(p.site_alloc, (unsigned long long)n_lines);        // the call's site count
            uint32_t written = 0;           // bytes of the region already in global memory (multiple of 16)
            uint32_t fill = 0;              // bytes in the staging buffer: region bytes [written, written + fill)
            uint32_t n_rows = 0;
            bool overflow = false;
            for (uint32_t g = 0; g < n_lines; g += 32) {
                const uint32_t j = g + lane;
                const bool mine = j < n_lines;
                const uint32_t off = starts[mine ? j : g];
                const uint64_t line_abs = tb + off;
                WinLine wl;
                wl.status = LINE_MALFORMED;
                bool fast = false;
#ifdef SID_WHATIF_NO_STAGE2
                wl.status = LINE_OK; wl.profile = 30 + (off & 3); wl.name_len = 4; wl.hdr_len = 13; wl.pos_canonical = true; fast = true;
#else
                if (win_ok) {
                    fast = parse_line_win<false>(txt, region_off, cw, nlw, n_bits, TILE_PAD + off, wl);
                    fast = fast && wl.pos_canonical;
                }
#endif
                uint64_t profile = wl.profile;
                int status = wl.status;
                uint32_t name_off = 0, name_len = wl.name_len, hdr_len = wl.hdr_len;
                int32_t pos = 0;
                if (!fast && mine) {
                    ParsedLine pl;
                    parse_line_slow(txt, abs0, tile_smem, p.text, p.text_len, line_abs, false, pl);
                    status = pl.status; profile = pl.profile; pos = pl.pos; name_off = pl.chrom_off; name_len = pl.chrom_len;
                    hdr_len = name_len + 1 + (uint32_t)digits_i32(pos);
                }
                __syncwarp();
                const bool good = mine && status == LINE_OK;
                if (mine && status != LINE_OK) report_error_at(p.error, line_abs, status);
                // ---- join; the lane that creates an entry classifies it
                uint32_t slot = 0;
                bool inserted = false;
#ifndef SID_WHATIF_NO_JOIN
                if (good) slot = table_join(p.table, profile, inserted);
#else
                slot = (uint32_t)(profile & 1023u);
#endif
                __syncwarp();
                if (inserted) classify_inserted(p.table, slot, profile, p.prior, p.error_threshold, p.alpha, p.het_only != 0);
                __syncwarp();
                // ---- the slot's suffix record (read at L2: it may have been written a moment ago by another SM)
                RowSrc rs;
                rs.line_off = TILE_PAD + off;
                rs.hdr_len = hdr_len;
                rs.name_len = name_len;
                rs.sfx_len = 0;
#pragma unroll
                for (int i = 0; i < 12; ++i) rs.sfx[i] = 0;
#ifdef SID_WHATIF_NO_SUFFIX_LOAD
                if (good) { rs.sfx_len = 28; rs.sfx[0] = 0x6D6F682Cu; rs.sfx[6] = 0x0A657565u; }
                if (false) {
#else
                if (good) {
#endif
                    const uint32_t* rec = reinterpret_cast<const uint32_t*>(p.table.suffix + (size_t)slot * SUFFIX_BYTES);
                    uint32_t last = ld_cg_ordered_u32(rec + 11);
                    for (uint32_t spin = 0; !((last >> 24) & SUFFIX_READY); ++spin) {        // another SM is still classifying it
                        if (spin > (1u << 22)) { report_error_at(p.error, line_abs, LINE_MALFORMED + 4); break; }
                        __nanosleep(200);
                        last = ld_cg_ordered_u32(rec + 11);
                    }
                    const uint4 v0 = __ldcg(reinterpret_cast<const uint4*>(rec)), v1 = __ldcg(reinterpret_cast<const uint4*>(rec) + 1);
                    const uint2 v2 = __ldcg(reinterpret_cast<const uint2*>(rec) + 4);
                    const uint32_t v3 = __ldcg(rec + 10);
                    rs.sfx[0] = v0.x; rs.sfx[1] = v0.y; rs.sfx[2] = v0.z; rs.sfx[3] = v0.w;
                    rs.sfx[4] = v1.x; rs.sfx[5] = v1.y; rs.sfx[6] = v1.z; rs.sfx[7] = v1.w;
                    rs.sfx[8] = v2.x; rs.sfx[9] = v2.y; rs.sfx[10] = v3; rs.sfx[11] = last & 0x00FFFFFFu;
                    rs.sfx_len = (last >> 24) & 0x7Fu;
                }
                const uint32_t row_len = rs.sfx_len ? hdr_len + rs.sfx_len : 0u;
                n_rows += __popc(__ballot_sync(0xFFFFFFFFu, row_len != 0));
                // ---- offsets of the rows in the staging buffer
                uint32_t incl = row_len;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += o;
                }
                const uint32_t excl = incl - row_len;
                // rows are staged in batches of consecutive lanes that fit the buffer (ordinarily one batch of 32)
                uint32_t a_lane = 0;
                while (a_lane < 32) {
                    const uint32_t excl_a = __shfl_sync(0xFFFFFFFFu, excl, a_lane);
                    const bool fits = fill + (incl - excl_a) <= (uint32_t)ROW_STAGE - 16u;
                    const uint32_t nofit = __ballot_sync(0xFFFFFFFFu, !fits && (uint32_t)lane >= a_lane && row_len != 0);
                    const uint32_t b_lane = nofit ? (uint32_t)(__ffs((int)nofit) - 1) : 32u;
                    const bool in_batch = (uint32_t)lane >= a_lane && (uint32_t)lane < b_lane && row_len != 0;
                    const uint32_t d = fill + (excl - excl_a);
                    const uint32_t batch_bytes = __shfl_sync(0xFFFFFFFFu, incl, b_lane ? b_lane - 1 : 0) - excl_a;
                    if (b_lane == a_lane) {
                        // the row of lane a_lane alone exceeds the buffer (a name of kilobytes): straight to the region, byte by byte
                        const uint32_t len_a = __shfl_sync(0xFFFFFFFFu, row_len, a_lane);
                        if ((uint64_t)written + fill + len_a + 32 > p.region_cap) overflow = true;
                        if (!overflow) {
                            if ((uint32_t)lane < fill) reg[written + lane] = stage[lane];                       // fill < 16 here
                            if ((uint32_t)lane == a_lane)
                                row_bytewise(reg + written + fill, gsrc, line_abs + name_off, name_len, pos,
                                             reinterpret_cast<const uint8_t*>(p.table.suffix + (size_t)slot * SUFFIX_BYTES), rs.sfx_len);
                            __threadfence();
                            __syncwarp();
                            const uint32_t end = written + fill + len_a;
                            const uint32_t keep = end & 15u;
                            if ((uint32_t)lane < keep) stage[lane] = *((volatile uint8_t*)(reg + (end & ~15u) + lane));
                            written = end & ~15u;
                            fill = keep;
                            __syncwarp();
                        }
                        a_lane += 1;
                        continue;
                    }
                    // phase A: whole words; phase B: shared first words, commas, byte-wise rows
                    uint32_t hw = 0, sxw = 0;
                    if (in_batch && fast) {
                        hw = ((d & 3u) + hdr_len + 3u) >> 2;
                        sxw = (((d + hdr_len) & 3u) + rs.sfx_len + 3u) >> 2;
                    }
#pragma unroll
                    for (int dd = 16; dd; dd >>= 1) {
                        hw = max(hw, __shfl_xor_sync(0xFFFFFFFFu, hw, dd));
                        sxw = max(sxw, __shfl_xor_sync(0xFFFFFFFFu, sxw, dd));
                    }
                    uint32_t first = 0;
#ifndef SID_WHATIF_NO_ASSEMBLY
                    if (in_batch && fast) first = row_phase_a(txt, stage, d, rs, hw, sxw);
#endif
                    __syncwarp();
#ifndef SID_WHATIF_NO_ASSEMBLY
                    if (in_batch) {
                        if (fast) row_phase_b(stage, d, rs, first);
                        else row_bytewise(stage + d, gsrc, line_abs + name_off, name_len, pos,
                                          reinterpret_cast<const uint8_t*>(p.table.suffix + (size_t)slot * SUFFIX_BYTES), rs.sfx_len);
                    }
#endif
                    __syncwarp();
                    fill += batch_bytes;
                    // ---- whole 16-byte chunks go to the region; the tail stays for the next batch
                    const uint32_t n16 = fill >> 4;
                    if ((uint64_t)written + fill + 32 > p.region_cap) overflow = true;
                    if (!overflow) {
                        const uint4* sv = reinterpret_cast<const uint4*>(stage);
                        uint4* gv = reinterpret_cast<uint4*>(reg + written);
                        for (uint32_t i = lane; i < n16; i += 32) gv[i] = sv[i];
                    }
                    __syncwarp();
                    uint8_t tail_byte = 0;
                    const uint32_t keep = fill & 15u;
                    if (n16 && (uint32_t)lane < keep) tail_byte = stage[(n16 << 4) + lane];
                    __syncwarp();
                    if (n16 && (uint32_t)lane < keep) stage[lane] = tail_byte;
                    __syncwarp();
                    written += n16 << 4;
                    fill = keep;
                    a_lane = b_lane;
                }
            }
            // ---- the last partial chunk (bytes past the end are never read: the block table carries the byte count)
            if (fill && !overflow) {
                if (lane == 0) *reinterpret_cast<uint4*>(reg + written) = *reinterpret_cast<const uint4*>(stage);
            }
            if (overflow && lane == 0) report_error_at(p.error, tb + slice_off, LINE_ROWS_OVERFLOW);
            if (lane == 0) p.blk[region] = ((unsigned long long)n_rows << 32) | (unsigned long long)(written + fill);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_done[b]);
    }
}

// ---- ROWS form, second half: the regions laid end to end in file order ------------------------------------------
constexpr int RC_THREADS = 512;               // 16 warps, one region per warp at a time
constexpr int RC_REGIONS = 256;               // regions per CTA
constexpr int RC_STAGE = 4096 + 32;           // bytes of shared memory per warp (larger regions are copied byte-wise)

// part[c] = bytes << 0 of chunk c (RC_REGIONS regions), rows in part_rows[c]
__global__ void __launch_bounds__(256) k_rows_sums(const unsigned long long* blk, uint32_t n_regions, unsigned long long* part, unsigned long long* part_rows) {
    __shared__ unsigned long long s_b[8], s_r[8];
    const uint32_t r = blockIdx.x * RC_REGIONS + threadIdx.x;
    unsigned long long bytes = 0, rows = 0;
    if (threadIdx.x < RC_REGIONS && r < n_regions) { const unsigned long long e = blk[r]; bytes = e & 0xFFFFFFFFull; rows = e >> 32; }
#pragma unroll
    for (int d = 16; d; d >>= 1) { bytes += __shfl_xor_sync(0xFFFFFFFFu, bytes, d); rows += __shfl_xor_sync(0xFFFFFFFFu, rows, d); }
    if ((threadIdx.x & 31) == 0) { s_b[threadIdx.x >> 5] = bytes; s_r[threadIdx.x >> 5] = rows; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tb = 0, tr = 0;
        for (int w = 0; w < 8; ++w) { tb += s_b[w]; tr += s_r[w]; }
        part[blockIdx.x] = tb;
        part_rows[blockIdx.x] = tr;
    }
}

// exclusive scan of the chunk sums in place (one block); totals out
__global__ void __launch_bounds__(1024) k_rows_scan(unsigned long long* part, const unsigned long long* part_rows, uint32_t n_chunks,
                                                    unsigned long long* bytes_out, unsigned long long* rows_out) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    unsigned long long rows = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 1024) {
        const uint32_t c = c0 + threadIdx.x;
        const unsigned long long v = c < n_chunks ? part[c] : 0ull;
        if (c < n_chunks) rows += part_rows[c];
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        unsigned long long before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (c < n_chunks) part[c] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + incl;
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) rows += __shfl_xor_sync(0xFFFFFFFFu, rows, d);
    if (lane == 0) s_w[warp] = rows;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tr = 0;
        for (int w = 0; w < 32; ++w) tr += s_w[w];
        *bytes_out = s_carry;
        *rows_out = tr;
    }
}

__global__ void __launch_bounds__(RC_THREADS) k_rows_compact(const uint8_t* rows, uint32_t region_cap, const unsigned long long* blk, uint32_t n_regions,
                                                             const unsigned long long* part, uint8_t* out, uint64_t out_cap) {
    extern __shared__ __align__(16) uint8_t s_rc[];
    __shared__ uint32_t s_off[RC_REGIONS + 1];
    __shared__ uint32_t s_wsum[RC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t r0 = blockIdx.x * RC_REGIONS;
    // exclusive scan of the chunk's byte counts
    uint32_t v = 0;
    if (tid < RC_REGIONS && r0 + tid < n_regions) v = (uint32_t)blk[r0 + tid];
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += s_wsum[w];
    if (tid < RC_REGIONS) s_off[tid] = before + incl - v;
    if (tid == RC_REGIONS - 1) s_off[RC_REGIONS] = before + incl;
    __syncthreads();
    const unsigned long long chunk_base = part[blockIdx.x];
    uint8_t* st = s_rc + (size_t)warp * RC_STAGE;
    for (uint32_t k = warp; k < RC_REGIONS && r0 + k < n_regions; k += RC_THREADS / 32) {
        const uint32_t n = s_off[k + 1] - s_off[k];
        if (n == 0) continue;
        const unsigned long long dst0 = chunk_base + s_off[k];
        if (dst0 + n > out_cap) continue;                              // the host reports SIDGPU_ECAPACITY from the total
        const uint8_t* src = rows + (size_t)(r0 + k) * region_cap;
        uint8_t* dst = out + dst0;
        if (n + 64 > RC_STAGE) {                                       // an unusually large region: plain byte copy
            for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
            continue;
        }
        // stage the region at (destination offset mod 16) so that 16-byte chunks of the destination line up
        const uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
        const uint32_t n16 = (n + 15) >> 4;
        if (mis == 0) {
            const uint32_t body = n >> 4;
            const uint4* sv = reinterpret_cast<const uint4*>(src);
            uint4* gv = reinterpret_cast<uint4*>(dst);
            for (uint32_t i = lane; i < body; i += 32) gv[i] = __ldcs(sv + i);
            const uint32_t done = body << 4;
            if (done + lane < n) dst[done + lane] = src[done + lane];
            continue;
        }
        {
            const uint4* sv = reinterpret_cast<const uint4*>(src);
            uint4* tv = reinterpret_cast<uint4*>(st + 16);
            for (uint32_t i = lane; i < n16; i += 32) tv[i] = __ldcs(sv + i);
        }
        __syncwarp();
        // bytes of the region start at st + 16; destination-aligned chunk c (c >= 1) = region bytes [16c - mis, +16)
        const uint32_t head = 16u - mis;                               // bytes before the first aligned chunk
        if ((uint32_t)lane < min(head, n)) dst[lane] = st[16 + lane];
        if (n > head) {
            const uint32_t rest = n - head;
            const uint32_t body = rest >> 4;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(st + 16 + (head & ~3u));
            const uint32_t sh = (head & 3u) * 8u;
            uint4* gv = reinterpret_cast<uint4*>(dst + head);
            for (uint32_t i = lane; i < body; i += 32) {
                const uint32_t* q = w + 4 * i;
                const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
                gv[i] = make_uint4(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh), __funnelshift_r(a3, a4, sh));
            }
            const uint32_t done = head + (body << 4);
            if (done + lane < n) dst[done + lane] = st[16 + done + lane];
        }
        __syncwarp();
    }
}

#endif  // __CUDACC__

}  // namespace sid
