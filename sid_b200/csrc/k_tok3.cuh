// K1, streaming form: every warp is its own pipeline.
//   readFile call.cpp:11-20, parsePileupLine pileup.cpp:13-68, parseReadBases pileup.cpp:70-153,
//   callSiteMLError body call.cpp:238-285 + operator<< call.hpp:29-38 (ROWS), countUniqueProfiles pileup.cpp:169-196
//   (the join against the unique-profile table).
//
// What the ncu captures of k_tok2 showed (profiles/r2_*): the eight parse warps of a CTA share staged tiles, so they
// start every tile together, reach the table lookups together and wait for the slowest of them before the stage is
// handed back -- the latency of one warp is nobody's cover.  And the slice geometry wastes instructions: 79 units per
// slice are 2.47 warp iterations of stage 1 (three are issued), every warp re-classifies the `ext` bytes of its
// neighbour, and the line groups of stage 2 are 29.5 lines on 32 lanes with a second, nearly empty group now and then.
//
// Here a warp draws a CHUNK (64 KiB of text by default) from an atomic ticket and streams it through a private ring
// in shared memory, one 1 KiB BLOCK (32 units of 32 bytes: one unit per lane, no partial iteration, no overlap
// between warps) at a time:
//   * lane 0 keeps two bulk asynchronous copies (cp.async.bulk -> the warp's own mbarriers) in flight ahead of the
//     block being classified; nobody else touches the ring, so there is no barrier between warps at all;
//   * stage 1 classifies the block (classify_unit, parse_win.cuh) into the class ring and appends the line starts it
//     finds to a queue;
//   * as soon as 32 complete lines are queued (or the ring needs the room) stage 2 runs on them, one line per lane:
//     parse_line_win on the ring, join against the profile table, classification of new profiles by the inserting
//     lane, rows assembled in the warp's staging buffer and copied to the chunk's region with 16-byte stores.
// A chunk owns the lines whose first byte lies in it and reads on past its end to finish the last one (the sharding
// rule of include/sidgpu.h applied per chunk).  Lines the ring cannot hold (longer than two blocks) and lines outside
// the fast grammar go through the byte-wise tokenizer (parse.cuh) straight from global memory.
// k_rows3_scan / k_rows3_copy lay the chunks' regions end to end in file order.
#pragma once
#include "k_tok2.cuh"

namespace sid {

constexpr uint32_t T3_BLOCK = 1024;                       // bytes per block: 32 units
constexpr uint32_t T3_NBLK = 4;                           // blocks in the ring
constexpr uint32_t T3_RING = T3_BLOCK * T3_NBLK;          // bytes of text a warp holds
constexpr uint32_t T3_UNITS = T3_RING / 32;
constexpr uint32_t T3_QCAP = 128;                         // queued line starts (power of two)
constexpr uint32_t T3_AHEAD = 2;                          // blocks in flight ahead of the one being classified
#ifndef SID_TOK3_WARPS
#define SID_TOK3_WARPS 2
#endif
constexpr int T3_WARPS = SID_TOK3_WARPS;                  // warps per CTA; they never talk to each other
constexpr int T3_THREADS = 32 * T3_WARPS;

struct Tok3Warp {                                         // shared memory of one warp
    alignas(128) uint8_t text[T3_RING];
    alignas(16) uint32_t cw[T3_UNITS * CW_WORDS];
    uint32_t nlw[T3_UNITS];
    uint32_t queue[T3_QCAP];                              // line starts, bytes from the chunk's origin
    alignas(16) uint8_t stage[ROW_STAGE];
    alignas(8) uint64_t full[T3_NBLK];
};

struct Tok3Params {
    const uint8_t* text;
    uint64_t text_len, range_begin, range_end;
    uint64_t origin;                // absolute offset of chunk 0 (range_begin rounded down to 16)
    uint32_t n_chunks, chunk_bytes; // chunk_bytes: a multiple of T3_BLOCK
    uint8_t* rows;                  // region of chunk c at rows + c * region_cap
    uint32_t region_cap;            // multiple of 16
    double prior, error_threshold, alpha;
    int het_only;
    unsigned int* ticket;
    unsigned long long* site_alloc; // lines parsed by the call
    unsigned long long* blk;        // per chunk: rows << 32 | bytes
    unsigned long long* error;
    TableView table;
};

#if defined(__CUDACC__)

#ifndef SID_TOK3_CTAS
#define SID_TOK3_CTAS 9
#endif

// The byte-wise tokenizer on global memory, out of line: lines outside the fast grammar, lines the ring cannot hold.
__device__ __noinline__ void parse_line_global(const uint8_t* text, uint64_t text_len, uint64_t line_abs, ParsedLine& pl) {
    FlatSrc gsrc {text, text_len};
    parse_line(gsrc, line_abs, false, pl);
}

__global__ void __launch_bounds__(T3_THREADS, SID_TOK3_CTAS) k_tok3_rows(const Tok3Params p) {
    extern __shared__ __align__(128) uint8_t s_dyn3[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Tok3Warp& S = reinterpret_cast<Tok3Warp*>(s_dyn3)[warp];
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    if (lane == 0) {
        for (uint32_t b = 0; b < T3_NBLK; ++b) mbar_init(&S.full[b], 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t parity = 0;                    // bit s: phase parity the next wait on slot s expects
    uint32_t async_slots = 0;               // bit s: the block in slot s arrives by bulk copy
    const FlatSrc gsrc {p.text, p.text_len};
    uint8_t* const text = S.text;
    uint8_t* const stage = S.stage;

    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(p.ticket, 1u);
        chunk = __shfl_sync(FULL, chunk, 0);
        if (chunk >= p.n_chunks) break;
        const uint64_t cbeg = p.origin + (uint64_t)chunk * p.chunk_bytes;
        const uint64_t own_begin = cbeg > p.range_begin ? cbeg : p.range_begin;
        const uint64_t own_end = cbeg + p.chunk_bytes < p.range_end ? cbeg + p.chunk_bytes : p.range_end;
        uint8_t* const reg = p.rows + (size_t)chunk * p.region_cap;

        // ---- block loads: bulk copies for blocks that lie wholly inside the text, plain loads with '\n' padding else
        auto issue = [&](uint32_t k) {
            const uint32_t slot = k & (T3_NBLK - 1);
            const uint64_t babs = cbeg + (uint64_t)k * T3_BLOCK;
            if (babs + T3_BLOCK <= p.text_len) {
                async_slots |= 1u << slot;
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(&S.full[slot], T3_BLOCK);
                    bulk_load(text + slot * T3_BLOCK, p.text + babs, T3_BLOCK, &S.full[slot]);
                }
            } else {
                async_slots &= ~(1u << slot);
            }
        };
        // returns false on a lost copy
        auto land = [&](uint32_t k) -> bool {
            const uint32_t slot = k & (T3_NBLK - 1);
            if (async_slots & (1u << slot)) {
                const bool ok = mbar_wait<PARSE_SLEEP>(&S.full[slot], (parity >> slot) & 1u);
                parity ^= 1u << slot;
                return ok;
            }
            const uint64_t babs = cbeg + (uint64_t)k * T3_BLOCK;
            for (uint32_t i = lane; i < T3_BLOCK / 16; i += 32) {
                const uint64_t a = babs + 16ull * i;
                uint4 v;
                if (a + 16 <= p.text_len) {
                    v = __ldg(reinterpret_cast<const uint4*>(p.text + a));
                } else {
                    uint32_t w[4] = {0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au};
                    for (int b = 0; b < 16; ++b) {
                        const uint64_t q = a + b;
                        if (q < p.text_len) w[b >> 2] = (w[b >> 2] & ~(0xFFu << (8 * (b & 3)))) | ((uint32_t)p.text[q] << (8 * (b & 3)));
                    }
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
                reinterpret_cast<uint4*>(text + slot * T3_BLOCK)[i] = v;
            }
            __syncwarp();
            return true;
        };

        // ---- state of the chunk
        uint32_t q_head = 0, q_tail = 0;            // queue positions (free running)
        uint32_t last_start = 0, last_nl = 0;       // (offset from cbeg) + 1 of the last queued start / the last '\n' seen; 0: none
        uint32_t bad_lo = 0xFFFFFFFFu, bad_hi = 0;  // offsets [lo, hi) that hold a control byte: their lines go byte by byte
        uint32_t n_lines_chunk = 0;
        uint32_t written = 0, fill = 0, n_rows = 0;
        bool overflow = false, failed = false;
        uint32_t carry_nl = 1;
        if (cbeg > 0) carry_nl = p.text[cbeg - 1] == (uint8_t)'\n' ? 1u : 0u;

        // Stage 2 on the first n queued lines (n <= 32).  force_slow: the line is not in the ring as a whole.
        auto process = [&](uint32_t n, bool force_slow) {
            const bool mine = (uint32_t)lane < n;
            const uint32_t rel = S.queue[(q_head + (mine ? lane : 0)) & (T3_QCAP - 1)];
            // the line ends before the next queued start; the last one before the last '\n' seen
            uint32_t next_rel = __shfl_down_sync(FULL, rel, 1);
            if ((uint32_t)lane + 1 >= n) next_rel = (q_head + n != q_tail) ? S.queue[(q_head + n) & (T3_QCAP - 1)] : last_nl;
            const uint64_t line_abs = cbeg + rel;
            const uint32_t roff = rel & (T3_RING - 1);
            WinLine wl;
            wl.status = LINE_MALFORMED;
            wl.profile = 0; wl.name_len = 0; wl.hdr_len = 0; wl.pos_canonical = false;
            bool fast = false;
            const bool in_bad = rel < bad_hi && next_rel > bad_lo;
            if (!force_slow) {
                fast = parse_line_win<false, T3_UNITS>(text, 0, S.cw, S.nlw, 0, roff, wl);
                fast = fast && wl.pos_canonical && !in_bad;
            }
            uint64_t profile = wl.profile;
            int status = wl.status;
            uint32_t name_off = 0, name_len = wl.name_len, hdr_len = wl.hdr_len;
            int32_t pos = 0;
            if (!fast && mine) {
                ParsedLine pl;
                parse_line_global(p.text, p.text_len, line_abs, pl);
                status = pl.status; profile = pl.profile; pos = pl.pos; name_off = pl.chrom_off; name_len = pl.chrom_len;
                hdr_len = name_len + 1 + (uint32_t)digits_i32(pos);
            }
            __syncwarp();
            const bool good = mine && status == LINE_OK;
            if (mine && status != LINE_OK) report_error_at(p.error, line_abs, status);
            // ---- join; the lane that creates an entry classifies it
            uint32_t slot = 0;
            bool inserted = false;
            if (good) slot = table_join(p.table, profile, inserted);
            __syncwarp();
            if (inserted) classify_inserted(p.table, slot, profile, p.prior, p.error_threshold, p.alpha, p.het_only != 0);
            __syncwarp();
            RowSrc rs;
            rs.line_off = roff;
            rs.hdr_len = hdr_len;
            rs.name_len = name_len;
            rs.sfx_len = 0;
#pragma unroll
            for (int i = 0; i < 12; ++i) rs.sfx[i] = 0;
            if (good) {
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
            n_rows += __popc(__ballot_sync(FULL, row_len != 0));
            uint32_t incl = row_len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t excl = incl - row_len;
            // rows are staged in batches of consecutive lanes that fit the buffer (ordinarily one batch of 32)
            uint32_t a_lane = 0;
            while (a_lane < 32) {
                const uint32_t excl_a = __shfl_sync(FULL, excl, a_lane);
                const bool fits = fill + (incl - excl_a) <= (uint32_t)ROW_STAGE - 16u;
                const uint32_t nofit = __ballot_sync(FULL, !fits && (uint32_t)lane >= a_lane && row_len != 0);
                const uint32_t b_lane = nofit ? (uint32_t)(__ffs((int)nofit) - 1) : 32u;
                const bool in_batch = (uint32_t)lane >= a_lane && (uint32_t)lane < b_lane && row_len != 0;
                const uint32_t d = fill + (excl - excl_a);
                const uint32_t batch_bytes = __shfl_sync(FULL, incl, b_lane ? b_lane - 1 : 0) - excl_a;
                if (b_lane == a_lane) {
                    // the row of lane a_lane alone exceeds the buffer (a name of kilobytes): straight to the region, byte by byte
                    const uint32_t len_a = __shfl_sync(FULL, row_len, a_lane);
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
                uint32_t hw = 0, sxw = 0;
                if (in_batch && fast) {
                    hw = ((d & 3u) + hdr_len + 3u) >> 2;
                    sxw = (((d + hdr_len) & 3u) + rs.sfx_len + 3u) >> 2;
                }
                hw = __reduce_max_sync(FULL, hw);
                sxw = __reduce_max_sync(FULL, sxw);
                uint32_t first = 0;
                if (in_batch && fast) first = row_phase_a<T3_RING / 4 - 1>(text, stage, d, rs, hw, sxw);
                __syncwarp();
                if (in_batch) {
                    if (fast) row_phase_b(stage, d, rs, first);
                    else row_bytewise(stage + d, gsrc, line_abs + name_off, name_len, pos,
                                      reinterpret_cast<const uint8_t*>(p.table.suffix + (size_t)slot * SUFFIX_BYTES), rs.sfx_len);
                }
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
            q_head += n;
            n_lines_chunk += n;
        };

        // Runs stage 2 while the queue holds a full group, or lines that start below `floor` (their ring slots are
        // about to be reused), or -- with `drain` -- anything at all.
        auto service = [&](uint32_t floor, bool drain) {
            for (;;) {
                const uint32_t n_q = q_tail - q_head;
                if (n_q == 0) break;
                const bool open = !drain && last_start > last_nl;               // the last queued line has not ended yet
                const uint32_t complete = n_q - (open ? 1u : 0u);
                const uint32_t oldest = S.queue[q_head & (T3_QCAP - 1)];
                const bool force = drain || oldest < floor;
                if (!(complete >= 32u || force)) break;
                // complete == 0 here: one line longer than the ring holds, read from global memory byte by byte
                process(complete == 0 ? 1u : (complete < 32u ? complete : 32u), complete == 0);
            }
        };

        uint32_t issued = 0;
        for (; issued < T3_AHEAD; ++issued) issue(issued);
        for (uint32_t k = 0;; ++k) {
            if (k >= issued) { issue(k); ++issued; }
            if (!land(k)) {
                if (lane == 0) report_error_at(p.error, 0, LINE_MALFORMED + 4);
                failed = true;
                break;
            }
            const uint32_t slot = k & (T3_NBLK - 1);
            const uint64_t babs = cbeg + (uint64_t)k * T3_BLOCK;
            const uint32_t brel = k * T3_BLOCK;
            // ---- stage 1: one unit per lane
            uint32_t st;
            {
                const uint32_t ru = slot * 32 + lane;
                const uint8_t* up = text + ru * 32;
                const uint4 v0 = *reinterpret_cast<const uint4*>(up), v1 = *reinterpret_cast<const uint4*>(up + 16);
                const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                const UnitClasses c = classify_unit(w);
                uint4* rec = reinterpret_cast<uint4*>(S.cw + (size_t)ru * CW_WORDS);
                rec[0] = make_uint4(c.w[0], c.w[1], c.w[2], c.w[3]);
                rec[1] = make_uint4(c.w[4], c.w[5], c.w[6], c.w[7]);
                S.nlw[ru] = c.nl;
                const uint32_t nl = c.nl;
                uint32_t prev = __shfl_up_sync(FULL, nl >> 31, 1);
                if (lane == 0) prev = carry_nl;
                carry_nl = __shfl_sync(FULL, nl >> 31, 31);
                st = ((nl << 1) | prev) & ~nl;
                if (!(babs >= own_begin && babs + T3_BLOCK <= own_end)) {        // first / last block of the chunk, or past its end
                    const uint64_t first = babs + (uint64_t)lane * 32;
                    if (first + 32 <= own_begin || first >= own_end) st = 0;
                    else {
                        if (first < own_begin) st &= 0xFFFFFFFFu << (uint32_t)(own_begin - first);
                        if (first + 32 > own_end) st &= 0xFFFFFFFFu >> (32u - (uint32_t)(own_end - first));
                    }
                }
                const uint32_t nlpos = nl ? brel + (uint32_t)lane * 32 + (31u - (uint32_t)__clz((int)nl)) + 1u : 0u;
                const uint32_t m = __reduce_max_sync(FULL, nlpos);
                if (m) last_nl = m;
                if (__any_sync(FULL, c.bad != 0)) {
                    // control bytes: every line that touches this block goes through the byte-wise tokenizer
                    const uint32_t lo = q_head != q_tail ? S.queue[q_head & (T3_QCAP - 1)] : brel;
                    bad_lo = bad_lo < lo ? bad_lo : lo;
                    bad_hi = brel + T3_BLOCK;
                }
            }
            // ---- line starts into the queue; stage 2 whenever a group is ready
            const uint32_t floor = k >= 1 ? (k - 1) * T3_BLOCK : 0u;
            bool last_block = false;
            for (;;) {
                bool more = false;
                if (__any_sync(FULL, st != 0)) {
                    const uint32_t cnt = (uint32_t)__popc(st);
                    uint32_t incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t o = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += o;
                    }
                    const uint32_t total = __shfl_sync(FULL, incl, 31);
                    const uint32_t room = T3_QCAP - (q_tail - q_head);
                    uint32_t idx = incl - cnt, hi = 0;
                    while (st && idx < room) {
                        const uint32_t bit = (uint32_t)__ffs((int)st) - 1u;
                        st &= st - 1;
                        const uint32_t rel = brel + (uint32_t)lane * 32 + bit;
                        S.queue[(q_tail + idx) & (T3_QCAP - 1)] = rel;
                        hi = rel + 1;
                        ++idx;
                    }
                    q_tail += total < room ? total : room;
                    hi = __reduce_max_sync(FULL, hi);
                    if (hi) last_start = hi;
                    more = total > room;                    // the queue is full: run every complete line, then push the rest
                }
                __syncwarp();
                if (!more) {
                    // is this the last block the chunk needs?  (all of its lines have started and the last one has ended)
                    const bool open = q_tail != q_head && last_start > last_nl;
                    last_block = babs + T3_BLOCK >= own_end && !open;
                }
                service(more ? 0xFFFFFFFFu : floor, last_block);
                if (!more) break;
            }
            if (last_block) {
                // bulk copies still in flight for blocks this chunk will not read: let them land (keeps the phases in step)
                for (uint32_t j = k + 1; j < issued; ++j)
                    if (async_slots & (1u << (j & (T3_NBLK - 1)))) land(j);
                break;
            }
            // the slot of block k - 2 is free now: the copy for block k + 2 goes there
            if (issued < k + 1 + T3_AHEAD) {
                __syncwarp();
                issue(issued);
                ++issued;
            }
        }
        // ---- the last partial 16 bytes of the region, the chunk's entry in the block table
        if (fill && !overflow && !failed) {
            if (lane == 0) *reinterpret_cast<uint4*>(reg + written) = *reinterpret_cast<const uint4*>(stage);
        }
        if (lane == 0) {
            if (overflow) report_error_at(p.error, cbeg, LINE_ROWS_OVERFLOW);
            p.blk[chunk] = ((unsigned long long)n_rows << 32) | (unsigned long long)(written + fill);
            if (n_lines_chunk) atomicAdd(p.site_alloc, (unsigned long long)n_lines_chunk);
        }
        __syncwarp();
        if (failed) break;
    }
}

// ---- the chunks' regions laid end to end in file order -----------------------------------------------------------
// off[r] = bytes of the regions before r; totals out.  One CTA.
__global__ void __launch_bounds__(1024) k_rows3_scan(const unsigned long long* blk, uint32_t n_regions, unsigned long long* off,
                                                     unsigned long long* bytes_out, unsigned long long* rows_out) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    unsigned long long rows = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n_regions; c0 += 1024) {
        const uint32_t c = c0 + threadIdx.x;
        const unsigned long long e = c < n_regions ? blk[c] : 0ull;
        const unsigned long long v = e & 0xFFFFFFFFull;
        rows += e >> 32;
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
        if (c < n_regions) off[c] = before + incl - v;
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

constexpr int RC3_THREADS = 256;
constexpr uint32_t RC3_PIECE = 4096;                  // bytes of a region one warp moves at a time
constexpr uint32_t RC3_STAGE = RC3_PIECE + 64;        // shared memory per warp

// One warp per (region, piece): source 16-byte aligned, destination at any alignment; the piece is staged in shared
// memory so that both sides move in 16-byte accesses.
__global__ void __launch_bounds__(RC3_THREADS) k_rows3_copy(const uint8_t* rows, uint32_t region_cap, const unsigned long long* blk,
                                                            const unsigned long long* off, uint32_t n_regions, uint32_t pieces_per_region,
                                                            uint8_t* out, uint64_t out_cap) {
    extern __shared__ __align__(16) uint8_t s_rc3[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* st = s_rc3 + (size_t)warp * RC3_STAGE;
    const uint64_t n_items = (uint64_t)n_regions * pieces_per_region;
    const uint64_t stride = (uint64_t)gridDim.x * (RC3_THREADS / 32);
    for (uint64_t item = (uint64_t)blockIdx.x * (RC3_THREADS / 32) + warp; item < n_items; item += stride) {
        const uint32_t r = (uint32_t)(item / pieces_per_region), q = (uint32_t)(item % pieces_per_region);
        const uint32_t bytes = (uint32_t)blk[r] < region_cap ? (uint32_t)blk[r] : region_cap;
        const uint32_t p0 = q * RC3_PIECE;
        if (p0 >= bytes) continue;
        const uint32_t n = bytes - p0 < RC3_PIECE ? bytes - p0 : RC3_PIECE;
        const unsigned long long dst0 = off[r] + p0;
        if (dst0 + n > out_cap) continue;                              // the host reports SIDGPU_ECAPACITY from the total
        const uint8_t* src = rows + (size_t)r * region_cap + p0;
        uint8_t* dst = out + dst0;
        const uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
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
            const uint32_t n16 = (n + 15) >> 4;
            const uint4* sv = reinterpret_cast<const uint4*>(src);
            uint4* tv = reinterpret_cast<uint4*>(st + 16);
            for (uint32_t i = lane; i < n16; i += 32) tv[i] = __ldcs(sv + i);
        }
        __syncwarp();
        // bytes of the piece start at st + 16; destination-aligned chunk c (c >= 1) = piece bytes [16c - mis, +16)
        const uint32_t head = 16u - mis;
        if ((uint32_t)lane < (head < n ? head : n)) dst[lane] = st[16 + lane];
        if (n > head) {
            const uint32_t rest = n - head;
            const uint32_t body = rest >> 4;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(st + 16 + (head & ~3u));
            const uint32_t sh = (head & 3u) * 8u;
            uint4* gv = reinterpret_cast<uint4*>(dst + head);
            for (uint32_t i = lane; i < body; i += 32) {
                const uint32_t* qq = w + 4 * i;
                const uint32_t a0 = qq[0], a1 = qq[1], a2 = qq[2], a3 = qq[3], a4 = qq[4];
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
