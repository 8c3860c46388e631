// libsidgpu.so: the C ABI of include/sidgpu.h over the sm_100a kernels.
// Host-side orchestration only: buffers, streams, sessions, the Nelder-Mead driver.  All pileup
// parsing, profile building, likelihoods, p-values, histogramming and CSV formatting run in the
// kernels of k_*.cuh; there is no CPU implementation of any of them in this library.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/sidgpu.h"
#include "k_calls.cuh"
#include "k_lynch.cuh"
#include "k_quality.cuh"
#include "k_tokenize.cuh"
#include "k_tok2.cuh"
#include "k_reads.cuh"
#include "inflate.cuh"
#include "nelder_mead.hpp"

using namespace sid;

namespace {

std::string g_create_error;

// Device-resident counters every kernel of a ctx shares; mirrored in pinned host memory.
struct Control {
    unsigned int tok_ticket;
    unsigned int csv_ticket;
    unsigned long long n_sites;      // storage indices the running tokenizer call has handed out (TokParams::site_alloc)
    unsigned long long error;
    unsigned long long csv_bytes;
    unsigned long long csv_rows;
    unsigned int n_entries;
    unsigned int special_used;
    unsigned int table_overflow;
    unsigned int name_cursor;
    unsigned int name_overflow;
    unsigned int n_selected;
    unsigned int obj_done;
    unsigned int pad0;
    unsigned long long nd_acc[5];
    double objective;
    // merged (all shards) histogram table: its own counters
    unsigned int g_n_entries;
    unsigned int g_special_used;
    unsigned int g_overflow;
    unsigned int fit_barrier;        // arrival counter of k_lynch_fit's grid barrier
    double fit_out[6];               // pi, eps, fval, iterations, evaluations, converged
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

enum Phase { PHASE_IDLE = 0, PHASE_FEED = 1, PHASE_FINISHED = 2 };

}  // namespace

struct sidgpu_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    std::string err;
    uint64_t launches = 0;
    size_t max_chunk = 0;
    double avg_line_bytes = 80.0;    // running estimate, sizes the tokenizer's slices

    Control* d_ctl = nullptr;
    Control* h_ctl = nullptr;          // mapped pinned memory: the device writes it directly (k_publish_ctl)
    Control* h_ctl_dev = nullptr;      // its device-side address

    // unique-profile table
    int table_log2 = 0;
    TableView tab {};
    // chromosome names
    NameDict names {};
    // look-back status words of the CSV writer
    DevBuf csv_status;
    // site store: arrays indexed by storage index (dense, not in file order) and order[file index] = storage index
    DevBuf pos, slot, name_ref, profile, line_off, site_suffix, order;
    DevBuf fwd, v_fwd;               // forward-strand profile per site (k_tok2<..., STRANDS>) and its file-ordered copy
    bool want_fwd = false, fwd_valid = false;
    DevBuf qual_l;                   // quality sessions: two doubles per site from the tokenizer (k_tok2<..., QUAL>)
    DevBuf blk, blk_part;            // block table of the running tokenizer call and its per-chunk sums
    DevBuf rows_scratch, rows_part, rows_part_rows;   // fused row writer: per-slice regions, per-chunk byte / row sums
    double region_factor = 1.0;      // grows when a slice's rows did not fit its region
    DevBuf v_pos, v_slot, v_name_ref, v_profile, v_line_off;   // file-ordered copies behind sidgpu_sites_view
    uint64_t site_cap = 0;
    bool want_profile = false, want_line_off = false, want_site_suffix = false;
    bool qual_sums_valid = false;    // the last tokenizer call left the log-likelihood sums of `-m quality` in qual_l
    uint64_t n_sites_total = 0;      // sites in the store (accumulate mode) or in the last chunk (streaming)
    uint64_t chunk_begin = 0, chunk_sites = 0;

    // session
    sidgpu_params params {};
    int phase = PHASE_IDLE;
    bool streaming = false;          // per-chunk emission possible
    bool counting = false;
    uint32_t classified = 0;         // entries [0, classified) carry a class for this session
    const char* last_text = nullptr; // text of the most recent feed (quality needs it at emit time)
    uint64_t last_text_len = 0;
    double session_prior = -1;
    bool fit_done = false;
    sidgpu_fit fit {};
    double fit_nd[4] = {0.25, 0.25, 0.25, 0.25};

    // histogram
    DevBuf sort_keys, sort_vals, u_profile, u_count, u_logM, entry_to_unique, p_hom, p_het, adj_hom, adj_het, bh_c, bh_block;
    uint64_t n_unique = 0;
    bool hist_valid = false;
    bool global_hist = false;        // the histogram is the merged one of all shards (sidgpu_set_global_histogram)
    DevBuf g_e2u;
    TableView g_tab {};               // counting table of the merged histogram, kept from call to call
    int g_log2 = 0;
    DevBuf partials;
    DevBuf quality_lut;

    // sidgpu_call_host staging (two text buffers, two CSV buffers, their events), kept across calls
    DevBuf hp_text[2], hp_csv[2];
    DevBuf hp_comp[2], inf_blocks[2];  // BGZF input: compressed chunks and their member tables (bgzf_path.inl)
    DevBuf crc_tables;                 // k_crc32_members
    cudaStream_t inflate_stream = nullptr;              // the inflate of chunk i + 1 runs beside the calling kernels of chunk i
    cudaEvent_t ev_inflate[2] = {nullptr, nullptr};
    unsigned long long* d_inf = nullptr;                // per text buffer: inflate error word, end of the last whole line
    unsigned long long* h_inf = nullptr;                // mapped pinned mirror, written by k_publish_words
    unsigned long long* h_inf_dev = nullptr;
    cudaEvent_t hp_ev_in[2] = {nullptr, nullptr}, hp_ev_out[2] = {nullptr, nullptr};

    // optional per-kernel timing (sidgpu_profile): event pairs recorded around launches, resolved lazily
    bool profiling = false;
    struct Pending { cudaEvent_t a, b; int which; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> free_events;
    double kernel_ms[SIDGPU_N_TIMERS] = {0};
    uint64_t kernel_launches[SIDGPU_N_TIMERS] = {0};

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return ctx->fail(SIDGPU_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define TRY(call)                  \
    do {                           \
        int rc_ = (call);          \
        if (rc_ != SIDGPU_OK) return rc_; \
    } while (0)

int ensure(sidgpu_ctx* ctx, DevBuf& b, size_t bytes, bool keep = false) {
    if (bytes <= b.cap) return SIDGPU_OK;
    size_t want = std::max(bytes, b.cap + b.cap / 2);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&np, want);
    }
    if (e != cudaSuccess) return ctx->fail(SIDGPU_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    if (keep && b.p && b.cap) {
        e = cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(np); return ctx->fail(SIDGPU_ECUDA, "grow copy failed: %s", cudaGetErrorString(e)); }
    }
    if (b.p) cudaFree(b.p);
    b.p = np;
    b.cap = want;
    return SIDGPU_OK;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

template <class T>
T* ctl_field(sidgpu_ctx* ctx, T Control::*m) { return &(ctx->d_ctl->*m); }

// The control block travels to the host by stores from a one-warp kernel into mapped pinned memory, not
// by a cudaMemcpy: a copy would queue in the device-to-host copy engine behind the CSV chunks of the
// host path (milliseconds each), and the host loop that keeps the uploads going waits on this.
__global__ void k_publish_ctl(const Control* d, Control* h) {
    static_assert(sizeof(Control) % 8 == 0, "Control is copied in 8-byte words");
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(d);
    volatile unsigned long long* t = reinterpret_cast<volatile unsigned long long*>(h);
    for (unsigned i = threadIdx.x; i < sizeof(Control) / 8; i += blockDim.x) t[i] = s[i];
    __threadfence_system();
}

int sync_ctl(sidgpu_ctx* ctx) {
    k_publish_ctl<<<1, 32, 0, ctx->stream>>>(ctx->d_ctl, ctx->h_ctl_dev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

int check_launch(sidgpu_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ctx->fail(SIDGPU_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    ctx->launches++;
    return SIDGPU_OK;
}

// NVTX range around an ABI call: shows up as a named span in Nsight Systems timelines
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

enum { PROF_TOKENIZE = 0, PROF_CLASSIFY = 1, PROF_CSV = 2, PROF_ORDER = 3, PROF_FIT = 4, PROF_HIST = 5, PROF_QUALITY = 6, PROF_INFLATE = 7 };

cudaEvent_t take_event(sidgpu_ctx* ctx) {
    if (!ctx->free_events.empty()) { cudaEvent_t e = ctx->free_events.back(); ctx->free_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ProfScope {
    sidgpu_ctx* ctx;
    cudaEvent_t a = nullptr;
    int which;
    cudaStream_t stream;
    ProfScope(sidgpu_ctx* c, int w, cudaStream_t s = nullptr) : ctx(c), which(w), stream(s ? s : c->stream) {
        if (ctx->profiling) { a = take_event(ctx); cudaEventRecord(a, stream); }
    }
    ~ProfScope() {
        if (a) { cudaEvent_t b = take_event(ctx); cudaEventRecord(b, stream); ctx->pending.push_back({a, b, which}); }
    }
};

void resolve_profile(sidgpu_ctx* ctx) {
    for (auto& p : ctx->pending) {
        float ms = 0;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            ctx->kernel_ms[p.which] += ms;
            ctx->kernel_launches[p.which]++;
        }
        ctx->free_events.push_back(p.a);
        ctx->free_events.push_back(p.b);
    }
    ctx->pending.clear();
}

// ---- table -------------------------------------------------------------------------------------

void free_table(TableView& t) {
    cudaFree(t.keys); cudaFree(t.counts); cudaFree(t.entry_list); cudaFree(t.suffix);
    cudaFree(t.label); cudaFree(t.gt); cudaFree(t.hom); cudaFree(t.het);
    t = TableView {};
}

// aux: a counting-only table (keys, counts, entry list) with its own counters: the merged histogram of all shards
int alloc_table(sidgpu_ctx* ctx, int log2cap, TableView& t, bool aux = false) {
    t = TableView {};
    const size_t cap = (size_t)1 << log2cap, n = cap + 1;
    t.cap = (uint32_t)cap;
    t.mask = (uint32_t)(cap - 1);
    t.n_entries = ctl_field(ctx, aux ? &Control::g_n_entries : &Control::n_entries);
    t.special_used = ctl_field(ctx, aux ? &Control::g_special_used : &Control::special_used);
    t.overflow = ctl_field(ctx, aux ? &Control::g_overflow : &Control::table_overflow);
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void**)&t.keys, n * 8);
    A((void**)&t.counts, n * 8);
    A((void**)&t.entry_list, n * 4);
    if (!aux) {
        A((void**)&t.suffix, n * SUFFIX_BYTES);
        A((void**)&t.label, n);
        A((void**)&t.gt, n * 2);
        A((void**)&t.hom, n * 8);
        A((void**)&t.het, n * 8);
    }
    if (e != cudaSuccess) { free_table(t); return ctx->fail(SIDGPU_ENOMEM, "profile table of 2^%d slots: %s", log2cap, cudaGetErrorString(e)); }
    CK(cudaMemsetAsync(t.keys, 0xFF, n * 8, ctx->stream));
    CK(cudaMemsetAsync(t.counts, 0, n * 8, ctx->stream));
    if (!aux) {
        CK(cudaMemsetAsync(t.suffix, 0, n * SUFFIX_BYTES, ctx->stream));
        CK(cudaMemsetAsync(t.label, 255, n, ctx->stream));
    }
    return SIDGPU_OK;
}

__global__ void k_rehash(TableView from, TableView to, uint32_t n_entries, uint32_t* remap) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const uint32_t s = from.entry_list[e];
    const uint64_t key = from.keys[s];
    uint32_t h;
    if (key == TABLE_EMPTY) {
        h = to.cap;
        to.keys[h] = key;
    } else {
        h = table_hash(key) & to.mask;
        for (;;) {
            const unsigned long long old = atomicCAS(&to.keys[h], (unsigned long long)TABLE_EMPTY, (unsigned long long)key);
            if (old == TABLE_EMPTY) break;
            h = (h + 1) & to.mask;
        }
    }
    to.entry_list[e] = h;
    to.counts[h] = from.counts[s];
    to.label[h] = from.label[s];
    to.gt[2 * h] = from.gt[2 * s];
    to.gt[2 * h + 1] = from.gt[2 * s + 1];
    to.hom[h] = from.hom[s];
    to.het[h] = from.het[s];
    for (int i = 0; i < SUFFIX_BYTES; ++i) to.suffix[(size_t)h * SUFFIX_BYTES + i] = from.suffix[(size_t)s * SUFFIX_BYTES + i];
    remap[s] = h;
}

__global__ void k_remap_slots(uint32_t* slot, uint64_t n, const uint32_t* remap) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot[i] = remap[slot[i]];
}

// Grows the table by `shift` powers of two, keeping entries (and their order), counts and classes.
int grow_table(sidgpu_ctx* ctx, int shift, uint64_t stored_sites) {
    CK(cudaStreamSynchronize(ctx->stream));
    TRY(sync_ctl(ctx));
    const uint32_t n_entries = ctx->h_ctl->n_entries;
    TableView nt;
    TRY(alloc_table(ctx, ctx->table_log2 + shift, nt));
    uint32_t* remap = nullptr;
    CK(cudaMalloc((void**)&remap, ((size_t)ctx->tab.cap + 1) * 4));
    if (n_entries) {
        k_rehash<<<(n_entries + 255) / 256, 256, 0, ctx->stream>>>(ctx->tab, nt, n_entries, remap);
        TRY(check_launch(ctx, "k_rehash"));
        if (stored_sites) {
            k_remap_slots<<<(unsigned)((stored_sites + 255) / 256), 256, 0, ctx->stream>>>((uint32_t*)ctx->slot.p, stored_sites, remap);
            TRY(check_launch(ctx, "k_remap_slots"));
        }
    }
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::table_overflow), 0, sizeof(unsigned int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(remap);
    free_table(ctx->tab);
    ctx->tab = nt;
    ctx->table_log2 += shift;
    return SIDGPU_OK;
}

// Forgets the interned chromosome names (and a reported overflow): every session starts with an empty dictionary.
int reset_names(sidgpu_ctx* ctx) {
    CK(cudaMemsetAsync(ctx->names.slots, 0, ((size_t)ctx->names.mask + 1) * 8, ctx->stream));
    const unsigned int four = 4;             // pool offsets start at 4: 0 means "no name"
    CK(cudaMemcpyAsync(ctx->names.cursor, &four, sizeof four, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->names.overflow, 0, sizeof(unsigned int), ctx->stream));
    return SIDGPU_OK;
}

int reset_table(sidgpu_ctx* ctx) {
    const size_t n = (size_t)ctx->tab.cap + 1;
    CK(cudaMemsetAsync(ctx->tab.keys, 0xFF, n * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->tab.counts, 0, n * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->tab.suffix, 0, n * SUFFIX_BYTES, ctx->stream));
    CK(cudaMemsetAsync(ctx->tab.label, 255, n, ctx->stream));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::n_entries), 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::special_used), 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::table_overflow), 0, sizeof(unsigned int), ctx->stream));
    ctx->classified = 0;
    ctx->hist_valid = false;
    return SIDGPU_OK;
}

// ---- site store --------------------------------------------------------------------------------

int ensure_sites(sidgpu_ctx* ctx, uint64_t n, bool keep) {
    if (n > ctx->site_cap || (ctx->want_profile && ctx->profile.cap < n * 8) || (ctx->want_line_off && ctx->line_off.cap < n * 8) || (ctx->want_fwd && ctx->fwd.cap < n * 8) ||
        (ctx->want_site_suffix && (ctx->site_suffix.cap < n * SUFFIX_BYTES || ctx->qual_l.cap < n * 16))) {
        const uint64_t cap = std::max<uint64_t>(n, ctx->site_cap);
        if (cap >= 0xFFFFFFFFull) return ctx->fail(SIDGPU_ECAPACITY, "more than 2^32 sites in one store: feed smaller ranges");
        TRY(ensure(ctx, ctx->order, cap * 4, keep));
        TRY(ensure(ctx, ctx->pos, cap * 4, keep));
        TRY(ensure(ctx, ctx->slot, cap * 4, keep));
        TRY(ensure(ctx, ctx->name_ref, cap * 4, keep));
        if (ctx->want_profile) TRY(ensure(ctx, ctx->profile, cap * 8, keep));
        if (ctx->want_line_off) TRY(ensure(ctx, ctx->line_off, cap * 8, keep));
        if (ctx->want_fwd) TRY(ensure(ctx, ctx->fwd, cap * 8, keep));
        if (ctx->want_site_suffix) TRY(ensure(ctx, ctx->site_suffix, cap * SUFFIX_BYTES, keep));
        if (ctx->want_site_suffix) TRY(ensure(ctx, ctx->qual_l, cap * 16, keep));
        ctx->site_cap = cap;
    }
    return SIDGPU_OK;
}

const char* status_text(int st) {
    switch (st) {
        case LINE_MALFORMED: return "Malformed pileup line";
        case LINE_MISSING_MAPQ: return "Malformed pileup line or missing mapping qualities";
        case LINE_QUAL_SHORT: return "fewer quality characters than counted bases";
        default: return "internal tokenizer error";
    }
}

// Runs K1 over [range_begin, range_end): the sites go to storage indices [site_base, site_base + n) in
// no particular order, and order[site_base + f] is the storage index of the f-th line of the range.
// On return *n_out holds the number of sites the range produced.
int run_tokenizer(sidgpu_ctx* ctx, const char* d_text, size_t text_len, size_t range_begin, size_t range_end,
                  bool want_qual, bool use_table, uint64_t site_base, bool keep_sites, uint64_t* n_out, bool strict_qual = false) {
    if (((uintptr_t)d_text & 15) != 0) return ctx->fail(SIDGPU_EINVAL, "d_text must be 16-byte aligned");
    if (range_begin > range_end || range_end > text_len) return ctx->fail(SIDGPU_EINVAL, "bad range [%zu,%zu) for text of %zu bytes", range_begin, range_end, text_len);
    *n_out = 0;
    if (range_begin == range_end) return sync_ctl(ctx);      // callers read the host mirror of the control block afterwards
    const uint64_t tile0 = range_begin & ~(uint64_t)15;
    const uint64_t span = range_end - tile0;
    // a slice (one parse warp's share of a tile) should hold about 30 lines: 32 lanes, little overflow
    static const double slice_lines = getenv("SIDGPU_SLICE_LINES") ? atof(getenv("SIDGPU_SLICE_LINES")) : 31.0;   // tuning knob
    uint32_t slice = (uint32_t)(ctx->avg_line_bytes * slice_lines) & ~31u;
    slice = std::max<uint32_t>(SLICE_MIN, std::min<uint32_t>(SLICE_MAX, slice));
    // a parse warp also classifies the bytes after its slice that its last lines reach into
    uint32_t ext = ((uint32_t)(ctx->avg_line_bytes * 1.25) + 31u) & ~31u;
    ext = std::max<uint32_t>(128u, std::min<uint32_t>(2016u, ext));
    const uint64_t tile_bytes = (uint64_t)slice * TOK_PARSE_WARPS;
    const uint64_t n_tiles64 = (span + tile_bytes - 1) / tile_bytes;
    if (n_tiles64 > 0x7FFFFFFFull) return ctx->fail(SIDGPU_EINVAL, "range too large");
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    const uint32_t n_blocks = n_tiles * TOK_PARSE_WARPS;
    if (n_tiles64 * TOK_PARSE_WARPS > 0x7FFFFFFFull) return ctx->fail(SIDGPU_EINVAL, "range too large");
    const uint32_t n_chunks = (n_blocks + BLK_CHUNK - 1) / BLK_CHUNK;
    TRY(ensure(ctx, ctx->blk, (size_t)n_blocks * 8));
    TRY(ensure(ctx, ctx->blk_part, (size_t)n_chunks * 8));

    uint64_t guess = (range_end - range_begin) / 24 + 4096;
    for (int attempt = 0;; ++attempt) {
        TRY(ensure_sites(ctx, site_base + guess, keep_sites || attempt > 0 ? keep_sites : false));
        CK(cudaMemsetAsync(ctx->blk.p, 0, (size_t)n_blocks * 8, ctx->stream));
        CK(cudaMemsetAsync(ctl_field(ctx, &Control::tok_ticket), 0, sizeof(unsigned int), ctx->stream));
        CK(cudaMemsetAsync(ctl_field(ctx, &Control::n_sites), 0, sizeof(unsigned long long), ctx->stream));
        CK(cudaMemsetAsync(ctl_field(ctx, &Control::error), 0xFF, sizeof(unsigned long long), ctx->stream));
        TokParams p {};
        p.text = (const uint8_t*)d_text;
        p.text_len = text_len;
        p.range_begin = range_begin;
        p.range_end = range_end;
        p.tile0 = tile0;
        p.n_tiles = n_tiles;
        p.site_base = site_base;
        p.site_cap = ctx->site_cap;
        p.profile = ctx->want_profile ? (uint64_t*)ctx->profile.p : nullptr;
        p.pos = (int32_t*)ctx->pos.p;
        p.slot = (uint32_t*)ctx->slot.p;
        p.name_ref = (uint32_t*)ctx->name_ref.p;
        p.line_off = ctx->want_line_off ? (uint64_t*)ctx->line_off.p : nullptr;
        p.tile_ticket = ctl_field(ctx, &Control::tok_ticket);
        p.site_alloc = ctl_field(ctx, &Control::n_sites);
        p.blk = (unsigned long long*)ctx->blk.p;
        p.error = ctl_field(ctx, &Control::error);
        p.table = ctx->tab;
        p.names = ctx->names;
        p.use_table = use_table ? 1 : 0;
        p.count_profiles = 0;
        p.want_qual = want_qual ? 1 : 0;
        p.slice_bytes = slice;
        p.text_stride = tok_text_stride(slice, ext);
        p.tail_bytes = tok_tail_bytes(ext);
        p.lines_cap = slice / 8;
        p.ext_bytes = ext;
        p.words_cap = tok_words_cap(slice, ext);
        // two staged tiles per CTA (the copy of the next tile runs under the parsing of this one) unless the
        // slices are so long that a single stage doubles the CTAs an SM holds: then the CTAs cover each other's copies
        static const int force_stages = getenv("SIDGPU_TOK_STAGES") ? atoi(getenv("SIDGPU_TOK_STAGES")) : 0;     // tuning knob
        static const int k1_version = getenv("SIDGPU_K1") ? atoi(getenv("SIDGPU_K1")) : 2;                       // 1: round-1 kernel (A/B runs)
        if (k1_version != 1) {
            Tok2Params q {};
            q.text = p.text; q.text_len = p.text_len; q.range_begin = p.range_begin; q.range_end = p.range_end;
            q.tile0 = p.tile0; q.n_tiles = p.n_tiles;
            q.site_base = p.site_base; q.site_cap = p.site_cap; q.profile = p.profile; q.pos = p.pos; q.slot = p.slot;
            q.name_ref = p.name_ref; q.line_off = p.line_off; q.site_alloc = p.site_alloc;
            q.tile_ticket = p.tile_ticket; q.blk = p.blk; q.error = p.error; q.table = p.table; q.names = p.names;
            q.use_table = p.use_table; q.want_qual = p.want_qual; q.bytewise = (want_qual && strict_qual) ? 1 : 0;
            q.slice_bytes = slice; q.text_stride = p.text_stride; q.tail_bytes = p.tail_bytes; q.lines_cap = p.lines_cap;
            q.ext_bytes = ext; q.units_cap = tok2_units(slice, ext) + CW_PAD_UNITS;
            static const int deep_knob = getenv("SIDGPU_DEEP_LINES") ? atoi(getenv("SIDGPU_DEEP_LINES")) : -1;     // A/B knob: 0 off, 1 on
            const bool deep = deep_knob >= 0 ? deep_knob != 0 : ctx->avg_line_bytes >= 256.0;
            // quality sessions: the per-read sums are formed by the tokenizer (A/B knob: SIDGPU_QUALITY_K1=0 leaves them to k_quality)
            static const bool qual_knob = !(getenv("SIDGPU_QUALITY_K1") && atoi(getenv("SIDGPU_QUALITY_K1")) == 0);
            const bool qual = qual_knob && want_qual && !strict_qual && !deep && ctx->want_site_suffix && ctx->qual_l.p;
            q.qual_l = qual ? (double*)ctx->qual_l.p : nullptr;
            q.qual_lut = (const double*)ctx->quality_lut.p;
            ctx->qual_sums_valid = qual;
            // strand counts as a by-product of the tokenizer: the ordinary sites form only (else: k_strand_counts over the line offsets)
            const bool strands = ctx->want_fwd && !qual && !deep && !q.bytewise;
            q.fwd = strands ? (uint64_t*)ctx->fwd.p : nullptr;
            ctx->fwd_valid = strands;
            int s2 = 0, s1 = 0;
            if (strands) {
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s2, k_tok2<false, 2, false, false, true>, TOK_THREADS, tok2_dyn_smem(slice, ext, 2, false, true)) != cudaSuccess || s2 < 1) s2 = 1;
                s1 = 0;
            } else if (qual) {
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s2, k_tok2<false, 2, false, true>, TOK_THREADS, tok2_dyn_smem(slice, ext, 2, false)) != cudaSuccess || s2 < 1) s2 = 1;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s1, k_tok2<false, 1, false, true>, TOK_THREADS, tok2_dyn_smem(slice, ext, 1, false)) != cudaSuccess || s1 < 1) s1 = 1;
            } else if (deep) {
                // long lines: the instantiation whose stage 2 gives every 64-byte window of a bases field its own lane
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s2, k_tok2<false, 2, true>, TOK_THREADS, tok2_dyn_smem(slice, ext, 2, false)) != cudaSuccess || s2 < 1) s2 = 1;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s1, k_tok2<false, 1, true>, TOK_THREADS, tok2_dyn_smem(slice, ext, 1, false)) != cudaSuccess || s1 < 1) s1 = 1;
            } else {
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s2, k_tok2<false, 2>, TOK_THREADS, tok2_dyn_smem(slice, ext, 2, false)) != cudaSuccess || s2 < 1) s2 = 1;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s1, k_tok2<false, 1>, TOK_THREADS, tok2_dyn_smem(slice, ext, 1, false)) != cudaSuccess || s1 < 1) s1 = 1;
            }
            const int stages = force_stages == 1 || force_stages == 2 ? force_stages : (s1 > s2 ? 1 : 2);
            const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * (stages == 1 ? s1 : s2));
            const uint32_t dyn = tok2_dyn_smem(slice, ext, (uint32_t)stages, false, strands);
            {
                ProfScope prof(ctx, PROF_TOKENIZE);
                if (strands) {
                    k_tok2<false, 2, false, false, true><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
                } else if (qual) {
                    if (stages == 1) k_tok2<false, 1, false, true><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
                    else k_tok2<false, 2, false, true><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
                } else if (deep) {
                    if (stages == 1) k_tok2<false, 1, true><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
                    else k_tok2<false, 2, true><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
                } else {
                    if (stages == 1) k_tok2<false, 1><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
                    else k_tok2<false, 2><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
                }
                TRY(check_launch(ctx, "k_tok2"));
            }
        } else {
        int per_sm2 = 0, per_sm1 = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_tokenize<true, 2>, TOK_THREADS, tok_dyn_smem(slice, ext, 2)) != cudaSuccess || per_sm2 < 1) per_sm2 = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, k_tokenize<true, 1>, TOK_THREADS, tok_dyn_smem(slice, ext, 1)) != cudaSuccess || per_sm1 < 1) per_sm1 = 1;
        const int stages = force_stages == 1 || force_stages == 2 ? force_stages : (per_sm1 >= 2 * per_sm2 ? 1 : 2);
        const uint32_t dyn_smem = tok_dyn_smem(slice, ext, (uint32_t)stages);
        const int per_sm = stages == 1 ? per_sm1 : per_sm2;
        const int max_ctas = ctx->sm_count * per_sm;
        const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)max_ctas);
        {
            ProfScope prof(ctx, PROF_TOKENIZE);
            // strict_qual (sidgpu_tokenize with want_qual): the byte-wise grammar also validates the quality
            // columns; inside a quality session k_quality re-reads every line and reports them itself
            if (want_qual && strict_qual) {
                if (stages == 1) k_tokenize<false, 1><<<grid, TOK_THREADS, dyn_smem, ctx->stream>>>(p);
                else k_tokenize<false, 2><<<grid, TOK_THREADS, dyn_smem, ctx->stream>>>(p);
            } else {
                if (stages == 1) k_tokenize<true, 1><<<grid, TOK_THREADS, dyn_smem, ctx->stream>>>(p);
                else k_tokenize<true, 2><<<grid, TOK_THREADS, dyn_smem, ctx->stream>>>(p);
            }
            TRY(check_launch(ctx, "k_tokenize"));
        }
        }
        {
            ProfScope prof(ctx, PROF_ORDER);
            k_blk_sums<<<n_chunks, BLK_THREADS, 0, ctx->stream>>>((const unsigned long long*)ctx->blk.p, n_blocks, (unsigned long long*)ctx->blk_part.p);
            TRY(check_launch(ctx, "k_blk_sums"));
            k_blk_order<<<n_chunks, BLK_THREADS, 0, ctx->stream>>>((const unsigned long long*)ctx->blk.p, n_blocks, (const unsigned long long*)ctx->blk_part.p,
                                                                   site_base, (uint32_t*)ctx->order.p, ctx->site_cap);
            TRY(check_launch(ctx, "k_blk_order"));
        }
        TRY(sync_ctl(ctx));
        const Control& c = *ctx->h_ctl;
        if (c.name_overflow) return ctx->fail(SIDGPU_ECAPACITY, c.name_overflow == 2 ? "chromosome name longer than 65535 bytes" : "chromosome name dictionary is full");
        if (c.error != ~0ull) {
            const int st = (int)(c.error & 7);
            const unsigned long long off = c.error >> 3;
            if (st == LINE_MALFORMED + 5) {              // site store too small for this many lines: grow and retry
                guess = attempt == 0 ? (range_end - range_begin) / 8 + 4096 : (range_end - range_begin) / 2 + 4096;
                if (attempt >= 2) return ctx->fail(SIDGPU_EINTERNAL, "site store sizing failed");
                continue;
            }
            const int code = st == LINE_MISSING_MAPQ ? SIDGPU_EMISSING_MAPQ : st == LINE_QUAL_SHORT ? SIDGPU_EQUAL_SHORT
                             : st == LINE_MALFORMED ? SIDGPU_EMALFORMED : SIDGPU_EINTERNAL;
            return ctx->fail(code, "%s (line starting at byte %llu)", status_text(st), off);
        }
        if (c.table_overflow) {
            TRY(grow_table(ctx, 2, keep_sites ? site_base : 0));
            continue;
        }
        *n_out = c.n_sites;
        if (c.n_sites >= 64) ctx->avg_line_bytes = (double)(range_end - range_begin) / (double)c.n_sites;   // sizes the slices of the next call
        // keep the load factor below one half for the next chunk
        while ((uint64_t)c.n_entries * 2 > ctx->tab.cap) TRY(grow_table(ctx, 2, site_base + c.n_sites));
        return SIDGPU_OK;
    }
}

// K1 in its ROWS form over [range_begin, range_end) of a streaming `local` session, then the compaction of the
// slices' regions into d_out: text in, CSV rows out, nothing stored per site.
int run_tok2_rows(sidgpu_ctx* ctx, const char* d_text, size_t text_len, size_t range_begin, size_t range_end, char* d_out, size_t out_cap,
                  uint64_t* bytes_out, uint64_t* rows_out, uint64_t* n_out) {
    if (((uintptr_t)d_text & 15) != 0) return ctx->fail(SIDGPU_EINVAL, "d_text must be 16-byte aligned");
    if (range_begin > range_end || range_end > text_len) return ctx->fail(SIDGPU_EINVAL, "bad range [%zu,%zu) for text of %zu bytes", range_begin, range_end, text_len);
    *bytes_out = *rows_out = *n_out = 0;
    if (range_begin == range_end) return sync_ctl(ctx);
    const uint64_t tile0 = range_begin & ~(uint64_t)15;
    const uint64_t span = range_end - tile0;
    static const double slice_lines = getenv("SIDGPU_SLICE_LINES") ? atof(getenv("SIDGPU_SLICE_LINES")) : 31.0;
    static const int force_stages = getenv("SIDGPU_TOK_STAGES") ? atoi(getenv("SIDGPU_TOK_STAGES")) : 0;
    uint32_t slice = (uint32_t)(ctx->avg_line_bytes * slice_lines) & ~31u;
    slice = std::max<uint32_t>(SLICE_MIN, std::min<uint32_t>(SLICE_MAX, slice));
    uint32_t ext = ((uint32_t)(ctx->avg_line_bytes * 1.25) + 31u) & ~31u;
    ext = std::max<uint32_t>(128u, std::min<uint32_t>(2016u, ext));
    const uint64_t tile_bytes = (uint64_t)slice * TOK_PARSE_WARPS;
    const uint64_t n_tiles64 = (span + tile_bytes - 1) / tile_bytes;
    if (n_tiles64 * TOK_PARSE_WARPS > 0x7FFFFFFFull) return ctx->fail(SIDGPU_EINVAL, "range too large");
    const uint32_t n_tiles = (uint32_t)n_tiles64, n_regions = n_tiles * TOK_PARSE_WARPS;
    const uint32_t n_chunks = (n_regions + RC_REGIONS - 1) / RC_REGIONS;
    TRY(ensure(ctx, ctx->blk, (size_t)n_regions * 8));
    TRY(ensure(ctx, ctx->rows_part, (size_t)n_chunks * 8));
    TRY(ensure(ctx, ctx->rows_part_rows, (size_t)n_chunks * 8));
    for (int attempt = 0;; ++attempt) {
        // a region holds the rows of one slice: about 30 rows of up to 30 + 46 bytes when the slice was sized for 29.5 lines
        const double lines_per_slice = (double)slice / std::max(8.0, ctx->avg_line_bytes);
        uint64_t region_cap64 = (uint64_t)((lines_per_slice * 80.0 + 512.0) * ctx->region_factor);
        region_cap64 = std::min<uint64_t>((region_cap64 + 15) & ~(uint64_t)15, (uint64_t)1 << 24);
        const uint32_t region_cap = (uint32_t)region_cap64;
        TRY(ensure(ctx, ctx->rows_scratch, (size_t)n_regions * region_cap));
        CK(cudaMemsetAsync(ctl_field(ctx, &Control::tok_ticket), 0, sizeof(unsigned int), ctx->stream));
        CK(cudaMemsetAsync(ctl_field(ctx, &Control::n_sites), 0, sizeof(unsigned long long), ctx->stream));
        CK(cudaMemsetAsync(ctl_field(ctx, &Control::error), 0xFF, sizeof(unsigned long long), ctx->stream));
        CK(cudaMemsetAsync(ctl_field(ctx, &Control::csv_bytes), 0, 2 * sizeof(unsigned long long), ctx->stream));
        Tok2Params q {};
        q.text = (const uint8_t*)d_text; q.text_len = text_len; q.range_begin = range_begin; q.range_end = range_end;
        q.tile0 = tile0; q.n_tiles = n_tiles;
        q.site_alloc = ctl_field(ctx, &Control::n_sites);
        q.rows = (uint8_t*)ctx->rows_scratch.p; q.region_cap = region_cap;
        q.prior = ctx->session_prior; q.error_threshold = ctx->params.error_threshold; q.alpha = ctx->params.significance_level;
        q.het_only = ctx->params.het_only;
        q.tile_ticket = ctl_field(ctx, &Control::tok_ticket);
        q.blk = (unsigned long long*)ctx->blk.p;
        q.error = ctl_field(ctx, &Control::error);
        q.table = ctx->tab; q.names = ctx->names; q.use_table = 1;
        q.slice_bytes = slice; q.text_stride = tok_text_stride(slice, ext); q.tail_bytes = tok_tail_bytes(ext); q.lines_cap = slice / 8;
        q.ext_bytes = ext; q.units_cap = tok2_units(slice, ext) + CW_PAD_UNITS;
        int s2 = 0, s1 = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s2, k_tok2<true, 2>, TOK_THREADS, tok2_dyn_smem(slice, ext, 2, true)) != cudaSuccess || s2 < 1) s2 = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s1, k_tok2<true, 1>, TOK_THREADS, tok2_dyn_smem(slice, ext, 1, true)) != cudaSuccess || s1 < 1) s1 = 1;
        const int stages = force_stages == 1 || force_stages == 2 ? force_stages : (s1 > s2 ? 1 : 2);
        const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * (stages == 1 ? s1 : s2));
        const uint32_t dyn = tok2_dyn_smem(slice, ext, (uint32_t)stages, true);
        {
            ProfScope prof(ctx, PROF_TOKENIZE);
            if (stages == 1) k_tok2<true, 1><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
            else k_tok2<true, 2><<<grid, TOK_THREADS, dyn, ctx->stream>>>(q);
            TRY(check_launch(ctx, "k_tok2 (rows)"));
        }
        {
            ProfScope prof(ctx, PROF_CSV);
            k_rows_sums<<<n_chunks, 256, 0, ctx->stream>>>((const unsigned long long*)ctx->blk.p, n_regions, (unsigned long long*)ctx->rows_part.p,
                                                          (unsigned long long*)ctx->rows_part_rows.p);
            TRY(check_launch(ctx, "k_rows_sums"));
            k_rows_scan<<<1, 1024, 0, ctx->stream>>>((unsigned long long*)ctx->rows_part.p, (const unsigned long long*)ctx->rows_part_rows.p, n_chunks,
                                                     ctl_field(ctx, &Control::csv_bytes), ctl_field(ctx, &Control::csv_rows));
            TRY(check_launch(ctx, "k_rows_scan"));
            k_rows_compact<<<n_chunks, RC_THREADS, (RC_THREADS / 32) * RC_STAGE, ctx->stream>>>(
                (const uint8_t*)ctx->rows_scratch.p, region_cap, (const unsigned long long*)ctx->blk.p, n_regions,
                (const unsigned long long*)ctx->rows_part.p, (uint8_t*)d_out, out_cap);
            TRY(check_launch(ctx, "k_rows_compact"));
        }
        TRY(sync_ctl(ctx));
        const Control& c = *ctx->h_ctl;
        if (c.error != ~0ull) {
            const int st = (int)(c.error & 7);
            const unsigned long long off = c.error >> 3;
            if (st == LINE_ROWS_OVERFLOW) {
                if (attempt >= 6) return ctx->fail(SIDGPU_EINTERNAL, "row region sizing failed");
                ctx->region_factor *= 2.0;
                continue;
            }
            const int code = st == LINE_MALFORMED ? SIDGPU_EMALFORMED : SIDGPU_EINTERNAL;
            return ctx->fail(code, "%s (line starting at byte %llu)", status_text(st), off);
        }
        if (c.table_overflow) {
            TRY(grow_table(ctx, 2, 0));
            continue;
        }
        *n_out = c.n_sites;
        *bytes_out = c.csv_bytes;
        *rows_out = c.csv_rows;
        if (c.n_sites >= 64) ctx->avg_line_bytes = (double)(range_end - range_begin) / (double)c.n_sites;
        while ((uint64_t)c.n_entries * 2 > ctx->tab.cap) TRY(grow_table(ctx, 2, 0));
        ctx->classified = c.n_entries;                       // every entry was classified by the lane that inserted it
        if (c.csv_bytes > out_cap) return ctx->fail(SIDGPU_ECAPACITY, "CSV needs %llu bytes, buffer has %zu", c.csv_bytes, out_cap);
        return SIDGPU_OK;
    }
}

__global__ void k_insert_profiles(TableView t, const uint64_t* profiles, const uint64_t* weights, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = table_find_or_insert(t, profiles[i]);
    atomicAdd(&t.counts[slot], weights ? (unsigned long long)weights[i] : 1ull);
}

// For each local table entry: its row in a merged, lexicographically sorted profile list
// (0xFFFFFFFF: not present there, i.e. coverage < 4 -> dropped, call.cpp:66-70,131-140).
__global__ void k_map_entries(TableView t, uint32_t n_entries, const unsigned long long* sorted_profiles, uint32_t n, uint32_t* entry_to_unique) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const uint64_t key = profile_sort_key(t.keys[t.entry_list[e]]);
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (profile_sort_key(sorted_profiles[mid]) < key) lo = mid + 1; else hi = mid;
    }
    entry_to_unique[e] = (lo < n && profile_sort_key(sorted_profiles[lo]) == key) ? lo : 0xFFFFFFFFu;
}

template <class T>
__global__ void k_gather(T* out, const T* in, const uint32_t* order, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[order[i]];
}

__global__ void k_count_slots(const uint32_t* slot, uint64_t begin, uint64_t n, unsigned long long* counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[slot[begin + i]], 1ull);
}

LynchConsts host_lynch_consts(const double nd[4], double eps) { return lynch_consts(nd, eps); }

int classify_entries(sidgpu_ctx* ctx, uint32_t first, uint32_t last) {
    if (first >= last) return SIDGPU_OK;
    ClassifyParams p {};
    p.table = ctx->tab;
    p.first = first;
    p.last = last;
    p.method = ctx->params.method == SIDGPU_METHOD_LOCAL ? 0 : ctx->params.method == SIDGPU_METHOD_BAYES ? 1 : 2;
    p.prior = ctx->session_prior;
    p.error_threshold = ctx->params.error_threshold;
    p.alpha = ctx->params.significance_level;
    p.het_only = ctx->params.het_only;
    if (p.method != 0) {
        p.lynch = host_lynch_consts(ctx->fit_nd, ctx->fit.eps);
        p.pi = ctx->fit.pi;
        p.use_prior = ctx->params.estimate_prior;
        p.adj_hom = (const double*)ctx->adj_hom.p;
        p.adj_het = (const double*)ctx->adj_het.p;
        p.entry_to_unique = (const uint32_t*)ctx->entry_to_unique.p;
    }
    {
        ProfScope prof(ctx, PROF_CLASSIFY);
        k_classify<<<(last - first + 127) / 128, 128, 0, ctx->stream>>>(p);
    }
    return check_launch(ctx, "k_classify");
}

int sort_pairs(sidgpu_ctx* ctx, unsigned long long* keys, uint32_t* vals, uint32_t n_pow2) {
    const uint32_t blocks = n_pow2 / BITONIC_BLOCK;
    k_bitonic_local<<<blocks, BITONIC_BLOCK / 2, 0, ctx->stream>>>(keys, vals, n_pow2, 0, 0, 1);
    TRY(check_launch(ctx, "k_bitonic_local"));
    for (uint32_t k = BITONIC_BLOCK * 2; k <= n_pow2; k <<= 1) {
        for (uint32_t j = k >> 1; j >= BITONIC_BLOCK; j >>= 1) {
            k_bitonic_step<<<(n_pow2 + 255) / 256, 256, 0, ctx->stream>>>(keys, vals, n_pow2, j, k);
            TRY(check_launch(ctx, "k_bitonic_step"));
        }
        k_bitonic_local<<<blocks, BITONIC_BLOCK / 2, 0, ctx->stream>>>(keys, vals, n_pow2, BITONIC_BLOCK / 2, k, 0);
        TRY(check_launch(ctx, "k_bitonic_local"));
    }
    return SIDGPU_OK;
}

uint32_t pow2_at_least(uint64_t n) {
    uint32_t p = BITONIC_BLOCK;
    while (p < n) p <<= 1;
    return p;
}

// The histogram of table `tab` (its first n_entries entries): unique arrays in lexicographic order, logM, integer
// nucleotide sums; e2u[entry] = row of the entry in the unique arrays.
int build_histogram_of(sidgpu_ctx* ctx, const TableView& tab, uint32_t n_entries, uint32_t min_cov, DevBuf& e2u) {
    ProfScope prof(ctx, PROF_HIST);
    const uint32_t np2 = pow2_at_least(n_entries);
    TRY(ensure(ctx, ctx->sort_keys, (size_t)np2 * 8));
    TRY(ensure(ctx, ctx->sort_vals, (size_t)np2 * 4));
    TRY(ensure(ctx, e2u, (size_t)std::max<uint32_t>(n_entries, 1) * 4));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::n_selected), 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::nd_acc), 0, 5 * sizeof(unsigned long long), ctx->stream));
    ctx->n_unique = 0;
    if (n_entries) {
        k_fill_u32<<<(n_entries + 255) / 256, 256, 0, ctx->stream>>>((uint32_t*)e2u.p, n_entries, 0xFFFFFFFFu);
        TRY(check_launch(ctx, "k_fill_u32"));
        k_hist_select<<<(n_entries + 255) / 256, 256, 0, ctx->stream>>>(tab, n_entries, min_cov, (unsigned long long*)ctx->sort_keys.p,
                                                                         (uint32_t*)ctx->sort_vals.p, ctl_field(ctx, &Control::n_selected));
        TRY(check_launch(ctx, "k_hist_select"));
        TRY(sync_ctl(ctx));
        const uint32_t n_sel = ctx->h_ctl->n_selected;
        if (n_sel) {
            const uint32_t sp2 = pow2_at_least(n_sel);
            k_fill_pad<<<(sp2 - n_sel + 255) / 256 + 1, 256, 0, ctx->stream>>>((unsigned long long*)ctx->sort_keys.p, (uint32_t*)ctx->sort_vals.p, n_sel, sp2);
            TRY(check_launch(ctx, "k_fill_pad"));
            TRY(sort_pairs(ctx, (unsigned long long*)ctx->sort_keys.p, (uint32_t*)ctx->sort_vals.p, sp2));
            TRY(ensure(ctx, ctx->u_profile, (size_t)n_sel * 8));
            TRY(ensure(ctx, ctx->u_count, (size_t)n_sel * 8));
            TRY(ensure(ctx, ctx->u_logM, (size_t)n_sel * 8));
            k_hist_gather<<<(n_sel + 255) / 256, 256, 0, ctx->stream>>>(tab, (const uint32_t*)ctx->sort_vals.p, n_sel,
                                                                         (unsigned long long*)ctx->u_profile.p, (unsigned long long*)ctx->u_count.p,
                                                                         (double*)ctx->u_logM.p, (uint32_t*)e2u.p,
                                                                         ctl_field(ctx, &Control::nd_acc)[0]);
            TRY(check_launch(ctx, "k_hist_gather"));
            TRY(sync_ctl(ctx));
        }
        ctx->n_unique = n_sel;
    }
    const unsigned long long* acc = ctx->h_ctl->nd_acc;
    if (ctx->n_unique && acc[4] != 0) {
        for (int i = 0; i < 4; ++i) ctx->fit_nd[i] = (double)acc[i] / (double)acc[4];     // pileup.cpp:209-213
    } else {
        for (int i = 0; i < 4; ++i) ctx->fit_nd[i] = 0.25;                               // pileup.cpp:215
    }
    ctx->hist_valid = true;
    return SIDGPU_OK;
}

int build_histogram(sidgpu_ctx* ctx, uint32_t min_cov) {
    TRY(sync_ctl(ctx));
    ctx->global_hist = false;
    return build_histogram_of(ctx, ctx->tab, ctx->h_ctl->n_entries, min_cov, ctx->entry_to_unique);
}

int launch_objective(sidgpu_ctx* ctx, const double nd[4], double pi, double eps, double* d_out) {
    if (!ctx->hist_valid) return ctx->fail(SIDGPU_ESTATE, "no histogram: call sidgpu_histogram or sidgpu_finish first");
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((ctx->n_unique + OBJ_THREADS - 1) / OBJ_THREADS, (uint64_t)ctx->sm_count * 4));
    TRY(ensure(ctx, ctx->partials, (size_t)ctx->sm_count * 4 * 16));
    ObjParams p {};
    p.u_profile = (const unsigned long long*)ctx->u_profile.p;
    p.u_count = (const unsigned long long*)ctx->u_count.p;
    p.u_logM = (const double*)ctx->u_logM.p;
    p.n_unique = (uint32_t)ctx->n_unique;
    p.k = host_lynch_consts(nd, eps);
    p.log1m_pi = log1p(-pi);
    p.log_pi = log(pi);
    p.partials = (double*)ctx->partials.p;
    p.done_blocks = ctl_field(ctx, &Control::obj_done);
    p.out = d_out;
    k_lynch_objective<<<grid, OBJ_THREADS, 0, ctx->stream>>>(p);
    return check_launch(ctx, "k_lynch_objective");
}

int objective_value(sidgpu_ctx* ctx, const double nd[4], double pi, double eps, double* value) {
    if (pi < 0 || pi > 1 || eps < 0 || eps > 1) {        // lynch.cpp:41-43
        *value = std::numeric_limits<double>::max();
        return SIDGPU_OK;
    }
    double* d_obj = ctl_field(ctx, &Control::objective);
    TRY(launch_objective(ctx, nd, pi, eps, d_obj));
    TRY(sync_ctl(ctx));
    *value = ctx->h_ctl->objective;
    return SIDGPU_OK;
}

int run_fit_host(sidgpu_ctx* ctx, const double nd[4], sidgpu_fit* out);

// The whole Nelder-Mead fit as one cooperative kernel (k_lynch_fit); SIDGPU_FIT=host: the round-1 form, the simplex
// on the host and one launch + read-back per objective evaluation.
int run_fit(sidgpu_ctx* ctx, const double nd[4], sidgpu_fit* out) {
    static const bool host_loop = getenv("SIDGPU_FIT") && strcmp(getenv("SIDGPU_FIT"), "host") == 0;
    if (host_loop) return run_fit_host(ctx, nd, out);
    if (!ctx->hist_valid) return ctx->fail(SIDGPU_ESTATE, "no histogram: call sidgpu_histogram or sidgpu_finish first");
    ProfScope prof(ctx, PROF_FIT);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lynch_fit, OBJ_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    // one profile per thread while that fits on a quarter of the SMs (a small grid keeps the barrier cheap), then grid-stride
    const uint64_t want = (ctx->n_unique + OBJ_THREADS - 1) / OBJ_THREADS;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)ctx->sm_count * std::min(per_sm, 2)));
    TRY(ensure(ctx, ctx->partials, (size_t)std::max<unsigned>(grid, (unsigned)ctx->sm_count * 4) * 4 * sizeof(double)));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::fit_barrier), 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::error), 0xFF, sizeof(unsigned long long), ctx->stream));
    FitParams p {};
    p.u_profile = (const unsigned long long*)ctx->u_profile.p;
    p.u_count = (const unsigned long long*)ctx->u_count.p;
    p.u_logM = (const double*)ctx->u_logM.p;
    p.n_unique = (uint32_t)ctx->n_unique;
    for (int i = 0; i < 4; ++i) p.nd[i] = nd[i];
    p.x0[0] = p.x0[1] = 1e-3;                                        // lynch.cpp:8-10,20
    p.step[0] = p.step[1] = 1e-4;
    p.partials = (double*)ctx->partials.p;
    p.barrier = ctl_field(ctx, &Control::fit_barrier);
    p.out = ctl_field(ctx, &Control::fit_out)[0];
    p.error = ctl_field(ctx, &Control::error);
    void* args[] = {&p};
    CK(cudaLaunchCooperativeKernel((const void*)k_lynch_fit, dim3(grid), dim3(OBJ_THREADS), args, 0, ctx->stream));
    TRY(check_launch(ctx, "k_lynch_fit"));
    TRY(sync_ctl(ctx));
    if (ctx->h_ctl->error != ~0ull) return ctx->fail(SIDGPU_EINTERNAL, "k_lynch_fit: grid barrier timed out");
    const double* r = ctx->h_ctl->fit_out;
    out->pi = r[0];
    out->eps = r[1];
    out->fval = r[2];
    out->iterations = (int)r[3];
    out->evaluations = (int)r[4];
    out->converged = r[5] != 0.0;
    return SIDGPU_OK;
}

struct HostObjective {           // one device reduction + read-back per evaluation
    sidgpu_ctx* ctx;
    const double* nd;
    int* rc;
    SID_HD double operator()(double pi, double eps) {
        double v = 0;
#if !defined(__CUDA_ARCH__)
        v = std::numeric_limits<double>::quiet_NaN();
        if (*rc == SIDGPU_OK) *rc = objective_value(ctx, nd, pi, eps, &v);
#endif
        return v;
    }
};

int run_fit_host(sidgpu_ctx* ctx, const double nd[4], sidgpu_fit* out) {
    ProfScope prof(ctx, PROF_FIT);
    int rc = SIDGPU_OK;
    HostObjective f {ctx, nd, &rc};
    const double x0[2] = {1e-3, 1e-3}, step[2] = {1e-4, 1e-4};       // lynch.cpp:8-10,20
    NelderMeadResult r = nelder_mead_2d(f, x0, step);
    if (rc != SIDGPU_OK) return rc;
    out->pi = r.x[0];
    out->eps = r.x[1];
    out->fval = r.fval;
    out->iterations = r.iterations;
    out->evaluations = r.evaluations;
    out->converged = r.converged ? 1 : 0;
    return SIDGPU_OK;
}

int bh_adjust(sidgpu_ctx* ctx, const double* d_p, uint32_t n, double* d_adj) {
    if (n == 0) return SIDGPU_OK;
    const uint32_t np2 = pow2_at_least(n);
    TRY(ensure(ctx, ctx->sort_keys, (size_t)np2 * 8));
    TRY(ensure(ctx, ctx->sort_vals, (size_t)np2 * 4));
    const uint32_t blocks = (n + BH_THREADS - 1) / BH_THREADS;
    TRY(ensure(ctx, ctx->bh_c, (size_t)n * 8));
    TRY(ensure(ctx, ctx->bh_block, (size_t)blocks * 8));
    unsigned long long* keys = (unsigned long long*)ctx->sort_keys.p;
    uint32_t* vals = (uint32_t*)ctx->sort_vals.p;
    k_bh_keys<<<blocks, BH_THREADS, 0, ctx->stream>>>(d_p, n, keys, vals);
    TRY(check_launch(ctx, "k_bh_keys"));
    k_fill_pad<<<(np2 - n + 255) / 256 + 1, 256, 0, ctx->stream>>>(keys, vals, n, np2);
    TRY(check_launch(ctx, "k_fill_pad"));
    TRY(sort_pairs(ctx, keys, vals, np2));
    k_bh_scan1<<<blocks, BH_THREADS, 0, ctx->stream>>>(d_p, vals, n, (double*)ctx->bh_c.p, (double*)ctx->bh_block.p);
    TRY(check_launch(ctx, "k_bh_scan1"));
    k_bh_scan2<<<1, 1, 0, ctx->stream>>>((double*)ctx->bh_block.p, blocks);
    TRY(check_launch(ctx, "k_bh_scan2"));
    k_bh_scatter<<<blocks, BH_THREADS, 0, ctx->stream>>>((const double*)ctx->bh_c.p, (const double*)ctx->bh_block.p, vals, n, d_adj);
    return check_launch(ctx, "k_bh_scatter");
}

bool method_streams(const sidgpu_params& p) {
    return (p.method == SIDGPU_METHOD_LOCAL || p.method == SIDGPU_METHOD_QUALITY) && !p.estimate_prior;
}

int quality_rows(sidgpu_ctx* ctx, uint64_t site_begin, uint64_t n, uint8_t* rec_label = nullptr, char* rec_gt = nullptr,
                 double* rec_hom = nullptr, double* rec_het = nullptr) {
    if (!ctx->last_text) return ctx->fail(SIDGPU_ESTATE, "quality needs the text of the chunk: feed before emit");
    if (n == 0) return SIDGPU_OK;
    QualityParams q {};
    q.text = (const uint8_t*)ctx->last_text;
    q.text_len = ctx->last_text_len;
    q.line_off = (const uint64_t*)ctx->line_off.p;
    q.profile = (const uint64_t*)ctx->profile.p;
    q.qual_l = ctx->qual_sums_valid ? (const double*)ctx->qual_l.p : nullptr;
    // the rows of a whole chunk need no particular order here (K6 reads the suffixes through order[] afterwards)
    const bool whole = !rec_label && !rec_gt && !rec_hom && !rec_het && site_begin == 0 && n == ctx->n_sites_total;
    q.order = whole ? nullptr : (const uint32_t*)ctx->order.p;
    q.site_begin = site_begin;
    q.n_sites = n;
    q.lut = (const double*)ctx->quality_lut.p;
    q.prior = ctx->session_prior;
    q.alpha = ctx->params.significance_level;
    q.het_only = ctx->params.het_only;
    q.site_suffix = (char*)ctx->site_suffix.p;
    q.error = ctl_field(ctx, &Control::error);
    q.rec_label = rec_label;
    q.rec_gt = rec_gt;
    q.rec_hom = rec_hom;
    q.rec_het = rec_het;
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::error), 0xFF, sizeof(unsigned long long), ctx->stream));
    k_quality<<<(unsigned)((n + QUAL_THREADS - 1) / QUAL_THREADS), QUAL_THREADS, 0, ctx->stream>>>(q);
    return check_launch(ctx, "k_quality");
}

}  // namespace

// ================================================================================================
extern "C" {

const char* sidgpu_version(void) { return "sid-b200 0.1 (sm_100a)"; }

const char* sidgpu_last_error(const sidgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

uint64_t sidgpu_launch_count(const sidgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- the per-read side of pileup.hpp and the per-profile side of lynch.hpp / stats.hpp (k_reads.cuh)
int sidgpu_qualities(sidgpu_ctx* ctx, const char* d_quals, size_t n, uint8_t* d_out, uint64_t* n_out) {
    if (!ctx || !n_out || (n && (!d_quals || !d_out))) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    *n_out = 0;
    if (n == 0) return SIDGPU_OK;
    k_qualities<<<1, 256, 0, ctx->stream>>>((const uint8_t*)d_quals, n, d_out, ctl_field(ctx, &Control::csv_bytes));
    TRY(check_launch(ctx, "k_qualities"));
    TRY(sync_ctl(ctx));
    *n_out = ctx->h_ctl->csv_bytes;
    return SIDGPU_OK;
}

int sidgpu_read_counts(sidgpu_ctx* ctx, const char* d_text, size_t text_len, const uint64_t* d_line_off, uint64_t n_lines, int want_baseq,
                       int want_mapq, uint32_t* d_n_bases, uint32_t* d_n_bq, uint32_t* d_n_mq) {
    if (!ctx || (n_lines && (!d_text || !d_line_off || !d_n_bases || !d_n_bq || !d_n_mq))) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (n_lines == 0) return SIDGPU_OK;
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::error), 0xFF, sizeof(unsigned long long), ctx->stream));
    k_read_counts<<<(unsigned)((n_lines + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)d_text, text_len, d_line_off, n_lines, want_baseq,
                                                                              want_mapq, d_n_bases, d_n_bq, d_n_mq, ctl_field(ctx, &Control::error));
    TRY(check_launch(ctx, "k_read_counts"));
    TRY(sync_ctl(ctx));
    if (ctx->h_ctl->error != ~0ull) {
        const int st = (int)(ctx->h_ctl->error & 7);
        return ctx->fail(st == LINE_MISSING_MAPQ ? SIDGPU_EMISSING_MAPQ : SIDGPU_EMALFORMED, "%s (line starting at byte %llu)", status_text(st),
                         ctx->h_ctl->error >> 3);
    }
    return SIDGPU_OK;
}

int sidgpu_read_fill(sidgpu_ctx* ctx, const char* d_text, size_t text_len, const uint64_t* d_line_off, uint64_t n_lines,
                     const uint64_t* d_base_off, const uint64_t* d_bq_off, const uint64_t* d_mq_off, char* d_bases, uint8_t* d_strands,
                     uint8_t* d_bq, uint8_t* d_mq) {
    if (!ctx || (n_lines && (!d_text || !d_line_off))) return SIDGPU_EINVAL;
    if (((d_bases || d_strands) && !d_base_off) || (d_bq && !d_bq_off) || (d_mq && !d_mq_off)) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (n_lines == 0) return SIDGPU_OK;
    k_read_fill<<<(unsigned)((n_lines + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)d_text, text_len, d_line_off, n_lines, d_base_off,
                                                                            d_bq_off, d_mq_off, d_bases, d_strands, d_bq, d_mq);
    TRY(check_launch(ctx, "k_read_fill"));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

int sidgpu_strand_counts(sidgpu_ctx* ctx, const char* d_text, size_t text_len, const uint64_t* d_line_off, uint64_t n_lines, uint64_t* d_fwd,
                         uint64_t* d_rev) {
    if (!ctx || (n_lines && (!d_text || !d_line_off))) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (n_lines == 0) return SIDGPU_OK;
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::error), 0xFF, sizeof(unsigned long long), ctx->stream));
    k_strand_counts<<<(unsigned)((n_lines + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)d_text, text_len, d_line_off, n_lines,
                                                                                (unsigned long long*)d_fwd, (unsigned long long*)d_rev,
                                                                                ctl_field(ctx, &Control::error));
    TRY(check_launch(ctx, "k_strand_counts"));
    TRY(sync_ctl(ctx));
    if (ctx->h_ctl->error != ~0ull)
        return ctx->fail(SIDGPU_EMALFORMED, "%s (line starting at byte %llu)", status_text((int)(ctx->h_ctl->error & 7)), ctx->h_ctl->error >> 3);
    return SIDGPU_OK;
}

int sidgpu_profile_loglik(sidgpu_ctx* ctx, const uint64_t* d_profiles, uint64_t n, const double nd[4], double eps, double* d_log_hom,
                          double* d_log_het) {
    if (!ctx || !nd || (n && (!d_profiles || !d_log_hom || !d_log_het))) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return SIDGPU_OK;
    k_profile_loglik<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const unsigned long long*)d_profiles, n, host_lynch_consts(nd, eps),
                                                                           d_log_hom, d_log_het);
    TRY(check_launch(ctx, "k_profile_loglik"));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

int sidgpu_lr_test(sidgpu_ctx* ctx, const double* d_log_h0, const double* d_log_h1, uint64_t n, double* d_p) {
    if (!ctx || (n && (!d_log_h0 || !d_log_h1 || !d_p))) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return SIDGPU_OK;
    k_lr_test<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_log_h0, d_log_h1, n, d_p);
    TRY(check_launch(ctx, "k_lr_test"));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

int sidgpu_create(const sidgpu_config* cfg, sidgpu_ctx** out) {
    if (!out) return SIDGPU_EINVAL;
    *out = nullptr;
    sidgpu_config c {};
    if (cfg) c = *cfg;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libsidgpu has no CPU fallback)";
        return SIDGPU_ECUDA;
    }
    if (c.device < 0 || c.device >= n_dev) { g_create_error = "bad device ordinal"; return SIDGPU_EINVAL; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(c.device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, c.device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return SIDGPU_ECUDA;
    }
    if (prop.major != 10) {
        g_create_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                         "; libsidgpu carries sm_100a code only";
        return SIDGPU_ECUDA;
    }
    sidgpu_ctx* ctx = new sidgpu_ctx;
    ctx->device = c.device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_chunk = c.max_chunk_bytes ? c.max_chunk_bytes : ((size_t)256 << 20);
    auto bail = [&](int rc) { g_create_error = ctx->err; sidgpu_destroy(ctx); return rc; };
    if (c.stream) ctx->stream = (cudaStream_t)c.stream;
    else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(SIDGPU_ECUDA); }
        ctx->own_stream = true;
    }
    if (cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "stream creation failed"; return bail(SIDGPU_ECUDA); }
    if (cudaMalloc((void**)&ctx->d_ctl, sizeof(Control)) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_ctl, sizeof(Control), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->h_ctl_dev, ctx->h_ctl, 0) != cudaSuccess) {
        ctx->err = "control block allocation failed";
        return bail(SIDGPU_ENOMEM);
    }
    cudaMemsetAsync(ctx->d_ctl, 0, sizeof(Control), ctx->stream);
    ctx->table_log2 = c.table_log2 > 0 ? c.table_log2 : 20;
    int rc = alloc_table(ctx, ctx->table_log2, ctx->tab);
    if (rc != SIDGPU_OK) return bail(rc);
    // chromosome-name dictionary: 2^20 slots, 64 MiB of names
    const uint32_t dict_slots = 1u << 20, pool_cap = 64u << 20;
    if (cudaMalloc((void**)&ctx->names.slots, (size_t)dict_slots * 8) != cudaSuccess || cudaMalloc((void**)&ctx->names.pool, pool_cap) != cudaSuccess) {
        ctx->err = "name dictionary allocation failed";
        return bail(SIDGPU_ENOMEM);
    }
    cudaMemsetAsync(ctx->names.slots, 0, (size_t)dict_slots * 8, ctx->stream);
    cudaMemsetAsync(ctx->names.pool, 0, 16, ctx->stream);
    ctx->names.mask = dict_slots - 1;
    ctx->names.pool_cap = pool_cap;
    ctx->names.cursor = ctl_field(ctx, &Control::name_cursor);
    ctx->names.overflow = ctl_field(ctx, &Control::name_overflow);
    const unsigned int four = 4;
    cudaMemcpyAsync(ctx->names.cursor, &four, sizeof four, cudaMemcpyHostToDevice, ctx->stream);
    // quality lookup tables: log(1-e), log(e), log(1-2e/3), log(2e/3) with e = 10^(-q/10) (call.cpp:330-341)
    {
        std::vector<double> lut(4 * 256 + LOG_FACT_N);
        for (int q = 0; q < 256; ++q) {
            const double error = pow(10., q / -10.);
            lut[q] = log(1 - error);
            lut[256 + q] = log(error);
            lut[512 + q] = log(1 - 2. / 3. * error);
            lut[768 + q] = log(2. / 3. * error);
        }
        for (int n = 0; n < LOG_FACT_N; ++n) lut[1024 + n] = lgamma((double)n + 1.0);     // log n! for the binomial of call.cpp:347-349
        rc = ensure(ctx, ctx->quality_lut, lut.size() * 8);
        if (rc != SIDGPU_OK) return bail(rc);
        cudaMemcpyAsync(ctx->quality_lut.p, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    if ((e = cudaFuncSetAttribute(k_tokenize<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok_dyn_smem(SLICE_MAX, 2016))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tokenize<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok_dyn_smem(SLICE_MAX, 2016))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tokenize<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok_dyn_smem(SLICE_MAX, 2016))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tokenize<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok_dyn_smem(SLICE_MAX, 2016))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok2_dyn_smem(SLICE_MAX, 2016, 1, false))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<uint32_t>(227u << 10, tok2_dyn_smem(SLICE_MAX, 2016, 2, false)))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok2_dyn_smem(SLICE_MAX, 2016, 1, false))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<uint32_t>(227u << 10, tok2_dyn_smem(SLICE_MAX, 2016, 2, false)))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<false, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok2_dyn_smem(SLICE_MAX, 2016, 1, false))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<false, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<uint32_t>(227u << 10, tok2_dyn_smem(SLICE_MAX, 2016, 2, false)))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<false, 2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<uint32_t>(227u << 10, tok2_dyn_smem(SLICE_MAX, 2016, 2, false, true)))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok2_dyn_smem(SLICE_MAX, 2016, 1, true))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_tok2<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<uint32_t>(227u << 10, tok2_dyn_smem(SLICE_MAX, 2016, 2, true)))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_rows_compact, cudaFuncAttributeMaxDynamicSharedMemorySize, (RC_THREADS / 32) * RC_STAGE)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_csv, cudaFuncAttributeMaxDynamicSharedMemorySize, CSV_STAGE)) != cudaSuccess) {
        ctx->err = std::string("shared memory opt-in: ") + cudaGetErrorString(e);
        return bail(SIDGPU_ECUDA);
    }
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(SIDGPU_ECUDA); }
    if (c.max_sites) {
        rc = ensure_sites(ctx, c.max_sites, false);
        if (rc != SIDGPU_OK) return bail(rc);
    }
    *out = ctx;
    return SIDGPU_OK;
}

void sidgpu_destroy(sidgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    resolve_profile(ctx);
    for (cudaEvent_t e : ctx->free_events) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        release(ctx->hp_text[i]);
        release(ctx->hp_csv[i]);
        if (ctx->hp_ev_in[i]) cudaEventDestroy(ctx->hp_ev_in[i]);
        if (ctx->hp_ev_out[i]) cudaEventDestroy(ctx->hp_ev_out[i]);
    }
    free_table(ctx->tab);
    if (ctx->g_log2) free_table(ctx->g_tab);
    cudaFree(ctx->names.slots);
    cudaFree(ctx->names.pool);
    for (DevBuf* b : {&ctx->blk, &ctx->blk_part, &ctx->order, &ctx->v_pos, &ctx->v_slot, &ctx->v_name_ref, &ctx->v_profile, &ctx->v_line_off, &ctx->csv_status, &ctx->pos, &ctx->slot, &ctx->name_ref, &ctx->profile, &ctx->line_off,
                      &ctx->site_suffix, &ctx->rows_scratch, &ctx->rows_part, &ctx->rows_part_rows, &ctx->sort_keys, &ctx->sort_vals, &ctx->u_profile, &ctx->u_count, &ctx->u_logM,
                      &ctx->entry_to_unique, &ctx->g_e2u, &ctx->p_hom, &ctx->p_het, &ctx->adj_hom, &ctx->adj_het, &ctx->bh_c, &ctx->bh_block,
                      &ctx->partials, &ctx->quality_lut, &ctx->hp_comp[0], &ctx->hp_comp[1], &ctx->inf_blocks[0], &ctx->inf_blocks[1], &ctx->crc_tables, &ctx->fwd, &ctx->v_fwd})
        release(*b);
    if (ctx->d_ctl) cudaFree(ctx->d_ctl);
    if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
    if (ctx->inflate_stream) cudaStreamDestroy(ctx->inflate_stream);
    for (int i = 0; i < 2; ++i) if (ctx->ev_inflate[i]) cudaEventDestroy(ctx->ev_inflate[i]);
    if (ctx->d_inf) cudaFree(ctx->d_inf);
    if (ctx->h_inf) cudaFreeHost(ctx->h_inf);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int sidgpu_profile(sidgpu_ctx* ctx, int enable) {
    if (!ctx) return SIDGPU_EINVAL;
    resolve_profile(ctx);
    ctx->profiling = enable != 0;
    for (int i = 0; i < SIDGPU_N_TIMERS; ++i) { ctx->kernel_ms[i] = 0; ctx->kernel_launches[i] = 0; }
    return SIDGPU_OK;
}

int sidgpu_kernel_times(sidgpu_ctx* ctx, double ms[SIDGPU_N_TIMERS], uint64_t launches[SIDGPU_N_TIMERS]) {
    if (!ctx) return SIDGPU_EINVAL;
    CK(cudaStreamSynchronize(ctx->stream));
    resolve_profile(ctx);
    for (int i = 0; i < SIDGPU_N_TIMERS; ++i) {
        if (ms) ms[i] = ctx->kernel_ms[i];
        if (launches) launches[i] = ctx->kernel_launches[i];
    }
    return SIDGPU_OK;
}

int sidgpu_synchronize(sidgpu_ctx* ctx) {
    if (!ctx) return SIDGPU_EINVAL;
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

int sidgpu_malloc(sidgpu_ctx* ctx, size_t bytes, void** d_ptr) {
    if (!ctx || !d_ptr) return SIDGPU_EINVAL;
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) return ctx->fail(SIDGPU_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return SIDGPU_OK;
}
int sidgpu_free(sidgpu_ctx* ctx, void* d_ptr) {
    if (!ctx) return SIDGPU_EINVAL;
    CK(cudaFree(d_ptr));
    return SIDGPU_OK;
}
int sidgpu_malloc_host(sidgpu_ctx* ctx, size_t bytes, void** h_ptr) {
    if (!ctx || !h_ptr) return SIDGPU_EINVAL;
    cudaError_t e = cudaMallocHost(h_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) return ctx->fail(SIDGPU_ENOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    return SIDGPU_OK;
}
int sidgpu_free_host(sidgpu_ctx* ctx, void* h_ptr) {
    if (!ctx) return SIDGPU_EINVAL;
    CK(cudaFreeHost(h_ptr));
    return SIDGPU_OK;
}
int sidgpu_memcpy_h2d(sidgpu_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    if (!ctx) return SIDGPU_EINVAL;
    CK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}
int sidgpu_memcpy_d2d(sidgpu_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
    if (!ctx) return SIDGPU_EINVAL;
    CK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}
int sidgpu_memcpy_d2d_async(sidgpu_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
    if (!ctx) return SIDGPU_EINVAL;
    CK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return SIDGPU_OK;
}
int sidgpu_memcpy_d2h(sidgpu_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    if (!ctx) return SIDGPU_EINVAL;
    CK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

// ------------------------------------------------------------------------------------------- K1
int sidgpu_tokenize(sidgpu_ctx* ctx, const char* d_text, size_t text_len, size_t range_begin, size_t range_end,
                    int want_qual, sidgpu_sites_view* out) {
    Range nvtx_range("sidgpu_tokenize");
    if (!ctx || !out) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    ctx->phase = PHASE_IDLE;
    ctx->want_profile = true;
    ctx->want_line_off = want_qual != 0;
    ctx->want_fwd = want_qual == 3;
    TRY(reset_table(ctx));
    TRY(reset_names(ctx));
    uint64_t n = 0;
    // want_qual 2: the line offsets without the quality columns (a six-column file stays valid)
    const int rc_tok = run_tokenizer(ctx, d_text, text_len, range_begin, range_end, want_qual == 1, true, 0, false, &n, true);
    ctx->want_fwd = false;
    TRY(rc_tok);
    ctx->n_sites_total = n;
    ctx->chunk_begin = 0;
    ctx->chunk_sites = n;
    memset(out, 0, sizeof *out);
    out->n_sites = n;
    // the view is in file order: gather the stored columns through order[]
    const size_t nn = std::max<uint64_t>(n, 1);
    TRY(ensure(ctx, ctx->v_profile, nn * 8));
    TRY(ensure(ctx, ctx->v_pos, nn * 4));
    TRY(ensure(ctx, ctx->v_slot, nn * 4));
    TRY(ensure(ctx, ctx->v_name_ref, nn * 4));
    if (want_qual) TRY(ensure(ctx, ctx->v_line_off, nn * 8));
    if (want_qual == 3) {
        TRY(ensure(ctx, ctx->v_fwd, nn * 8));
        if (n && !ctx->fwd_valid) {
            // long lines (the DEEP tokenizer): the strands by a second walk over the lines, in storage order
            TRY(ensure(ctx, ctx->fwd, nn * 8));
            TRY(sidgpu_strand_counts(ctx, d_text, text_len, (const uint64_t*)ctx->line_off.p, n, (uint64_t*)ctx->fwd.p, nullptr));
        }
    }
    if (n) {
        const unsigned g = (unsigned)((n + 255) / 256);
        const uint32_t* ord = (const uint32_t*)ctx->order.p;
        k_gather<<<g, 256, 0, ctx->stream>>>((uint64_t*)ctx->v_profile.p, (const uint64_t*)ctx->profile.p, ord, n);
        k_gather<<<g, 256, 0, ctx->stream>>>((int32_t*)ctx->v_pos.p, (const int32_t*)ctx->pos.p, ord, n);
        k_gather<<<g, 256, 0, ctx->stream>>>((uint32_t*)ctx->v_slot.p, (const uint32_t*)ctx->slot.p, ord, n);
        k_gather<<<g, 256, 0, ctx->stream>>>((uint32_t*)ctx->v_name_ref.p, (const uint32_t*)ctx->name_ref.p, ord, n);
        if (want_qual) k_gather<<<g, 256, 0, ctx->stream>>>((uint64_t*)ctx->v_line_off.p, (const uint64_t*)ctx->line_off.p, ord, n);
        if (want_qual == 3) k_gather<<<g, 256, 0, ctx->stream>>>((uint64_t*)ctx->v_fwd.p, (const uint64_t*)ctx->fwd.p, ord, n);
        TRY(check_launch(ctx, "k_gather"));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    out->d_profile = (const uint64_t*)ctx->v_profile.p;
    out->d_pos = (const int32_t*)ctx->v_pos.p;
    out->d_slot = (const uint32_t*)ctx->v_slot.p;
    out->d_line_off = want_qual ? (const uint64_t*)ctx->v_line_off.p : nullptr;
    out->d_fwd = want_qual == 3 ? (const uint64_t*)ctx->v_fwd.p : nullptr;
    out->d_name_ref = (const uint32_t*)ctx->v_name_ref.p;
    out->d_names = ctx->names.pool;
    out->names_bytes = ctx->h_ctl->name_cursor;
    return SIDGPU_OK;
}

// ------------------------------------------------------------------------------------- sessions
int sidgpu_begin(sidgpu_ctx* ctx, const sidgpu_params* params) {
    if (!ctx || !params) return SIDGPU_EINVAL;
    if (params->method < 0 || params->method > 3) return ctx->fail(SIDGPU_EINVAL, "unknown method %d", params->method);
    CK(cudaSetDevice(ctx->device));
    ctx->params = *params;
    ctx->streaming = method_streams(*params);
    ctx->counting = !ctx->streaming;
    ctx->want_profile = params->method == SIDGPU_METHOD_QUALITY;      // k_quality takes the counts from the tokenizer
    ctx->want_fwd = params->want_strands != 0;
    ctx->want_line_off = params->method == SIDGPU_METHOD_QUALITY || ctx->want_fwd;      // (strands of long lines: k_strand_counts over the offsets)
    ctx->want_site_suffix = params->method == SIDGPU_METHOD_QUALITY;
    ctx->session_prior = params->prior;
    ctx->fit_done = false;
    ctx->global_hist = false;
    ctx->fit = sidgpu_fit {};
    ctx->n_sites_total = 0;
    ctx->chunk_begin = ctx->chunk_sites = 0;
    ctx->last_text = nullptr;
    TRY(reset_table(ctx));
    TRY(reset_names(ctx));
    ctx->phase = PHASE_FEED;
    return SIDGPU_OK;
}

int sidgpu_feed(sidgpu_ctx* ctx, const char* d_text, size_t text_len, size_t range_begin, size_t range_end, uint64_t* n_sites_out) {
    Range nvtx_range("sidgpu_feed");
    if (!ctx) return SIDGPU_EINVAL;
    const bool second_pass = ctx->phase == PHASE_FINISHED && ctx->params.method == SIDGPU_METHOD_QUALITY;
    if (ctx->phase != PHASE_FEED && !second_pass) return ctx->fail(SIDGPU_ESTATE, "sidgpu_feed outside a session");
    CK(cudaSetDevice(ctx->device));
    const bool is_quality = ctx->params.method == SIDGPU_METHOD_QUALITY;
    const bool chunk_local = ctx->streaming || second_pass || (is_quality && ctx->phase == PHASE_FEED);
    // quality with -R: the first pass only needs the histogram; sites are re-fed after the fit
    const uint64_t base = chunk_local ? 0 : ctx->n_sites_total;
    uint64_t n = 0;
    const bool use_table = !(is_quality && (ctx->streaming || second_pass));
    TRY(run_tokenizer(ctx, d_text, text_len, range_begin, range_end, is_quality, use_table, base, !chunk_local, &n));
    ctx->last_text = d_text;
    ctx->last_text_len = text_len;
    ctx->chunk_begin = base;
    ctx->chunk_sites = n;
    ctx->n_sites_total = chunk_local ? n : base + n;
    if (ctx->want_fwd && !ctx->fwd_valid && n) {
        // the tokenizer form of this chunk does not count strands (long lines, quality sessions): a walk over the chunk's lines
        TRY(sidgpu_strand_counts(ctx, d_text, text_len, (const uint64_t*)ctx->line_off.p + base, n, (uint64_t*)ctx->fwd.p + base, nullptr));
    }
    if (ctx->counting && ctx->phase == PHASE_FEED && n) {
        k_count_slots<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t*)ctx->slot.p, base, n, ctx->tab.counts);
        TRY(check_launch(ctx, "k_count_slots"));
        ctx->hist_valid = false;
    }
    if (ctx->streaming && ctx->params.method == SIDGPU_METHOD_LOCAL) {
        const uint32_t n_entries = ctx->h_ctl->n_entries;
        TRY(classify_entries(ctx, ctx->classified, n_entries));
        ctx->classified = n_entries;
    }
    if (n_sites_out) *n_sites_out = n;
    return SIDGPU_OK;
}

int sidgpu_feed_rows(sidgpu_ctx* ctx, const char* d_text, size_t text_len, size_t range_begin, size_t range_end, char* d_out, size_t out_cap,
                     uint64_t* bytes_out, uint64_t* rows_out, uint64_t* n_sites_out) {
    Range nvtx_range("sidgpu_feed_rows");
    if (!ctx || (out_cap && !d_out)) return SIDGPU_EINVAL;
    if (ctx->phase != PHASE_FEED) return ctx->fail(SIDGPU_ESTATE, "sidgpu_feed_rows outside a session");
    if (!(ctx->streaming && ctx->params.method == SIDGPU_METHOD_LOCAL))
        return ctx->fail(SIDGPU_EINVAL, "sidgpu_feed_rows is for `local` sessions without -R (rows need no genome-wide step); use sidgpu_feed + sidgpu_emit_csv");
    CK(cudaSetDevice(ctx->device));
    uint64_t bytes = 0, rows = 0, n = 0;
    const int rc = run_tok2_rows(ctx, d_text, text_len, range_begin, range_end, d_out, out_cap, &bytes, &rows, &n);
    if (rc == SIDGPU_ECAPACITY && bytes_out) *bytes_out = bytes;       // the size the caller has to offer
    if (rc != SIDGPU_OK) return rc;
    ctx->last_text = d_text;
    ctx->last_text_len = text_len;
    ctx->chunk_begin = 0;
    ctx->chunk_sites = 0;
    ctx->n_sites_total = 0;                 // no site store: rows of this chunk exist only in d_out
    if (bytes_out) *bytes_out = bytes;
    if (rows_out) *rows_out = rows;
    if (n_sites_out) *n_sites_out = n;
    return SIDGPU_OK;
}

int sidgpu_finish(sidgpu_ctx* ctx) {
    Range nvtx_range("sidgpu_finish");
    if (!ctx) return SIDGPU_EINVAL;
    if (ctx->phase != PHASE_FEED) return ctx->fail(SIDGPU_ESTATE, "sidgpu_finish outside a session");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->streaming) {
        if (!ctx->global_hist) TRY(build_histogram(ctx, 4));            // call.cpp:66-70,149-153,224-229,296-301
        if (ctx->params.fit_given) {
            ctx->fit.pi = ctx->params.fit_pi;
            ctx->fit.eps = ctx->params.fit_eps;
            ctx->fit.converged = 1;
            for (int i = 0; i < 4; ++i) ctx->fit_nd[i] = ctx->params.fit_nd[i];
        } else {
            TRY(run_fit(ctx, ctx->fit_nd, &ctx->fit));
        }
        ctx->fit_done = true;
        TRY(sync_ctl(ctx));
        const uint32_t n_entries = ctx->h_ctl->n_entries;
        const int m = ctx->params.method;
        if (m == SIDGPU_METHOD_LOCAL || m == SIDGPU_METHOD_QUALITY) {
            ctx->session_prior = ctx->fit.pi;                           // call.cpp:233,305
        }
        if (m == SIDGPU_METHOD_LIKELIHOOD_RATIO) {
            const uint32_t nu = (uint32_t)ctx->n_unique;
            TRY(ensure(ctx, ctx->p_hom, (size_t)std::max(nu, 1u) * 8));
            TRY(ensure(ctx, ctx->p_het, (size_t)std::max(nu, 1u) * 8));
            TRY(ensure(ctx, ctx->adj_hom, (size_t)std::max(nu, 1u) * 8));
            TRY(ensure(ctx, ctx->adj_het, (size_t)std::max(nu, 1u) * 8));
            if (nu) {
                k_lr_pvalues<<<(nu + 255) / 256, 256, 0, ctx->stream>>>((const unsigned long long*)ctx->u_profile.p, nu,
                                                                        host_lynch_consts(ctx->fit_nd, ctx->fit.eps),
                                                                        ctx->params.estimate_prior, ctx->fit.pi,
                                                                        (double*)ctx->p_hom.p, (double*)ctx->p_het.p);
                TRY(check_launch(ctx, "k_lr_pvalues"));
                TRY(bh_adjust(ctx, (const double*)ctx->p_hom.p, nu, (double*)ctx->adj_hom.p));
                TRY(bh_adjust(ctx, (const double*)ctx->p_het.p, nu, (double*)ctx->adj_het.p));
            }
        }
        if (m != SIDGPU_METHOD_QUALITY) {
            TRY(classify_entries(ctx, 0, n_entries));
            ctx->classified = n_entries;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->phase = PHASE_FINISHED;
    return SIDGPU_OK;
}

int sidgpu_finish_global(sidgpu_ctx* ctx, const uint64_t* h_profiles_sorted, uint64_t n_global) {
    Range nvtx_range("sidgpu_finish_global");
    if (!ctx || (n_global && !h_profiles_sorted)) return SIDGPU_EINVAL;
    if (ctx->phase != PHASE_FEED) return ctx->fail(SIDGPU_ESTATE, "sidgpu_finish_global outside a session");
    if (ctx->params.method != SIDGPU_METHOD_LIKELIHOOD_RATIO) return ctx->fail(SIDGPU_EINVAL, "sidgpu_finish_global is for likelihood_ratio sessions");
    if (!ctx->params.fit_given) return ctx->fail(SIDGPU_ESTATE, "sidgpu_finish_global needs sidgpu_set_fit first");
    if (n_global > 0x7FFFFFFFull) return ctx->fail(SIDGPU_EINVAL, "too many unique profiles");
    CK(cudaSetDevice(ctx->device));
    ctx->fit.pi = ctx->params.fit_pi;
    ctx->fit.eps = ctx->params.fit_eps;
    ctx->fit.converged = 1;
    for (int i = 0; i < 4; ++i) ctx->fit_nd[i] = ctx->params.fit_nd[i];
    ctx->fit_done = true;
    TRY(sync_ctl(ctx));
    const uint32_t n_entries = ctx->h_ctl->n_entries;
    const uint32_t nu = (uint32_t)n_global;
    TRY(ensure(ctx, ctx->u_profile, (size_t)std::max(nu, 1u) * 8));
    TRY(ensure(ctx, ctx->p_hom, (size_t)std::max(nu, 1u) * 8));
    TRY(ensure(ctx, ctx->p_het, (size_t)std::max(nu, 1u) * 8));
    TRY(ensure(ctx, ctx->adj_hom, (size_t)std::max(nu, 1u) * 8));
    TRY(ensure(ctx, ctx->adj_het, (size_t)std::max(nu, 1u) * 8));
    TRY(ensure(ctx, ctx->entry_to_unique, (size_t)std::max<uint32_t>(n_entries, 1) * 4));
    if (nu) {
        CK(cudaMemcpyAsync(ctx->u_profile.p, h_profiles_sorted, (size_t)nu * 8, cudaMemcpyHostToDevice, ctx->stream));
        k_lr_pvalues<<<(nu + 255) / 256, 256, 0, ctx->stream>>>((const unsigned long long*)ctx->u_profile.p, nu,
                                                                host_lynch_consts(ctx->fit_nd, ctx->fit.eps),
                                                                ctx->params.estimate_prior, ctx->fit.pi,
                                                                (double*)ctx->p_hom.p, (double*)ctx->p_het.p);
        TRY(check_launch(ctx, "k_lr_pvalues"));
        TRY(bh_adjust(ctx, (const double*)ctx->p_hom.p, nu, (double*)ctx->adj_hom.p));
        TRY(bh_adjust(ctx, (const double*)ctx->p_het.p, nu, (double*)ctx->adj_het.p));
    }
    if (n_entries) {
        k_map_entries<<<(n_entries + 255) / 256, 256, 0, ctx->stream>>>(ctx->tab, n_entries, (const unsigned long long*)ctx->u_profile.p, nu,
                                                                         (uint32_t*)ctx->entry_to_unique.p);
        TRY(check_launch(ctx, "k_map_entries"));
        TRY(classify_entries(ctx, 0, n_entries));
        ctx->classified = n_entries;
    }
    ctx->n_unique = nu;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->phase = PHASE_FINISHED;
    return SIDGPU_OK;
}

int sidgpu_emit_csv(sidgpu_ctx* ctx, uint64_t site_begin, uint64_t n_sites, char* d_out, size_t out_cap, uint64_t* bytes_out, uint64_t* rows_out) {
    Range nvtx_range("sidgpu_emit_csv");
    if (!ctx) return SIDGPU_EINVAL;
    if (ctx->phase == PHASE_IDLE) return ctx->fail(SIDGPU_ESTATE, "sidgpu_emit_csv outside a session");
    if (!ctx->streaming && ctx->phase != PHASE_FINISHED) return ctx->fail(SIDGPU_ESTATE, "this method needs sidgpu_finish before rows can be emitted");
    if (site_begin + n_sites > ctx->n_sites_total) return ctx->fail(SIDGPU_EINVAL, "sites [%llu,+%llu) not in the store (%llu)", (unsigned long long)site_begin, (unsigned long long)n_sites, (unsigned long long)ctx->n_sites_total);
    CK(cudaSetDevice(ctx->device));
    if (bytes_out) *bytes_out = 0;
    if (rows_out) *rows_out = 0;
    if (n_sites == 0) return SIDGPU_OK;
    const bool is_quality = ctx->params.method == SIDGPU_METHOD_QUALITY;
    if (is_quality) TRY(quality_rows(ctx, site_begin, n_sites));
    const uint32_t n_tiles = (uint32_t)((n_sites + CSV_TILE - 1) / CSV_TILE);
    TRY(ensure(ctx, ctx->csv_status, (size_t)n_tiles * 8));
    CK(cudaMemsetAsync(ctx->csv_status.p, 0, (size_t)n_tiles * 8, ctx->stream));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::csv_ticket), 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::csv_bytes), 0, 2 * sizeof(unsigned long long), ctx->stream));
    CsvParams p {};
    p.site_begin = site_begin;
    p.n_sites = n_sites;
    p.order = (const uint32_t*)ctx->order.p;
    p.pos = (const int32_t*)ctx->pos.p;
    p.slot = (const uint32_t*)ctx->slot.p;
    p.name_ref = (const uint32_t*)ctx->name_ref.p;
    p.site_suffix = is_quality ? (const char*)ctx->site_suffix.p : nullptr;
    p.table = ctx->tab;
    p.pool = ctx->names.pool;
    p.out = d_out;
    p.out_cap = out_cap;
    p.ticket = ctl_field(ctx, &Control::csv_ticket);
    p.status = (unsigned long long*)ctx->csv_status.p;
    p.bytes_out = ctl_field(ctx, &Control::csv_bytes);
    p.rows_out = ctl_field(ctx, &Control::csv_rows);
    p.error = ctl_field(ctx, &Control::error);
    if (!is_quality) CK(cudaMemsetAsync(ctl_field(ctx, &Control::error), 0xFF, sizeof(unsigned long long), ctx->stream));
    p.n_tiles = n_tiles;
    int csv_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&csv_per_sm, k_csv, CSV_THREADS, CSV_STAGE) != cudaSuccess || csv_per_sm < 1) csv_per_sm = 1;
    const unsigned grid = (unsigned)std::min<uint64_t>((n_tiles + CSV_WARPS - 1) / CSV_WARPS, (uint64_t)ctx->sm_count * csv_per_sm);
    {
        ProfScope prof(ctx, PROF_CSV);
        k_csv<<<grid, CSV_THREADS, CSV_STAGE, ctx->stream>>>(p);
    }
    TRY(check_launch(ctx, "k_csv"));
    TRY(sync_ctl(ctx));
    if (ctx->h_ctl->error != ~0ull) {
        const int st = (int)(ctx->h_ctl->error & 7);
        const int code = st == LINE_MISSING_MAPQ ? SIDGPU_EMISSING_MAPQ : st == LINE_QUAL_SHORT ? SIDGPU_EQUAL_SHORT
                         : st == LINE_MALFORMED ? SIDGPU_EMALFORMED : SIDGPU_EINTERNAL;
        return ctx->fail(code, "%s (line starting at byte %llu)", status_text(st), ctx->h_ctl->error >> 3);
    }
    if (bytes_out) *bytes_out = ctx->h_ctl->csv_bytes;
    if (rows_out) *rows_out = ctx->h_ctl->csv_rows;
    if (ctx->h_ctl->csv_bytes > out_cap) return ctx->fail(SIDGPU_ECAPACITY, "CSV needs %llu bytes, buffer has %zu", ctx->h_ctl->csv_bytes, out_cap);
    return SIDGPU_OK;
}

int sidgpu_emit_records(sidgpu_ctx* ctx, uint64_t site_begin, uint64_t n_sites, uint8_t* d_label, char* d_gt, double* d_hom, double* d_het) {
    Range nvtx_range("sidgpu_emit_records");
    if (!ctx) return SIDGPU_EINVAL;
    if (ctx->phase == PHASE_IDLE) return ctx->fail(SIDGPU_ESTATE, "sidgpu_emit_records outside a session");
    if (!ctx->streaming && ctx->phase != PHASE_FINISHED) return ctx->fail(SIDGPU_ESTATE, "this method needs sidgpu_finish first");
    if (site_begin + n_sites > ctx->n_sites_total) return ctx->fail(SIDGPU_EINVAL, "site range not in the store");
    if (n_sites == 0) return SIDGPU_OK;
    CK(cudaSetDevice(ctx->device));
    if (ctx->params.method == SIDGPU_METHOD_QUALITY) {
        // per-site results (call.cpp:309-370): the quality kernel writes them beside the row text
        TRY(quality_rows(ctx, site_begin, n_sites, d_label, d_gt, d_hom, d_het));
        TRY(sync_ctl(ctx));
        if (ctx->h_ctl->error != ~0ull) {
            const int st = (int)(ctx->h_ctl->error & 7);
            const int code = st == LINE_MISSING_MAPQ ? SIDGPU_EMISSING_MAPQ : st == LINE_QUAL_SHORT ? SIDGPU_EQUAL_SHORT
                             : st == LINE_MALFORMED ? SIDGPU_EMALFORMED : SIDGPU_EINTERNAL;
            return ctx->fail(code, "%s (line starting at byte %llu)", status_text(st), ctx->h_ctl->error >> 3);
        }
        return SIDGPU_OK;
    }
    RecordParams p {site_begin, n_sites, (const uint32_t*)ctx->order.p, (const uint32_t*)ctx->slot.p, ctx->tab, d_label, d_gt, d_hom, d_het,
                    nullptr, nullptr, nullptr, nullptr};
    k_records<<<(unsigned)((n_sites + 255) / 256), 256, 0, ctx->stream>>>(p);
    TRY(check_launch(ctx, "k_records"));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

int sidgpu_emit_columns(sidgpu_ctx* ctx, uint64_t site_begin, uint64_t n_sites, const sidgpu_columns* cols) {
    if (!ctx || !cols) return SIDGPU_EINVAL;
    if (ctx->phase == PHASE_IDLE) return ctx->fail(SIDGPU_ESTATE, "sidgpu_emit_columns outside a session");
    if (!ctx->streaming && ctx->phase != PHASE_FINISHED) return ctx->fail(SIDGPU_ESTATE, "this method needs sidgpu_finish first");
    if (site_begin + n_sites > ctx->n_sites_total) return ctx->fail(SIDGPU_EINVAL, "site range not in the store");
    if (n_sites == 0) return SIDGPU_OK;
    if (cols->d_fwd && !ctx->params.want_strands) return ctx->fail(SIDGPU_EINVAL, "d_fwd needs a session begun with want_strands");
    CK(cudaSetDevice(ctx->device));
    const bool is_quality = ctx->params.method == SIDGPU_METHOD_QUALITY;
    if (is_quality && cols->d_profile) {
        // quality sessions keep the counts per site, not in the table
        k_gather<<<(unsigned)((n_sites + 255) / 256), 256, 0, ctx->stream>>>((uint64_t*)cols->d_profile, (const uint64_t*)ctx->profile.p,
                                                                            (const uint32_t*)ctx->order.p + site_begin, n_sites);
        TRY(check_launch(ctx, "k_gather"));
    }
    // quality: the call is per site, not per profile: k_quality writes label / genotype / confidences itself
    if (is_quality) TRY(sidgpu_emit_records(ctx, site_begin, n_sites, cols->d_label, cols->d_gt, cols->d_hom_conf, cols->d_het_conf));
    RecordParams p {site_begin, n_sites, (const uint32_t*)ctx->order.p, (const uint32_t*)ctx->slot.p, ctx->tab,
                    is_quality ? nullptr : cols->d_label, is_quality ? nullptr : cols->d_gt, is_quality ? nullptr : cols->d_hom_conf,
                    is_quality ? nullptr : cols->d_het_conf,
                    (const int32_t*)ctx->pos.p, (const uint32_t*)ctx->name_ref.p, cols->d_pos, cols->d_name_ref,
                    (const unsigned long long*)ctx->fwd.p, is_quality ? nullptr : (unsigned long long*)cols->d_profile,
                    (unsigned long long*)cols->d_fwd};
    k_records<<<(unsigned)((n_sites + 255) / 256), 256, 0, ctx->stream>>>(p);
    TRY(check_launch(ctx, "k_records"));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

int sidgpu_names(sidgpu_ctx* ctx, const char** d_names, uint64_t* names_bytes) {
    if (!ctx || !d_names || !names_bytes) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    TRY(sync_ctl(ctx));
    *d_names = ctx->names.pool;
    *names_bytes = ctx->h_ctl->name_cursor;
    return SIDGPU_OK;
}

// ------------------------------------------------------------------------------------------- K3/K4
int sidgpu_histogram(sidgpu_ctx* ctx, uint32_t min_coverage, sidgpu_unique_view* out) {
    Range nvtx_range("sidgpu_histogram");
    if (!ctx || !out) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    TRY(build_histogram(ctx, min_coverage));
    out->n_unique = ctx->n_unique;
    out->d_profile = (const uint64_t*)ctx->u_profile.p;
    out->d_count = (const uint64_t*)ctx->u_count.p;
    for (int i = 0; i < 4; ++i) out->nd[i] = ctx->fit_nd[i];
    for (int i = 0; i < 5; ++i) out->nd_sums[i] = ctx->n_unique ? ctx->h_ctl->nd_acc[i] : 0;
    return SIDGPU_OK;
}

static int count_unique_impl(sidgpu_ctx* ctx, const uint64_t* d_profiles, const uint64_t* d_counts, uint64_t n, uint32_t min_coverage,
                             sidgpu_unique_view* out) {
    if (!ctx || !out || (n && !d_profiles)) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    ctx->phase = PHASE_IDLE;
    for (;;) {
        TRY(reset_table(ctx));
        if (n) {
            k_insert_profiles<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->tab, d_profiles, d_counts, n);
            TRY(check_launch(ctx, "k_insert_profiles"));
        }
        TRY(sync_ctl(ctx));
        if (ctx->h_ctl->table_overflow || (uint64_t)ctx->h_ctl->n_entries * 2 > ctx->tab.cap) {
            TRY(grow_table(ctx, 2, 0));
            continue;
        }
        break;
    }
    return sidgpu_histogram(ctx, min_coverage, out);
}

int sidgpu_count_unique(sidgpu_ctx* ctx, const uint64_t* d_profiles, uint64_t n, uint32_t min_coverage, sidgpu_unique_view* out) {
    return count_unique_impl(ctx, d_profiles, nullptr, n, min_coverage, out);
}

int sidgpu_count_unique_weighted(sidgpu_ctx* ctx, const uint64_t* d_profiles, const uint64_t* d_counts, uint64_t n,
                                 uint32_t min_coverage, sidgpu_unique_view* out) {
    if (n && !d_counts) return SIDGPU_EINVAL;
    return count_unique_impl(ctx, d_profiles, d_counts, n, min_coverage, out);
}

int sidgpu_set_global_histogram(sidgpu_ctx* ctx, const uint64_t* d_profiles, const uint64_t* d_counts, uint64_t n) {
    Range nvtx_range("sidgpu_set_global_histogram");
    if (!ctx || (n && (!d_profiles || !d_counts))) return SIDGPU_EINVAL;
    if (ctx->phase != PHASE_FEED || ctx->streaming) return ctx->fail(SIDGPU_ESTATE, "sidgpu_set_global_histogram needs an open session with a genome-wide step");
    if (n > 0x3FFFFFFFull) return ctx->fail(SIDGPU_EINVAL, "too many histogram entries");
    CK(cudaSetDevice(ctx->device));
    int log2cap = 12;
    while (((uint64_t)1 << log2cap) < 2 * n + 2) ++log2cap;
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::g_n_entries), 0, 3 * sizeof(unsigned int), ctx->stream));
    if (log2cap > ctx->g_log2) {
        if (ctx->g_log2) free_table(ctx->g_tab);
        ctx->g_log2 = 0;
        TRY(alloc_table(ctx, log2cap, ctx->g_tab, true));
        ctx->g_log2 = log2cap;
    } else {
        // the table of the previous exchange: only the slots that can be probed need clearing
        log2cap = ctx->g_log2;
        CK(cudaMemsetAsync(ctx->g_tab.keys, 0xFF, ((size_t)ctx->g_tab.cap + 1) * 8, ctx->stream));
        CK(cudaMemsetAsync(ctx->g_tab.counts, 0, ((size_t)ctx->g_tab.cap + 1) * 8, ctx->stream));
    }
    const TableView& g = ctx->g_tab;
    int rc = SIDGPU_OK;
    if (n) {
        k_insert_profiles<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(g, d_profiles, d_counts, n);
        rc = check_launch(ctx, "k_insert_profiles");
    }
    if (rc == SIDGPU_OK) rc = sync_ctl(ctx);
    if (rc == SIDGPU_OK && ctx->h_ctl->g_overflow) rc = ctx->fail(SIDGPU_EINTERNAL, "merged histogram table overflow");
    // entries with count 0 (padding of the all-gather) and coverage < 4 (call.cpp:66-70) are left out by the selection
    if (rc == SIDGPU_OK) rc = build_histogram_of(ctx, g, ctx->h_ctl->g_n_entries, 4, ctx->g_e2u);
    if (rc == SIDGPU_OK) {
        const uint32_t n_entries = ctx->h_ctl->n_entries;
        rc = ensure(ctx, ctx->entry_to_unique, (size_t)std::max<uint32_t>(n_entries, 1) * 4);
        if (rc == SIDGPU_OK && n_entries) {
            // this rank's own profiles -> their rows in the merged list
            k_map_entries<<<(n_entries + 255) / 256, 256, 0, ctx->stream>>>(ctx->tab, n_entries, (const unsigned long long*)ctx->u_profile.p,
                                                                             (uint32_t)ctx->n_unique, (uint32_t*)ctx->entry_to_unique.p);
            rc = check_launch(ctx, "k_map_entries");
        }
    }
    if (rc == SIDGPU_OK) ctx->global_hist = true;
    return rc;
}

int sidgpu_set_fit(sidgpu_ctx* ctx, double pi, double eps, const double nd[4]) {
    if (!ctx || !nd) return SIDGPU_EINVAL;
    if (ctx->phase != PHASE_FEED) return ctx->fail(SIDGPU_ESTATE, "sidgpu_set_fit outside a session");
    ctx->params.fit_given = 1;
    ctx->params.fit_pi = pi;
    ctx->params.fit_eps = eps;
    for (int i = 0; i < 4; ++i) ctx->params.fit_nd[i] = nd[i];
    return SIDGPU_OK;
}

int sidgpu_lynch_objective_partial(sidgpu_ctx* ctx, const double nd[4], double pi, double eps, double* d_out) {
    if (!ctx || !nd || !d_out) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (pi < 0 || pi > 1 || eps < 0 || eps > 1) {
        const double big = std::numeric_limits<double>::max();
        CK(cudaMemcpyAsync(d_out, &big, sizeof big, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return SIDGPU_OK;
    }
    return launch_objective(ctx, nd, pi, eps, d_out);
}

int sidgpu_lynch_objective(sidgpu_ctx* ctx, const double nd[4], double pi, double eps, double* value) {
    if (!ctx || !nd || !value) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    return objective_value(ctx, nd, pi, eps, value);
}

int sidgpu_lynch_fit(sidgpu_ctx* ctx, const double nd[4], sidgpu_fit* out) {
    Range nvtx_range("sidgpu_lynch_fit");
    if (!ctx || !nd || !out) return SIDGPU_EINVAL;
    CK(cudaSetDevice(ctx->device));
    return run_fit(ctx, nd, out);
}

int sidgpu_session_fit(sidgpu_ctx* ctx, sidgpu_fit* out, double nd[4], uint64_t* n_unique) {
    if (!ctx) return SIDGPU_EINVAL;
    if (!ctx->fit_done) return ctx->fail(SIDGPU_ESTATE, "no fit in this session");
    if (out) *out = ctx->fit;
    if (nd) for (int i = 0; i < 4; ++i) nd[i] = ctx->fit_nd[i];
    if (n_unique) *n_unique = ctx->n_unique;
    return SIDGPU_OK;
}

int sidgpu_bh_adjust(sidgpu_ctx* ctx, const double* d_p, uint64_t n, double* d_adjusted) {
    if (!ctx || (n && (!d_p || !d_adjusted))) return SIDGPU_EINVAL;
    if (n > 0x7FFFFFFFull) return ctx->fail(SIDGPU_EINVAL, "too many p-values");
    CK(cudaSetDevice(ctx->device));
    TRY(bh_adjust(ctx, d_p, (uint32_t)n, d_adjusted));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

int sidgpu_format_g(sidgpu_ctx* ctx, const double* d_values, uint64_t n, char* d_out16) {
    if (!ctx) return SIDGPU_EINVAL;
    if (n == 0) return SIDGPU_OK;
    CK(cudaSetDevice(ctx->device));
    k_format_g<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_values, n, d_out16);
    TRY(check_launch(ctx, "k_format_g"));
    CK(cudaStreamSynchronize(ctx->stream));
    return SIDGPU_OK;
}

}  // extern "C"

#include "host_path.inl"
#include "bgzf_path.inl"
#include "host_io.inl"
