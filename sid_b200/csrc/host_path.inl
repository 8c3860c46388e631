// sidgpu_call_host: the host-buffer path behind the `sid` binary and the call.hpp wrappers
// (replaces the ifstream + readFile + vector<OutputRecord> + operator<< chain of sid.cpp:85-105).
// Text goes to the device in line-aligned chunks through two device buffers (copy stream) while
// the previous chunk is processed (compute stream); CSV comes back through two device buffers
// (copy-out stream).  With pinned h_text / h_csv the three overlap; pageable memory works too.

namespace {

// The staging buffers and events live in the ctx (sidgpu_ctx::hp_*) and are reused from call to call.
struct HostPath {
    sidgpu_ctx* ctx;
    DevBuf* text;
    DevBuf* csv;
    cudaEvent_t* ev_in;
    cudaEvent_t* ev_out;
};

// End (exclusive) of the chunk starting at `start`: the last line end within max_chunk, or the
// end of the first line when a single line is longer than that.
size_t chunk_end(const char* t, size_t len, size_t start, size_t max_chunk) {
    size_t end = std::min(len, start + max_chunk);
    if (end == len) return end;
    const void* nl = memrchr(t + start, '\n', end - start);
    if (nl) return (size_t)((const char*)nl - t) + 1;
    const void* fw = memchr(t + end, '\n', len - end);
    return fw ? (size_t)((const char*)fw - t) + 1 : len;
}

// State of one pass of text in / rows out over the ctx's staging buffers.
struct HostIo {
    sidgpu_ctx* ctx;
    HostPath hp;
    const char* h_text = nullptr;
    size_t text_len = 0;
    char* h_csv = nullptr;
    size_t csv_cap = 0;
    uint64_t total_sites = 0, total_rows = 0, out_off = 0;
    bool out_overflow = false;

    explicit HostIo(sidgpu_ctx* c) : ctx(c), hp {c, c->hp_text, c->hp_csv, c->hp_ev_in, c->hp_ev_out} {}

    int init() {
        for (int i = 0; i < 2; ++i) {
            if (!hp.ev_in[i]) CK(cudaEventCreateWithFlags(&hp.ev_in[i], cudaEventDisableTiming));
            if (!hp.ev_out[i]) CK(cudaEventCreateWithFlags(&hp.ev_out[i], cudaEventDisableTiming));
        }
        return SIDGPU_OK;
    }
    int upload(int b, size_t start, size_t end) {
        const size_t n = end - start;
        TRY(ensure(ctx, hp.text[b], ((n + 15) & ~(size_t)15) + 16));
        CK(cudaMemcpyAsync(hp.text[b].p, h_text + start, n, cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaEventRecord(hp.ev_in[b], ctx->copy_in));
        return SIDGPU_OK;
    }
    int emit_range(int b, uint64_t site_begin, uint64_t count) {
        if (count == 0) return SIDGPU_OK;
        CK(cudaEventSynchronize(hp.ev_out[b]));                     // the previous D2H out of csv[b] is done
        uint64_t bytes = 0, rows = 0;
        size_t want = std::max<size_t>(hp.csv[b].cap, (size_t)count * 48 + 4096);
        for (;;) {
            TRY(ensure(ctx, hp.csv[b], want));
            const int rc = sidgpu_emit_csv(ctx, site_begin, count, (char*)hp.csv[b].p, hp.csv[b].cap, &bytes, &rows);
            if (rc == SIDGPU_ECAPACITY && bytes + 4096 > want) { want = (size_t)bytes + 4096; continue; }
            if (rc != SIDGPU_OK) return rc;
            break;
        }
        if (out_off + bytes <= csv_cap && h_csv) {
            CK(cudaMemcpyAsync(h_csv + out_off, hp.csv[b].p, bytes, cudaMemcpyDeviceToHost, ctx->copy_out));
            CK(cudaEventRecord(hp.ev_out[b], ctx->copy_out));
        } else {
            out_overflow = true;
        }
        out_off += bytes;
        total_rows += rows;
        return SIDGPU_OK;
    }
    // the fused form of feed + emit_range for `local` sessions: K1 writes the rows of the chunk itself
    int rows_chunk(int b, const char* d_text, size_t len, uint64_t* n_sites) {
        CK(cudaEventSynchronize(hp.ev_out[b]));                     // the previous D2H out of csv[b] is done
        uint64_t bytes = 0, rows = 0;
        size_t want = std::max<size_t>(hp.csv[b].cap, len + len / 2 + 4096);
        for (;;) {
            TRY(ensure(ctx, hp.csv[b], want));
            const int rc = sidgpu_feed_rows(ctx, d_text, len, 0, len, (char*)hp.csv[b].p, hp.csv[b].cap, &bytes, &rows, n_sites);
            if (rc == SIDGPU_ECAPACITY && bytes + 4096 > want) { want = (size_t)bytes + 4096; continue; }
            if (rc != SIDGPU_OK) return rc;
            break;
        }
        if (out_off + bytes <= csv_cap && h_csv) {
            CK(cudaMemcpyAsync(h_csv + out_off, hp.csv[b].p, bytes, cudaMemcpyDeviceToHost, ctx->copy_out));
            CK(cudaEventRecord(hp.ev_out[b], ctx->copy_out));
        } else {
            out_overflow = true;
        }
        out_off += bytes;
        total_rows += rows;
        return SIDGPU_OK;
    }
    // one streamed pass over the text; `emit` says whether rows are produced chunk by chunk
    int pass(bool emit) {
        if (text_len == 0) return SIDGPU_OK;
        const size_t max_chunk = ctx->max_chunk;
        size_t start = 0, end = chunk_end(h_text, text_len, 0, max_chunk);
        TRY(upload(0, start, end));
        for (int i = 0; start < text_len; ++i) {
            const int b = i & 1;
            const size_t next_start = end;
            size_t next_end = next_start;
            if (next_start < text_len) {
                next_end = chunk_end(h_text, text_len, next_start, max_chunk);
                TRY(upload(b ^ 1, next_start, next_end));           // overlaps with the work on chunk i
            }
            CK(cudaStreamWaitEvent(ctx->stream, hp.ev_in[b], 0));
            uint64_t n = 0;
            // K1 can write the rows of a `local` chunk itself (sidgpu_feed_rows); measured slower than site store + K6
            // (2.93 vs 1.97 ms per 20 M sites, profiles/README.md), so it is opt-in: SIDGPU_FUSED_ROWS=1
            static const bool fused = getenv("SIDGPU_FUSED_ROWS") && atoi(getenv("SIDGPU_FUSED_ROWS")) != 0;
            if (emit && fused && ctx->streaming && ctx->params.method == SIDGPU_METHOD_LOCAL && ctx->phase == PHASE_FEED) {
                TRY(rows_chunk(b, (const char*)hp.text[b].p, end - start, &n));       // one kernel from text to rows
                total_sites += n;
            } else {
                TRY(sidgpu_feed(ctx, (const char*)hp.text[b].p, end - start, 0, end - start, &n));
                total_sites += n;
                if (emit) TRY(emit_range(b, 0, n));
            }
            start = next_start;
            end = next_end;
        }
        return SIDGPU_OK;
    }
    // the rows of every stored site, in blocks
    int emit_store() {
        const uint64_t step = (uint64_t)8 << 20;                    // sites per emitted block
        int b = 0;
        for (uint64_t s = 0; s < ctx->n_sites_total; s += step, b ^= 1) {
            TRY(emit_range(b, s, std::min<uint64_t>(step, ctx->n_sites_total - s)));
        }
        return SIDGPU_OK;
    }
    int drain() {
        CK(cudaStreamSynchronize(ctx->copy_out));
        CK(cudaStreamSynchronize(ctx->stream));
        return SIDGPU_OK;
    }
};

}  // namespace

extern "C" int sidgpu_call_host(sidgpu_ctx* ctx, const sidgpu_params* params, const char* h_text, size_t text_len,
                                char* h_csv, size_t csv_cap, uint64_t* csv_bytes, uint64_t* n_sites, uint64_t* n_rows) {
    if (!ctx || !params || (text_len && !h_text)) return SIDGPU_EINVAL;
    Range nvtx_range("sidgpu_call_host");
    CK(cudaSetDevice(ctx->device));
    HostIo io(ctx);
    TRY(io.init());
    io.h_text = h_text;
    io.text_len = text_len;
    io.h_csv = h_csv;
    io.csv_cap = csv_cap;
    TRY(sidgpu_begin(ctx, params));
    if (ctx->streaming) {
        TRY(io.pass(true));
    } else {
        TRY(io.pass(false));
        TRY(sidgpu_finish(ctx));
        if (params->method == SIDGPU_METHOD_QUALITY) {
            io.total_sites = 0;
            TRY(io.pass(true));                                     // second pass with the fitted prior
        } else {
            TRY(io.emit_store());
        }
    }
    TRY(io.drain());
    if (csv_bytes) *csv_bytes = io.out_off;
    if (n_sites) *n_sites = io.total_sites;
    if (n_rows) *n_rows = io.total_rows;
    if (io.out_overflow) return ctx->fail(SIDGPU_ECAPACITY, "CSV needs %llu bytes, buffer has %zu", (unsigned long long)io.out_off, csv_cap);
    return SIDGPU_OK;
}

// The two halves of the above for sessions that keep their sites (bayes, likelihood_ratio, local -R) and are
// driven from outside between them: several shards of one genome that share a fit (host/sid_host.cpp).
extern "C" int sidgpu_feed_host(sidgpu_ctx* ctx, const char* h_text, size_t text_len, uint64_t* n_sites) {
    if (!ctx || (text_len && !h_text)) return SIDGPU_EINVAL;
    if (ctx->phase != PHASE_FEED) return ctx->fail(SIDGPU_ESTATE, "sidgpu_feed_host outside a session");
    if (ctx->streaming) return ctx->fail(SIDGPU_EINVAL, "sidgpu_feed_host is for sessions with a genome-wide fit; use sidgpu_call_host");
    CK(cudaSetDevice(ctx->device));
    HostIo io(ctx);
    TRY(io.init());
    io.h_text = h_text;
    io.text_len = text_len;
    TRY(io.pass(false));
    TRY(io.drain());
    if (n_sites) *n_sites = io.total_sites;
    return SIDGPU_OK;
}

// Text in, rows out chunk by chunk on an open session: a streaming session (local / quality without -R), or the
// second pass of `quality -R` after sidgpu_finish (the first pass, sidgpu_feed_host, only built the histogram).
extern "C" int sidgpu_stream_host(sidgpu_ctx* ctx, const char* h_text, size_t text_len, char* h_csv, size_t csv_cap,
                                  uint64_t* csv_bytes, uint64_t* n_sites, uint64_t* n_rows) {
    if (!ctx || (text_len && !h_text)) return SIDGPU_EINVAL;
    const bool second_pass = ctx->phase == PHASE_FINISHED && ctx->params.method == SIDGPU_METHOD_QUALITY;
    if (!((ctx->streaming && ctx->phase == PHASE_FEED) || second_pass))
        return ctx->fail(SIDGPU_ESTATE, "sidgpu_stream_host needs a streaming session or the second pass of quality -R");
    CK(cudaSetDevice(ctx->device));
    HostIo io(ctx);
    TRY(io.init());
    io.h_text = h_text;
    io.text_len = text_len;
    io.h_csv = h_csv;
    io.csv_cap = csv_cap;
    TRY(io.pass(true));
    TRY(io.drain());
    if (csv_bytes) *csv_bytes = io.out_off;
    if (n_sites) *n_sites = io.total_sites;
    if (n_rows) *n_rows = io.total_rows;
    if (io.out_overflow) return ctx->fail(SIDGPU_ECAPACITY, "CSV needs %llu bytes, buffer has %zu", (unsigned long long)io.out_off, csv_cap);
    return SIDGPU_OK;
}

extern "C" int sidgpu_emit_host(sidgpu_ctx* ctx, char* h_csv, size_t csv_cap, uint64_t* csv_bytes, uint64_t* n_rows) {
    if (!ctx) return SIDGPU_EINVAL;
    if (ctx->phase != PHASE_FINISHED || ctx->streaming || ctx->params.method == SIDGPU_METHOD_QUALITY)
        return ctx->fail(SIDGPU_ESTATE, "sidgpu_emit_host needs a finished session that kept its sites (quality -R: sidgpu_stream_host)");
    CK(cudaSetDevice(ctx->device));
    HostIo io(ctx);
    TRY(io.init());
    io.h_csv = h_csv;
    io.csv_cap = csv_cap;
    TRY(io.emit_store());
    TRY(io.drain());
    if (csv_bytes) *csv_bytes = io.out_off;
    if (n_rows) *n_rows = io.total_rows;
    if (io.out_overflow) return ctx->fail(SIDGPU_ECAPACITY, "CSV needs %llu bytes, buffer has %zu", (unsigned long long)io.out_off, csv_cap);
    return SIDGPU_OK;
}
