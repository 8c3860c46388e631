// sidgpu_call_io: the streaming form of sidgpu_call_host for inputs and outputs that do not fit in host memory
// (replaces the ifstream -> readFile -> vector<OutputRecord> -> operator<< chain of sid.cpp:85-105 for files of any
// size).  Three threads share a ring of pinned text slots and a ring of pinned CSV slots:
//   reader  fills text slots through the caller's read callback and cuts them at the last line end (the rest of the
//           slot opens the next one);
//   caller  uploads slot i + 1 while the kernels work on slot i, queues the device -> host copy of the rows;
//   writer  waits for that copy and hands the rows to the caller's write callback, in file order.
// Sessions with a genome-wide step (bayes, likelihood_ratio, -R) feed every slot first (their sites stay in HBM),
// fit, and then stream the rows of the site store out the same way; `quality -R` reads the text a second time
// (rewind callback).

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

struct IoPipe {
    static constexpr int NT = 3, NC = 3;
    sidgpu_ctx* ctx;
    const sidgpu_io* io;
    size_t slot_cap = 0;
    struct TextSlot {
        char* p = nullptr;
        size_t cap = 0, len = 0;
        // BGZF input (bgzf = true): the slot holds compressed members; their table (pinned) and the bytes of text they inflate to
        sidgpu_bgzf_block* blocks = nullptr;
        size_t n_blocks = 0, text_len = 0;
        bool last = false;
    } ts[NT];
    bool bgzf = false;                      // the read callback delivers a BGZF file: members are inflated on the device
    static constexpr size_t MAX_BLOCKS = (size_t)1 << 16;
    size_t text_cap = 0;                    // bgzf: text bytes one chunk may inflate to
    BgzfChunks chunks {nullptr, 0};         // bgzf: the inflate chains of the two text buffers (bgzf_path.inl)
    struct CsvSlot { char* p = nullptr; size_t cap = 0, len = 0; cudaEvent_t ev = nullptr; } cs[NC];
    std::mutex m;
    std::condition_variable cv;
    uint64_t t_prod = 0, t_cons = 0;        // text slots filled by the reader / released by the caller
    uint64_t c_prod = 0, c_cons = 0;        // CSV slots queued by the caller / written by the writer
    bool eof = false;                       // the reader has produced its last slot
    bool stop_writer = false;
    int error = SIDGPU_OK;                  // first error of the reader or the writer
    std::string errmsg;
    std::vector<char> carry;                // the unfinished line at the end of the previous slot
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    uint64_t total_sites = 0, total_rows = 0, total_bytes = 0;
    // SIDGPU_IO_TIMING=1: where the caller's thread spends its time (seconds), printed to stderr at the end of the call
    double t_wait_text = 0, t_upload = 0, t_inflate = 0, t_feed = 0, t_rows = 0, t_queue = 0, t_init = 0;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    void set_error(int code, const std::string& msg) {
        std::lock_guard<std::mutex> g(m);
        if (error == SIDGPU_OK) { error = code; errmsg = msg; }
        cv.notify_all();
    }

    // ---- reader thread: one pass over the input
    void reader() {
        bool input_done = false, src_eof = false;
        while (!input_done) {
            uint64_t k;
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [&] { return error != SIDGPU_OK || t_prod - t_cons < (uint64_t)NT; });
                if (error != SIDGPU_OK) return;
                k = t_prod;
            }
            TextSlot& s = ts[k % NT];
            size_t have = 0;
            if (bgzf) {
                // whole members only: what the header walk does not accept opens the next slot
                if (!carry.empty()) { memcpy(s.p, carry.data(), carry.size()); have = carry.size(); carry.clear(); }
                while (have < s.cap && !src_eof) {
                    const int64_t got = io->read(io->user, s.p + have, s.cap - have);
                    if (got < 0) { set_error(SIDGPU_EINVAL, "read callback failed"); return; }
                    if (got == 0) src_eof = true;
                    have += (size_t)got;
                }
                size_t consumed = 0;
                if (sidgpu_bgzf_scan(s.p, have, s.blocks, MAX_BLOCKS, text_cap, &s.n_blocks, &consumed, &s.text_len) != SIDGPU_OK) {
                    set_error(SIDGPU_EINVAL, "not a BGZF file (bad member header)");
                    return;
                }
                if (consumed == 0 && have > 0) {
                    // a full slot always holds whole members (they are at most 64 KiB); at the end of the file it is a cut one
                    set_error(SIDGPU_EINVAL, src_eof ? "truncated BGZF member at the end of the file" : "BGZF member larger than a slot");
                    return;
                }
                carry.assign(s.p + consumed, s.p + have);
                s.len = consumed;
                input_done = src_eof && carry.empty();
                s.last = input_done;
                {
                    std::lock_guard<std::mutex> g(m);
                    ++t_prod;
                    if (input_done) eof = true;
                }
                cv.notify_all();
                continue;
            }
            for (;;) {
                if (carry.size() + 16 > s.cap || have == s.cap) {
                    // a single line longer than the slot: grow it (pinned memory; this is rare)
                    const size_t ncap = std::max(s.cap * 2, carry.size() * 2 + 4096);
                    char* np = nullptr;
                    if (cudaHostAlloc((void**)&np, ncap, cudaHostAllocDefault) != cudaSuccess) { set_error(SIDGPU_ENOMEM, "pinned text slot"); return; }
                    if (have) memcpy(np, s.p, have);
                    if (s.p) cudaFreeHost(s.p);
                    s.p = np;
                    s.cap = ncap;
                }
                if (!carry.empty()) {
                    memcpy(s.p, carry.data(), carry.size());
                    have = carry.size();
                    carry.clear();
                }
                while (have < s.cap && !input_done) {
                    const int64_t got = io->read(io->user, s.p + have, s.cap - have);
                    if (got < 0) { set_error(SIDGPU_EINVAL, "read callback failed"); return; }
                    if (got == 0) input_done = true;
                    have += (size_t)got;
                }
                if (input_done) break;
                // cut at the last line end; what follows opens the next slot
                const void* nl = memrchr(s.p, '\n', have);
                if (nl) {
                    const size_t keep = (size_t)((const char*)nl - s.p) + 1;
                    carry.assign(s.p + keep, s.p + have);
                    have = keep;
                    break;
                }
                // no line end in a full slot: grow and keep reading
            }
            s.len = have;
            {
                std::lock_guard<std::mutex> g(m);
                ++t_prod;
                if (input_done) eof = true;
            }
            cv.notify_all();
        }
    }

    // ---- writer thread
    void writer() {
        cudaSetDevice(ctx->device);
        for (;;) {
            uint64_t k;
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [&] { return c_cons < c_prod || stop_writer; });
                if (c_cons == c_prod) return;
                k = c_cons;
            }
            CsvSlot& s = cs[k % NC];
            bool ok = cudaEventSynchronize(s.ev) == cudaSuccess;
            if (!ok) set_error(SIDGPU_ECUDA, "device -> host copy of the rows failed");
            if (ok && s.len && error == SIDGPU_OK && io->write(io->user, s.p, s.len) != 0) set_error(SIDGPU_EINVAL, "write callback failed");
            {
                std::lock_guard<std::mutex> g(m);
                ++c_cons;
            }
            cv.notify_all();
        }
    }

    // waits for text slot i; returns false at the end of the input (or on an error)
    bool wait_text(uint64_t i, bool block) {
        std::unique_lock<std::mutex> g(m);
        if (block) cv.wait(g, [&] { return error != SIDGPU_OK || t_prod > i || eof; });
        return error == SIDGPU_OK && t_prod > i;
    }
    void release_text() {
        { std::lock_guard<std::mutex> g(m); ++t_cons; }
        cv.notify_all();
    }

    int upload(uint64_t i) {
        const int b = (int)(i & 1);
        TextSlot& s = ts[i % NT];
        if (bgzf) {                                                 // the compressed members and their table
            TRY(ensure(ctx, ctx->hp_comp[b], ((s.len + 15) & ~(size_t)15) + 16));
            TRY(ensure(ctx, ctx->inf_blocks[b], std::max<size_t>(1, s.n_blocks) * sizeof(sidgpu_bgzf_block)));
            if (s.len) CK(cudaMemcpyAsync(ctx->hp_comp[b].p, s.p, s.len, cudaMemcpyHostToDevice, ctx->copy_in));
            if (s.n_blocks) CK(cudaMemcpyAsync(ctx->inf_blocks[b].p, s.blocks, s.n_blocks * sizeof(sidgpu_bgzf_block), cudaMemcpyHostToDevice, ctx->copy_in));
            CK(cudaEventRecord(ev_in[b], ctx->copy_in));
            return SIDGPU_OK;
        }
        TRY(ensure(ctx, ctx->hp_text[b], ((s.len + 15) & ~(size_t)15) + 16));
        if (s.len) CK(cudaMemcpyAsync(ctx->hp_text[b].p, s.p, s.len, cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaEventRecord(ev_in[b], ctx->copy_in));
        return SIDGPU_OK;
    }

    // queues the copy of `bytes` rows in hp_csv[b] to the next pinned CSV slot
    int queue_rows(int b, uint64_t bytes, uint64_t rows) {
        uint64_t k;
        {
            std::unique_lock<std::mutex> g(m);
            cv.wait(g, [&] { return error != SIDGPU_OK || c_prod - c_cons < (uint64_t)NC; });
            if (error != SIDGPU_OK) return error;
            k = c_prod;
        }
        CsvSlot& s = cs[k % NC];
        if (bytes > s.cap) {
            if (s.p) cudaFreeHost(s.p);
            s.p = nullptr;
            s.cap = (size_t)bytes + (size_t)bytes / 4 + 4096;
            CK(cudaHostAlloc((void**)&s.p, s.cap, cudaHostAllocDefault));
        }
        s.len = (size_t)bytes;
        if (bytes) CK(cudaMemcpyAsync(s.p, ctx->hp_csv[b].p, bytes, cudaMemcpyDeviceToHost, ctx->copy_out));
        CK(cudaEventRecord(s.ev, ctx->copy_out));
        CK(cudaEventRecord(ev_out[b], ctx->copy_out));
        total_bytes += bytes;
        total_rows += rows;
        { std::lock_guard<std::mutex> g(m); ++c_prod; }
        cv.notify_all();
        return SIDGPU_OK;
    }

    // rows of the chunk (d_text, len) of a streaming session (or of the second pass of quality -R): fed at once, written
    // in pieces of at most ROW_STEP sites through the two device buffers (a chunk of inflated BGZF members is several
    // times the size of a text slot; the pinned CSV slots stay the size of a piece)
    static constexpr uint64_t ROW_STEP = (uint64_t)1 << 20;
    int rows_of_chunk(int b0, const char* d_text, size_t len, uint64_t* n_sites) {
        TRY(sidgpu_feed(ctx, d_text, len, 0, len, n_sites));
        int b = b0;
        for (uint64_t s0 = 0; s0 < *n_sites || s0 == 0; s0 += ROW_STEP, b ^= 1) {
            const uint64_t count = std::min<uint64_t>(ROW_STEP, *n_sites - s0);
            CK(cudaEventSynchronize(ev_out[b]));                    // the previous copy out of hp_csv[b] is done
            uint64_t bytes = 0, rows = 0;
            size_t want = std::max<size_t>(ctx->hp_csv[b].cap, (size_t)count * 48 + 4096);
            for (;;) {
                TRY(ensure(ctx, ctx->hp_csv[b], want));
                const int rc = sidgpu_emit_csv(ctx, s0, count, (char*)ctx->hp_csv[b].p, ctx->hp_csv[b].cap, &bytes, &rows);
                if (rc == SIDGPU_ECAPACITY && bytes + 4096 > want) { want = (size_t)bytes + 4096; continue; }
                if (rc != SIDGPU_OK) return rc;
                break;
            }
            const double q0 = now();
            const int rc = queue_rows(b, bytes, rows);
            t_queue += now() - q0;
            if (rc != SIDGPU_OK) return rc;
            if (*n_sites == 0) break;
        }
        return SIDGPU_OK;
    }

    // one pass over the input; emit: rows per chunk (streaming session / second pass), else the chunks are only fed
    int pass(bool emit) {
        {
            std::lock_guard<std::mutex> g(m);
            t_prod = t_cons = 0;
            eof = false;
            carry.clear();
        }
        uint64_t launched = 0;                                      // bgzf: chunks whose inflate chain has been queued
        if (bgzf) {
            chunks.ctx = ctx;
            chunks.text_cap = text_cap;
            const int rc0 = chunks.prepare();
            if (rc0 != SIDGPU_OK) return rc0;
        }
        std::thread rd([this] { reader(); });
        int rc = SIDGPU_OK;
        uint64_t uploaded = 0;
        for (uint64_t i = 0; rc == SIDGPU_OK; ++i) {
            if (uploaded <= i) {
                const double w0 = now();
                const bool have = wait_text(i, true);
                t_wait_text += now() - w0;
                if (!have) break;
                const double u0 = now();
                rc = upload(i);
                t_upload += now() - u0;
                if (rc != SIDGPU_OK) break;
                uploaded = i + 1;
            }
            if (wait_text(i + 1, false)) {                          // the next slot is already there: its copy runs under this chunk's kernels
                rc = upload(i + 1);
                if (rc != SIDGPU_OK) break;
                uploaded = i + 2;
            }
            const int b = (int)(i & 1);
            size_t len = ts[i % NT].len;
            if (cudaStreamWaitEvent(ctx->stream, ev_in[b], 0) != cudaSuccess) { rc = ctx->fail(SIDGPU_ECUDA, "cudaStreamWaitEvent failed"); break; }
            if (bgzf) {
                // the chain of chunk i was queued while chunk i - 1 was being called (or is queued now: the first chunk,
                // or the reader was late); the chain of chunk i + 1 goes out as soon as this one says where its lines end
                const double f0 = now();
                if (launched <= i) {
                    rc = chunks.launch(b, ev_in[b], ts[i % NT].n_blocks, ts[i % NT].text_len, ts[i % NT].last);
                    launched = i + 1;
                }
                if (rc == SIDGPU_OK) rc = chunks.finish(b, &len);
                if (rc == SIDGPU_OK && uploaded > i + 1 && launched <= i + 1) {
                    const TextSlot& nx = ts[(i + 1) % NT];
                    rc = chunks.launch(b ^ 1, ev_in[b ^ 1], nx.n_blocks, nx.text_len, nx.last);
                    launched = i + 2;
                }
                t_inflate += now() - f0;
                if (rc != SIDGPU_OK) break;
            }
            uint64_t n = 0;
            if (len) {
                const double f0 = now();
                if (emit) rc = rows_of_chunk(b, (const char*)ctx->hp_text[b].p, len, &n);
                else rc = sidgpu_feed(ctx, (const char*)ctx->hp_text[b].p, len, 0, len, &n);
                t_feed += now() - f0;
                total_sites += n;
            }
            release_text();                                         // the kernels are done with it (both calls synchronise), so is the copy
        }
        if (rc != SIDGPU_OK) set_error(rc, ctx->err);
        rd.join();
        std::lock_guard<std::mutex> g(m);
        if (rc == SIDGPU_OK && error != SIDGPU_OK) { rc = error; ctx->err = errmsg; }
        return rc;
    }

    // the rows of every stored site, in blocks
    int emit_store() {
        const uint64_t step = (uint64_t)8 << 20;
        int b = 0;
        for (uint64_t s = 0; s < ctx->n_sites_total; s += step, b ^= 1) {
            const uint64_t count = std::min<uint64_t>(step, ctx->n_sites_total - s);
            CK(cudaEventSynchronize(ev_out[b]));
            uint64_t bytes = 0, rows = 0;
            size_t want = std::max<size_t>(ctx->hp_csv[b].cap, (size_t)count * 48 + 4096);
            for (;;) {
                TRY(ensure(ctx, ctx->hp_csv[b], want));
                const int rc = sidgpu_emit_csv(ctx, s, count, (char*)ctx->hp_csv[b].p, ctx->hp_csv[b].cap, &bytes, &rows);
                if (rc == SIDGPU_ECAPACITY && bytes + 4096 > want) { want = (size_t)bytes + 4096; continue; }
                if (rc != SIDGPU_OK) return rc;
                break;
            }
            TRY(queue_rows(b, bytes, rows));
        }
        return SIDGPU_OK;
    }

    int init() {
        slot_cap = std::min<size_t>(ctx->max_chunk, (size_t)64 << 20);      // the text slots are pinned: three of 64 MiB at most
        text_cap = std::max<size_t>(4 * slot_cap, (size_t)1 << 17);
        for (int i = 0; i < NT; ++i) {
            ts[i].cap = bgzf ? std::max<size_t>(slot_cap, (size_t)1 << 17) : slot_cap;
            CK(cudaHostAlloc((void**)&ts[i].p, ts[i].cap, cudaHostAllocDefault));
            if (bgzf) CK(cudaHostAlloc((void**)&ts[i].blocks, MAX_BLOCKS * sizeof(sidgpu_bgzf_block), cudaHostAllocDefault));
        }
        for (int i = 0; i < NC; ++i) CK(cudaEventCreateWithFlags(&cs[i].ev, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        }
        return SIDGPU_OK;
    }
    void destroy() {
        for (int i = 0; i < NT; ++i) { if (ts[i].p) cudaFreeHost(ts[i].p); if (ts[i].blocks) cudaFreeHost(ts[i].blocks); }
        for (int i = 0; i < NC; ++i) { if (cs[i].p) cudaFreeHost(cs[i].p); if (cs[i].ev) cudaEventDestroy(cs[i].ev); }
        for (int i = 0; i < 2; ++i) { if (ev_in[i]) cudaEventDestroy(ev_in[i]); if (ev_out[i]) cudaEventDestroy(ev_out[i]); }
    }
};

}  // namespace

static int call_io(sidgpu_ctx* ctx, const sidgpu_params* params, const sidgpu_io* io, bool bgzf, uint64_t* csv_bytes, uint64_t* n_sites,
                   uint64_t* n_rows) {
    if (!ctx || !params || !io || !io->read || !io->write) return SIDGPU_EINVAL;
    Range nvtx_range(bgzf ? "sidgpu_call_io_bgzf" : "sidgpu_call_io");
    CK(cudaSetDevice(ctx->device));
    InflateDrain drain_inflate {ctx};           // (bgzf) no inflate chain outlives the call, whatever way it ends
    IoPipe pipe;
    pipe.ctx = ctx;
    pipe.io = io;
    pipe.bgzf = bgzf;
    const double i0 = IoPipe::now();
    int rc = pipe.init();
    pipe.t_init = IoPipe::now() - i0;
    std::thread wr;
    if (rc == SIDGPU_OK) wr = std::thread([&pipe] { pipe.writer(); });
    if (rc == SIDGPU_OK) rc = sidgpu_begin(ctx, params);
    if (rc == SIDGPU_OK) {
        if (ctx->streaming) {
            rc = pipe.pass(true);
        } else {
            rc = pipe.pass(false);
            if (rc == SIDGPU_OK) rc = sidgpu_finish(ctx);
            if (rc == SIDGPU_OK) {
                if (params->method == SIDGPU_METHOD_QUALITY) {
                    if (!io->rewind || io->rewind(io->user) != 0) rc = ctx->fail(SIDGPU_EINVAL, "quality with -R reads the text twice: the input must be rewindable");
                    else { pipe.total_sites = 0; rc = pipe.pass(true); }
                } else {
                    rc = pipe.emit_store();
                }
            }
        }
    }
    if (wr.joinable()) {
        { std::lock_guard<std::mutex> g(pipe.m); pipe.stop_writer = true; }
        pipe.cv.notify_all();
        wr.join();
    }
    cudaStreamSynchronize(ctx->copy_out);
    cudaStreamSynchronize(ctx->stream);
    if (rc == SIDGPU_OK && pipe.error != SIDGPU_OK) { rc = pipe.error; ctx->err = pipe.errmsg; }
    pipe.destroy();
    if (getenv("SIDGPU_IO_TIMING"))
        fprintf(stderr, "# sidgpu_call_io: init %.3f, wait for text %.3f, upload %.3f, inflate %.3f, feed+rows %.3f (of which queueing rows %.3f) s\n",
                pipe.t_init, pipe.t_wait_text, pipe.t_upload, pipe.t_inflate, pipe.t_feed, pipe.t_queue);
    if (csv_bytes) *csv_bytes = pipe.total_bytes;
    if (n_sites) *n_sites = pipe.total_sites;
    if (n_rows) *n_rows = pipe.total_rows;
    return rc;
}

extern "C" int sidgpu_call_io(sidgpu_ctx* ctx, const sidgpu_params* params, const sidgpu_io* io, uint64_t* csv_bytes, uint64_t* n_sites,
                              uint64_t* n_rows) {
    return call_io(ctx, params, io, false, csv_bytes, n_sites, n_rows);
}

extern "C" int sidgpu_call_io_bgzf(sidgpu_ctx* ctx, const sidgpu_params* params, const sidgpu_io* io, uint64_t* csv_bytes, uint64_t* n_sites,
                                   uint64_t* n_rows) {
    return call_io(ctx, params, io, true, csv_bytes, n_sites, n_rows);
}
