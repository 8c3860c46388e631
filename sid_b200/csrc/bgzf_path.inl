// BGZF input inflated on the device (inflate.cuh): the header walk on the host, the kernel launch, and the text chunk
// bookkeeping the streaming host needs (where the last whole line ends).

extern "C" int sidgpu_bgzf_scan(const void* h_comp, size_t len, sidgpu_bgzf_block* blocks, size_t max_blocks, size_t text_cap,
                                size_t* n_blocks, size_t* consumed, size_t* text_bytes) {
    if ((!h_comp && len) || !blocks || !n_blocks || !consumed || !text_bytes) return SIDGPU_EINVAL;
    const unsigned char* p = (const unsigned char*)h_comp;
    size_t q = 0, n = 0, out = 0;
    while (q + 18 <= len && n < max_blocks) {
        const unsigned char* h = p + q;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return SIDGPU_EINVAL;         // not a BGZF block header
        const unsigned xlen = h[10] | (h[11] << 8);
        if (q + 12 + xlen > len) break;                                 // header cut by the window
        unsigned bsize = 0;
        bool found = false;
        for (size_t e = 12; e + 4 <= 12 + (size_t)xlen;) {
            const unsigned slen = h[e + 2] | (h[e + 3] << 8);
            if (h[e] == 'B' && h[e + 1] == 'C' && slen == 2 && e + 6 <= 12 + (size_t)xlen) { bsize = h[e + 4] | (h[e + 5] << 8); found = true; }
            e += 4 + slen;
        }
        if (!found) return SIDGPU_EINVAL;                               // a gzip member without the BGZF size field
        const size_t block_len = (size_t)bsize + 1;
        if (block_len < 12 + xlen + 8) return SIDGPU_EINVAL;
        if (q + block_len > len) break;                                 // block cut by the window
        const unsigned char* t = h + block_len - 4;
        const uint32_t isize = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        const uint32_t crc = (uint32_t)t[-4] | ((uint32_t)t[-3] << 8) | ((uint32_t)t[-2] << 16) | ((uint32_t)t[-1] << 24);
        if (isize > 65536) return SIDGPU_EINVAL;
        if (out + isize > text_cap) break;
        if (isize) {                                                    // the end-of-file marker and other empty members carry no text
            blocks[n].c_off = q + 12 + xlen;
            blocks[n].out_off = out;
            blocks[n].c_len = (uint32_t)(block_len - 12 - xlen - 8);
            blocks[n].isize = isize;
            blocks[n].crc = crc;
            blocks[n].reserved = 0;
            ++n;
            out += isize;
        }
        q += block_len;
    }
    *n_blocks = n;
    *consumed = q;
    *text_bytes = out;
    return SIDGPU_OK;
}

namespace {

const char* inflate_error_text(int code) {
    switch (code) {
        case sid::INF_BAD_BLOCK_TYPE: return "bad deflate block type";
        case sid::INF_BAD_STORED: return "bad stored block";
        case sid::INF_BAD_LENGTHS: return "bad code lengths";
        case sid::INF_BAD_SYMBOL: return "bad symbol";
        case sid::INF_BAD_DISTANCE: return "bad distance";
        case sid::INF_OUTPUT_OVERRUN: return "more text than the member's trailer says";
        case sid::INF_INPUT_OVERRUN: return "deflate stream runs past the member";
        case sid::INF_SIZE_MISMATCH: return "less text than the member's trailer says";
        case sid::INF_CRC_MISMATCH: return "the text does not have the CRC-32 of the member's trailer";
        default: return "damaged member";
    }
}

// The tables of k_crc32_members, built once per ctx.
int ensure_crc_tables(sidgpu_ctx* ctx) {
    if (ctx->crc_tables.p) return SIDGPU_OK;
    sid::CrcTables h;
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        h.byte_table[i] = c;
    }
    for (int j = 0; j < 32; ++j) {
        uint32_t r = 1u << j;
        for (uint32_t k = 0; k < sid::CRC_PIECE; ++k) r = h.byte_table[r & 0xFFu] ^ (r >> 8);
        h.advance_2k[j] = r;
    }
    TRY(ensure(ctx, ctx->crc_tables, sizeof h));
    CK(cudaMemcpyAsync(ctx->crc_tables.p, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));          // `h` lives on this stack frame
    return SIDGPU_OK;
}

__global__ void k_publish_words(const unsigned long long* d, unsigned long long* h, unsigned n) {
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) ((volatile unsigned long long*)h)[i] = d[i];
    __threadfence_system();
}

// Stream, events and result words of the inflate chain, created on first use.
int ensure_inflate_state(sidgpu_ctx* ctx) {
    if (ctx->inflate_stream) return SIDGPU_OK;
    CK(cudaStreamCreateWithFlags(&ctx->inflate_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&ctx->ev_inflate[i], cudaEventDisableTiming));
    CK(cudaMalloc((void**)&ctx->d_inf, 4 * sizeof(unsigned long long)));
    CK(cudaHostAlloc((void**)&ctx->h_inf, 4 * sizeof(unsigned long long), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&ctx->h_inf_dev, ctx->h_inf, 0));
    return ensure_crc_tables(ctx);
}

// Queues on `stream`: the inflation of n members (table on the device) into d_text, the CRC-32 check of their text, the
// search for the end of the last whole line of text_base[0, total) when `cut`, and the publication of the two result
// words of buffer b (error: ~0 = none; line end) to the host.
int launch_inflate(sidgpu_ctx* ctx, cudaStream_t stream, int b, const uint8_t* d_comp, const sid::BgzfBlock* d_blocks, size_t n, uint8_t* d_text,
                   const uint8_t* text_base, size_t total, bool cut) {
    TRY(ensure_inflate_state(ctx));
    unsigned long long* words = ctx->d_inf + 2 * b;
    CK(cudaMemsetAsync(words, 0xFF, sizeof(unsigned long long), stream));
    CK(cudaMemsetAsync(words + 1, 0, sizeof(unsigned long long), stream));
    if (n) {
        static int per_sm = 0;
        if (per_sm == 0) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_inflate_bgzf, INF_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        }
        const size_t ctas = (n + INF_WARPS - 1) / INF_WARPS;
        const unsigned grid = (unsigned)std::min<size_t>(ctas, (size_t)ctx->sm_count * (size_t)per_sm);
        ProfScope prof(ctx, PROF_INFLATE, stream);
        k_inflate_bgzf<<<grid, INF_WARPS * 32, 0, stream>>>(d_comp, d_blocks, (uint32_t)n, d_text, words);
        TRY(check_launch(ctx, "k_inflate_bgzf"));
        // every member's text against the CRC-32 of its trailer (what zcat checks)
        const unsigned crc_grid = (unsigned)std::min<size_t>((n + CRC_WARPS - 1) / CRC_WARPS, (size_t)ctx->sm_count * 8);
        k_crc32_members<<<crc_grid, CRC_WARPS * 32, 0, stream>>>(d_text, d_blocks, (uint32_t)n, (const sid::CrcTables*)ctx->crc_tables.p, words);
        TRY(check_launch(ctx, "k_crc32_members"));
    }
    if (cut && total) {
        k_last_line_end<<<1, 256, 0, stream>>>(text_base, total, words + 1);
        TRY(check_launch(ctx, "k_last_line_end"));
    }
    k_publish_words<<<1, 32, 0, stream>>>(words, ctx->h_inf_dev + 2 * b, 2);
    TRY(check_launch(ctx, "k_publish_words"));
    CK(cudaEventRecord(ctx->ev_inflate[b], stream));
    return SIDGPU_OK;
}

// Waits for the chain of buffer b; *line_end (optional) receives the end of the last whole line.
int finish_inflate(sidgpu_ctx* ctx, int b, size_t* line_end) {
    CK(cudaEventSynchronize(ctx->ev_inflate[b]));
    const unsigned long long e = ((volatile unsigned long long*)ctx->h_inf)[2 * b];
    if (e != ~0ull) return ctx->fail(SIDGPU_EINVAL, "could not inflate BGZF member %llu: %s", e >> 4, inflate_error_text((int)(e & 15)));
    if (line_end) *line_end = (size_t)((volatile unsigned long long*)ctx->h_inf)[2 * b + 1];
    return SIDGPU_OK;
}

// Whoever queues inflate chains waits for the last one on every way out (an error may leave the chain of the next chunk in
// flight; the buffers it writes belong to the ctx and are reused by the next call).
struct InflateDrain {
    sidgpu_ctx* ctx;
    ~InflateDrain() { if (ctx->inflate_stream) cudaStreamSynchronize(ctx->inflate_stream); }
};

// A streamed BGZF input, chunk by chunk through the two text buffers of the ctx.  Chunk i lies compressed in hp_comp[i & 1]
// (table in inf_blocks[i & 1]); its text goes behind the unfinished line the chunk before left at the front of
// hp_text[i & 1].  launch(i) queues its inflate chain on the inflate stream -- for chunk i + 1 right after finish(i), so that
// it runs beside the calling kernels of chunk i -- and finish(i) says how many bytes of whole lines there are and moves
// what follows them to the front of the other buffer.
struct BgzfChunks {
    sidgpu_ctx* ctx;
    size_t text_cap;            // most text a chunk inflates to
    size_t tail[2] = {0, 0};    // bytes of an unfinished line at the front of hp_text[b]
    size_t total[2] = {0, 0};   // tail + text of the chunk in flight in buffer b
    bool cut[2] = {false, false};

    int prepare() {             // before the first launch of a pass
        tail[0] = tail[1] = 0;
        TRY(ensure_inflate_state(ctx));
        return ensure(ctx, ctx->hp_text[0], ((text_cap + 15) & ~(size_t)15) + 32);
    }
    int launch(int b, cudaEvent_t uploaded, size_t n_blocks, size_t text_len, bool last) {
        total[b] = tail[b] + text_len;
        cut[b] = !last && total[b] != 0;
        TRY(ensure(ctx, ctx->hp_text[b], ((total[b] + 15) & ~(size_t)15) + 32, tail[b] != 0));       // (sized by finish() already: a no-op)
        CK(cudaStreamWaitEvent(ctx->inflate_stream, uploaded, 0));
        return launch_inflate(ctx, ctx->inflate_stream, b, (const uint8_t*)ctx->hp_comp[b].p, (const sid::BgzfBlock*)ctx->inf_blocks[b].p, n_blocks,
                              (uint8_t*)ctx->hp_text[b].p + tail[b], (const uint8_t*)ctx->hp_text[b].p, total[b], cut[b]);
    }
    int finish(int b, size_t* keep) {
        size_t line_end = 0;
        TRY(finish_inflate(ctx, b, &line_end));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_inflate[b], 0));
        *keep = cut[b] ? line_end : total[b];
        const size_t rest = total[b] - *keep;
        tail[b ^ 1] = rest;
        if (cut[b]) {
            TRY(ensure(ctx, ctx->hp_text[b ^ 1], ((rest + text_cap + 15) & ~(size_t)15) + 32));
            if (rest) CK(cudaMemcpyAsync(ctx->hp_text[b ^ 1].p, (const char*)ctx->hp_text[b].p + *keep, rest, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        return SIDGPU_OK;
    }
};

}  // namespace

// sidgpu_call_host for a BGZF file that lies in host memory: compressed bytes in (straight from h_comp: pin it for speed),
// CSV rows out.  One thread: chunk i + 1 is scanned and uploaded while the device works on chunk i.
extern "C" int sidgpu_call_host_bgzf(sidgpu_ctx* ctx, const sidgpu_params* params, const void* h_comp, size_t comp_len, char* h_csv,
                                     size_t csv_cap, uint64_t* csv_bytes, uint64_t* n_sites, uint64_t* n_rows) {
    if (!ctx || !params || (comp_len && !h_comp)) return SIDGPU_EINVAL;
    Range nvtx_range("sidgpu_call_host_bgzf");
    CK(cudaSetDevice(ctx->device));
    InflateDrain drain_inflate {ctx};
    HostIo io(ctx);
    TRY(io.init());
    io.h_csv = h_csv;
    io.csv_cap = csv_cap;
    const unsigned char* comp = (const unsigned char*)h_comp;
    const size_t text_cap = std::max<size_t>(4 * std::min<size_t>(ctx->max_chunk, (size_t)64 << 20), (size_t)1 << 17);
    const size_t max_blocks = (size_t)1 << 16;
    std::vector<sidgpu_bgzf_block> table[2] = {std::vector<sidgpu_bgzf_block>(max_blocks), std::vector<sidgpu_bgzf_block>(max_blocks)};
    struct Chunk { size_t off = 0, len = 0, n_blocks = 0, text_len = 0; bool last = false; };
    auto scan_upload = [&](int b, size_t off, Chunk& c) -> int {
        c.off = off;
        size_t used = 0;
        if (sidgpu_bgzf_scan(comp + off, comp_len - off, table[b].data(), max_blocks, text_cap, &c.n_blocks, &used, &c.text_len) != SIDGPU_OK)
            return ctx->fail(SIDGPU_EINVAL, "not a BGZF file (bad member header at byte %zu)", off);
        if (used == 0 && off < comp_len) return ctx->fail(SIDGPU_EINVAL, "truncated BGZF member at the end of the buffer");
        c.len = used;
        c.last = off + used >= comp_len;
        TRY(ensure(ctx, ctx->hp_comp[b], ((used + 15) & ~(size_t)15) + 16));
        TRY(ensure(ctx, ctx->inf_blocks[b], std::max<size_t>(1, c.n_blocks) * sizeof(sidgpu_bgzf_block)));
        if (used) CK(cudaMemcpyAsync(ctx->hp_comp[b].p, comp + off, used, cudaMemcpyHostToDevice, ctx->copy_in));
        if (c.n_blocks) CK(cudaMemcpyAsync(ctx->inf_blocks[b].p, table[b].data(), c.n_blocks * sizeof(sidgpu_bgzf_block), cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaEventRecord(io.hp.ev_in[b], ctx->copy_in));
        return SIDGPU_OK;
    };
    BgzfChunks chunks {ctx, text_cap};
    auto pass = [&](bool emit) -> int {
        TRY(chunks.prepare());
        Chunk cur, nxt;
        TRY(scan_upload(0, 0, cur));
        TRY(chunks.launch(0, io.hp.ev_in[0], cur.n_blocks, cur.text_len, cur.last));
        for (int i = 0;; ++i) {
            const int b = i & 1;
            if (!cur.last) TRY(scan_upload(b ^ 1, cur.off + cur.len, nxt));       // its copy overlaps with the work on chunk i
            size_t keep = 0;
            TRY(chunks.finish(b, &keep));
            if (!cur.last) TRY(chunks.launch(b ^ 1, io.hp.ev_in[b ^ 1], nxt.n_blocks, nxt.text_len, nxt.last));      // beside the kernels below
            if (keep) {
                uint64_t n = 0;
                TRY(sidgpu_feed(ctx, (const char*)ctx->hp_text[b].p, keep, 0, keep, &n));
                io.total_sites += n;
                int bb = b;
                if (emit) for (uint64_t s0 = 0; s0 < n; s0 += (uint64_t)1 << 20, bb ^= 1) TRY(io.emit_range(bb, s0, std::min<uint64_t>((uint64_t)1 << 20, n - s0)));
            }
            if (cur.last) break;
            cur = nxt;
        }
        return SIDGPU_OK;
    };
    TRY(sidgpu_begin(ctx, params));
    if (ctx->streaming) {
        TRY(pass(true));
    } else {
        TRY(pass(false));
        TRY(sidgpu_finish(ctx));
        if (params->method == SIDGPU_METHOD_QUALITY) {
            io.total_sites = 0;
            TRY(pass(true));                                        // second pass with the fitted prior
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

extern "C" int sidgpu_inflate_bgzf(sidgpu_ctx* ctx, const void* d_comp, size_t comp_len, const sidgpu_bgzf_block* h_blocks, size_t n_blocks,
                                   char* d_text, size_t text_cap) {
    if (!ctx || (n_blocks && (!d_comp || !h_blocks || !d_text)) || ((uintptr_t)d_comp & 3)) return SIDGPU_EINVAL;
    Range nvtx_range("sidgpu_inflate_bgzf");
    CK(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < n_blocks; ++i)
        if (h_blocks[i].c_off + h_blocks[i].c_len > comp_len || h_blocks[i].out_off + h_blocks[i].isize > text_cap) return SIDGPU_EINVAL;
    if (n_blocks == 0) return SIDGPU_OK;
    TRY(ensure(ctx, ctx->inf_blocks[0], n_blocks * sizeof(sidgpu_bgzf_block)));
    CK(cudaMemcpyAsync(ctx->inf_blocks[0].p, h_blocks, n_blocks * sizeof(sidgpu_bgzf_block), cudaMemcpyHostToDevice, ctx->stream));
    TRY(launch_inflate(ctx, ctx->stream, 0, (const uint8_t*)d_comp, (const sid::BgzfBlock*)ctx->inf_blocks[0].p, n_blocks, (uint8_t*)d_text,
                       (const uint8_t*)d_text, 0, false));
    return finish_inflate(ctx, 0, nullptr);
}
