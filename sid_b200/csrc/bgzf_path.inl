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

// Queues the inflation of n members (table on the device) on ctx->stream; the error word is Control::error (reset here).
int launch_inflate(sidgpu_ctx* ctx, const uint8_t* d_comp, const sid::BgzfBlock* d_blocks, size_t n, uint8_t* d_text) {
    CK(cudaMemsetAsync(ctl_field(ctx, &Control::error), 0xFF, sizeof(unsigned long long), ctx->stream));
    if (n == 0) return SIDGPU_OK;
    TRY(ensure_crc_tables(ctx));
    static int per_sm = 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_inflate_bgzf, INF_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const size_t ctas = (n + INF_WARPS - 1) / INF_WARPS;
    const unsigned grid = (unsigned)std::min<size_t>(ctas, (size_t)ctx->sm_count * (size_t)per_sm);
    ProfScope prof(ctx, PROF_INFLATE);
    k_inflate_bgzf<<<grid, INF_WARPS * 32, 0, ctx->stream>>>(d_comp, d_blocks, (uint32_t)n, d_text, ctl_field(ctx, &Control::error));
    TRY(check_launch(ctx, "k_inflate_bgzf"));
    // every member's text against the CRC-32 of its trailer (what zcat checks)
    const unsigned crc_grid = (unsigned)std::min<size_t>((n + CRC_WARPS - 1) / CRC_WARPS, (size_t)ctx->sm_count * 8);
    k_crc32_members<<<crc_grid, CRC_WARPS * 32, 0, ctx->stream>>>(d_text, d_blocks, (uint32_t)n, (const sid::CrcTables*)ctx->crc_tables.p,
                                                                  ctl_field(ctx, &Control::error));
    return check_launch(ctx, "k_crc32_members");
}

int inflate_failed(sidgpu_ctx* ctx) {        // after sync_ctl
    if (ctx->h_ctl->error == ~0ull) return SIDGPU_OK;
    return ctx->fail(SIDGPU_EINVAL, "could not inflate BGZF member %llu: %s", ctx->h_ctl->error >> 4, inflate_error_text((int)(ctx->h_ctl->error & 15)));
}

}  // namespace

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
    TRY(launch_inflate(ctx, (const uint8_t*)d_comp, (const sid::BgzfBlock*)ctx->inf_blocks[0].p, n_blocks, (uint8_t*)d_text));
    TRY(sync_ctl(ctx));
    return inflate_failed(ctx);
}
