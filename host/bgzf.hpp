// Block-parallel reader for BGZF (the blocked gzip that `bgzip` and samtools write): SURVEY.md 8f row 1, host half.
// The reference's pipeline stores its pileups as plain `gzip -c` streams (scripts/prepare-data.sh:14) and inflates them
// with `zcat` into a temporary file (scripts/sid-pipeline/run-sid.sh:15): one core at 0.4 GB/s.  A plain gzip stream is
// serial by construction; a BGZF file is a sequence of independent gzip members of at most 64 KiB, each of which says in
// its header how long it is (extra subfield 'B','C': BSIZE) and in its trailer how long its text is (ISIZE).  So the
// reader walks the block headers, lays the blocks' texts end to end in the caller's buffer and inflates them on as many
// threads as it is given -- straight into the pinned slot of the streaming host (sidgpu_call_io's read callback).
// Header only, host code, zlib.
#pragma once
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace bgzf {

// Is this the beginning of a BGZF file?  (gzip magic, deflate, FEXTRA, and a 'B','C' subfield of length 2)
inline bool looks_like(const unsigned char* p, size_t n) {
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
    const unsigned xlen = p[10] | (p[11] << 8);
    size_t q = 12;
    while (q + 4 <= 12 + (size_t)xlen && q + 4 <= n) {
        const unsigned slen = p[q + 2] | (p[q + 3] << 8);
        if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2) return true;
        q += 4 + slen;
    }
    return false;
}

class Reader {
public:
    Reader(int fd, int threads) : fd_(fd), threads_(std::max(1, threads)) {}
    const std::string& error() const { return err_; }
    void rewind() { pos_ = 0; eof_ = false; spill_.clear(); spill_off_ = 0; }

    // Fills dst with the text of as many whole blocks as fit into cap; when not even the next block fits, with the first
    // cap bytes of it (the rest waits for the next call).  Returns the number of bytes, 0 at the end of the file, -1 on
    // a damaged file (error() says what).
    int64_t read(char* dst, size_t cap) {
        if (cap == 0) return 0;
        if (spill_off_ < spill_.size()) {
            const size_t n = std::min(cap, spill_.size() - spill_off_);
            std::memcpy(dst, spill_.data() + spill_off_, n);
            spill_off_ += n;
            return (int64_t)n;
        }
        if (eof_) return 0;
        blocks_.clear();
        // ---- walk the block headers: compressed bytes of the batch go to cbuf_
        const size_t want = std::max<size_t>(cap / 2, (size_t)1 << 20);       // pileup text deflates to a quarter or less
        if (cbuf_.size() < want + 65536) cbuf_.resize(want + 65536);
        const ssize_t got = pread_all(cbuf_.data(), cbuf_.size(), pos_);
        if (got < 0) return fail("read error");
        if (got == 0) { eof_ = true; return 0; }
        size_t q = 0, out = 0;
        while (q + 18 <= (size_t)got) {
            const unsigned char* h = cbuf_.data() + q;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return fail("not a BGZF block header");
            const unsigned xlen = h[10] | (h[11] << 8);
            if (q + 12 + xlen > (size_t)got) break;                            // header cut by the window: next call
            unsigned bsize = 0;
            bool found = false;
            for (size_t e = 12; e + 4 <= 12 + (size_t)xlen;) {
                const unsigned slen = h[e + 2] | (h[e + 3] << 8);
                if (h[e] == 'B' && h[e + 1] == 'C' && slen == 2) { bsize = h[e + 4] | (h[e + 5] << 8); found = true; }
                e += 4 + slen;
            }
            if (!found) return fail("gzip member without a BGZF size field");
            const size_t block_len = (size_t)bsize + 1;
            if (block_len < 12 + xlen + 8) return fail("BGZF block shorter than its header");
            if (q + block_len > (size_t)got) break;                            // block cut by the window
            const unsigned char* t = h + block_len - 8;
            const uint32_t crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            const uint32_t isize = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
            if (isize > 65536) return fail("BGZF block claims more than 64 KiB of text");
            if (out + isize > cap) {                                           // the caller's buffer is full
                if (out == 0) {                                                // not even one block fits: through the spill buffer
                    spill_.resize(isize);
                    spill_off_ = 0;
                    const Block b {q + 12 + xlen, block_len - 12 - xlen - 8, 0, isize, crc};
                    z_stream zs;
                    std::memset(&zs, 0, sizeof zs);
                    if (inflateInit2(&zs, -15) != Z_OK || !inflate_block(zs, b, spill_.data())) { inflateEnd(&zs); return fail("damaged BGZF block (inflate or CRC)"); }
                    inflateEnd(&zs);
                    pos_ += (off_t)block_len;
                    return read(dst, cap);
                }
                break;
            }
            blocks_.push_back({q + 12 + xlen, block_len - 12 - xlen - 8, out, isize, crc});
            out += isize;
            q += block_len;
        }
        if (q == 0) {
            if ((size_t)got < cbuf_.size()) return fail("truncated BGZF block at the end of the file");
            return fail("BGZF block larger than the read window");
        }
        pos_ += (off_t)q;
        // ---- inflate the blocks side by side, each into its place
        const int nt = (int)std::min<size_t>((size_t)threads_, std::max<size_t>(1, blocks_.size() / 4));
        std::vector<int> bad((size_t)nt, 0);
        auto work = [&](int t) {
            z_stream zs;
            std::memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { bad[(size_t)t] = 1; return; }
            for (size_t i = (size_t)t; i < blocks_.size(); i += (size_t)nt)
                if (!inflate_block(zs, blocks_[i], dst)) { bad[(size_t)t] = 1; break; }
            inflateEnd(&zs);
        };
        if (nt == 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; ++t) pool.emplace_back(work, t);
            for (auto& th : pool) th.join();
        }
        for (int b : bad) if (b) return fail("damaged BGZF block (inflate or CRC)");
        if (out == 0) return read(dst, cap);                                   // only empty blocks in this window: go on
        return (int64_t)out;
    }

private:
    struct Block { size_t c_off, c_len, out_off; uint32_t isize, crc; };
    int fd_;
    int threads_;
    off_t pos_ = 0;
    bool eof_ = false;
    std::string err_;
    std::vector<unsigned char> cbuf_;
    std::vector<Block> blocks_;
    std::vector<char> spill_;           // a block that did not fit into the caller's buffer
    size_t spill_off_ = 0;

    bool inflate_block(z_stream& zs, const Block& b, char* dst) {
        if (b.isize == 0) return true;                                         // the end-of-file marker and other empty blocks
        inflateReset(&zs);
        zs.next_in = cbuf_.data() + b.c_off;
        zs.avail_in = (uInt)b.c_len;
        zs.next_out = (Bytef*)dst + b.out_off;
        zs.avail_out = b.isize;
        const int rc = inflate(&zs, Z_FINISH);
        return rc == Z_STREAM_END && zs.avail_out == 0 &&
               (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)dst + b.out_off, b.isize) == b.crc;
    }

    int64_t fail(const char* what) { err_ = what; return -1; }
    ssize_t pread_all(unsigned char* p, size_t n, off_t at) {
        size_t have = 0;
        while (have < n) {
            const ssize_t g = pread(fd_, p + have, n - have, at + (off_t)have);
            if (g < 0) return -1;
            if (g == 0) break;
            have += (size_t)g;
        }
        return (ssize_t)have;
    }
};

}  // namespace bgzf
