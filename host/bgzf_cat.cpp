// bgzf_cat FILE [THREADS [BUFFER_BYTES]]: the text of a BGZF file on stdout, through bgzf::Reader -- the reader `sid` uses for
// blocked-gzip input, without the GPU.  Exit 1 with a message on a damaged file.
#include <fcntl.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "bgzf.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: bgzf_cat FILE [THREADS [BUFFER_BYTES]]\n"); return 2; }
    const int fd = open(argv[1], O_RDONLY);
    if (fd < 0) { std::fprintf(stderr, "Could not open file: %s\n", argv[1]); return 1; }
    const int threads = argc > 2 ? std::atoi(argv[2]) : 4;
    const size_t cap = argc > 3 ? (size_t)std::atol(argv[3]) : ((size_t)8 << 20);
    unsigned char head[64];
    const ssize_t got = pread(fd, head, sizeof head, 0);
    if (got < 18 || !bgzf::looks_like(head, (size_t)got)) { std::fprintf(stderr, "not a BGZF file\n"); return 1; }
    bgzf::Reader r(fd, threads);
    std::vector<char> buf(cap);
    for (;;) {
        const int64_t n = r.read(buf.data(), buf.size());
        if (n < 0) { std::fprintf(stderr, "bgzf: %s\n", r.error().c_str()); return 1; }
        if (n == 0) break;
        if (std::fwrite(buf.data(), 1, (size_t)n, stdout) != (size_t)n) return 1;
    }
    return 0;
}
