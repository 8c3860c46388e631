// `sid [flags] input_file` -- the reference's command line (sid.cpp:11-110) over the GPU path.
// Same flags and defaults (-m METHOD, -r PRIOR, -R, -p LEVEL, -E ERROR, -h), same CSV on stdout,
// same `# ...` lines on stderr, same exit codes.  Extra long options: --device N, --devices A,B,.. (one
// position shard per GPU, one shared fit), --chunk-mb N, --read-threads N (parallel preads per chunk, default min(16, cores)),
// --het-only (rows labelled het only: the pipeline's `grep ',het,'`, scripts/sid-pipeline/run-sid.sh:16-17;
// the header line is kept).  With one GPU the input streams through pinned chunks and the rows are written as they
// arrive (sidCallFile): memory use does not grow with the file.  A gzip-compressed input (as the pipeline stores its
// pileups, scripts/prepare-data.sh:14) is inflated on the fly, under the GPU work, instead of `zcat` to a temporary
// file (scripts/sid-pipeline/run-sid.sh:15): a plain gzip stream by zlib on the reader thread; a blocked one (BGZF, what
// `bgzip` writes) on the device, one warp per member, so that only the compressed bytes cross the link (--host-inflate:
// by the reader's threads instead, host/bgzf.hpp).
#include <fcntl.h>
#include <getopt.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "call.hpp"

struct GlobalOptions {                               // sid.cpp:11-17
    std::string method {"local"};
    bool estimate_prior = false;
    double snp_prior = -1;
    double significance_level = 0.05;
    double site_error_threshold = 0.1;
};

static void help() {                                 // sid.cpp:27-37 (same layout: one line per flag, sorted by letter)
    std::cout << "sid [flags] input_file" << '\n';
    std::cout << "\t-E ERROR\tMaximum allowed site error rate for 'local' method. Default: 0.1\n";
    std::cout << "\t-R\tEstimate SNP prior from data, applicable for methods 'likelihood_ratio', 'local', 'quality'. Conflicts -r.\n";
    std::cout << "\t-h\tPrint this help message\n";
    std::cout << "\t-m METHOD\tSelect the method to use for SNP calling: 'likelihood_ratio' , 'bayes', 'local' or 'quality', default: local\n";
    std::cout << "\t-p LEVEL\tSignificance level for statistical tests, only applicable for methods 'likelihood_ratio', 'local'. Default: 0.05\n";
    std::cout << "\t-r PRIOR\tUse the given prior for SNPs, applicable for methods 'local', 'quality'. Conflicts -R. Default: no prior\n";
}

int main(int argc, char** argv) {
    GlobalOptions o;
    int device = 0;
    size_t chunk_mb = 0;
    bool het_only = false, host_inflate = false;
    std::vector<int> devices;
    int read_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    static const option LONG[] = {{"device", required_argument, nullptr, 1000}, {"chunk-mb", required_argument, nullptr, 1001}, {"het-only", no_argument, nullptr, 1002}, {"devices", required_argument, nullptr, 1003}, {"read-threads", required_argument, nullptr, 1004}, {"host-inflate", no_argument, nullptr, 1005}, {nullptr, 0, nullptr, 0}};
    int flag;
    while ((flag = getopt_long(argc, argv, "E:Rhm:p:r:", LONG, nullptr)) != -1) {      // optstring as built by sid.cpp:60-69
        switch (flag) {
            case 'h': help(); break;
            case 'm': o.method = optarg; break;
            case 'r': o.snp_prior = atof(optarg); break;
            case 'R': o.estimate_prior = true; break;
            case 'p': o.significance_level = atof(optarg); break;
            case 'E': o.site_error_threshold = atof(optarg); break;
            case 1000: device = atoi(optarg); break;
            case 1001: chunk_mb = (size_t)atol(optarg); break;
            case 1002: het_only = true; break;
            case 1004: read_threads = std::max(1, atoi(optarg)); break;
            case 1005: host_inflate = true; break;
            case 1003:
                for (const char* q = optarg; *q;) {
                    devices.push_back(atoi(q));
                    while (*q && *q != ',') ++q;
                    if (*q == ',') ++q;
                }
                break;
            default: exit(EXIT_FAILURE);             // sid.cpp:80-82
        }
    }
    if (optind >= argc) {                            // sid.cpp:106-109
        std::cerr << "No file name given!" << std::endl;
        exit(EXIT_FAILURE);
    }
    const char* path = argv[optind];
    const int fd = open(path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {             // sid.cpp:86-89
        std::cerr << "Could not open file: " << path << std::endl;
        exit(EXIT_FAILURE);
    }
    size_t len = (size_t)st.st_size;
    unsigned char magic[2] = {0, 0};
    const bool gz = len >= 18 && pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    static const char* HEADER = "chrom,pos,label,gt,hom_conf,het_conf,conf_type";     // sid.cpp:102
    std::ios::sync_with_stdio(false);
    const char* text = "";
    void* map = nullptr;
    std::vector<char> inflated;
    try {
        if (devices.size() == 1) device = devices[0];
        sidSetDevice(device, chunk_mb << 20);
        sidSetHetOnly(het_only);
        sidSetHostInflate(host_inflate);
        if (devices.size() > 1) {
            // one position shard per GPU: the shards are cut from the whole text, which therefore has to be in memory
            if (gz) {
                gzFile z = gzdopen(dup(fd), "rb");
                if (!z) { std::cerr << "Could not open file: " << path << std::endl; exit(EXIT_FAILURE); }
                gzbuffer(z, 1u << 20);
                inflated.resize(std::max<size_t>(len * 4, (size_t)1 << 20));
                size_t have = 0;
                for (;;) {
                    if (have == inflated.size()) inflated.resize(inflated.size() * 2);
                    const int got = gzread(z, inflated.data() + have, (unsigned)std::min<size_t>(inflated.size() - have, (size_t)1 << 30));
                    if (got < 0) { std::cerr << "Could not inflate file: " << path << std::endl; exit(EXIT_FAILURE); }
                    if (got == 0) break;
                    have += (size_t)got;
                }
                gzclose(z);
                len = have;
                text = inflated.data();
            } else if (len) {
                map = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
                if (map == MAP_FAILED) { std::cerr << "Could not open file: " << path << std::endl; exit(EXIT_FAILURE); }
                text = (const char*)map;
            }
            sidCallToStreamSharded(o.method, text, len, o.estimate_prior, o.snp_prior, o.site_error_threshold, o.significance_level, devices,
                                   std::cout, std::cerr, HEADER);
            std::cout.flush();
        } else {
            // one GPU: the file streams through pinned chunks, rows go to stdout as they arrive (call.hpp: sidCallFile)
            sidCallFile(o.method, fd, gz, o.estimate_prior, o.snp_prior, o.site_error_threshold, o.significance_level, STDOUT_FILENO,
                        std::cerr, HEADER, read_threads);
        }
    } catch (const std::invalid_argument& e) {
        // the reference lets the exception escape: terminate() -> abort, exit status 134
        std::cerr << "terminate called after throwing an instance of 'std::invalid_argument'\n  what():  " << e.what() << std::endl;
        std::abort();
    } catch (const std::exception& e) {
        std::cerr << "sid: " << e.what() << std::endl;
        return 2;
    }
    if (map) munmap(map, len);
    close(fd);
    return 0;
}
