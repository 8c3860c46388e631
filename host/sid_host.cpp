// Host side of the drop-in: the reference's free functions (call.hpp / pileup.hpp) implemented over
// the C ABI of libsidgpu.so.  No parsing, likelihood or formatting code lives here -- this file only
// moves bytes: istream -> pinned host buffer -> sidgpu_call_host -> CSV rows -> OutputRecord.
#include <cstdio>
#include <cstring>
#include <iomanip>
#include <iterator>
#include <memory>
#include <sstream>
#include <algorithm>
#include <chrono>
#include <limits>
#include <stdexcept>
#include <thread>

#include <unistd.h>
#include <zlib.h>

#include "bgzf.hpp"

#include "../include/sidgpu.h"
#include "../sid_b200/csrc/nelder_mead.hpp"      // the simplex driver the library itself uses (host code, header only)
#include "call.hpp"
#include "lynch.hpp"
#include "stats.hpp"

namespace {

int g_device = 0;
size_t g_chunk = 0;
bool g_het_only = false;
bool g_host_inflate = false;         // BGZF input: inflate on the host threads (bgzf.hpp) instead of on the device

struct Ctx {
    sidgpu_ctx* h = nullptr;
    Ctx() {
        sidgpu_config cfg {};
        cfg.device = g_device;
        cfg.max_chunk_bytes = g_chunk;
        if (sidgpu_create(&cfg, &h) != SIDGPU_OK) throw std::runtime_error(std::string("sidgpu: ") + sidgpu_last_error(nullptr));
    }
    ~Ctx() { sidgpu_destroy(h); }
};

Ctx& ctx() {
    static std::unique_ptr<Ctx> c;
    if (!c) c.reset(new Ctx);
    return *c;
}

[[noreturn]] void raise(int rc) {
    const std::string msg = sidgpu_last_error(ctx().h);
    // pileup.cpp:9-10: the reference throws std::invalid_argument with exactly these texts
    if (rc == SIDGPU_EMALFORMED) throw std::invalid_argument("Malformed pileup line");
    if (rc == SIDGPU_EMISSING_MAPQ) throw std::invalid_argument("Malformed pileup line or missing mapping qualities");
    if (rc == SIDGPU_EQUAL_SHORT) throw std::invalid_argument("Malformed pileup line: " + msg);
    throw std::runtime_error("sidgpu error " + std::to_string(rc) + ": " + msg);
}

void check(int rc) { if (rc != SIDGPU_OK) raise(rc); }

struct Pinned {
    char* p = nullptr;
    size_t cap = 0;
    ~Pinned() { if (p) sidgpu_free_host(ctx().h, p); }
    void reserve(size_t n) {
        if (n <= cap) return;
        char* np = nullptr;
        check(sidgpu_malloc_host(ctx().h, n, (void**)&np));
        if (p) { std::memcpy(np, p, cap); sidgpu_free_host(ctx().h, p); }
        p = np;
        cap = n;
    }
};

// Drains a stream into pinned memory (the reference's getline loop, call.cpp:13, reads it all too).
size_t slurp(std::istream& in, Pinned& buf) {
    size_t len = 0;
    buf.reserve(1 << 20);
    while (in) {
        if (len == buf.cap) buf.reserve(buf.cap * 2);
        in.read(buf.p + len, (std::streamsize)(buf.cap - len));
        len += (size_t)in.gcount();
    }
    return len;
}

int method_id(const std::string& m) {
    if (m == "local") return SIDGPU_METHOD_LOCAL;
    if (m == "bayes") return SIDGPU_METHOD_BAYES;
    if (m == "likelihood_ratio") return SIDGPU_METHOD_LIKELIHOOD_RATIO;
    if (m == "quality") return SIDGPU_METHOD_QUALITY;
    return -1;
}

struct Rows {
    Pinned csv;
    uint64_t bytes = 0, sites = 0, rows = 0;
};

void run(int method, const char* text, size_t len, bool estimate_prior, double prior, double error_threshold,
         double significance_level, Rows& out, SidRunInfo& info) {
    sidgpu_params p {};
    p.method = method;
    p.estimate_prior = estimate_prior ? 1 : 0;
    p.prior = prior;
    p.error_threshold = error_threshold;
    p.significance_level = significance_level;
    p.het_only = g_het_only ? 1 : 0;
    out.csv.reserve(len + len / 2 + 4096);
    for (;;) {
        const int rc = sidgpu_call_host(ctx().h, &p, text, len, out.csv.p, out.csv.cap, &out.bytes, &out.sites, &out.rows);
        if (rc == SIDGPU_ECAPACITY && out.bytes > out.csv.cap) { out.csv.reserve(out.bytes + 4096); continue; }
        check(rc);
        break;
    }
    info.n_sites = out.sites;
    info.n_rows = out.rows;
    sidgpu_fit fit {};
    double nd[4];
    uint64_t nu = 0;
    if (sidgpu_session_fit(ctx().h, &fit, nd, &nu) == SIDGPU_OK) {
        info.has_fit = true;
        info.heterozygosity = fit.pi;
        info.error_rate = fit.eps;
        info.iterations = fit.iterations;
        info.converged = fit.converged != 0;
        info.unique_profiles = nu;
    }
}

// The `# ...` lines the reference writes to std::cerr (call.cpp:72,78-80,155,161-163; optimization.hpp:70,76)
void log_fit(int method, const SidRunInfo& info, std::ostream& log) {
    if (!info.has_fit) return;
    const bool lynch_method = method == SIDGPU_METHOD_BAYES || method == SIDGPU_METHOD_LIKELIHOOD_RATIO;
    if (lynch_method) log << "# unique profiles: " << info.unique_profiles << std::endl;
    if (info.converged) log << "# GSL function minimization converged in " << info.iterations << " iterations." << std::endl;
    else log << "# Error: GSL function minimization did not converge in " << info.iterations << " iterations!" << std::endl;
    if (lynch_method) {
        std::ios_base::fmtflags f = log.flags();
        log << std::scientific;
        log << "# heterozygosity: " << info.heterozygosity << std::endl;
        log << "# error: " << info.error_rate << std::endl;
        log.flags(f);
    }
}

std::vector<OutputRecord> to_records(const Rows& r) {
    std::vector<OutputRecord> out;
    out.reserve(r.rows);
    const char* p = r.csv.p;
    const char* end = p + r.bytes;
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        if (!nl) nl = end;
        // split at the LAST six commas: the chromosome name may itself contain commas
        const char* f[7];
        int k = 6;
        f[6] = nl;
        for (const char* q = nl - 1; q >= p && k > 0; --q) if (*q == ',') f[--k] = q;
        if (k == 0) {
            OutputRecord rec;
            rec.chromosome_name.assign(p, f[0]);
            rec.position = std::atoi(std::string(f[0] + 1, f[1]).c_str());
            rec.classification.label.assign(f[1] + 1, f[2]);
            rec.classification.genotype.assign(f[2] + 1, f[3]);
            rec.classification.confidence_homozygous = std::strtod(std::string(f[3] + 1, f[4]).c_str(), nullptr);
            rec.classification.confidence_heterozygous = std::strtod(std::string(f[4] + 1, f[5]).c_str(), nullptr);
            rec.classification.confidence_type.assign(f[5] + 1, nl);
            out.push_back(std::move(rec));
        }
        p = nl + 1;
    }
    return out;
}

std::vector<OutputRecord> call_stream(int method, std::istream& in, bool estimate_prior, double prior, double error_threshold,
                                      double significance_level) {
    Pinned text;
    const size_t len = slurp(in, text);
    Rows rows;
    SidRunInfo info;
    run(method, text.p, len, estimate_prior, prior, error_threshold, significance_level, rows, info);
    log_fit(method, info, std::cerr);
    return to_records(rows);
}

// Device-side record of one parsed text, copied back.
struct Parsed {
    std::vector<uint64_t> profile;
    std::vector<int32_t> pos;
    std::vector<uint32_t> name_ref;
    std::vector<char> names;
};

Parsed tokenize(const char* text, size_t len, bool want_qual) {
    Parsed r;
    char* d = nullptr;
    check(sidgpu_malloc(ctx().h, ((len + 15) & ~(size_t)15) + 16, (void**)&d));
    struct Free { char* d; ~Free() { sidgpu_free(ctx().h, d); } } guard {d};
    if (len) check(sidgpu_memcpy_h2d(ctx().h, d, text, len));
    sidgpu_sites_view v {};
    check(sidgpu_tokenize(ctx().h, d, len, 0, len, want_qual ? 1 : 0, &v));
    r.profile.resize(v.n_sites);
    r.pos.resize(v.n_sites);
    r.name_ref.resize(v.n_sites);
    r.names.resize(v.names_bytes);
    if (v.n_sites) {
        check(sidgpu_memcpy_d2h(ctx().h, r.profile.data(), v.d_profile, v.n_sites * 8));
        check(sidgpu_memcpy_d2h(ctx().h, r.pos.data(), v.d_pos, v.n_sites * 4));
        check(sidgpu_memcpy_d2h(ctx().h, r.name_ref.data(), v.d_name_ref, v.n_sites * 4));
    }
    if (v.names_bytes) check(sidgpu_memcpy_d2h(ctx().h, r.names.data(), v.d_names, v.names_bytes));
    return r;
}

profile_t unpack(uint64_t p) {
    return {uint16_t(p), uint16_t(p >> 16), uint16_t(p >> 32), uint16_t(p >> 48)};
}

}  // namespace

void sidSetDevice(int device, size_t max_chunk_bytes) {
    g_device = device;
    g_chunk = max_chunk_bytes;
}

void sidSetHetOnly(bool het_only) { g_het_only = het_only; }
void sidSetHostInflate(bool on) { g_host_inflate = on; }

// ---- call.hpp:40-43 ------------------------------------------------------------------------------
std::vector<OutputRecord> callSiteMLError(std::istream& in, const bool estimate_prior, double prior, double error_threshold,
                                          const double significance_level) {
    return call_stream(SIDGPU_METHOD_LOCAL, in, estimate_prior, prior, error_threshold, significance_level);
}
std::vector<OutputRecord> callBayes(std::istream& in) { return call_stream(SIDGPU_METHOD_BAYES, in, false, -1, 0.1, 0.05); }
std::vector<OutputRecord> callLikelihoodRatio(std::istream& in, const bool use_prior, const double significance_level) {
    return call_stream(SIDGPU_METHOD_LIKELIHOOD_RATIO, in, use_prior, -1, 0.1, significance_level);
}
std::vector<OutputRecord> callQualityBasedSimple(std::istream& in, const bool estimate_prior, double prior, const double significance_level) {
    return call_stream(SIDGPU_METHOD_QUALITY, in, estimate_prior, prior, 0.1, significance_level);
}

SidRunInfo sidCallToStream(const std::string& method, const char* text, size_t len, bool estimate_prior, double prior,
                           double error_threshold, double significance_level, std::ostream& out, std::ostream& log,
                           const char* header) {
    SidRunInfo info;
    const int m = method_id(method);
    if (m < 0) {                                    // sid.cpp:92-100: unknown methods print the header only
        if (header) out << header << std::endl;
        return info;
    }
    Rows rows;
    Pinned staged;
    staged.reserve(len + 16);
    std::memcpy(staged.p, text, len);               // pinned staging: the H2D copies overlap with the kernels
    run(m, staged.p, len, estimate_prior, prior, error_threshold, significance_level, rows, info);
    log_fit(m, info, log);
    if (header) out << header << std::endl;
    out.write(rows.csv.p, (std::streamsize)rows.bytes);
    return info;
}

namespace {

// ---- callbacks of sidgpu_call_io over file descriptors
struct FileIo {
    int fd_in = -1, fd_out = -1;
    bool gzip = false;
    gzFile gz = nullptr;
    bgzf::Reader* blocks = nullptr;     // a BGZF file: its blocks are inflated side by side (bgzf.hpp)
    off_t offset = 0;                   // next byte of a regular file
    bool seekable = false;
    int threads = 1;
    const char* header = nullptr;
    bool header_done = false;
    bool write_failed = false;

    static int64_t read_cb(void* user, char* dst, size_t cap) {
        FileIo& f = *(FileIo*)user;
        if (f.blocks) return f.blocks->read(dst, cap);
        if (f.gzip) {
            const int got = gzread(f.gz, dst, (unsigned)std::min<size_t>(cap, (size_t)1 << 30));
            return got < 0 ? -1 : got;
        }
        if (!f.seekable || f.threads <= 1 || cap < ((size_t)8 << 20)) {
            const ssize_t got = f.seekable ? pread(f.fd_in, dst, cap, f.offset) : read(f.fd_in, dst, cap);
            if (got > 0) f.offset += got;
            return got < 0 ? -1 : (int64_t)got;
        }
        // a regular file: the slot is filled by parallel preads (one kernel copy per thread)
        const size_t part = (cap / (size_t)f.threads + 4095) & ~(size_t)4095;
        std::vector<ssize_t> got((size_t)f.threads, 0);
        std::vector<std::thread> workers;
        for (int t = 0; t < f.threads; ++t) {
            const size_t lo = std::min(cap, part * (size_t)t), hi = std::min(cap, part * (size_t)(t + 1));
            workers.emplace_back([&, t, lo, hi]() {
                size_t done = 0;
                while (lo + done < hi) {
                    const ssize_t g = pread(f.fd_in, dst + lo + done, hi - lo - done, f.offset + (off_t)(lo + done));
                    if (g < 0) { got[(size_t)t] = -1; return; }
                    if (g == 0) break;
                    done += (size_t)g;
                }
                got[(size_t)t] = (ssize_t)done;
            });
        }
        for (auto& w : workers) w.join();
        size_t total = 0;
        for (int t = 0; t < f.threads; ++t) {
            if (got[(size_t)t] < 0) return -1;
            total += (size_t)got[(size_t)t];
            const size_t lo = std::min(cap, part * (size_t)t), hi = std::min(cap, part * (size_t)(t + 1));
            if ((size_t)got[(size_t)t] < hi - lo) break;       // end of the file inside this part
        }
        f.offset += (off_t)total;
        return (int64_t)total;
    }
    bool put(const char* p, size_t n) {
        while (n) {
            const ssize_t w = write(fd_out, p, n);
            if (w < 0) { write_failed = true; return false; }
            p += w;
            n -= (size_t)w;
        }
        return true;
    }
    bool put_header() {
        if (header_done || !header) return true;
        header_done = true;
        return put(header, std::strlen(header)) && put("\n", 1);
    }
    static int write_cb(void* user, const char* rows, size_t n) {
        FileIo& f = *(FileIo*)user;
        return f.put_header() && f.put(rows, n) ? 0 : 1;
    }
    static int rewind_cb(void* user) {
        FileIo& f = *(FileIo*)user;
        if (f.blocks) { f.blocks->rewind(); return 0; }
        if (f.gzip) return gzrewind(f.gz) == 0 ? 0 : 1;
        if (!f.seekable) return 1;
        f.offset = 0;
        return 0;
    }
};

// lexicographic order of std::array<uint16_t,4> (pileup.cpp:179-182) for a packed profile
uint64_t profile_sort_key(uint64_t p) {
    return ((p & 0xFFFFull) << 48) | (((p >> 16) & 0xFFFFull) << 32) | (((p >> 32) & 0xFFFFull) << 16) | (p >> 48);
}

struct Shard {
    sidgpu_ctx* h = nullptr;
    char* csv = nullptr;
    size_t cap = 0;
    uint64_t bytes = 0, sites = 0, rows = 0;
    int rc = SIDGPU_OK;
    std::string err;
    void fail(int code) { rc = code; err = h ? sidgpu_last_error(h) : sidgpu_last_error(nullptr); }
};

template <class F>
void on_every_shard(std::vector<Shard>& shard, F f) {
    std::vector<std::thread> workers;
    for (size_t k = 0; k < shard.size(); ++k) workers.emplace_back([&, k]() { if (shard[k].rc == SIDGPU_OK) f(k, shard[k]); });
    for (auto& w : workers) w.join();
}

}  // namespace

SidRunInfo sidCallToStreamSharded(const std::string& method, const char* text, size_t len, bool estimate_prior, double prior,
                                  double error_threshold, double significance_level, const std::vector<int>& devices, std::ostream& out,
                                  std::ostream& log, const char* header) {
    SidRunInfo info;
    const int m = method_id(method);
    if (m < 0) {                                    // sid.cpp:92-100: unknown methods print the header only
        if (header) out << header << std::endl;
        return info;
    }
    const size_t n = devices.size();
    if (n == 0) throw std::runtime_error("no devices given");
    const bool streams = (m == SIDGPU_METHOD_LOCAL || m == SIDGPU_METHOD_QUALITY) && !estimate_prior;
    // shard k owns the lines whose first byte lies in [cut[k], cut[k+1]): cuts are moved to line starts
    std::vector<size_t> cut(n + 1, len);
    cut[0] = 0;
    for (size_t k = 1; k < n; ++k) {
        size_t c = std::max(cut[k - 1], len / n * k);
        if (c > 0 && c < len && text[c - 1] != '\n') {
            const void* nl = std::memchr(text + c, '\n', len - c);
            c = nl ? (size_t)((const char*)nl - text) + 1 : len;
        }
        cut[k] = std::min(c, len);
    }
    sidgpu_params p {};
    p.method = m;
    p.estimate_prior = estimate_prior ? 1 : 0;
    p.prior = prior;
    p.error_threshold = error_threshold;
    p.significance_level = significance_level;
    p.het_only = g_het_only ? 1 : 0;
    std::vector<Shard> shard(n);
    auto alloc_csv = [](Shard& s, size_t cap) {
        if (s.csv) { sidgpu_free_host(s.h, s.csv); s.csv = nullptr; }
        s.cap = cap;
        const int rc = sidgpu_malloc_host(s.h, cap, (void**)&s.csv);
        if (rc != SIDGPU_OK) s.fail(rc);
        return rc == SIDGPU_OK;
    };
    // ---- every shard: its own ctx, its text in
    on_every_shard(shard, [&](size_t k, Shard& s) {
        sidgpu_config cfg {};
        cfg.device = devices[k];
        cfg.max_chunk_bytes = g_chunk;
        int rc = sidgpu_create(&cfg, &s.h);
        if (rc != SIDGPU_OK) { s.fail(rc); return; }
        const size_t bytes_in = cut[k + 1] - cut[k];
        if (streams) {                              // no exchange between shards: the whole call at once
            size_t cap = bytes_in + bytes_in / 2 + 4096;
            for (;;) {
                if (!alloc_csv(s, cap)) return;
                rc = sidgpu_call_host(s.h, &p, text + cut[k], bytes_in, s.csv, s.cap, &s.bytes, &s.sites, &s.rows);
                if (rc == SIDGPU_ECAPACITY && s.bytes > s.cap) { cap = s.bytes + 4096; continue; }
                break;
            }
        } else {
            rc = sidgpu_begin(s.h, &p);
            if (rc == SIDGPU_OK) rc = sidgpu_feed_host(s.h, text + cut[k], bytes_in, &s.sites);
        }
        if (rc != SIDGPU_OK) s.fail(rc);
    });
    auto first_error = [&]() {
        for (auto& s : shard) if (s.rc != SIDGPU_OK) return &s;      // first failing shard in file order
        return (Shard*)nullptr;
    };
    if (!streams && !first_error()) {
        // ---- one fit for all shards: integer nucleotide sums, then Nelder-Mead on the sum of the shards' objectives
        //      (estimateProfileGenotypeLikelihoods lynch.cpp:17-35; the objective is linear in the profile counts)
        std::vector<sidgpu_unique_view> view(n);
        uint64_t sums[5] = {0, 0, 0, 0, 0};
        for (size_t k = 0; k < n && !first_error(); ++k) {
            const int rc = sidgpu_histogram(shard[k].h, 4, &view[k]);
            if (rc != SIDGPU_OK) { shard[k].fail(rc); break; }
            for (int i = 0; i < 5; ++i) sums[i] += view[k].nd_sums[i];
        }
        double nd[4] = {0.25, 0.25, 0.25, 0.25};                     // pileup.cpp:209-215
        if (sums[4]) for (int i = 0; i < 4; ++i) nd[i] = (double)sums[i] / (double)sums[4];
        // the merged unique profiles (countUniqueProfiles of the whole genome): BH of likelihood_ratio ranks over them
        std::vector<uint64_t> merged;
        for (size_t k = 0; k < n && !first_error(); ++k) {
            std::vector<uint64_t> prof(view[k].n_unique);
            if (!prof.empty()) {
                const int rc = sidgpu_memcpy_d2h(shard[k].h, prof.data(), view[k].d_profile, prof.size() * 8);
                if (rc != SIDGPU_OK) { shard[k].fail(rc); break; }
            }
            merged.insert(merged.end(), prof.begin(), prof.end());
        }
        std::sort(merged.begin(), merged.end(), [](uint64_t a, uint64_t b) { return profile_sort_key(a) < profile_sort_key(b); });
        merged.erase(std::unique(merged.begin(), merged.end()), merged.end());
        if (!first_error()) {
            auto objective = [&](double pi, double eps) {
                if (pi < 0 || pi > 1 || eps < 0 || eps > 1) return std::numeric_limits<double>::max();     // lynch.cpp:41-43
                double f = 0;
                for (size_t k = 0; k < n; ++k) {
                    double fk = 0;
                    const int rc = sidgpu_lynch_objective(shard[k].h, nd, pi, eps, &fk);
                    if (rc != SIDGPU_OK) { shard[k].fail(rc); return std::numeric_limits<double>::max(); }
                    f += fk;
                }
                return f;
            };
            const double x0[2] = {1e-3, 1e-3}, step[2] = {1e-4, 1e-4};                                    // lynch.cpp:8-10,20
            const sid::NelderMeadResult r = sid::nelder_mead_2d(objective, x0, step);
            info.has_fit = true;
            info.heterozygosity = r.x[0];
            info.error_rate = r.x[1];
            info.iterations = r.iterations;
            info.converged = r.converged;
            info.unique_profiles = merged.size();
            for (size_t k = 0; k < n && !first_error(); ++k) {
                const int rc = sidgpu_set_fit(shard[k].h, r.x[0], r.x[1], nd);
                if (rc != SIDGPU_OK) shard[k].fail(rc);
            }
        }
        // ---- classification with the shared fit, rows out
        on_every_shard(shard, [&](size_t k, Shard& s) {
            int rc = m == SIDGPU_METHOD_LIKELIHOOD_RATIO ? sidgpu_finish_global(s.h, merged.data(), merged.size()) : sidgpu_finish(s.h);
            size_t cap = (size_t)s.sites * 48 + 4096;
            while (rc == SIDGPU_OK) {
                if (!alloc_csv(s, cap)) return;
                if (m == SIDGPU_METHOD_QUALITY)     // quality -R keeps no sites: its second pass over the shard's text
                    rc = sidgpu_stream_host(s.h, text + cut[k], cut[k + 1] - cut[k], s.csv, s.cap, &s.bytes, &s.sites, &s.rows);
                else
                rc = sidgpu_emit_host(s.h, s.csv, s.cap, &s.bytes, &s.rows);
                if (rc == SIDGPU_ECAPACITY && s.bytes > s.cap) { cap = s.bytes + 4096; rc = SIDGPU_OK; continue; }
                break;
            }
            if (rc != SIDGPU_OK) s.fail(rc);
        });
    }
    int rc = SIDGPU_OK;
    std::string err;
    if (Shard* bad = first_error()) { rc = bad->rc; err = bad->err; }
    if (rc == SIDGPU_OK) {
        log_fit(m, info, log);
        if (header) out << header << std::endl;
        for (auto& s : shard) {
            out.write(s.csv, (std::streamsize)s.bytes);
            info.n_sites += s.sites;
            info.n_rows += s.rows;
        }
    }
    for (auto& s : shard) {
        if (s.h && s.csv) sidgpu_free_host(s.h, s.csv);
        if (s.h) sidgpu_destroy(s.h);
    }
    if (rc == SIDGPU_EMALFORMED) throw std::invalid_argument("Malformed pileup line");
    if (rc == SIDGPU_EMISSING_MAPQ) throw std::invalid_argument("Malformed pileup line or missing mapping qualities");
    if (rc == SIDGPU_EQUAL_SHORT) throw std::invalid_argument("Malformed pileup line: " + err);
    if (rc != SIDGPU_OK) throw std::runtime_error("sidgpu error " + std::to_string(rc) + ": " + err);
    return info;
}

SidRunInfo sidCallFile(const std::string& method, int fd_in, bool gzip, bool estimate_prior, double prior, double error_threshold,
                       double significance_level, int fd_out, std::ostream& log, const char* header, int read_threads) {
    SidRunInfo info;
    FileIo f;
    f.fd_in = fd_in;
    f.fd_out = fd_out;
    f.gzip = gzip;
    f.header = header;
    f.threads = std::max(1, read_threads);
    f.seekable = !gzip && lseek(fd_in, 0, SEEK_CUR) != (off_t)-1;
    const int m = method_id(method);
    if (m < 0) {                                    // sid.cpp:92-100: unknown methods print the header only
        f.put_header();
        return info;
    }
    std::unique_ptr<bgzf::Reader> blocks;
    bool device_inflate = false;
    if (gzip) {
        unsigned char head[64];
        const ssize_t got = pread(fd_in, head, sizeof head, 0);       // fails on a pipe: plain gzip stream then
        if (got >= 18 && bgzf::looks_like(head, (size_t)got)) {
            if (g_host_inflate) {
                blocks.reset(new bgzf::Reader(fd_in, f.threads));
                f.blocks = blocks.get();
            } else {
                // the file's own bytes go to the device (parallel preads like a plain text file); it inflates the members
                device_inflate = true;
                f.gzip = false;
                f.seekable = true;
            }
        } else {
            f.gz = gzdopen(dup(fd_in), "rb");
            if (!f.gz) throw std::runtime_error("could not open the gzip stream");
            gzbuffer(f.gz, 1u << 20);
        }
    }
    sidgpu_params p {};
    p.method = m;
    p.estimate_prior = estimate_prior ? 1 : 0;
    p.prior = prior;
    p.error_threshold = error_threshold;
    p.significance_level = significance_level;
    p.het_only = g_het_only ? 1 : 0;
    sidgpu_io io {FileIo::read_cb, FileIo::write_cb, FileIo::rewind_cb, &f};
    uint64_t bytes = 0;
    const auto t0 = std::chrono::steady_clock::now();
    sidgpu_ctx* h = ctx().h;                        // creates the ctx on first use: CUDA context, kernels, tables
    const auto t1 = std::chrono::steady_clock::now();
    const int rc = device_inflate ? sidgpu_call_io_bgzf(h, &p, &io, &bytes, &info.n_sites, &info.n_rows)
                                  : sidgpu_call_io(h, &p, &io, &bytes, &info.n_sites, &info.n_rows);
    const auto t2 = std::chrono::steady_clock::now();
    if (getenv("SID_TIMING")) {                     // where the wall clock and the memory of a run go (bench.py cli_e2e)
        long hwm_kb = -1;                           // VmHWM of this process image (ru_maxrss would carry the parent's peak over fork)
        if (FILE* st = std::fopen("/proc/self/status", "r")) {
            char line[256];
            while (std::fgets(line, sizeof line, st)) if (std::sscanf(line, "VmHWM: %ld kB", &hwm_kb) == 1) break;
            std::fclose(st);
        }
        log << "# timing: device setup " << std::chrono::duration<double>(t1 - t0).count() << " s, streaming " << bytes << " CSV bytes "
            << std::chrono::duration<double>(t2 - t1).count() << " s, peak RSS " << hwm_kb / 1024 << " MB" << std::endl;
    }
    if (f.gz) gzclose(f.gz);
    if (rc != SIDGPU_OK) {
        if (f.write_failed) throw std::runtime_error("could not write the rows");
        if (f.blocks && !f.blocks->error().empty()) throw std::runtime_error("could not inflate the file: " + f.blocks->error());
        raise(rc);
    }
    sidgpu_fit fit {};
    double nd[4];
    uint64_t nu = 0;
    if (sidgpu_session_fit(ctx().h, &fit, nd, &nu) == SIDGPU_OK) {
        info.has_fit = true;
        info.heterozygosity = fit.pi;
        info.error_rate = fit.eps;
        info.iterations = fit.iterations;
        info.converged = fit.converged != 0;
        info.unique_profiles = nu;
    }
    log_fit(m, info, log);
    f.put_header();                                 // an input without a single row still prints the header (sid.cpp:102)
    return info;
}

// ---- pileup.hpp / call.hpp:12 ----------------------------------------------------------------------
void fillReadVectors(std::vector<PileupLine>& lines, const char* text, size_t len, const std::vector<uint64_t>& starts,
                     bool parse_base_qualities, bool parse_mapping_qualities);

std::vector<PileupLine> parsePileupText(const char* text, size_t len, bool parse_base_qualities, bool parse_mapping_qualities) {
    const Parsed p = tokenize(text, len, false);
    std::vector<PileupLine> out(p.pos.size());
    std::vector<uint64_t> starts;
    starts.reserve(out.size());
    // reference base: the third column; recovered from the text only to fill the struct (1 byte per line)
    size_t line = 0, i = 0;
    while (i < len && line < out.size()) {
        const char* nl = (const char*)std::memchr(text + i, '\n', len - i);
        const size_t e = nl ? (size_t)(nl - text) : len;
        if (e > i) {
            PileupLine& l = out[line];
            const uint32_t r = p.name_ref[line];
            const uint32_t nlen = (uint8_t)p.names[r] | ((uint32_t)(uint8_t)p.names[r + 1] << 8);
            l.chromosome_name.assign(p.names.data() + r + 2, nlen);
            l.position = p.pos[line];
            l.base_counts = unpack(p.profile[line]);
            size_t q = i;
            for (int col = 0; col < 2; ++col) {     // skip two columns
                while (q < e && (text[q] == ' ' || text[q] == '\t')) ++q;
                while (q < e && text[q] != ' ' && text[q] != '\t') ++q;
            }
            while (q < e && (text[q] == ' ' || text[q] == '\t')) ++q;
            if (q < e) l.reference_base = text[q];
            starts.push_back(i);
            ++line;
        }
        i = e + 1;
    }
    // the per-read vectors (pileup.cpp:42-44: bases and strands always; :53-66: the qualities when asked)
    fillReadVectors(out, text, len, starts, parse_base_qualities, parse_mapping_qualities);
    return out;
}

std::vector<PileupLine> readFile(std::istream& in, bool parse_base_qualities, bool parse_mapping_qualities) {
    std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    return parsePileupText(text.data(), text.size(), parse_base_qualities, parse_mapping_qualities);
}

PileupLine parsePileupLine(char* line, bool parse_base_qualities, bool parse_mapping_qualities) {
    const size_t n = std::strlen(line);
    std::vector<PileupLine> v = parsePileupText(line, n, parse_base_qualities, parse_mapping_qualities);
    if (v.empty()) throw std::invalid_argument("Malformed pileup line");
    return v[0];
}

ReadStack parseReadBases(const char* read_bases, char reference, int coverage) {
    (void)coverage;                                 // the reference only uses it as a reserve() hint (pileup.cpp:72-73)
    ReadStack r;
    r.counts = {0, 0, 0, 0};
    if (read_bases[0] == '\0') return r;            // an empty bases string is not a valid column; counts are zero
    const std::string line = std::string("x\t1\t") + reference + "\t0\t" + read_bases + "\n";
    const std::vector<PileupLine> v = parsePileupText(line.data(), line.size(), false, false);
    if (!v.empty()) {
        r.counts = v[0].base_counts;
        r.bases = v[0].bases;
        r.strands = v[0].strands;
    }
    return r;
}

std::vector<UniqueProfile> countUniqueProfiles(const std::vector<PileupLine>& pileup) {
    std::vector<UniqueProfile> out;
    if (pileup.empty()) return out;
    std::vector<uint64_t> prof(pileup.size());
    for (size_t i = 0; i < pileup.size(); ++i) {
        const profile_t& c = pileup[i].base_counts;
        prof[i] = (uint64_t)c[0] | ((uint64_t)c[1] << 16) | ((uint64_t)c[2] << 32) | ((uint64_t)c[3] << 48);
    }
    uint64_t* d = nullptr;
    check(sidgpu_malloc(ctx().h, prof.size() * 8, (void**)&d));
    struct Free { void* d; ~Free() { sidgpu_free(ctx().h, d); } } guard {d};
    check(sidgpu_memcpy_h2d(ctx().h, d, prof.data(), prof.size() * 8));
    sidgpu_unique_view v {};
    check(sidgpu_count_unique(ctx().h, d, prof.size(), 0, &v));
    std::vector<uint64_t> up(v.n_unique), uc(v.n_unique);
    if (v.n_unique) {
        check(sidgpu_memcpy_d2h(ctx().h, up.data(), v.d_profile, v.n_unique * 8));
        check(sidgpu_memcpy_d2h(ctx().h, uc.data(), v.d_count, v.n_unique * 8));
    }
    out.reserve(v.n_unique);
    for (size_t i = 0; i < up.size(); ++i) out.emplace_back(unpack(up[i]), (uint32_t)uc[i]);
    return out;
}

std::array<double, 4> computeNucleotideDistribution(const std::vector<UniqueProfile>& profiles) {
    if (profiles.empty()) return {0.25, 0.25, 0.25, 0.25};
    // weight every profile by its count: expand to (profile, count) pairs on the device
    std::vector<uint64_t> prof(profiles.size()), cnt(profiles.size());
    for (size_t i = 0; i < profiles.size(); ++i) {
        const profile_t& c = profiles[i].profile;
        prof[i] = (uint64_t)c[0] | ((uint64_t)c[1] << 16) | ((uint64_t)c[2] << 32) | ((uint64_t)c[3] << 48);
        cnt[i] = profiles[i].count;
    }
    uint64_t *dp = nullptr, *dc = nullptr;
    check(sidgpu_malloc(ctx().h, prof.size() * 8, (void**)&dp));
    check(sidgpu_malloc(ctx().h, cnt.size() * 8, (void**)&dc));
    struct Free { void *a, *b; ~Free() { sidgpu_free(ctx().h, a); sidgpu_free(ctx().h, b); } } guard {dp, dc};
    check(sidgpu_memcpy_h2d(ctx().h, dp, prof.data(), prof.size() * 8));
    check(sidgpu_memcpy_h2d(ctx().h, dc, cnt.data(), cnt.size() * 8));
    sidgpu_unique_view v {};
    check(sidgpu_count_unique_weighted(ctx().h, dp, dc, prof.size(), 0, &v));
    return {v.nd[0], v.nd[1], v.nd[2], v.nd[3]};
}

// ---- the per-read vectors and the Lynch / statistics mirrors ---------------------------------------------------------
namespace {

struct Dev {                                        // a device allocation that goes away with the scope
    void* p = nullptr;
    explicit Dev(size_t bytes) { check(sidgpu_malloc(ctx().h, std::max<size_t>(bytes, 16), &p)); }
    ~Dev() { sidgpu_free(ctx().h, p); }
    Dev(const Dev&) = delete;
    Dev& operator=(const Dev&) = delete;
    template <class T> T* as() const { return (T*)p; }
};

template <class T>
std::vector<uint64_t> offsets_of(const std::vector<T>& counts) {          // exclusive prefix sums (index arithmetic)
    std::vector<uint64_t> off(counts.size() + 1, 0);
    for (size_t i = 0; i < counts.size(); ++i) off[i + 1] = off[i] + counts[i];
    return off;
}

std::vector<uint64_t> pack_profiles(const std::vector<UniqueProfile>& profiles) {
    std::vector<uint64_t> p(profiles.size());
    for (size_t i = 0; i < profiles.size(); ++i) {
        const profile_t& c = profiles[i].profile;
        p[i] = (uint64_t)c[0] | ((uint64_t)c[1] << 16) | ((uint64_t)c[2] << 32) | ((uint64_t)c[3] << 48);
    }
    return p;
}

}  // namespace

// Fills bases / strands (always) and the quality vectors (when asked) of the lines of `text`; starts[i] = offset of line i.
void fillReadVectors(std::vector<PileupLine>& lines, const char* text, size_t len, const std::vector<uint64_t>& starts,
                     bool parse_base_qualities, bool parse_mapping_qualities) {
    const size_t n = lines.size();
    if (n == 0) return;
    Dev d_text(((len + 15) & ~(size_t)15) + 16), d_off(n * 8), d_nb(n * 4), d_nq(n * 4), d_nm(n * 4);
    check(sidgpu_memcpy_h2d(ctx().h, d_text.p, text, len));
    check(sidgpu_memcpy_h2d(ctx().h, d_off.p, starts.data(), n * 8));
    check(sidgpu_read_counts(ctx().h, d_text.as<char>(), len, d_off.as<uint64_t>(), n, parse_base_qualities ? 1 : 0, parse_mapping_qualities ? 1 : 0,
                             d_nb.as<uint32_t>(), d_nq.as<uint32_t>(), d_nm.as<uint32_t>()));
    std::vector<uint32_t> nb(n), nq(n), nm(n);
    check(sidgpu_memcpy_d2h(ctx().h, nb.data(), d_nb.p, n * 4));
    check(sidgpu_memcpy_d2h(ctx().h, nq.data(), d_nq.p, n * 4));
    check(sidgpu_memcpy_d2h(ctx().h, nm.data(), d_nm.p, n * 4));
    const std::vector<uint64_t> ob = offsets_of(nb), oq = offsets_of(nq), om = offsets_of(nm);
    Dev d_ob(n * 8), d_oq(n * 8), d_om(n * 8), d_bases(ob[n]), d_str(ob[n]), d_bq(oq[n]), d_mq(om[n]);
    check(sidgpu_memcpy_h2d(ctx().h, d_ob.p, ob.data(), n * 8));
    check(sidgpu_memcpy_h2d(ctx().h, d_oq.p, oq.data(), n * 8));
    check(sidgpu_memcpy_h2d(ctx().h, d_om.p, om.data(), n * 8));
    check(sidgpu_read_fill(ctx().h, d_text.as<char>(), len, d_off.as<uint64_t>(), n, d_ob.as<uint64_t>(), d_oq.as<uint64_t>(), d_om.as<uint64_t>(),
                           d_bases.as<char>(), d_str.as<uint8_t>(), parse_base_qualities ? d_bq.as<uint8_t>() : nullptr,
                           parse_mapping_qualities ? d_mq.as<uint8_t>() : nullptr));
    std::vector<char> bases(ob[n]);
    std::vector<uint8_t> str(ob[n]), bq(parse_base_qualities ? oq[n] : 0), mq(parse_mapping_qualities ? om[n] : 0);
    if (ob[n]) {
        check(sidgpu_memcpy_d2h(ctx().h, bases.data(), d_bases.p, ob[n]));
        check(sidgpu_memcpy_d2h(ctx().h, str.data(), d_str.p, ob[n]));
    }
    if (!bq.empty()) check(sidgpu_memcpy_d2h(ctx().h, bq.data(), d_bq.p, bq.size()));
    if (!mq.empty()) check(sidgpu_memcpy_d2h(ctx().h, mq.data(), d_mq.p, mq.size()));
    for (size_t i = 0; i < n; ++i) {
        PileupLine& l = lines[i];
        l.bases.assign(bases.begin() + (ptrdiff_t)ob[i], bases.begin() + (ptrdiff_t)ob[i + 1]);
        l.strands.resize(nb[i]);
        for (uint32_t k = 0; k < nb[i]; ++k) l.strands[k] = str[ob[i] + k] != 0;
        if (parse_base_qualities) l.base_qualities.assign(bq.begin() + (ptrdiff_t)oq[i], bq.begin() + (ptrdiff_t)oq[i + 1]);
        if (parse_mapping_qualities) l.mapping_qualities.assign(mq.begin() + (ptrdiff_t)om[i], mq.begin() + (ptrdiff_t)om[i + 1]);
    }
}

std::vector<uint8_t> parseQualities(const char* base_qualities, int coverage) {
    (void)coverage;                                 // a reserve() hint in the reference (pileup.cpp:157)
    const size_t n = std::strlen(base_qualities);
    std::vector<uint8_t> out;
    if (n == 0) return out;
    Dev d_in(n), d_out(n);
    check(sidgpu_memcpy_h2d(ctx().h, d_in.p, base_qualities, n));
    uint64_t m = 0;
    check(sidgpu_qualities(ctx().h, d_in.as<char>(), n, d_out.as<uint8_t>(), &m));
    out.resize(m);
    if (m) check(sidgpu_memcpy_d2h(ctx().h, out.data(), d_out.p, m));
    return out;
}

// ---- lynch.hpp:44-46 ---------------------------------------------------------------------------------------------------
ProfileGenotypeLikelihoods estimateProfileGenotypeLikelihoods(const std::vector<UniqueProfile>& profiles, const std::array<double, 4> nd) {
    ProfileGenotypeLikelihoods r {0, 0, {}};
    const size_t n = profiles.size();
    if (n == 0) return r;
    const std::vector<uint64_t> prof = pack_profiles(profiles);
    std::vector<uint64_t> cnt(n);
    for (size_t i = 0; i < n; ++i) cnt[i] = profiles[i].count;
    Dev d_p(n * 8), d_c(n * 8), d_h(n * 8), d_t(n * 8);
    check(sidgpu_memcpy_h2d(ctx().h, d_p.p, prof.data(), n * 8));
    check(sidgpu_memcpy_h2d(ctx().h, d_c.p, cnt.data(), n * 8));
    sidgpu_unique_view v {};
    check(sidgpu_count_unique_weighted(ctx().h, d_p.as<uint64_t>(), d_c.as<uint64_t>(), n, 0, &v));     // the histogram the fit runs on
    sidgpu_fit fit {};
    check(sidgpu_lynch_fit(ctx().h, nd.data(), &fit));
    r.heterozygosity = fit.pi;
    r.error_rate = fit.eps;
    check(sidgpu_profile_loglik(ctx().h, d_p.as<uint64_t>(), n, nd.data(), fit.eps, d_h.as<double>(), d_t.as<double>()));
    std::vector<double> lh(n), lt(n);
    check(sidgpu_memcpy_d2h(ctx().h, lh.data(), d_h.p, n * 8));
    check(sidgpu_memcpy_d2h(ctx().h, lt.data(), d_t.p, n * 8));
    r.profile_likelihoods.resize(n);
    for (size_t i = 0; i < n; ++i) r.profile_likelihoods[i] = {expl((long double)lh[i]), expl((long double)lt[i])};
    return r;
}

double compoundLikelihood(double pi, double epsilon, const std::vector<UniqueProfile>& profiles, const std::array<double, 4> nd) {
    const size_t n = profiles.size();
    if (n == 0) return (pi < 0 || pi > 1 || epsilon < 0 || epsilon > 1) ? std::numeric_limits<double>::max() : 0.0;
    const std::vector<uint64_t> prof = pack_profiles(profiles);
    std::vector<uint64_t> cnt(n);
    for (size_t i = 0; i < n; ++i) cnt[i] = profiles[i].count;
    Dev d_p(n * 8), d_c(n * 8);
    check(sidgpu_memcpy_h2d(ctx().h, d_p.p, prof.data(), n * 8));
    check(sidgpu_memcpy_h2d(ctx().h, d_c.p, cnt.data(), n * 8));
    sidgpu_unique_view v {};
    check(sidgpu_count_unique_weighted(ctx().h, d_p.as<uint64_t>(), d_c.as<uint64_t>(), n, 0, &v));
    double f = 0;
    check(sidgpu_lynch_objective(ctx().h, nd.data(), pi, epsilon, &f));
    return f;
}

// ---- stats.hpp:8,11 ----------------------------------------------------------------------------------------------------
double likelihoodRatioTest(long double l_H0, long double l_H1) {
    // the device works on log-likelihoods (-inf: l == 0)
    const double a = l_H0 > 0 ? (double)logl(l_H0) : -std::numeric_limits<double>::infinity();
    const double b = l_H1 > 0 ? (double)logl(l_H1) : -std::numeric_limits<double>::infinity();
    Dev d_a(8), d_b(8), d_p(8);
    check(sidgpu_memcpy_h2d(ctx().h, d_a.p, &a, 8));
    check(sidgpu_memcpy_h2d(ctx().h, d_b.p, &b, 8));
    check(sidgpu_lr_test(ctx().h, d_a.as<double>(), d_b.as<double>(), 1, d_p.as<double>()));
    double p = 0;
    check(sidgpu_memcpy_d2h(ctx().h, &p, d_p.p, 8));
    return p;
}

std::vector<double> adjustBenjaminiHochberg(const std::vector<double>& p_values) {
    std::vector<double> out(p_values.size());
    if (p_values.empty()) return out;
    const size_t n = p_values.size();
    Dev d_p(n * 8), d_a(n * 8);
    check(sidgpu_memcpy_h2d(ctx().h, d_p.p, p_values.data(), n * 8));
    check(sidgpu_bh_adjust(ctx().h, d_p.as<double>(), n, d_a.as<double>()));
    check(sidgpu_memcpy_d2h(ctx().h, out.data(), d_a.p, n * 8));
    return out;
}
