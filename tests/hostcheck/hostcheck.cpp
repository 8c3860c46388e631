// TEST INFRASTRUCTURE ONLY.  Compiles the SID_HD (host+device) arithmetic of sid_b200/csrc/*.cuh
// with g++ so that the CPU test suite can check the exact code the kernels run -- tokenizer state
// machine, %g formatter, per-profile calls, Lynch terms -- against the oracle without a GPU.
// Nothing here is reachable from libsidgpu.so.
#include <cstdint>
#include <cstring>

#include "../../sid_b200/csrc/calls.cuh"
#include "../../sid_b200/csrc/fmt.cuh"
#include "../../sid_b200/csrc/parse.cuh"
#include "../../sid_b200/csrc/k_quality.cuh"
#include <cmath>
#include <vector>
#ifdef SID_HAVE_FAST
#include "parse_swar.hpp"
#include "../../sid_b200/csrc/parse_bits.cuh"
#endif

using namespace sid;

extern "C" {

struct hc_line {
    int32_t status, pos;
    uint64_t profile;
    uint32_t n_bases, chrom_off, chrom_len, bases_off, bases_len, bq_off, bq_len, mq_off, mq_len;
    int32_t ref;
};

static void fill(const ParsedLine& p, hc_line* o) {
    o->status = p.status; o->pos = p.pos; o->profile = p.profile; o->n_bases = p.n_bases;
    o->chrom_off = p.chrom_off; o->chrom_len = p.chrom_len; o->bases_off = p.bases_off; o->bases_len = p.bases_len;
    o->bq_off = p.bq_off; o->bq_len = p.bq_len; o->mq_off = p.mq_off; o->mq_len = p.mq_len; o->ref = (uint8_t)p.ref;
}

void hc_parse_line(const uint8_t* text, uint64_t len, uint64_t p, int want_qual, hc_line* out) {
    FlatSrc src {text, len};
    ParsedLine pl;
    parse_line(src, p, want_qual != 0, pl);
    fill(pl, out);
}

// quality_fields (the field scan k_quality uses once the tokenizer has stored the profile) against
// parse_line(want_qual) on every line of a text.  Returns the number of lines, or -(k+1) when line k differs.
int64_t hc_compare_quality_fields(const uint8_t* text, uint64_t len) {
    FlatSrc src {text, len};
    int64_t k = 0;
    for (uint64_t p = 0; p < len; ++p) {
        if (text[p] == '\n' || (p > 0 && text[p - 1] != '\n')) continue;
        ParsedLine a, b;
        parse_line(src, p, true, a);
        quality_fields(WordSrc {text, len}, p, a.profile, b);
        const bool same = a.status == b.status && a.ref == b.ref && a.chrom_off == b.chrom_off && a.chrom_len == b.chrom_len &&
                          a.bases_off == b.bases_off && a.bases_len == b.bases_len && a.bq_off == b.bq_off && a.bq_len == b.bq_len &&
                          a.mq_off == b.mq_off && a.mq_len == b.mq_len && (a.status == LINE_MALFORMED || a.n_bases == b.n_bases);
        if (!same) return -(k + 1);
        ++k;
    }
    return k;
}

// call_quality exactly as k_quality runs it (field scan with the profile given, then the per-read sums)
// over every line of a 7-column text.  Returns the number of lines, or -(k+1) when line k is not LINE_OK.
int64_t hc_call_quality(const uint8_t* text, uint64_t len, double prior, double alpha, int32_t* label, char* gt, double* hom, double* het) {
    std::vector<double> lut(4 * 256 + LOG_FACT_N);
    for (int n = 0; n < LOG_FACT_N; ++n) lut[1024 + n] = lgamma((double)n + 1.0);
    for (int q = 0; q < 256; ++q) {                       // same expressions as sidgpu_create (call.cpp:330-341)
        const double error = pow(10., q / -10.);
        lut[q] = log(1 - error);
        lut[256 + q] = log(error);
        lut[512 + q] = log(1 - 2. / 3. * error);
        lut[768 + q] = log(2. / 3. * error);
    }
    FlatSrc src {text, len};
    int64_t k = 0;
    for (uint64_t p = 0; p < len; ++p) {
        if (text[p] == '\n' || (p > 0 && text[p - 1] != '\n')) continue;
        ParsedLine full, pl;
        parse_line(src, p, true, full);
        quality_fields(WordSrc {text, len}, p, full.profile, pl);
        if (pl.status != LINE_OK) return -(k + 1);
        const CallResult r = call_quality(text, p, pl, lut.data(), prior, alpha);
        label[k] = r.label; gt[2 * k] = r.gt0; gt[2 * k + 1] = r.gt1; hom[k] = r.hom; het[k] = r.het;
        ++k;
    }
    return k;
}

#ifdef SID_HAVE_FAST
// The SWAR tokenizer the kernel uses.  Returns 1 when the fast grammar accepted the line.
int hc_parse_line_fast(const uint8_t* text, uint64_t len, uint64_t p, hc_line* out) {
    FastLine fl;
    memset(out, 0, sizeof *out);
    if (!parse_line_fast(text, len, p, fl)) return 0;
    out->status = fl.status; out->pos = fl.pos; out->profile = fl.profile;
    out->chrom_off = fl.chrom_off; out->chrom_len = fl.chrom_len;
    return 1;
}

// Parses every line of a text with both tokenizers.  Returns the number of lines, or -(k+1) when
// line k differs; *n_fast receives how many lines the fast grammar accepted.
int64_t hc_compare_parsers(const uint8_t* text, uint64_t len, uint64_t* n_fast) {
    FlatSrc src {text, len};
    int64_t k = 0;
    uint64_t fast = 0;
    for (uint64_t p = 0; p < len; ++p) {
        if (text[p] == '\n' || (p > 0 && text[p - 1] != '\n')) continue;
        ParsedLine a;
        parse_line(src, p, false, a);
        FastLine b;
        if (parse_line_fast(text, len, p, b)) {
            ++fast;
            if (a.status != LINE_OK || b.status != LINE_OK || a.profile != b.profile || a.pos != b.pos ||
                a.chrom_off != b.chrom_off || a.chrom_len != b.chrom_len) return -(k + 1);
        }
        ++k;
    }
    if (n_fast) *n_fast = fast;
    return k;
}

// Same comparison for the bit-parallel tokenizer (parse_bits.cuh).
int64_t hc_compare_parsers_bits(const uint8_t* text, uint64_t len, uint64_t* n_fast) {
    FlatSrc src {text, len};
    int64_t k = 0;
    uint64_t fast = 0;
    for (uint64_t p = 0; p < len; ++p) {
        if (text[p] == '\n' || (p > 0 && text[p - 1] != '\n')) continue;
        ParsedLine a;
        parse_line(src, p, false, a);
        FastLine b;
        if (parse_line_bits_host(text, len, p, b)) {
            ++fast;
            if (a.status != LINE_OK || b.status != LINE_OK || a.profile != b.profile || a.pos != b.pos ||
                a.chrom_off != b.chrom_off || a.chrom_len != b.chrom_len) return -(k + 1);
        }
        ++k;
    }
    if (n_fast) *n_fast = fast;
    return k;
}

// classify32 on one 32-byte unit: out[0..9] = term nl a c g t dot caret pm high
void hc_classify32(const uint8_t* bytes, uint32_t* out) {
    uint32_t w[8];
    memcpy(w, bytes, 32);
    const ClassWords k = classify32(w);
    out[0] = k.term; out[1] = k.nl; out[2] = k.a; out[3] = k.c; out[4] = k.g; out[5] = k.t; out[6] = k.dot; out[7] = k.caret;
    out[8] = k.pm; out[9] = k.high;
}
#endif

int hc_fmt_g6(double x, char* out) { int n = fmt_g6(x, out); out[n] = 0; return n; }
int hc_fmt_i32(int32_t v, char* out) { int n = fmt_i32(v, out); out[n] = 0; return n; }
int hc_fmt_i32_fast(int32_t v, char* out) { int n = fmt_i32_fast(v, out); out[n] = 0; return n; }
int hc_digits_i32(int32_t v) { return digits_i32(v); }

void hc_major_alleles(uint64_t profile, int* f, int* s) { major_alleles(profile, *f, *s); }

void hc_call_local(uint64_t profile, double prior, double E, double alpha, int* label, char* gt, double* hom, double* het) {
    CallResult r = call_local(profile, prior, E, alpha);
    *label = r.label; gt[0] = r.gt0; gt[1] = r.gt1; *hom = r.hom; *het = r.het;
}

void hc_call_bayes(uint64_t profile, const double nd[4], double pi, double eps, int* label, char* gt, double* hom, double* het) {
    LynchConsts k = lynch_consts(nd, eps);
    CallResult r = call_bayes(profile, k, pi);
    *label = r.label; gt[0] = r.gt0; gt[1] = r.gt1; *hom = r.hom; *het = r.het;
}

void hc_lr_pvalues(uint64_t profile, const double nd[4], int use_prior, double pi, double eps, double* p_hom, double* p_het) {
    LynchConsts k = lynch_consts(nd, eps);
    lr_pvalues(profile, k, use_prior != 0, pi, *p_hom, *p_het);
}

void hc_lynch_loglik(uint64_t profile, const double nd[4], double eps, double* lhom, double* lhet, double* logM) {
    LynchConsts k = lynch_consts(nd, eps);
    lynch_loglik(profile, k, *lhom, *lhet);
    *logM = log_multinomial(profile);
}

// compoundLikelihood over a histogram, sequentially with compensated accumulation
double hc_lynch_objective(uint64_t n, const uint64_t* profiles, const uint64_t* counts, const double nd[4], double pi, double eps) {
    if (pi < 0 || pi > 1 || eps < 0 || eps > 1) return 1.7976931348623157e308;
    LynchConsts k = lynch_consts(nd, eps);
    const double l1 = log1p(-pi), l2 = log(pi);
    CompSum acc;
    acc.init();
    for (uint64_t i = 0; i < n; ++i) {
        double t;
        if (lynch_term(profiles[i], log_multinomial(profiles[i]), k, l1, l2, t)) acc.add(t * (double)counts[i]);
    }
    return -acc.value();
}

int hc_format_suffix(int label, const char* gt, double hom, double het, int probability, char* out) {
    CallResult r;
    r.label = (uint8_t)label; r.gt0 = gt[0]; r.gt1 = gt[1]; r.hom = hom; r.het = het;
    int n = format_suffix(r, probability != 0, out);
    out[n] = 0;
    return n;
}

}  // extern "C"
