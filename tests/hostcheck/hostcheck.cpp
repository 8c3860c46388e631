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
#include "../../sid_b200/csrc/inflate.cuh"
#include <cmath>
#include <vector>
#ifdef SID_HAVE_FAST
#include "parse_swar.hpp"
#include "../../sid_b200/csrc/parse_bits.cuh"
#include "../../sid_b200/csrc/parse_win.cuh"
#include "../../sid_b200/csrc/parse_units.cuh"
#include "../../sid_b200/csrc/row_assemble.cuh"
#endif

using namespace sid;

extern "C" {

// inflate.cuh: one raw deflate stream on one thread (the walk the kernel's lane 0 does, matches copied in place).
// `in` must be readable 8 bytes past in_len.
int hc_inflate_member(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len) {
    static thread_local InflateTables t;
    return inflate_member_serial(in, in_len, out, out_len, t);
}

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

// The same calls formed the way a quality session forms them on the device now: the tokenizer's window walk sums the
// per-read terms (quality_sums_win), k_quality only finishes.  Every line the window path accepts must give
// BIT-IDENTICAL doubles to call_quality (same terms, same order).  Returns the number of lines, or -(k+1) when line k
// differs; *n_win counts the lines the window path took.
int64_t hc_call_quality_win(const uint8_t* text, uint64_t len, double prior, double alpha, uint64_t* n_win) {
    std::vector<double> lut(4 * 256 + LOG_FACT_N);
    for (int n = 0; n < LOG_FACT_N; ++n) lut[1024 + n] = lgamma((double)n + 1.0);
    for (int q = 0; q < 256; ++q) {
        const double error = pow(10., q / -10.);
        lut[q] = log(1 - error);
        lut[256 + q] = log(error);
        lut[512 + q] = log(1 - 2. / 3. * error);
        lut[768 + q] = log(2. / 3. * error);
    }
    FlatSrc src {text, len};
    int64_t k = 0;
    uint64_t taken = 0;
    for (uint64_t p = 0; p < len; ++p) {
        if (text[p] == '\n' || (p > 0 && text[p - 1] != '\n')) continue;
        CallResult w;
        if (quality_line_win_host(text, len, p, lut.data(), prior, alpha, w)) {
            ++taken;
            ParsedLine full, pl;
            parse_line(src, p, true, full);
            quality_fields(WordSrc {text, len}, p, full.profile, pl);
            if (pl.status != LINE_OK) return -(k + 1);                  // the window path must refuse what the byte-wise code reports
            const CallResult r = call_quality(text, p, pl, lut.data(), prior, alpha);
            if (double_bits(r.hom) != double_bits(w.hom) || double_bits(r.het) != double_bits(w.het) || r.label != w.label ||
                r.gt0 != w.gt0 || r.gt1 != w.gt1) return -(k + 1);
        }
        ++k;
    }
    if (n_win) *n_win = taken;
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

#ifdef SID_HAVE_FAST
// The window tokenizer (parse_win.cuh) against the byte-wise one on every line of a text.  Returns the number of
// lines, or -(k+1) when line k differs; *n_fast receives how many lines the fast grammar accepted.  Checks the
// profile, the name, the position (WANT_POS form) and, for the row writer, that "canonical" means exactly that the
// digits in the text are what printf("%d") prints for the position.
int64_t hc_compare_parsers_win(const uint8_t* text, uint64_t len, uint64_t* n_fast) {
    FlatSrc src {text, len};
    int64_t k = 0;
    uint64_t fast = 0;
    for (uint64_t p = 0; p < len; ++p) {
        if (text[p] == '\n' || (p > 0 && text[p - 1] != '\n')) continue;
        ParsedLine a;
        parse_line(src, p, false, a);
        WinLine b, c;
        const bool fb = parse_line_win_host<true>(text, len, p, b), fc = parse_line_win_host<false>(text, len, p, c);
        if (fb != fc) return -(k + 1);
        if (fb) {
            ++fast;
            if (a.status != LINE_OK || b.status != LINE_OK || a.profile != b.profile || a.pos != b.pos || c.profile != a.profile ||
                a.chrom_off != 0 || a.chrom_len != b.name_len || c.name_len != b.name_len || b.hdr_len != c.hdr_len ||
                b.pos_canonical != c.pos_canonical) return -(k + 1);
            char digits[16];
            const int nd = fmt_i32(a.pos, digits);
            const bool same_text = b.hdr_len == a.chrom_len + 1 + (uint32_t)nd && memcmp(text + p + a.chrom_len + 1, digits, nd) == 0;
            if (same_text != b.pos_canonical) return -(k + 1);
        }
        ++k;
    }
    if (n_fast) *n_fast = fast;
    return k;
}

// Stage 2 by units (parse_units.cuh: what the kernel runs on ordinary lines) against the byte-wise parser and against the
// window form: the same lines accepted, the same results.  Same return convention.
int64_t hc_compare_parsers_units(const uint8_t* text, uint64_t len, uint64_t* n_fast) {
    FlatSrc src {text, len};
    int64_t k = 0;
    uint64_t fast = 0;
    for (uint64_t p = 0; p < len; ++p) {
        if (text[p] == '\n' || (p > 0 && text[p - 1] != '\n')) continue;
        ParsedLine a;
        parse_line(src, p, false, a);
        WinLine b, c, w;
        uint64_t fwd = 0;
        const bool fb = parse_line_units_host<true>(text, len, p, b, &fwd), fc = parse_line_units_host<false>(text, len, p, c);
        const bool fw = parse_line_win_host<true>(text, len, p, w);
        if (fb != fc) return -(k + 1);
        if (fb) {
            ++fast;
            // the forward-strand profile against the byte-wise walk (pileup.cpp:78-124)
            BasesState st;
            st.init((uint8_t)a.ref);
            uint32_t fc4[4] = {0, 0, 0, 0};
            for (uint32_t i = 0; i < a.bases_len; ++i) {
                const uint8_t ch = text[p + a.bases_off + i];
                const int idx = st.feed(ch);
                if (idx < 0) continue;
                const uint8_t seen = ch == '.' ? st.dot_as : ch == ',' ? st.comma_as : ch;
                if (!(seen & 0x20u)) ++fc4[idx];
            }
            if (fwd != pack_profile(fc4[0], fc4[1], fc4[2], fc4[3])) return -(k + 1);
            if (a.status != LINE_OK || b.status != LINE_OK || a.profile != b.profile || a.pos != b.pos || c.profile != a.profile ||
                a.chrom_off != 0 || a.chrom_len != b.name_len || c.name_len != b.name_len || b.hdr_len != c.hdr_len ||
                b.pos_canonical != c.pos_canonical) return -(k + 1);
            if (fw && (w.hdr_len != b.hdr_len || w.pos_canonical != b.pos_canonical)) return -(k + 1);
        }
        ++k;
    }
    if (n_fast) *n_fast = fast;
    return k;
}

// The window-per-lane form of the window tokenizer (win_header + win_field_end + win_window, what the kernel runs on
// deep pileups) against the byte-wise one on every line of a text.  Same return convention; *n_fast counts accepted lines.
int64_t hc_compare_parsers_coop(const uint8_t* text, uint64_t len, uint64_t* n_fast) {
    FlatSrc src {text, len};
    int64_t k = 0;
    uint64_t fast = 0;
    for (uint64_t p = 0; p < len; ++p) {
        if (text[p] == '\n' || (p > 0 && text[p - 1] != '\n')) continue;
        ParsedLine a;
        parse_line(src, p, false, a);
        WinLine b;
        if (parse_line_win_host<true, true>(text, len, p, b)) {
            ++fast;
            if (a.status != LINE_OK || b.status != LINE_OK || a.profile != b.profile || a.pos != b.pos || a.chrom_off != 0 ||
                a.chrom_len != b.name_len) return -(k + 1);
        }
        ++k;
    }
    if (n_fast) *n_fast = fast;
    return k;
}

// The two-phase row assembly of the fused CSV writer, as a warp would run it: phase A of all `n` lanes in the order
// `order` gives (a warp-wide store with overlapping words has no defined winner), then phase B.  text: staged text
// (4-byte aligned), line_off/hdr_len/name_len/sfx_len per lane, sfx 48 bytes per lane; rows are laid end to end from
// offset d0 of `stage`.  Returns the offset after the last row.
uint32_t hc_assemble_rows(const uint8_t* text, uint8_t* stage, uint32_t d0, uint32_t n, const uint32_t* order, const uint32_t* line_off,
                          const uint32_t* hdr_len, const uint32_t* name_len, const uint8_t* sfx, const uint32_t* sfx_len) {
    RowSrc rows[32];
    uint32_t d[33], first[32], hw = 0, sxw = 0;
    d[0] = d0;
    for (uint32_t j = 0; j < n; ++j) {
        rows[j].line_off = line_off[j]; rows[j].hdr_len = hdr_len[j]; rows[j].name_len = name_len[j]; rows[j].sfx_len = sfx_len[j];
        memcpy(rows[j].sfx, sfx + 48 * j, 48);
        const uint32_t len = sfx_len[j] ? hdr_len[j] + sfx_len[j] : 0;
        d[j + 1] = d[j] + len;
        if (len) {
            const uint32_t a = d[j] & 3u, a2 = (d[j] + hdr_len[j]) & 3u;
            hw = hw > ((a + hdr_len[j] + 3) >> 2) ? hw : ((a + hdr_len[j] + 3) >> 2);
            sxw = sxw > ((a2 + sfx_len[j] + 3) >> 2) ? sxw : ((a2 + sfx_len[j] + 3) >> 2);
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t j = order[i];
        if (d[j + 1] > d[j]) first[j] = row_phase_a(text, stage, d[j], rows[j], hw, sxw);
    }
    for (uint32_t j = 0; j < n; ++j)
        if (d[j + 1] > d[j]) row_phase_b(stage, d[j], rows[j], first[j]);
    return d[n];
}

// classify_unit on one 32-byte unit: out[0..9] = term base p1 p2 dot caret pm digit nl bad
void hc_classify_unit(const uint8_t* bytes, uint32_t* out) {
    uint32_t w[8];
    memcpy(w, bytes, 32);
    const UnitClasses k = classify_unit(w);
    for (int c = 0; c < CW_WORDS; ++c) out[c] = k.w[c];
    out[8] = k.nl;
    out[9] = k.bad;
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
