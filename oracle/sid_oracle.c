/* TEST INFRASTRUCTURE ONLY -- see sid_oracle.h.  Plain-C restatement of the reference path;
 * every function cites the reference file:line it follows.  Never linked into the product. */
#define _POSIX_C_SOURCE 200809L
#include "sid_oracle.h"

#include <ctype.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- pileup.cpp */

/* pileup.cpp:70-153 parseReadBases; strands_out (optional): 1 for an upper-case character, 0 for a lower-case one (:85-124) */
static size_t read_bases(const char* s, char reference, uint16_t counts[4], char* bases_out, uint8_t* strands_out);
size_t orc_parse_read_bases(const char* s, char reference, uint16_t counts[4], char* bases_out) {
    return read_bases(s, reference, counts, bases_out, NULL);
}
/* ReadStack::strands (pileup.hpp:15) summed by letter: fwd[i] / rev[i] = counted bases i on the forward / reverse strand */
size_t orc_strand_counts(const char* s, char reference, uint16_t fwd[4], uint16_t rev[4]) {
    uint16_t counts[4];
    size_t cap = strlen(s) + 1;
    char* b = (char*)malloc(cap);
    uint8_t* st = (uint8_t*)malloc(cap);
    size_t nb = read_bases(s, reference, counts, b, st);
    memset(fwd, 0, 8);
    memset(rev, 0, 8);
    for (size_t k = 0; k < nb; ++k) {
        int idx = b[k] == 'A' ? 0 : b[k] == 'C' ? 1 : b[k] == 'G' ? 2 : 3;
        if (st[k]) ++fwd[idx]; else ++rev[idx];
    }
    free(b);
    free(st);
    return nb;
}
static size_t read_bases(const char* s, char reference, uint16_t counts[4], char* bases_out, uint8_t* strands_out) {
    size_t nb = 0;
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    size_t len = strlen(s);
    for (size_t i = 0; i < len; ++i) {
        char base = s[i];
        if (base == '.') {                       /* :78-80 */
            base = (char)toupper((unsigned char)reference);
        } else if (base == ',') {                /* :81-83 */
            base = (char)tolower((unsigned char)reference);
        }
        int idx = -1;
        switch (base) {                          /* :84-124 */
            case 'a': case 'A': idx = 0; break;
            case 'c': case 'C': idx = 1; break;
            case 'g': case 'G': idx = 2; break;
            case 't': case 'T': idx = 3; break;
            case '^': ++i; break;                /* :125-127 skip next char */
            case '+':
            case '-': {                          /* :128-147 */
                unsigned char nx = (unsigned char)s[i + 1];
                if (!(nx >= '0' && nx <= '9')) break;
                char* after;
                unsigned long length = (unsigned long)strtol(s + i + 1, &after, 10);
                if ((size_t)-1 - length < i) {
                    i = (size_t)-1;
                } else {
                    i = (size_t)(after - s) + length - 1;
                }
                break;
            }
            default: break;                      /* :148 */
        }
        if (idx >= 0) {
            if (bases_out) bases_out[nb] = "ACGT"[idx];
            if (strands_out) strands_out[nb] = (base >= 'A' && base <= 'Z') ? 1 : 0;     /* :85-124: the upper-case cases push 1 */
            ++nb;
            ++counts[idx];                       /* uint16_t: wraps mod 65536 like the reference */
        }
    }
    return nb;
}

/* pileup.cpp:155-167 parseQualities */
size_t orc_parse_qualities(const char* q, uint8_t* out) {
    size_t n = 0;
    for (; *q != '\0' && *q != '\t' && *q != '\n'; ++q) {
        uint8_t quality = (uint8_t)(*q - 33);
        if (quality < 1) quality = 1;
        if (out) out[n] = quality;
        ++n;
    }
    return n;
}

/* pileup.cpp:13-68 parsePileupLine */
int orc_parse_line(char* line, int want_bq, int want_mq, orc_line* out) {
    static const char* SEP = " \t";              /* :11 */
    char* save = NULL;
    memset(out, 0, sizeof *out);
    char* chrom = strtok_r(line, SEP, &save);    /* :17 */
    if (chrom == NULL) return ORC_MALFORMED;     /* reference: std::string(nullptr) -> terminate */
    out->chrom = chrom;
    char* position = strtok_r(NULL, SEP, &save); /* :20-24 */
    if (position == NULL) return ORC_MALFORMED;
    out->pos = atoi(position);
    char* reference = strtok_r(NULL, SEP, &save);/* :26-30 */
    if (reference == NULL || strlen(reference) != 1) return ORC_MALFORMED;
    out->ref = reference[0];
    char* coverage = strtok_r(NULL, SEP, &save); /* :32-36 (value only used as a reserve hint) */
    if (coverage == NULL) return ORC_MALFORMED;
    char* bases = strtok_r(NULL, SEP, &save);    /* :38-42 */
    if (bases == NULL) return ORC_MALFORMED;
    out->bases = bases;
    orc_parse_read_bases(bases, out->ref, out->counts, NULL);
    char* bq = strtok_r(NULL, SEP, &save);       /* :49 */
    if (want_bq) {                               /* :52-57 */
        if (bq == NULL) return ORC_REFERENCE_UB; /* reference dereferences a null pointer here */
        out->bq = bq;
    }
    if (want_mq) {                               /* :60-66 */
        char* mq = strtok_r(NULL, SEP, &save);
        if (mq == NULL) return ORC_MALFORMED_OR_MISSING;
        out->mq = mq;
    }
    return ORC_OK;
}

static int cmp_profile(const void* a, const void* b) {
    const uint16_t* x = (const uint16_t*)a;
    const uint16_t* y = (const uint16_t*)b;
    for (int i = 0; i < 4; ++i) {
        if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    }
    return 0;
}

/* pileup.cpp:169-196 countUniqueProfiles (lexicographic order of std::array<uint16_t,4>) */
orc_unique_profile* orc_count_unique(const uint16_t* profiles, size_t n, size_t* n_unique) {
    *n_unique = 0;
    if (n == 0) return NULL;
    uint16_t* sorted = (uint16_t*)malloc(n * 8);
    memcpy(sorted, profiles, n * 8);
    qsort(sorted, n, 8, cmp_profile);
    orc_unique_profile* u = (orc_unique_profile*)malloc(n * sizeof *u);
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        if (k > 0 && cmp_profile(u[k - 1].profile, sorted + 4 * i) == 0) {
            u[k - 1].count += 1;
        } else {
            memcpy(u[k].profile, sorted + 4 * i, 8);
            u[k].count = 1;
            u[k].coverage = (uint32_t)((int)u[k].profile[0] + u[k].profile[1] + u[k].profile[2] + u[k].profile[3]); /* pileup.hpp:37-39 */
            ++k;
        }
    }
    free(sorted);
    *n_unique = k;
    return u;
}

/* pileup.cpp:198-217 computeNucleotideDistribution */
void orc_nucleotide_distribution(const orc_unique_profile* u, size_t n, double nd[4]) {
    uint64_t acc[4] = {0, 0, 0, 0};
    uint64_t total = 0;
    for (size_t k = 0; k < n; ++k) {
        total += (uint32_t)(u[k].count * u[k].coverage);       /* :202 uint32 product, then widened */
        for (int i = 0; i < 4; ++i) acc[i] += (uint32_t)(u[k].count * (uint32_t)u[k].profile[i]); /* :204 */
    }
    if (total != 0) {
        for (int i = 0; i < 4; ++i) nd[i] = (double)acc[i] / (double)total;
    } else {
        nd[0] = nd[1] = nd[2] = nd[3] = 0.25;
    }
}

/* ---------------------------------------------------------------- call.cpp helpers */

/* call.cpp:52-60 getMajorAlleleIndices: std::sort of 4 indices ascending by count; for 4 elements
 * libstdc++ runs insertion sort (stable), so ties keep index order and the HIGHER index wins. */
void orc_major_alleles(const uint16_t profile[4], int* first, int* second) {
    int idx[4] = {0, 1, 2, 3};
    for (int i = 1; i < 4; ++i) {
        int v = idx[i];
        int j = i;
        while (j > 0 && profile[v] < profile[idx[j - 1]]) { idx[j] = idx[j - 1]; --j; }
        idx[j] = v;
    }
    *first = idx[3];
    *second = idx[2];
}

/* ---------------------------------------------------------------- lynch.hpp */

/* lynch.hpp:11-31 MemoizedLogGamma: log_gamma(0) = 0, otherwise gsl_sf_lngamma(x) -> lgamma */
static double log_gamma_i(int x) {
    if (x == 0) return 0;
    return lgamma((double)x);
}

/* lynch.hpp:48-55 */
long double orc_multinomial_coefficient(const orc_unique_profile* p) {
    return expl(log_gamma_i((int)p->coverage + 1) - log_gamma_i(p->profile[0] + 1) - log_gamma_i(p->profile[1] + 1)
                - log_gamma_i(p->profile[2] + 1) - log_gamma_i(p->profile[3] + 1));
}

/* lynch.hpp:57-74 */
long double orc_het_likelihood_nd(const orc_unique_profile* p, double e, const double nd[4]) {
    long double L = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = i + 1; j < 4; ++j) {
            L += nd[i] * nd[j]
                 * powl((1 - 2. / 3. * e) / 2., p->profile[i] + p->profile[j])
                 * powl(e / 3., p->coverage - p->profile[i] - p->profile[j]);
        }
    }
    long double s = 0;
    for (int i = 0; i < 4; ++i) s += nd[i] * nd[i];
    L /= (1 - s);
    return orc_multinomial_coefficient(p) * L;
}

/* lynch.hpp:76-80 */
long double orc_het_likelihood_ref(const orc_unique_profile* p, double e, int r0, int r1) {
    return orc_multinomial_coefficient(p)
           * powl((1 - 2. / 3. * e) / 2., p->profile[r0] + p->profile[r1])
           * powl(e / 3., p->coverage - p->profile[r0] - p->profile[r1]);
}

/* lynch.hpp:82-90 */
long double orc_hom_likelihood_nd(const orc_unique_profile* p, double e, const double nd[4]) {
    long double L = 0;
    for (int i = 0; i < 4; ++i) {
        L += nd[i] * powl(1 - e, p->profile[i]) * powl(e / 3., p->coverage - p->profile[i]);
    }
    return orc_multinomial_coefficient(p) * L;
}

/* lynch.hpp:92-96 */
long double orc_hom_likelihood_ref(const orc_unique_profile* p, double e, int ref) {
    return orc_multinomial_coefficient(p) * powl(1 - e, p->profile[ref]) * powl(e / 3., p->coverage - p->profile[ref]);
}

/* ---------------------------------------------------------------- stats.cpp */

/* gsl_cdf_chisq_Q(x, 1) restated by the identity Q_{chi2,1}(x) = erfc(sqrt(x/2)) */
static double chisq_q1(double x) {
    if (!(x > 0)) return 1.0;
    return erfc(sqrt(x / 2.0));
}

/* stats.cpp:29-37 likelihoodRatioTest */
double orc_lrt(long double l_h0, long double l_h1) {
    if (l_h0 != 0) {
        long double chisq = -2 * (logl(l_h0) - logl(fmaxl(l_h0, l_h1)));
        return chisq_q1((double)chisq);
    }
    return chisq_q1(DBL_MAX);
}

typedef struct { double v; size_t i; } bh_item;
static int cmp_desc(const void* a, const void* b) {
    const bh_item* x = (const bh_item*)a;
    const bh_item* y = (const bh_item*)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return 0;
}

/* stats.cpp:58-80 adjustBenjaminiHochberg (the order among tied p-values does not change the result) */
void orc_bh(const double* p, size_t n, double* adjusted) {
    if (n == 0) return;
    bh_item* s = (bh_item*)malloc(n * sizeof *s);
    for (size_t i = 0; i < n; ++i) { s[i].v = p[i]; s[i].i = i; }
    qsort(s, n, sizeof *s, cmp_desc);
    size_t m = n;
    adjusted[s[0].i] = p[s[0].i];
    for (size_t i = 1; i < n; ++i) {
        double cand = p[s[i].i] * (double)m / (double)(m - i);
        double prev = adjusted[s[i - 1].i];
        adjusted[s[i].i] = cand < prev ? cand : prev;   /* std::min(prev, cand) */
    }
    for (size_t i = 0; i < n; ++i) if (adjusted[i] > 1) adjusted[i] = 1.0;
    free(s);
}

/* ---------------------------------------------------------------- lynch.cpp */

/* lynch.cpp:37-61 compoundLikelihood */
double orc_compound_likelihood(const orc_unique_profile* u, size_t n, const double nd[4], double pi, double eps) {
    if (pi < 0 || pi > 1 || eps < 0 || eps > 1) return DBL_MAX;
    long double logLikelihood = 0;
    for (size_t k = 0; k < n; ++k) {
        long double L = (1. - pi) * orc_hom_likelihood_nd(&u[k], eps, nd) + pi * orc_het_likelihood_nd(&u[k], eps, nd);
        if (L > 0) logLikelihood += logl(L) * u[k].count;
    }
    if (isinf(logLikelihood)) {
        logLikelihood = logLikelihood > 0 ? LDBL_MAX : -LDBL_MAX;
    }
    return (double)(-logLikelihood);
}

/* optimization.hpp:51-89 around gsl nmsimplex2 (restated, N = 2; see gslshim.cpp for the notes) */
typedef struct {
    const orc_unique_profile* u; size_t n; const double* nd; int evals;
} nm_ctx;
static double nm_f(nm_ctx* c, const double x[2]) {
    c->evals++;
    return orc_compound_likelihood(c->u, c->n, c->nd, x[0], x[1]);
}
typedef struct { double X[3][2]; double Y[3]; double c[2]; double S2; } nm_state;
static void nm_center(nm_state* s) {
    for (int j = 0; j < 2; ++j) s->c[j] = (s->X[0][j] + s->X[1][j] + s->X[2][j]) / 3.0;
}
static double nm_size(nm_state* s) {
    double ss = 0;
    for (int k = 0; k < 3; ++k) {
        double t = 0;
        for (int j = 0; j < 2; ++j) { double d = s->X[k][j] - s->c[j]; t += d * d; }
        ss += t;
    }
    s->S2 = ss / 3.0;
    return sqrt(s->S2);
}
static double nm_move(double coeff, const nm_state* s, int corner, double xc[2], nm_ctx* c) {
    const double P = 3.0;
    double alpha = (1 - coeff) * P / (P - 1.0);
    double beta = (P * coeff - 1.0) / (P - 1.0);
    for (int j = 0; j < 2; ++j) xc[j] = alpha * s->c[j] + beta * s->X[corner][j];
    return nm_f(c, xc);
}
static void nm_update(nm_state* s, int i, const double x[2], double val) {
    const double P = 3.0;
    double d2 = 0, xmcd = 0;
    for (int j = 0; j < 2; ++j) {
        double delta = x[j] - s->X[i][j];
        double xmc = s->X[i][j] - s->c[j];
        d2 += delta * delta;
        xmcd += xmc * delta;
    }
    double d = sqrt(d2);
    s->S2 += (2.0 / P) * xmcd + ((P - 1.0) / P) * (d * d / P);
    for (int j = 0; j < 2; ++j) {
        s->c[j] -= (1.0 / P) * s->X[i][j];
        s->c[j] += (1.0 / P) * x[j];
        s->X[i][j] = x[j];
    }
    s->Y[i] = val;
}

/* lynch.cpp:17-24 (start (1e-3,1e-3), steps 1e-4) + optimization.hpp:51-89 (size < 1e-5, <= 1000 iterations) */
void orc_estimate(const orc_unique_profile* u, size_t n, const double nd[4], double* pi, double* eps,
                  int* iterations, int* evaluations, int* converged) {
    nm_ctx c = {u, n, nd, 0};
    nm_state s;
    const double x0[2] = {1e-3, 1e-3}, step[2] = {1e-4, 1e-4};
    for (int k = 0; k < 3; ++k) { s.X[k][0] = x0[0]; s.X[k][1] = x0[1]; }
    s.X[1][0] += step[0];
    s.X[2][1] += step[1];
    for (int k = 0; k < 3; ++k) s.Y[k] = nm_f(&c, s.X[k]);
    nm_center(&s);
    double size = nm_size(&s);
    double best[2] = {x0[0], x0[1]};
    int i = 0, status = -2;
    do {
        ++i;
        int hi = 0, s_hi = 1, lo = 0;
        double dhi = s.Y[0], dlo = s.Y[0], ds_hi = s.Y[1];
        for (int k = 1; k < 3; ++k) {
            double v = s.Y[k];
            if (v < dlo) { dlo = v; lo = k; }
            else if (v > dhi) { ds_hi = dhi; s_hi = hi; dhi = v; hi = k; }
            else if (v > ds_hi) { ds_hi = v; s_hi = k; }
        }
        double xc[2], xc2[2];
        int bad = 0;
        double val = nm_move(-1.0, &s, hi, xc, &c);
        if (isfinite(val) && val < s.Y[lo]) {
            double val2 = nm_move(-2.0, &s, hi, xc2, &c);
            if (isfinite(val2) && val2 < s.Y[lo]) nm_update(&s, hi, xc2, val2);
            else nm_update(&s, hi, xc, val);
        } else if (!isfinite(val) || val > s.Y[s_hi]) {
            if (isfinite(val) && val <= s.Y[hi]) nm_update(&s, hi, xc, val);
            double val2 = nm_move(0.5, &s, hi, xc2, &c);
            if (isfinite(val2) && val2 <= s.Y[hi]) {
                nm_update(&s, hi, xc2, val2);
            } else {
                for (int k = 0; k < 3; ++k) {
                    if (k == lo) continue;
                    for (int j = 0; j < 2; ++j) s.X[k][j] = 0.5 * (s.X[k][j] + s.X[lo][j]);
                    s.Y[k] = nm_f(&c, s.X[k]);
                    if (!isfinite(s.Y[k])) bad = 1;
                }
                nm_center(&s);
                nm_size(&s);
            }
        } else {
            nm_update(&s, hi, xc, val);
        }
        if (bad) { status = 9; break; }          /* optimization.hpp:62-64: iterate() failed -> stop */
        lo = 0;
        for (int k = 1; k < 3; ++k) if (s.Y[k] < s.Y[lo]) lo = k;
        best[0] = s.X[lo][0];
        best[1] = s.X[lo][1];
        size = s.S2 > 0 ? sqrt(s.S2) : nm_size(&s);
        status = size < 1e-5 ? 0 : -2;           /* optimization.hpp:66-67 */
    } while (status == -2 && i < 1000);          /* optimization.hpp:72 */
    *pi = best[0];
    *eps = best[1];
    if (iterations) *iterations = i;
    if (evaluations) *evaluations = c.evals;
    if (converged) *converged = status != -2;
}

/* ---------------------------------------------------------------- call.cpp: the four methods */

typedef struct {
    uint32_t chrom_off; uint16_t chrom_len; int32_t pos; uint16_t counts[4];
    uint32_t bases_off, bq_off, mq_off; char ref;
} site_t;

static int profile_find(const orc_unique_profile* u, size_t n, const uint16_t* p, size_t* idx) {
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        int c = cmp_profile(u[mid].profile, p);
        if (c == 0) { *idx = mid; return 1; }
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    return 0;
}

static size_t drop_low_coverage(orc_unique_profile* u, size_t n) { /* call.cpp:66-70,149-153,224-229,296-301 */
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) if (!(u[i].coverage < 4)) u[k++] = u[i];
    return k;
}

static void result_alloc(orc_result* r, size_t n_sites) {
    size_t m = n_sites ? n_sites : 1;
    r->chrom_off = (uint32_t*)malloc(m * sizeof(uint32_t));
    r->chrom_len = (uint16_t*)malloc(m * sizeof(uint16_t));
    r->pos = (int32_t*)malloc(m * sizeof(int32_t));
    r->label = (uint8_t*)malloc(m);
    r->gt = (char*)malloc(2 * m);
    r->hom_conf = (double*)malloc(m * sizeof(double));
    r->het_conf = (double*)malloc(m * sizeof(double));
    r->profiles = (uint16_t*)malloc(m * 8);
}

void orc_free_result(orc_result* r) {
    free(r->chrom_off); free(r->chrom_len); free(r->pos); free(r->label); free(r->gt);
    free(r->hom_conf); free(r->het_conf); free(r->profiles);
    memset(r, 0, sizeof *r);
}

typedef struct { uint8_t label; char gt[2]; double hom, het; } class_t;

static void emit(orc_result* r, const site_t* s, const class_t* c) {
    size_t k = r->n++;
    r->chrom_off[k] = s->chrom_off;
    r->chrom_len[k] = s->chrom_len;
    r->pos[k] = s->pos;
    r->label[k] = c->label;
    r->gt[2 * k] = c->gt[0];
    r->gt[2 * k + 1] = c->gt[1];
    r->hom_conf[k] = c->hom;
    r->het_conf[k] = c->het;
}

int orc_call(const char* text, size_t len, int method, int estimate_prior, double prior,
             double error_threshold, double alpha, orc_result* out) {
    memset(out, 0, sizeof *out);
    out->heterozygosity = NAN;
    out->error_rate = NAN;
    const int want_q = method == ORC_QUALITY;
    char* buf = (char*)malloc(len + 1);
    memcpy(buf, text, len);
    buf[len] = '\0';

    /* call.cpp:11-20 readFile: getline on '\n', empty lines skipped */
    size_t n_lines = 0;
    for (size_t i = 0; i < len; ++i) if (buf[i] == '\n') ++n_lines;
    site_t* sites = (site_t*)malloc((n_lines + 2) * sizeof *sites);
    size_t n = 0;
    int status = ORC_OK;
    for (size_t start = 0; start < len && status == ORC_OK;) {
        char* nl = (char*)memchr(buf + start, '\n', len - start);
        size_t end = nl ? (size_t)(nl - buf) : len;
        buf[end] = '\0';
        if (end > start) {
            orc_line pl;
            status = orc_parse_line(buf + start, want_q, want_q, &pl);
            if (status == ORC_OK) {
                site_t* s = &sites[n++];
                s->chrom_off = (uint32_t)(pl.chrom - buf);
                s->chrom_len = (uint16_t)strlen(pl.chrom);
                s->pos = pl.pos;
                s->ref = pl.ref;
                memcpy(s->counts, pl.counts, 8);
                s->bases_off = (uint32_t)(pl.bases - buf);
                s->bq_off = pl.bq ? (uint32_t)(pl.bq - buf) : 0;
                s->mq_off = pl.mq ? (uint32_t)(pl.mq - buf) : 0;
            }
        }
        start = end + 1;
    }
    if (status != ORC_OK) { free(buf); free(sites); return status; }

    out->n_sites = n;
    result_alloc(out, n);
    for (size_t i = 0; i < n; ++i) memcpy(out->profiles + 4 * i, sites[i].counts, 8);
    out->conf_type = method == ORC_BAYES ? 1 : 0;

    size_t n_all = 0;
    orc_unique_profile* all = NULL;      /* every unique profile (call.cpp:215) */
    orc_unique_profile* cov4 = NULL;     /* coverage >= 4 only */
    size_t n_cov4 = 0;
    double nd[4];
    const int need_hist = method != ORC_QUALITY || estimate_prior;
    const int need_fit = method == ORC_BAYES || method == ORC_LIKELIHOOD_RATIO || estimate_prior;
    if (need_hist) {
        all = orc_count_unique(out->profiles, n, &n_all);
        if (need_fit) {
            cov4 = (orc_unique_profile*)malloc((n_all ? n_all : 1) * sizeof *cov4);
            if (n_all) memcpy(cov4, all, n_all * sizeof *cov4);
            n_cov4 = drop_low_coverage(cov4, n_all);
            out->n_unique = n_cov4;
            orc_nucleotide_distribution(cov4, n_cov4, nd);
            orc_estimate(cov4, n_cov4, nd, &out->heterozygosity, &out->error_rate, &out->iterations,
                         &out->evaluations, &out->converged);
        } else {
            out->n_unique = n_all;
        }
    }

    if (method == ORC_LOCAL) {
        /* call.cpp:213-289 callSiteMLError */
        if (estimate_prior) prior = out->heterozygosity;            /* :233 */
        class_t* cls = (class_t*)malloc((n_all ? n_all : 1) * sizeof *cls);
        for (size_t k = 0; k < n_all; ++k) {
            const orc_unique_profile* p = &all[k];
            int f, s;
            orc_major_alleles(p->profile, &f, &s);                   /* :240 */
            double error1 = (double)(p->coverage - p->profile[f]) / (double)p->coverage;   /* :243 */
            if (error1 > error_threshold) error1 = error_threshold;
            long double l1 = orc_hom_likelihood_ref(p, error1, f);  /* :247 */
            double error2 = 1.5 * (double)(p->coverage - p->profile[f] - p->profile[s]) / (double)p->coverage; /* :250 */
            if (error2 > error_threshold) error2 = error_threshold;
            long double l2 = orc_het_likelihood_ref(p, error2, f, s); /* :254 */
            if (prior > 0) { l1 *= (1 - prior); l2 *= prior; }      /* :256-259 */
            double p1 = orc_lrt(l2, l1);                            /* :261 */
            double p2 = orc_lrt(l1, l2);                            /* :262 */
            cls[k].label = 0;
            cls[k].gt[0] = cls[k].gt[1] = "ACGT"[f];
            if (l2 > l1 && p2 < alpha) { cls[k].label = 1; cls[k].gt[1] = "ACGT"[s]; } /* :266-269 */
            cls[k].hom = p1;
            cls[k].het = p2;
        }
        for (size_t i = 0; i < n; ++i) {                             /* :276-285 */
            size_t idx;
            if (profile_find(all, n_all, sites[i].counts, &idx)) emit(out, &sites[i], &cls[idx]);
        }
        free(cls);
    } else if (method == ORC_BAYES || method == ORC_LIKELIHOOD_RATIO) {
        const double pi = out->heterozygosity, eps = out->error_rate;
        class_t* cls = (class_t*)malloc((n_cov4 ? n_cov4 : 1) * sizeof *cls);
        double* p_hom = (double*)malloc((n_cov4 ? n_cov4 : 1) * sizeof(double));
        double* p_het = (double*)malloc((n_cov4 ? n_cov4 : 1) * sizeof(double));
        double* a_hom = (double*)malloc((n_cov4 ? n_cov4 : 1) * sizeof(double));
        double* a_het = (double*)malloc((n_cov4 ? n_cov4 : 1) * sizeof(double));
        for (size_t k = 0; k < n_cov4; ++k) {
            long double L_hom = orc_hom_likelihood_nd(&cov4[k], eps, nd);   /* lynch.cpp:28-31 */
            long double L_het = orc_het_likelihood_nd(&cov4[k], eps, nd);
            int f, s;
            orc_major_alleles(cov4[k].profile, &f, &s);
            cls[k].label = 0;
            cls[k].gt[0] = cls[k].gt[1] = "ACGT"[f];
            if (method == ORC_BAYES) {                               /* call.cpp:176-194 */
                long double ah = L_hom * (1 - pi);
                long double at = L_het * pi;
                long double ph = ah / (ah + at);
                long double pt = at / (ah + at);
                if (pt > ph) { cls[k].label = 1; cls[k].gt[1] = "ACGT"[s]; }
                cls[k].hom = (double)ph;
                cls[k].het = (double)pt;
            } else {                                                 /* call.cpp:93-103 */
                if (estimate_prior) { L_het *= pi; L_hom *= 1 - pi; }
                p_hom[k] = orc_lrt(L_het, L_hom);
                p_het[k] = orc_lrt(L_hom, L_het);
            }
        }
        if (method == ORC_LIKELIHOOD_RATIO) {                        /* call.cpp:105-127 */
            orc_bh(p_hom, n_cov4, a_hom);
            orc_bh(p_het, n_cov4, a_het);
            for (size_t k = 0; k < n_cov4; ++k) {
                int f, s;
                orc_major_alleles(cov4[k].profile, &f, &s);
                if (a_het[k] < alpha) { cls[k].label = 1; cls[k].gt[1] = "ACGT"[s]; }
                cls[k].hom = a_hom[k];
                cls[k].het = a_het[k];
            }
        }
        for (size_t i = 0; i < n; ++i) {                             /* call.cpp:131-140,199-208 */
            size_t idx;
            if (profile_find(cov4, n_cov4, sites[i].counts, &idx)) emit(out, &sites[i], &cls[idx]);
        }
        free(cls); free(p_hom); free(p_het); free(a_hom); free(a_het);
    } else if (method == ORC_QUALITY) {
        /* call.cpp:291-372 callQualityBasedSimple */
        if (estimate_prior) prior = out->heterozygosity;            /* :305 */
        size_t cap = 0;
        char* bases = NULL; uint8_t* bq = NULL; uint8_t* mq = NULL;
        for (size_t i = 0; i < n && status == ORC_OK; ++i) {
            const site_t* st = &sites[i];
            size_t blen = strlen(buf + st->bases_off), qlen = strlen(buf + st->bq_off), mlen = strlen(buf + st->mq_off);
            size_t need = blen > qlen ? blen : qlen;
            if (mlen > need) need = mlen;
            if (need + 1 > cap) {
                cap = 2 * need + 64;
                bases = (char*)realloc(bases, cap); bq = (uint8_t*)realloc(bq, cap); mq = (uint8_t*)realloc(mq, cap);
            }
            uint16_t counts[4];
            size_t nb = orc_parse_read_bases(buf + st->bases_off, st->ref, counts, bases);
            size_t nq = orc_parse_qualities(buf + st->bq_off, bq);
            size_t nm = orc_parse_qualities(buf + st->mq_off, mq);
            if (nq < nb || nm < nb) { status = ORC_REFERENCE_UB; break; }   /* :330-331 reads past the vectors */
            int ref0, ref1;
            orc_major_alleles(counts, &ref0, &ref1);                 /* :311-319 (same stable sort) */
            long double lh = 0, lt = 0;
            for (size_t j = 0; j < nb; ++j) {                        /* :329-342 */
                uint8_t mn = bq[j] < mq[j] ? bq[j] : mq[j];
                double error = pow(10., mn / -10.);
                if (bases[j] == "ACGT"[ref0]) lh += log(1 - error); else lh += log(error);
                if (bases[j] == "ACGT"[ref0] || bases[j] == "ACGT"[ref1]) lt += log(1 - 2. / 3. * error);
                else lt += log(2. / 3. * error);
            }
            int nn = counts[ref0] + counts[ref1];                    /* :347-349 */
            int kk = counts[ref1];
            double logbinom = log_gamma_i(nn + 1) - log_gamma_i(nn - kk + 1) - log_gamma_i(kk + 1);
            lt += logbinom - nn * logl(2);
            long double pp1 = expl(lh);                              /* :352-357 */
            long double pp2 = expl(lt);
            if (prior > 0) { pp1 *= (1 - prior); pp2 *= prior; }
            class_t c;
            c.hom = orc_lrt(pp2, pp1);                               /* :359-360 */
            c.het = orc_lrt(pp1, pp2);
            c.label = 0;
            c.gt[0] = c.gt[1] = "ACGT"[ref0];
            if (c.het < alpha) { c.label = 1; c.gt[1] = "ACGT"[ref1]; }  /* :364-367 */
            emit(out, st, &c);
        }
        free(bases); free(bq); free(mq);
    }
    free(all); free(cov4); free(sites); free(buf);
    if (status != ORC_OK) { orc_free_result(out); }
    return status;
}

/* call.hpp:29-38 operator<< and sid.cpp:102-105 */
size_t orc_write_csv(const char* text, const orc_result* r, char* out, size_t cap) {
    static const char HEADER[] = "chrom,pos,label,gt,hom_conf,het_conf,conf_type\n";
    size_t w = 0;
    char row[256];
    size_t hl = sizeof HEADER - 1;
    if (w + hl <= cap) memcpy(out + w, HEADER, hl);
    w += hl;
    for (size_t k = 0; k < r->n; ++k) {
        size_t cl = r->chrom_len[k];
        if (w + cl <= cap) memcpy(out + w, text + r->chrom_off[k], cl);
        w += cl;
        int m = snprintf(row, sizeof row, ",%d,%s,%c%c,%g,%g,%s\n", r->pos[k], r->label[k] ? "het" : "hom",
                         r->gt[2 * k], r->gt[2 * k + 1], r->hom_conf[k], r->het_conf[k],
                         r->conf_type ? "probability" : "p_value");
        if (w + (size_t)m <= cap) memcpy(out + w, row, (size_t)m);
        w += (size_t)m;
    }
    return w;
}
