/* Synthetic `samtools mpileup` text generator (bench/test tooling, not on the product path).
 *
 * Shape follows SURVEY.md 8(d): 6 columns `chrom pos ref depth bases quals` (+ a 7th mapq column
 * when seven_columns != 0), tab separated, '\n' terminated.
 *   ref ~ U{A,C,G,T}; depth ~ Poisson(lambda); the site is heterozygous with probability het
 *   (second allele uniform over the other three bases); each read picks one of the two alleles
 *   50/50 and is mis-read with probability err to a uniform other base; strand 50/50 ('.' ','
 *   for a reference match, upper/lower case otherwise); '^'+chr(33+U[0,60]) before a base with
 *   probability start; '$' after a base with probability start; after a base, with probability
 *   indel, '+' or '-', a length U[1,11] and that many random bases in strand case;
 *   depth 0 => bases "*" and quals "*".
 *   quals: chr(33 + clip(round(N(35,5)), 2, 41)); mapq: chr(33 + U[20,60]).
 * The random stream is counter based: every site draws from SplitMix64 seeded by a hash of
 * (seed, absolute site index), so any shard can be generated independently and identically.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    uint64_t seed;
    double lambda;       /* mean depth */
    double het;          /* heterozygosity pi0 */
    double err;          /* per-read error eps0 */
    double start;        /* read start / read end marker probability s */
    double indel;        /* indel probability d */
    int seven_columns;   /* also emit the mapping-quality column */
    int n_chroms;        /* number of chromosomes (>=1) */
    const char* const* chrom_names;
    const uint64_t* chrom_lengths; /* sites per chromosome; the last one absorbs the remainder */
} pileup_gen_params;

#define QTAB 4096
typedef struct {
    uint64_t* pois;  /* pois[k] = floor(P(depth <= k) * 2^64), saturating */
    int n_pois;
    char qtab[QTAB];
    uint32_t thr_start, thr_indel; /* 16-bit thresholds */
    uint32_t thr_err;              /* 24-bit threshold */
    uint64_t thr_het;              /* 64-bit threshold */
} gen_tables;

static inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline uint64_t site_key(uint64_t seed, uint64_t site) {
    /* full avalanche of (seed, site): neighbouring sites must not land on shifted copies of one stream */
    uint64_t k = (site + 1) * 0xD6E8FEB86659FD93ULL;
    k ^= k >> 32; k *= 0xD6E8FEB86659FD93ULL; k ^= k >> 32;
    k += seed * 0xD1342543DE82EF95ULL + 0x2545F4914F6CDD1DULL;
    k ^= k >> 29; k *= 0xBF58476D1CE4E5B9ULL; k ^= k >> 32; k *= 0x94D049BB133111EBULL; k ^= k >> 29;
    return k;
}

static uint64_t prob_to_u64(long double p) {
    if (p <= 0) return 0;
    if (p >= 1) return UINT64_MAX;
    long double v = p * 18446744073709551616.0L;
    if (v >= 18446744073709551615.0L) return UINT64_MAX;
    return (uint64_t)v;
}

static void build_tables(const pileup_gen_params* p, gen_tables* t) {
    int kmax = (int)(p->lambda + 12.0 * sqrt(p->lambda + 1.0) + 24.0);
    t->n_pois = kmax + 1;
    t->pois = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)t->n_pois);
    long double cdf = 0;
    for (int k = 0; k <= kmax; ++k) {
        long double logp = -(long double)p->lambda + k * logl((long double)p->lambda) - lgammal((long double)k + 1.0L);
        if (p->lambda <= 0) logp = (k == 0) ? 0 : -INFINITY;
        cdf += expl(logp);
        t->pois[k] = prob_to_u64(cdf);
    }
    t->pois[kmax] = UINT64_MAX;
    /* quality table: pmf of clip(round(N(35,5)),2,41) laid out over QTAB slots */
    double cum = 0;
    int filled = 0;
    for (int q = 2; q <= 41; ++q) {
        double lo = (q == 2) ? -1e9 : (q - 0.5), hi = (q == 41) ? 1e9 : (q + 0.5);
        double pq = 0.5 * (erf((hi - 35.0) / (5.0 * sqrt(2.0))) - erf((lo - 35.0) / (5.0 * sqrt(2.0))));
        cum += pq;
        int upto = (int)floor(cum * QTAB + 0.5);
        if (q == 41) upto = QTAB;
        for (; filled < upto && filled < QTAB; ++filled) t->qtab[filled] = (char)(33 + q);
    }
    for (; filled < QTAB; ++filled) t->qtab[filled] = (char)(33 + 41);
    t->thr_start = (uint32_t)floor(p->start * 65536.0 + 0.5);
    t->thr_indel = (uint32_t)floor(p->indel * 65536.0 + 0.5);
    t->thr_err = (uint32_t)floor(p->err * 16777216.0 + 0.5);
    t->thr_het = prob_to_u64((long double)p->het);
}

static inline char* put_uint(char* o, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *o++ = tmp[--n];
    return o;
}

static const char UP[4] = {'A', 'C', 'G', 'T'};
static const char LOW[4] = {'a', 'c', 'g', 't'};

/* Upper bound on the bytes one site can take at depth d. */
static inline size_t site_bound(size_t chrom_len, int depth, int seven) {
    /* per read: ^q (2) + base (1) + $ (1) + indel (1 + 2 + 11) = 18, qual 1, mapq 1 */
    return chrom_len + 1 + 20 + 1 + 1 + 1 + 11 + 1 + (size_t)(depth ? depth : 1) * (18 + 1 + (seven ? 1 : 0)) + 4;
}

static char* gen_site(char* o, const pileup_gen_params* p, const gen_tables* t, uint64_t site,
                      const char* chrom, size_t chrom_len, uint64_t pos) {
    uint64_t s = site_key(p->seed, site);
    uint64_t r0 = splitmix64(&s);
    int ref = (int)(r0 & 3);
    uint64_t ud = splitmix64(&s);
    int lo = 0, hi = t->n_pois - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (ud < t->pois[mid]) hi = mid; else lo = mid + 1; }
    int depth = lo;
    int a1 = ref, a2 = ref;
    if (splitmix64(&s) < t->thr_het) a2 = (ref + 1 + (int)((r0 >> 2) % 3)) & 3;

    memcpy(o, chrom, chrom_len); o += chrom_len;
    *o++ = '\t';
    o = put_uint(o, pos);
    *o++ = '\t';
    *o++ = UP[ref];
    *o++ = '\t';
    o = put_uint(o, (uint64_t)depth);
    *o++ = '\t';
    if (depth == 0) {
        *o++ = '*'; *o++ = '\t'; *o++ = '*';
        if (p->seven_columns) { *o++ = '\t'; *o++ = '*'; }
        *o++ = '\n';
        return o;
    }
    char* q = o; /* quals are written after the bases; remember the draws instead of re-drawing */
    uint64_t s2 = s ^ 0xA5A5A5A55A5A5A5AULL; /* independent stream for the quality columns */
    for (int r = 0; r < depth; ++r) {
        uint64_t a = splitmix64(&s);
        int allele = (a & 1) ? a2 : a1;
        int fwd = (int)((a >> 1) & 1);
        uint32_t e24 = (uint32_t)((a >> 8) & 0xFFFFFF);
        uint32_t t_start = (uint32_t)((a >> 32) & 0xFFFF);
        uint32_t t_end = (uint32_t)((a >> 48) & 0xFFFF);
        int base = allele;
        if (e24 < t->thr_err) base = (allele + 1 + (int)(((a >> 2) & 0x3F) % 3)) & 3;
        if (t_start < t->thr_start) {
            uint64_t b = splitmix64(&s);
            *o++ = '^';
            *o++ = (char)(33 + (int)(b % 61));
        }
        if (base == ref) *o++ = fwd ? '.' : ',';
        else *o++ = fwd ? UP[base] : LOW[base];
        uint64_t c = splitmix64(&s);
        if ((uint32_t)(c & 0xFFFF) < t->thr_indel) {
            int len = 1 + (int)((c >> 16) % 11);
            *o++ = ((c >> 24) & 1) ? '+' : '-';
            o = put_uint(o, (uint64_t)len);
            uint64_t d = splitmix64(&s);
            for (int k = 0; k < len; ++k) { int bb = (int)((d >> (2 * k)) & 3); *o++ = fwd ? UP[bb] : LOW[bb]; }
        }
        if (t_end < t->thr_start) *o++ = '$';
    }
    (void)q;
    *o++ = '\t';
    for (int r = 0; r < depth; ++r) {
        uint64_t b = splitmix64(&s2);
        *o++ = t->qtab[b & (QTAB - 1)];
    }
    if (p->seven_columns) {
        *o++ = '\t';
        for (int r = 0; r < depth; ++r) {
            uint64_t b = splitmix64(&s2);
            *o++ = (char)(33 + 20 + (int)(b % 41));
        }
    }
    *o++ = '\n';
    return o;
}

static void locate(const pileup_gen_params* p, uint64_t site, int* chrom, uint64_t* pos) {
    uint64_t acc = 0;
    for (int c = 0; c < p->n_chroms; ++c) {
        uint64_t len = p->chrom_lengths ? p->chrom_lengths[c] : UINT64_MAX;
        if (c == p->n_chroms - 1 || site < acc + len) { *chrom = c; *pos = site - acc + 1; return; }
        acc += len;
    }
    *chrom = 0; *pos = site + 1;
}

#define BLOCK_SITES 16384

/* Generates sites [site_begin, site_begin + n_sites) into out (capacity cap).  Returns the number of
 * bytes the text takes; if that exceeds cap (or out is NULL) nothing beyond cap is written and the
 * caller should retry with a larger buffer.  Thread count: n_threads (<=0: OpenMP default). */
size_t pileup_gen(char* out, size_t cap, const pileup_gen_params* p, uint64_t site_begin, uint64_t n_sites,
                  int n_threads) {
    gen_tables t;
    build_tables(p, &t);
    uint64_t n_blocks = (n_sites + BLOCK_SITES - 1) / BLOCK_SITES;
    char** bufs = (char**)calloc((size_t)n_blocks ? (size_t)n_blocks : 1, sizeof(char*));
    size_t* lens = (size_t*)calloc((size_t)n_blocks + 1, sizeof(size_t));
    size_t max_chrom = 0;
    for (int c = 0; c < p->n_chroms; ++c) { size_t l = strlen(p->chrom_names[c]); if (l > max_chrom) max_chrom = l; }
    size_t per_site = site_bound(max_chrom, t.n_pois, p->seven_columns);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < (int64_t)n_blocks; ++b) {
        uint64_t s0 = site_begin + (uint64_t)b * BLOCK_SITES;
        uint64_t s1 = s0 + BLOCK_SITES;
        if (s1 > site_begin + n_sites) s1 = site_begin + n_sites;
        /* typical lines are far below the bound: grow a block buffer on demand */
        size_t bcap = (size_t)(s1 - s0) * (size_t)(max_chrom + 32 + 2.4 * (p->lambda + 1)) + per_site;
        char* buf = (char*)malloc(bcap);
        char* o = buf;
        for (uint64_t s = s0; s < s1; ++s) {
            if ((size_t)(o - buf) + per_site > bcap) {
                size_t used = (size_t)(o - buf);
                bcap = bcap * 2 + per_site;
                buf = (char*)realloc(buf, bcap);
                o = buf + used;
            }
            int c; uint64_t pos;
            locate(p, s, &c, &pos);
            o = gen_site(o, p, &t, s, p->chrom_names[c], strlen(p->chrom_names[c]), pos);
        }
        bufs[b] = buf;
        lens[b + 1] = (size_t)(o - buf);
    }
    for (uint64_t b = 0; b < n_blocks; ++b) lens[b + 1] += lens[b];
    size_t total = lens[n_blocks];
    if (out && total <= cap) {
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < (int64_t)n_blocks; ++b) memcpy(out + lens[b], bufs[b], lens[b + 1] - lens[b]);
    }
    for (uint64_t b = 0; b < n_blocks; ++b) free(bufs[b]);
    free(bufs); free(lens); free(t.pois);
    return total;
}

#ifdef PILEUP_GEN_MAIN
#include <stdio.h>
/* pileup_gen N lambda het err start indel seed seven [chrom] > file */
int main(int argc, char** argv) {
    if (argc < 9) {
        fprintf(stderr, "usage: %s n_sites lambda het err start indel seed seven_columns [chrom]\n", argv[0]);
        return 2;
    }
    const char* name = argc > 9 ? argv[9] : "chr1";
    pileup_gen_params p;
    memset(&p, 0, sizeof p);
    uint64_t n = strtoull(argv[1], 0, 10);
    p.lambda = atof(argv[2]); p.het = atof(argv[3]); p.err = atof(argv[4]);
    p.start = atof(argv[5]); p.indel = atof(argv[6]); p.seed = strtoull(argv[7], 0, 10);
    p.seven_columns = atoi(argv[8]);
    p.n_chroms = 1; p.chrom_names = &name; p.chrom_lengths = 0;
    const uint64_t step = 1u << 20;
    for (uint64_t s = 0; s < n; s += step) {
        uint64_t m = n - s < step ? n - s : step;
        size_t need = pileup_gen(0, 0, &p, s, m, 0);
        char* buf = (char*)malloc(need);
        pileup_gen(buf, need, &p, s, m, 0);
        fwrite(buf, 1, need, stdout);
        free(buf);
    }
    return 0;
}
#endif
