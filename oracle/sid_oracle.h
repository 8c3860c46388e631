/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's genotype-calling path.
 *
 * Plain-C restatement of EvolBioInf/sid (call.cpp, pileup.cpp, lynch.cpp/.hpp, stats.cpp,
 * optimization.hpp), one function per reference function, each citing the file:line it follows.
 * It exists to CHECK the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library never links it.
 *
 * Arithmetic follows the reference literally, including x87 80-bit long double (gcc on x86-64).
 * Third-party arithmetic not under /root/reference: GNU GSL (libgsl, version unpinned by
 * configure.ac:14-16).  gsl_sf_lngamma -> lgamma, gsl_cdf_chisq_Q(x,1) -> erfc(sqrt(x/2)),
 * nmsimplex2 -> Nelder-Mead restated from its published description.  PARITY UNPINNED at that
 * boundary (no GSL here, no reference test pins it); everything else is pinned against the
 * reference's own test vectors and against oracle/_ref (the reference compiled from its sources).
 */
#ifndef SID_ORACLE_H
#define SID_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_OK = 0, ORC_MALFORMED = 1, ORC_MALFORMED_OR_MISSING = 2, ORC_REFERENCE_UB = 3 };
enum { ORC_LOCAL = 0, ORC_BAYES = 1, ORC_LIKELIHOOD_RATIO = 2, ORC_QUALITY = 3 };

typedef struct {
    uint16_t profile[4];
    uint32_t count;
    uint32_t coverage;
} orc_unique_profile; /* pileup.hpp:32-40 */

typedef struct {
    size_t n;            /* output records (sites dropped by the coverage<4 rule are absent) */
    size_t n_sites;      /* parsed input lines */
    uint32_t* chrom_off; /* byte offset of the chromosome name inside the input text */
    uint16_t* chrom_len;
    int32_t* pos;
    uint8_t* label;      /* 0 "hom", 1 "het" */
    char* gt;            /* 2 chars per record */
    double* hom_conf;
    double* het_conf;
    int conf_type;       /* 0 "p_value", 1 "probability" */
    uint16_t* profiles;  /* 4 per parsed site (n_sites entries) */
    size_t n_unique;     /* unique profiles seen by the Lynch fit / histogram (after coverage filter) */
    double heterozygosity, error_rate; /* fitted (pi, eps), NaN when no fit ran */
    int iterations;      /* Nelder-Mead iterations, objective evaluations */
    int evaluations;
    int converged;
} orc_result;

/* pileup.cpp:70-153.  `bases` is NUL terminated.  bases_out (may be NULL) receives upper-cased
 * counted bases; returns how many bases were counted. */
size_t orc_parse_read_bases(const char* bases, char reference, uint16_t counts[4], char* bases_out);
/* pileup.cpp:155-167 */
size_t orc_parse_qualities(const char* q, uint8_t* out);
/* pileup.cpp:13-68; mutates `line` like strtok_r does.  Pointers point into `line`. */
typedef struct {
    const char* chrom;
    int32_t pos;
    char ref;
    uint16_t counts[4];
    const char* bases;
    const char* bq;
    const char* mq;
} orc_line;
int orc_parse_line(char* line, int want_bq, int want_mq, orc_line* out);

/* pileup.cpp:169-196: sort + run-length encode; returns malloc'd array */
orc_unique_profile* orc_count_unique(const uint16_t* profiles, size_t n, size_t* n_unique);
/* pileup.cpp:198-217 (including the uint32 product of count*coverage) */
void orc_nucleotide_distribution(const orc_unique_profile* u, size_t n, double nd[4]);
/* call.cpp:52-60 */
void orc_major_alleles(const uint16_t profile[4], int* first, int* second);
/* lynch.hpp:48-96 */
long double orc_multinomial_coefficient(const orc_unique_profile* p);
long double orc_hom_likelihood_ref(const orc_unique_profile* p, double e, int ref);
long double orc_het_likelihood_ref(const orc_unique_profile* p, double e, int r0, int r1);
long double orc_hom_likelihood_nd(const orc_unique_profile* p, double e, const double nd[4]);
long double orc_het_likelihood_nd(const orc_unique_profile* p, double e, const double nd[4]);
/* stats.cpp:29-37 */
double orc_lrt(long double l_h0, long double l_h1);
/* stats.cpp:58-80 */
void orc_bh(const double* p, size_t n, double* adjusted);
/* lynch.cpp:37-61 */
double orc_compound_likelihood(const orc_unique_profile* u, size_t n, const double nd[4], double pi, double eps);
/* lynch.cpp:17-35 + optimization.hpp:51-89 */
void orc_estimate(const orc_unique_profile* u, size_t n, const double nd[4], double* pi, double* eps,
                  int* iterations, int* evaluations, int* converged);

/* call.cpp:62-372 (the four methods) over a whole pileup text (not modified; copied internally). */
int orc_call(const char* text, size_t len, int method, int estimate_prior, double prior,
             double error_threshold, double significance_level, orc_result* out);
void orc_free_result(orc_result* r);
/* call.hpp:29-38 + sid.cpp:102-105: header line + one row per record, doubles as %g.
 * Returns bytes needed; writes at most cap bytes. */
size_t orc_write_csv(const char* text, const orc_result* r, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
