// TEST INFRASTRUCTURE ONLY -- C-callable handles onto the UNMODIFIED reference functions, so the
// tests can pin the C restatement (oracle/sid_oracle.c) against the reference itself.
// Linked with the reference's own translation units compiled from /root/reference (see Makefile);
// output goes to oracle/_ref/libsidref.so, which is git-ignored.  Nothing here is shipped.
#include <cstdint>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "call.hpp"
#include "lynch.hpp"
#include "pileup.hpp"
#include "stats.hpp"

extern "C" {

// pileup.cpp:70-153
void ref_parse_read_bases(const char* bases, char ref, uint16_t out[4], int* n_bases) {
    ReadStack s = parseReadBases(bases, ref, 0);
    for (int i = 0; i < 4; ++i) out[i] = s.counts[i];
    if (n_bases) *n_bases = int(s.bases.size());
}

// pileup.cpp:13-68 ; returns 0 ok, 1 "Malformed pileup line", 2 "... or missing mapping qualities"
int ref_parse_line(const char* line, int want_bq, int want_mq, char* chrom, size_t chrom_cap, int* pos,
                   char* refbase, uint16_t counts[4], char* bases, uint8_t* bq, uint8_t* mq,
                   int* n_bases, int* n_bq, int* n_mq, size_t cap) {
    std::string copy(line);
    try {
        PileupLine p = parsePileupLine(&copy[0], want_bq != 0, want_mq != 0);
        std::strncpy(chrom, p.chromosome_name.c_str(), chrom_cap - 1);
        chrom[chrom_cap - 1] = 0;
        *pos = p.position;
        *refbase = p.reference_base;
        for (int i = 0; i < 4; ++i) counts[i] = p.base_counts[i];
        *n_bases = int(p.bases.size());
        *n_bq = int(p.base_qualities.size());
        *n_mq = int(p.mapping_qualities.size());
        for (size_t i = 0; i < p.bases.size() && i < cap; ++i) bases[i] = p.bases[i];
        for (size_t i = 0; i < p.base_qualities.size() && i < cap; ++i) bq[i] = p.base_qualities[i];
        for (size_t i = 0; i < p.mapping_qualities.size() && i < cap; ++i) mq[i] = p.mapping_qualities[i];
        return 0;
    } catch (const std::invalid_argument& e) {
        return std::strstr(e.what(), "missing mapping") ? 2 : 1;
    }
}

static std::vector<UniqueProfile> makeProfiles(size_t n, const uint16_t* prof, const uint32_t* count) {
    std::vector<UniqueProfile> v;
    v.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        profile_t p {prof[4 * i], prof[4 * i + 1], prof[4 * i + 2], prof[4 * i + 3]};
        v.emplace_back(p, count[i]);
    }
    return v;
}

// pileup.cpp:169-196 on n sites given as profiles; returns number of unique profiles written
size_t ref_count_unique(size_t n, const uint16_t* prof, uint16_t* uprof, uint32_t* ucount) {
    std::vector<PileupLine> lines(n);
    for (size_t i = 0; i < n; ++i)
        lines[i].base_counts = {prof[4 * i], prof[4 * i + 1], prof[4 * i + 2], prof[4 * i + 3]};
    auto u = countUniqueProfiles(lines);
    for (size_t i = 0; i < u.size(); ++i) {
        for (int k = 0; k < 4; ++k) uprof[4 * i + k] = u[i].profile[k];
        ucount[i] = u[i].count;
    }
    return u.size();
}

// pileup.cpp:198-217
void ref_nucleotide_distribution(size_t n, const uint16_t* prof, const uint32_t* count, double nd[4]) {
    auto d = computeNucleotideDistribution(makeProfiles(n, prof, count));
    for (int i = 0; i < 4; ++i) nd[i] = d[i];
}

// lynch.cpp:37-61
double ref_compound_likelihood(size_t n, const uint16_t* prof, const uint32_t* count, const double nd[4],
                               double pi, double eps) {
    auto profiles = makeProfiles(n, prof, count);
    struct { const std::vector<UniqueProfile>& profiles; const std::array<double, 4> nd; } params {
        profiles, {nd[0], nd[1], nd[2], nd[3]}};
    double xv[2] = {pi, eps};
    gsl_vector v {2, xv};
    return compoundLikelihood(&v, &params);
}

// lynch.cpp:17-35
void ref_estimate(size_t n, const uint16_t* prof, const uint32_t* count, const double nd[4], double* pi,
                  double* eps, long double* L_hom, long double* L_het) {
    auto profiles = makeProfiles(n, prof, count);
    auto est = estimateProfileGenotypeLikelihoods(profiles, {nd[0], nd[1], nd[2], nd[3]});
    *pi = est.heterozygosity;
    *eps = est.error_rate;
    for (size_t i = 0; i < n; ++i) {
        if (L_hom) L_hom[i] = est.profile_likelihoods[i].L_homozygous;
        if (L_het) L_het[i] = est.profile_likelihoods[i].L_heterozygous;
    }
}

// lynch.hpp:57-96 (distribution-weighted overloads)
void ref_mixture_likelihoods(const uint16_t prof[4], const double nd[4], double eps, long double* L_hom,
                             long double* L_het) {
    UniqueProfile p({prof[0], prof[1], prof[2], prof[3]}, 1);
    std::array<double, 4> d {nd[0], nd[1], nd[2], nd[3]};
    *L_hom = homozygousLikelihood(p, eps, d);
    *L_het = heterozygousLikelihood(p, eps, d);
}

// stats.cpp:29-37
double ref_lrt(long double l_h0, long double l_h1) { return likelihoodRatioTest(l_h0, l_h1); }

// stats.cpp:58-80
void ref_bh(size_t n, const double* p, double* out) {
    std::vector<double> in(p, p + n);
    auto adj = adjustBenjaminiHochberg(in);
    for (size_t i = 0; i < n; ++i) out[i] = adj[i];
}

// call.cpp:52-60 is file-local in the reference (not in call.hpp) but has external linkage
}
std::pair<int, int> getMajorAlleleIndices(const UniqueProfile p);
extern "C" void ref_major_alleles(const uint16_t prof[4], int* first, int* second) {
    UniqueProfile p({prof[0], prof[1], prof[2], prof[3]}, 1);
    auto m = getMajorAlleleIndices(p);
    *first = m.first;
    *second = m.second;
}
