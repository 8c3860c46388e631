// Per-profile genotype calls in double precision, log space.
// The reference multiplies x87 long double powers (lynch.hpp:48-96) and feeds ratios to
// likelihoodRatioTest (stats.cpp:29-37); the multinomial coefficient cancels in every ratio, and
// gsl_cdf_chisq_Q(x, 1) == erfc(sqrt(x/2)), so each p-value is erfc(sqrt(log l_H1 - log l_H0)).
#pragma once
#include <math.h>

#include "common.cuh"
#include "fmt.cuh"

namespace sid {

struct CallResult {
    uint8_t label;   // 0 "hom", 1 "het", 255 dropped (coverage < 4 under bayes / likelihood_ratio)
    char gt0, gt1;
    double hom, het; // hom_conf, het_conf
};

SID_HD double neg_inf() { return bits_double(0xFFF0000000000000ull); }

// powl(y, k) in log space: k * log(y), with powl(anything, 0) == 1 (also for NaN and 0).
SID_HD double klog(uint32_t k, double logy) { return k == 0 ? 0.0 : (double)k * logy; }

// getMajorAlleleIndices (call.cpp:52-60): stable ascending sort of the 4 counts, so among equal
// counts the higher A<C<G<T index ranks higher.
SID_HD void major_alleles(uint64_t profile, int& first, int& second) {
    uint32_t c[4] = {profile_count(profile, 0), profile_count(profile, 1), profile_count(profile, 2),
                     profile_count(profile, 3)};
    int f = 0;
    for (int i = 1; i < 4; ++i) if (c[i] >= c[f]) f = i;
    int s = f == 0 ? 1 : 0;
    for (int i = 0; i < 4; ++i) if (i != f && c[i] >= c[s]) s = i;
    first = f;
    second = s;
}

// likelihoodRatioTest(l_H0, l_H1) (stats.cpp:29-37) on log-likelihoods; -inf stands for l == 0.
SID_HD double lrt_log(double log_h0, double log_h1) {
    if (log_h0 == neg_inf()) return 0.0;            // stats.cpp:35: Q(DBL_MAX) == 0
    const double d = log_h1 - log_h0;
    if (!(d > 0)) return 1.0;                       // chisq <= 0 -> Q == 1
    return erfc(sqrt(d));
}

SID_HD char base_char(int i) { return (char)("ACGT"[i]); }

// callSiteMLError body (call.cpp:238-273), `-m local`.
SID_HD CallResult call_local(uint64_t profile, double prior, double error_threshold, double alpha) {
    int f, s;
    major_alleles(profile, f, s);
    const uint32_t n = profile_coverage(profile);
    const uint32_t nf = profile_count(profile, f), ns = profile_count(profile, s);
    double e1 = (double)(n - nf) / (double)n;                       // call.cpp:243 (0/0 -> NaN at depth 0)
    if (e1 > error_threshold) e1 = error_threshold;
    double e2 = 1.5 * (double)(n - nf - ns) / (double)n;            // call.cpp:250
    if (e2 > error_threshold) e2 = error_threshold;
    double ll1 = klog(nf, log(1.0 - e1)) + klog(n - nf, log(e1 / 3.0));                               // lynch.hpp:92-96
    double ll2 = klog(nf + ns, log((1.0 - 2. / 3. * e2) / 2.0)) + klog(n - nf - ns, log(e2 / 3.0));   // lynch.hpp:76-80
    if (prior > 0) {                                                // call.cpp:256-259
        ll1 += log1p(-prior);
        ll2 += log(prior);
    }
    CallResult r;
    r.hom = lrt_log(ll2, ll1);                                      // call.cpp:261
    r.het = lrt_log(ll1, ll2);                                      // call.cpp:262
    r.label = 0;
    r.gt0 = r.gt1 = base_char(f);
    if (ll2 > ll1 && r.het < alpha) { r.label = 1; r.gt1 = base_char(s); }   // call.cpp:266-269
    return r;
}

// Constants of one (nucleotide distribution, epsilon) pair for the mixture likelihoods.
struct LynchConsts {
    double lnd[4];     // log nd_i
    double lpair[6];   // log(nd_i nd_j), i<j in the order 01 02 03 12 13 23
    double A, B, C;    // log(1-eps), log(eps/3), log((1-2eps/3)/2)
    double lnorm;      // log(1 - sum nd_i^2)   (lynch.hpp:68-72)
};

SID_HD LynchConsts lynch_consts(const double nd[4], double eps) {
    LynchConsts k;
    for (int i = 0; i < 4; ++i) k.lnd[i] = log(nd[i]);
    int t = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) k.lpair[t++] = log(nd[i] * nd[j]);
    k.A = log(1.0 - eps);
    k.B = log(eps / 3.0);
    k.C = log((1.0 - 2. / 3. * eps) / 2.0);
    double s = 0;
    for (int i = 0; i < 4; ++i) s += nd[i] * nd[i];
    k.lnorm = log(1.0 - s);
    return k;
}

// log multinomialCoefficient (lynch.hpp:48-55)
SID_HD double log_multinomial(uint64_t profile) {
    const uint32_t n = profile_coverage(profile);
    double v = lgamma((double)n + 1.0);
    for (int i = 0; i < 4; ++i) v -= lgamma((double)profile_count(profile, i) + 1.0);
    return v;
}

// log of homozygousLikelihood / heterozygousLikelihood with the distribution-weighted overloads
// (lynch.hpp:57-74, 82-90), WITHOUT the multinomial coefficient.
SID_HD void lynch_loglik(uint64_t profile, const LynchConsts& k, double& lhom, double& lhet) {
    const uint32_t n = profile_coverage(profile);
    uint32_t c[4] = {profile_count(profile, 0), profile_count(profile, 1), profile_count(profile, 2),
                     profile_count(profile, 3)};
    double h[4], m = neg_inf();
    for (int i = 0; i < 4; ++i) {
        h[i] = k.lnd[i] + klog(c[i], k.A) + klog(n - c[i], k.B);
        if (h[i] > m) m = h[i];
    }
    if (m == neg_inf() || m != m) lhom = m;
    else {
        double s = 0;
        for (int i = 0; i < 4; ++i) s += exp(h[i] - m);
        lhom = m + log(s);
    }
    double t[6];
    m = neg_inf();
    int q = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) {
            t[q] = k.lpair[q] + klog(c[i] + c[j], k.C) + klog(n - c[i] - c[j], k.B);
            if (t[q] > m) m = t[q];
            ++q;
        }
    if (m == neg_inf() || m != m) lhet = m;
    else {
        double s = 0;
        for (int i = 0; i < 6; ++i) s += exp(t[i] - m);
        lhet = m + log(s) - k.lnorm;
    }
}

// One term of compoundLikelihood (lynch.cpp:46-52): log((1-pi) L_hom + pi L_het) for one profile,
// logM included.  Returns false when the reference would skip the term (L <= 0).
SID_HD bool lynch_term(uint64_t profile, double logM, const LynchConsts& k, double log1m_pi, double log_pi,
                       double& out) {
    double lhom, lhet;
    lynch_loglik(profile, k, lhom, lhet);
    const double a = log1m_pi + lhom, b = log_pi + lhet;
    const double m = a > b ? a : b;
    if (!(m > neg_inf())) return false;             // both -inf or NaN: L <= 0 (lynch.cpp:49)
    const double lo = a > b ? b : a;
    out = logM + m + log1p(exp(lo - m));
    return true;
}

// callBayes body (call.cpp:176-194)
SID_HD CallResult call_bayes(uint64_t profile, const LynchConsts& k, double pi) {
    CallResult r;
    int f, s;
    major_alleles(profile, f, s);
    r.gt0 = r.gt1 = base_char(f);
    r.label = 0;
    if (profile_coverage(profile) < 4) { r.label = 255; r.hom = r.het = 0; return r; }   // call.cpp:149-153
    double lhom, lhet;
    lynch_loglik(profile, k, lhom, lhet);
    const double a = lhom + log1p(-pi), b = lhet + log(pi);
    r.hom = 1.0 / (1.0 + exp(b - a));
    r.het = 1.0 / (1.0 + exp(a - b));
    if (r.het > r.hom) { r.label = 1; r.gt1 = base_char(s); }
    return r;
}

// callLikelihoodRatio per-profile p-values before the Benjamini-Hochberg step (call.cpp:93-103)
SID_HD void lr_pvalues(uint64_t profile, const LynchConsts& k, bool use_prior, double pi, double& p_hom,
                       double& p_het) {
    double lhom, lhet;
    lynch_loglik(profile, k, lhom, lhet);
    if (use_prior) { lhet += log(pi); lhom += log1p(-pi); }
    p_hom = lrt_log(lhet, lhom);
    p_het = lrt_log(lhom, lhet);
}

// The text every site with this profile prints after "chrom,pos": ",label,gt,hom,het,type\n"
// (call.hpp:31-36).  At most 48 bytes.  Dropped profiles give length 0.
SID_HD int format_suffix(const CallResult& r, bool probability, char* out) {
    if (r.label == 255) return 0;
    int n = 0;
    out[n++] = ',';
    out[n++] = 'h';
    if (r.label) { out[n++] = 'e'; out[n++] = 't'; } else { out[n++] = 'o'; out[n++] = 'm'; }
    out[n++] = ',';
    out[n++] = r.gt0;
    out[n++] = r.gt1;
    out[n++] = ',';
    n += fmt_g6(r.hom, out + n);
    out[n++] = ',';
    n += fmt_g6(r.het, out + n);
    out[n++] = ',';
    const char* t = probability ? "probability" : "p_value";
    for (int i = 0; t[i]; ++i) out[n++] = t[i];
    out[n++] = '\n';
    return n;
}

// Neumaier compensated accumulation: stands in for the reference's long double running sums.
struct CompSum {
    double s, c;
    SID_HD void init() { s = 0; c = 0; }
    SID_HD void add(double x) {               // Knuth's branch-free two-sum: t + e == s + x exactly
        const double t = s + x;
        const double z = t - s;
        c += (s - (t - z)) + (x - z);
        s = t;
    }
    SID_HD double value() const { return s + c; }
};

}  // namespace sid
