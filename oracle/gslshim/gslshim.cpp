// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Stand-in for the 13 GNU GSL symbols the reference imports (nm -u of its objects):
//   gsl_sf_lngamma            lynch.hpp:18,26
//   gsl_cdf_chisq_Q           stats.cpp:33,35
//   gsl_vector_{alloc,free,get,set}, gsl_multimin_fminimizer_{alloc,free,set,iterate,size},
//   gsl_multimin_fminimizer_nmsimplex2, gsl_multimin_test_size      optimization.hpp:37-95
//
// GSL is absent from this image and un-vendored/unpinned in the reference (configure.ac:14-16),
// so these are restatements of the *published* algorithms:
//   * lngamma(x)            -> C99 lgamma (the reference only calls it with integer x >= 1)
//   * chisq_Q(x, nu=1)      -> erfc(sqrt(x/2))   (exact identity for one degree of freedom)
//   * nmsimplex2            -> Nelder-Mead with the O(N) centre/size update of the "simplex2"
//                              variant: reflect(-1) / expand(-2) / contract(0.5) / shrink(0.5).
// PARITY UNPINNED against real GSL: no GSL build and no reference test covers this boundary.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "gsl/gsl_cdf.h"
#include "gsl/gsl_multimin.h"
#include "gsl/gsl_sf_gamma.h"

extern "C" {

double gsl_sf_lngamma(double x) { return std::lgamma(x); }

double gsl_cdf_chisq_Q(double x, double nu) {
    if (nu != 1.0) { std::abort(); }  // the reference only ever asks for 1 degree of freedom
    if (!(x > 0)) { return 1.0; }
    return std::erfc(std::sqrt(x / 2.0));
}

gsl_vector* gsl_vector_alloc(size_t n) {
    gsl_vector* v = new gsl_vector;
    v->size = n;
    v->data = new double[n]();
    return v;
}
void gsl_vector_free(gsl_vector* v) {
    if (v) { delete[] v->data; delete v; }
}
double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i]; }
void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i] = x; }

}  // extern "C"

namespace {

struct SimplexState {
    size_t n = 0;                       // dimension; the simplex has P = n + 1 corners
    std::vector<std::vector<double>> X; // corners
    std::vector<double> Y;              // f at the corners
    std::vector<double> center;         // mean of all corners
    double S2 = 0;                      // mean squared distance of the corners to the centre
    unsigned long count = 0;
};

double eval(gsl_multimin_function* f, const std::vector<double>& x) {
    gsl_vector v {x.size(), const_cast<double*>(x.data())};
    return f->f(&v, f->params);
}

void computeCenter(SimplexState& s) {
    const size_t P = s.n + 1;
    for (size_t j = 0; j < s.n; ++j) {
        double acc = 0;
        for (size_t k = 0; k < P; ++k) acc += s.X[k][j];
        s.center[j] = acc / double(P);
    }
}

double computeSize(SimplexState& s) {
    const size_t P = s.n + 1;
    double ss = 0;
    for (size_t k = 0; k < P; ++k) {
        double t = 0;
        for (size_t j = 0; j < s.n; ++j) {
            const double d = s.X[k][j] - s.center[j];
            t += d * d;
        }
        ss += t;
    }
    s.S2 = ss / double(P);
    return std::sqrt(s.S2);
}

// xc = (1-coeff)*P/(P-1) * centre + (P*coeff-1)/(P-1) * X[corner]
double tryCornerMove(double coeff, const SimplexState& s, size_t corner, std::vector<double>& xc,
                     gsl_multimin_function* f) {
    const double P = double(s.n + 1);
    const double alpha = (1 - coeff) * P / (P - 1.0);
    const double beta = (P * coeff - 1.0) / (P - 1.0);
    for (size_t j = 0; j < s.n; ++j) xc[j] = alpha * s.center[j] + beta * s.X[corner][j];
    return eval(f, xc);
}

void updatePoint(SimplexState& s, size_t i, const std::vector<double>& x, double val) {
    const double P = double(s.n + 1);
    double d2 = 0, xmcd = 0;
    for (size_t j = 0; j < s.n; ++j) {
        const double delta = x[j] - s.X[i][j];
        const double xmc = s.X[i][j] - s.center[j];
        d2 += delta * delta;
        xmcd += xmc * delta;
    }
    const double d = std::sqrt(d2);
    s.S2 += (2.0 / P) * xmcd + ((P - 1.0) / P) * (d * d / P);
    for (size_t j = 0; j < s.n; ++j) {
        s.center[j] -= (1.0 / P) * s.X[i][j];
        s.center[j] += (1.0 / P) * x[j];
    }
    s.X[i] = x;
    s.Y[i] = val;
}

int contractByBest(SimplexState& s, size_t best, gsl_multimin_function* f) {
    const size_t P = s.n + 1;
    int status = GSL_SUCCESS;
    for (size_t i = 0; i < P; ++i) {
        if (i == best) continue;
        for (size_t j = 0; j < s.n; ++j) s.X[i][j] = 0.5 * (s.X[i][j] + s.X[best][j]);
        s.Y[i] = eval(f, s.X[i]);
        if (!std::isfinite(s.Y[i])) status = GSL_EBADFUNC;
    }
    computeCenter(s);
    computeSize(s);
    return status;
}

const gsl_multimin_fminimizer_type NMSIMPLEX2 {"nmsimplex2"};

}  // namespace

extern "C" {

const gsl_multimin_fminimizer_type* gsl_multimin_fminimizer_nmsimplex2 = &NMSIMPLEX2;

gsl_multimin_fminimizer* gsl_multimin_fminimizer_alloc(const gsl_multimin_fminimizer_type* T, size_t n) {
    gsl_multimin_fminimizer* m = new gsl_multimin_fminimizer;
    m->type = T;
    m->f = nullptr;
    m->fval = 0;
    m->x = gsl_vector_alloc(n);
    m->size = 0;
    SimplexState* s = new SimplexState;
    s->n = n;
    s->X.assign(n + 1, std::vector<double>(n, 0.0));
    s->Y.assign(n + 1, 0.0);
    s->center.assign(n, 0.0);
    m->state = s;
    return m;
}

void gsl_multimin_fminimizer_free(gsl_multimin_fminimizer* m) {
    if (!m) return;
    delete static_cast<SimplexState*>(m->state);
    gsl_vector_free(m->x);
    delete m;
}

int gsl_multimin_fminimizer_set(gsl_multimin_fminimizer* m, gsl_multimin_function* f,
                                const gsl_vector* x, const gsl_vector* step_size) {
    SimplexState& s = *static_cast<SimplexState*>(m->state);
    m->f = f;
    for (size_t j = 0; j < s.n; ++j) m->x->data[j] = x->data[j];
    // corner 0 is x0, corner i+1 is x0 + step_i * e_i
    for (size_t j = 0; j < s.n; ++j) s.X[0][j] = x->data[j];
    s.Y[0] = eval(f, s.X[0]);
    if (!std::isfinite(s.Y[0])) return GSL_EBADFUNC;
    for (size_t i = 0; i < s.n; ++i) {
        s.X[i + 1] = s.X[0];
        s.X[i + 1][i] += step_size->data[i];
        s.Y[i + 1] = eval(f, s.X[i + 1]);
        if (!std::isfinite(s.Y[i + 1])) return GSL_EBADFUNC;
    }
    computeCenter(s);
    m->size = computeSize(s);
    s.count++;
    return GSL_SUCCESS;
}

int gsl_multimin_fminimizer_iterate(gsl_multimin_fminimizer* m) {
    SimplexState& s = *static_cast<SimplexState*>(m->state);
    gsl_multimin_function* f = m->f;
    const size_t P = s.n + 1;
    std::vector<double> xc(s.n), xc2(s.n);

    // highest, second highest and lowest corner
    size_t hi = 0, s_hi = 1, lo = 0;
    double dhi = s.Y[0], dlo = s.Y[0], ds_hi = s.Y[1];
    for (size_t i = 1; i < P; ++i) {
        const double val = s.Y[i];
        if (val < dlo) { dlo = val; lo = i; }
        else if (val > dhi) { ds_hi = dhi; s_hi = hi; dhi = val; hi = i; }
        else if (val > ds_hi) { ds_hi = val; s_hi = i; }
    }

    int status = GSL_SUCCESS;
    double val = tryCornerMove(-1.0, s, hi, xc, f);  // reflect
    if (std::isfinite(val) && val < s.Y[lo]) {
        const double val2 = tryCornerMove(-2.0, s, hi, xc2, f);  // expand
        if (std::isfinite(val2) && val2 < s.Y[lo]) updatePoint(s, hi, xc2, val2);
        else updatePoint(s, hi, xc, val);
    } else if (!std::isfinite(val) || val > s.Y[s_hi]) {
        if (std::isfinite(val) && val <= s.Y[hi]) updatePoint(s, hi, xc, val);
        const double val2 = tryCornerMove(0.5, s, hi, xc2, f);  // contract
        if (std::isfinite(val2) && val2 <= s.Y[hi]) updatePoint(s, hi, xc2, val2);
        else status = contractByBest(s, lo, f);  // shrink towards the best corner
    } else {
        updatePoint(s, hi, xc, val);
    }
    if (status != GSL_SUCCESS) return status;

    lo = 0;
    for (size_t i = 1; i < P; ++i) if (s.Y[i] < s.Y[lo]) lo = i;
    for (size_t j = 0; j < s.n; ++j) m->x->data[j] = s.X[lo][j];
    m->fval = s.Y[lo];
    m->size = s.S2 > 0 ? std::sqrt(s.S2) : computeSize(s);
    return GSL_SUCCESS;
}

double gsl_multimin_fminimizer_size(const gsl_multimin_fminimizer* m) { return m->size; }

int gsl_multimin_test_size(double size, double epsabs) {
    return size < epsabs ? GSL_SUCCESS : GSL_CONTINUE;
}

}  // extern "C"
