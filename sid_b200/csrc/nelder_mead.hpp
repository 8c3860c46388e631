// Host-side Nelder-Mead driver for the Lynch fit: the product's counterpart of
// FunctionMinimizer<2>::run (optimization.hpp:51-89) around GSL's nmsimplex2 minimiser.
// The simplex bookkeeping is a handful of scalar operations per iteration and stays on the host;
// every objective evaluation is a device reduction (K4).  GSL itself is not available, so the
// update rules follow the published algorithm: reflect (-1), expand (-2), contract (0.5), shrink
// towards the best corner (0.5), size = sqrt(mean squared distance to the centroid), maintained
// incrementally.  Stop rule of the reference: size < 1e-5, at most 1000 iterations.
// The same code runs on the host (one device reduction per evaluation, sidgpu_lynch_objective) and inside
// k_lynch_fit, where every thread of a cooperative grid walks the identical trajectory and an evaluation is a
// grid-wide reduction: the whole fit is then one kernel launch.
#pragma once
#include <cmath>
#if !defined(__CUDACC__)
#include <functional>
#endif

#include "common.cuh"

namespace sid {

struct NelderMeadResult {
    double x[2];
    double fval;
    int iterations;
    int evaluations;
    bool converged;
};

SID_HD bool nm_finite(double v) { return v - v == 0.0; }      // false for NaN and +-inf

#if defined(__CUDACC__)
#pragma nv_exec_check_disable       // F is a host lambda in the host instantiation, a device functor in the kernel's
#endif
template <class F>
SID_HD NelderMeadResult nelder_mead_2d(F&& f, const double x0[2], const double step[2], double size_eps = 1e-5,
                                       int max_iterations = 1000) {
    constexpr int P = 3;
    double X[P][2], Y[P], c[2], S2 = 0;
    int evals = 0;
    auto eval = [&](const double* x) { ++evals; return f(x[0], x[1]); };
    auto center = [&]() {
        for (int j = 0; j < 2; ++j) c[j] = (X[0][j] + X[1][j] + X[2][j]) / 3.0;
    };
    auto full_size = [&]() {
        double ss = 0;
        for (int k = 0; k < P; ++k) {
            double t = 0;
            for (int j = 0; j < 2; ++j) { const double d = X[k][j] - c[j]; t += d * d; }
            ss += t;
        }
        S2 = ss / P;
        return sqrt(S2);
    };
    auto move = [&](double coeff, int corner, double* xc) {
        const double alpha = (1 - coeff) * P / (P - 1.0);
        const double beta = (P * coeff - 1.0) / (P - 1.0);
        for (int j = 0; j < 2; ++j) xc[j] = alpha * c[j] + beta * X[corner][j];
        return eval(xc);
    };
    auto update = [&](int i, const double* x, double val) {
        double d2 = 0, xmcd = 0;
        for (int j = 0; j < 2; ++j) {
            const double delta = x[j] - X[i][j];
            const double xmc = X[i][j] - c[j];
            d2 += delta * delta;
            xmcd += xmc * delta;
        }
        const double d = sqrt(d2);
        S2 += (2.0 / P) * xmcd + ((P - 1.0) / P) * (d * d / P);
        for (int j = 0; j < 2; ++j) {
            c[j] -= (1.0 / P) * X[i][j];
            c[j] += (1.0 / P) * x[j];
            X[i][j] = x[j];
        }
        Y[i] = val;
    };

    for (int k = 0; k < P; ++k) { X[k][0] = x0[0]; X[k][1] = x0[1]; }
    X[1][0] += step[0];
    X[2][1] += step[1];
    for (int k = 0; k < P; ++k) Y[k] = eval(X[k]);
    center();
    double size = full_size();

    NelderMeadResult r {{x0[0], x0[1]}, Y[0], 0, 0, false};
    int it = 0;
    bool go = true, failed = false;
    while (go) {
        ++it;
        int hi = 0, s_hi = 1, lo = 0;
        double dhi = Y[0], dlo = Y[0], ds_hi = Y[1];
        for (int k = 1; k < P; ++k) {
            const double v = Y[k];
            if (v < dlo) { dlo = v; lo = k; }
            else if (v > dhi) { ds_hi = dhi; s_hi = hi; dhi = v; hi = k; }
            else if (v > ds_hi) { ds_hi = v; s_hi = k; }
        }
        double xc[2], xc2[2];
        const double val = move(-1.0, hi, xc);
        if (nm_finite(val) && val < Y[lo]) {
            const double val2 = move(-2.0, hi, xc2);
            if (nm_finite(val2) && val2 < Y[lo]) update(hi, xc2, val2); else update(hi, xc, val);
        } else if (!nm_finite(val) || val > Y[s_hi]) {
            if (nm_finite(val) && val <= Y[hi]) update(hi, xc, val);
            const double val2 = move(0.5, hi, xc2);
            if (nm_finite(val2) && val2 <= Y[hi]) {
                update(hi, xc2, val2);
            } else {
                for (int k = 0; k < P; ++k) {
                    if (k == lo) continue;
                    for (int j = 0; j < 2; ++j) X[k][j] = 0.5 * (X[k][j] + X[lo][j]);
                    Y[k] = eval(X[k]);
                    if (!nm_finite(Y[k])) failed = true;
                }
                center();
                full_size();
            }
        } else {
            update(hi, xc, val);
        }
        if (failed) break;                       // optimization.hpp:62-64
        lo = 0;
        for (int k = 1; k < P; ++k) if (Y[k] < Y[lo]) lo = k;
        r.x[0] = X[lo][0];
        r.x[1] = X[lo][1];
        r.fval = Y[lo];
        size = S2 > 0 ? sqrt(S2) : full_size();
        r.converged = size < size_eps;            // optimization.hpp:66-67
        go = !r.converged && it < max_iterations; // optimization.hpp:72
    }
    r.iterations = it;
    r.evaluations = evals;
    if (failed) r.converged = true;               // the reference only reports GSL_CONTINUE as failure
    return r;
}

}  // namespace sid
