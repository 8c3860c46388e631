"""Host-side Nelder-Mead driver, the Python twin of sid_b200/csrc/nelder_mead.hpp (the product's
counterpart of FunctionMinimizer<2>::run, optimization.hpp:51-89, around GSL's nmsimplex2).  Used when
the objective has to cross ranks: every evaluation is a device reduction on each GPU followed by one
all-reduce, and every rank runs this loop on identical values."""
import math


def nelder_mead_2d(f, x0=(1e-3, 1e-3), step=(1e-4, 1e-4), size_eps=1e-5, max_iterations=1000):
    """Returns dict(x, fval, iterations, evaluations, converged).  Start point and steps default to
    lynch.cpp:8-10,20; stop rule to optimization.hpp:26,66-72."""
    P = 3
    evals = [0]

    def ev(x):
        evals[0] += 1
        return f(x[0], x[1])

    X = [[x0[0], x0[1]] for _ in range(P)]
    X[1][0] += step[0]
    X[2][1] += step[1]
    Y = [ev(x) for x in X]
    c = [0.0, 0.0]
    S2 = [0.0]

    def center():
        for j in range(2):
            c[j] = (X[0][j] + X[1][j] + X[2][j]) / 3.0

    def full_size():
        ss = 0.0
        for k in range(P):
            t = 0.0
            for j in range(2):
                d = X[k][j] - c[j]
                t += d * d
            ss += t
        S2[0] = ss / P
        return math.sqrt(S2[0])

    def move(coeff, corner):
        alpha = (1 - coeff) * P / (P - 1.0)
        beta = (P * coeff - 1.0) / (P - 1.0)
        xc = [alpha * c[j] + beta * X[corner][j] for j in range(2)]
        return xc, ev(xc)

    def update(i, x, val):
        d2 = xmcd = 0.0
        for j in range(2):
            delta = x[j] - X[i][j]
            xmc = X[i][j] - c[j]
            d2 += delta * delta
            xmcd += xmc * delta
        d = math.sqrt(d2)
        S2[0] += (2.0 / P) * xmcd + ((P - 1.0) / P) * (d * d / P)
        for j in range(2):
            c[j] -= (1.0 / P) * X[i][j]
            c[j] += (1.0 / P) * x[j]
            X[i][j] = x[j]
        Y[i] = val

    center()
    full_size()
    best, fval, it, converged, failed = list(x0), Y[0], 0, False, False
    while True:
        it += 1
        hi, s_hi, lo = 0, 1, 0
        dhi = dlo = Y[0]
        ds_hi = Y[1]
        for k in range(1, P):
            v = Y[k]
            if v < dlo:
                dlo, lo = v, k
            elif v > dhi:
                ds_hi, s_hi, dhi, hi = dhi, hi, v, k
            elif v > ds_hi:
                ds_hi, s_hi = v, k
        xc, val = move(-1.0, hi)
        if math.isfinite(val) and val < Y[lo]:
            xc2, val2 = move(-2.0, hi)
            if math.isfinite(val2) and val2 < Y[lo]:
                update(hi, xc2, val2)
            else:
                update(hi, xc, val)
        elif not math.isfinite(val) or val > Y[s_hi]:
            if math.isfinite(val) and val <= Y[hi]:
                update(hi, xc, val)
            xc2, val2 = move(0.5, hi)
            if math.isfinite(val2) and val2 <= Y[hi]:
                update(hi, xc2, val2)
            else:
                for k in range(P):
                    if k == lo:
                        continue
                    for j in range(2):
                        X[k][j] = 0.5 * (X[k][j] + X[lo][j])
                    Y[k] = ev(X[k])
                    if not math.isfinite(Y[k]):
                        failed = True
                center()
                full_size()
        else:
            update(hi, xc, val)
        if failed:
            break
        lo = min(range(P), key=lambda k: Y[k])
        best, fval = list(X[lo]), Y[lo]
        size = math.sqrt(S2[0]) if S2[0] > 0 else full_size()
        converged = size < size_eps
        if converged or it >= max_iterations:
            break
    return {"x": best, "fval": fval, "iterations": it, "evaluations": evals[0], "converged": converged or failed}
