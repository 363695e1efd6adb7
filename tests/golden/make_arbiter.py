"""Generates tests/golden/gple_arbiter_v1.npz: a 40-digit (mpmath) evaluation of the reference's FORMULAS for the
quantities whose double-precision evaluation is ill-conditioned -- the predictive variance, the cutoff prediction
inside the cubic band, and the whole complex chain (A = K^-1 conj(K~), P, Q, v, LOOCV error, prediction, variance) --
on small seeded inputs (SURVEY.md section 8c iv).

It is the arbiter between the three double-precision evaluations (the reference itself = oracle/_ref, the oracle
restatement, the CUDA path): tests/test_arbiter.py measures each one's distance to these values.  Formulas:
gple/kernel.cpp:38-85, 217-242, 279-335, 481-522; kernel.h:301-332; complex_kernel.cpp:134-164, 262-286, 594-646.
Run from the repo root (a few minutes):  python tests/golden/make_arbiter.py
"""
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gaussian_process_liouville_equation_b200 import synthetic as syn  # noqa: E402

mp.mp.dps = 40
THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])


def M(a):
    return mp.matrix(a.tolist())


def gauss(XL, XR, lx, lp):
    G = mp.matrix(len(XL), len(XR))
    for i in range(len(XL)):
        for j in range(len(XR)):
            dx = (mp.mpf(float(XL[i, 0])) - mp.mpf(float(XR[j, 0]))) / lx
            dp = (mp.mpf(float(XL[i, 1])) - mp.mpf(float(XR[j, 1]))) / lp
            G[i, j] = mp.exp(-(dx * dx + dp * dp) / 2)
    return G


def coincide(XL, XR):
    D = mp.matrix(len(XL), len(XR))
    for i in range(len(XL)):
        for j in range(len(XR)):
            D[i, j] = 1 if (XL[i] == XR[j]).all() else 0
    return D


def gate(f_abs2, var):
    """kernel.h:301-332"""
    if f_abs2 >= 4 * var:
        return mp.mpf(1)
    if f_abs2 <= var:
        return mp.mpf(0)
    a = mp.sqrt(f_abs2) / mp.sqrt(var)
    return (3 * 2 - 2 * a - 1) * (a - 1) ** 2 / (2 - 1) ** 3


def to_np(v, cplx=False):
    if isinstance(v, mp.matrix):
        if cplx:
            return np.array([[complex(v[i, j]) for j in range(v.cols)] for i in range(v.rows)]).squeeze()
        return np.array([[float(v[i, j]) for j in range(v.cols)] for i in range(v.rows)]).squeeze()
    return complex(v) if cplx else float(v)


def real_chain(th, X, y, Xq):
    sf, lx, lp, sn = [mp.mpf(float(t)) for t in th]
    n = len(X)
    K = sf * sf * (gauss(X, X, lx, lp) + sn * sn * mp.eye(n))
    s = mp.mpf(10) / max(abs(mp.mpf(float(t))) for t in y.real)
    lab = M(y.real.reshape(-1, 1)) * s
    Kinv = mp.inverse(K)
    v = Kinv * lab
    err = sum((v[i] / Kinv[i, i]) ** 2 for i in range(n))
    pop = 2 * mp.pi * sf * sf * lx * lp * sum(v) / s
    K1 = (sf * sf * mp.sqrt(lx * lp)) ** 2 * gauss(X, X, mp.sqrt(2) * lx, mp.sqrt(2) * lp)
    purity = 2 * mp.pi * mp.pi * (v.T * K1 * v)[0] / (s * s)
    Ks = sf * sf * (gauss(Xq, X, lx, lp) + sn * sn * coincide(Xq, X))
    f = Ks * v
    prior = sf * sf * (1 + sn * sn)
    T = Ks * Kinv
    var = mp.matrix(len(Xq), 1)
    cut = mp.matrix(len(Xq), 1)
    for i in range(len(Xq)):
        var[i] = prior - sum(T[i, j] * Ks[i, j] for j in range(n))
        cut[i] = f[i] * gate(f[i] ** 2, var[i]) / s
    return dict(r_error=to_np(err), r_population=to_np(pop), r_purity=to_np(purity), r_v=to_np(v), r_pred=to_np(f), r_var=to_np(var), r_cutoff=to_np(cut), r_kinv_diag=np.array([float(Kinv[i, i]) for i in range(n)]))


def complex_chain(th, X, y, Xq):
    sg, sr, lrx, lrp, si, lix, lip, sn = [mp.mpf(float(t)) for t in th]
    n = len(X)
    lcx, lcp = mp.sqrt((lrx ** 2 + lix ** 2) / 2), mp.sqrt((lrp ** 2 + lip ** 2) / 2)
    sc = mp.sqrt(sr * si * (2 * lrx * lix / (lrx ** 2 + lix ** 2)) * (2 * lrp * lip / (lrp ** 2 + lip ** 2)))

    def blocks(XL, XR, delta):
        KR, KI, KC = sr * sr * gauss(XL, XR, lrx, lrp), si * si * gauss(XL, XR, lix, lip), sc * sc * gauss(XL, XR, lcx, lcp)
        return sg * sg * (KR + KI + sn * sn * delta), sg * sg * (KR - KI + 2j * KC)

    K, Kt = blocks(X, X, mp.eye(n))
    s = mp.mpf(10) / max(abs(mp.mpc(complex(t))) for t in y)
    lab = mp.matrix([[mp.mpc(complex(t))] for t in y]) * s
    A = mp.inverse(K) * Kt.apply(mp.conj)
    P = mp.inverse(K - Kt * A)
    Q = -A * P
    v = P * lab + (Q * lab).apply(mp.conj)
    err = mp.mpf(0)
    for i in range(n):
        p, q = P[i, i], Q[i, i]
        d = (p * v[i] - mp.conj(q * v[i])) / (mp.re(p) ** 2 - abs(q) ** 2)
        err += abs(d) ** 2
    Ks, Kts = blocks(Xq, X, coincide(Xq, X))
    f = Ks * v + Kts * v.apply(mp.conj)
    prior = sg * sg * (sr * sr + si * si + sn * sn)
    m = len(Xq)
    var, cut = mp.matrix(m, 1), mp.matrix(m, 1)
    Pc, Qc = P.apply(mp.conj), Q.apply(mp.conj)
    for i in range(m):
        k, kt = Ks[i, :], Kts[i, :]
        kT, ktH = k.T, kt.T.apply(mp.conj)
        val = prior - (k * P * kT)[0] - (kt * Pc * ktH)[0] - (kt * Q * kT)[0] - (k * Qc * ktH)[0]
        var[i] = mp.re(val)
        cut[i] = f[i] * gate(abs(f[i]) ** 2, var[i]) / s
    return dict(c_error=to_np(err), c_v=to_np(v, True), c_P=to_np(P, True), c_Q=to_np(Q, True), c_pred=to_np(f, True), c_var=to_np(var), c_cutoff=to_np(cut, True))


def main():
    n, m, centre = 96, 160, (-0.8, syn.P0)
    th = np.array([1.0, 0.9 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 1e-2])
    out = dict(theta_r=th, theta_c=THETA_C)
    X0, y0 = syn.training_set(71, 0, n, centre)
    X1, y1 = syn.training_set(71, 1, n, centre)
    # queries around the rim of the point cloud; the (double-precision) oracle is only used to PICK inputs whose gate lies
    # inside the cubic band (about a third of the set), wide open (a third) and closed (a third)
    from oracle import oracle as orc  # noqa: E402

    def pick(theta, X, y, cplx, stream):
        rng = syn.rng(71, int(cplx), stream)
        cand = np.column_stack([centre[0] + syn.SIGMA_X * rng.uniform(-5.0, 5.0, 6000), centre[1] + syn.SIGMA_P * rng.uniform(-5.0, 5.0, 6000)])
        k = (orc.TrainingComplexKernel if cplx else orc.TrainingKernel)(theta, X, y, True, True, False)
        r = k.predict(cand)
        f2 = np.abs(r["pred"]) ** 2
        band = np.flatnonzero((f2 > 1.05 * r["var"]) & (f2 < 3.8 * r["var"]))[: m // 3]
        one = np.flatnonzero(f2 > 4.5 * r["var"])[: m // 3]
        zero = np.flatnonzero(f2 < 0.9 * r["var"])[: m - len(band) - len(one)]
        assert len(band) >= 20, len(band)
        sel = cand[np.concatenate([band, one, zero])]
        sel[5] = X[9]  # one exact coincidence (delta_kernel, kernel.cpp:16-29)
        return sel

    Xq, Xqc = pick(th, X0, y0, False, 3), pick(THETA_C, X1, y1, True, 4)
    out.update(X0=X0, y0=y0, X1=X1, y1=y1, Xq=Xq, Xqc=Xqc)
    out.update(real_chain(th, X0, y0, Xq))
    print("real chain done", flush=True)
    out.update(complex_chain(THETA_C, X1, y1, Xqc))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gple_arbiter_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
