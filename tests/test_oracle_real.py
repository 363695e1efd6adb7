"""Pins the CPU oracle's real-kernel chain (oracle/gple_oracle.hpp) WITHOUT the reference's help:
the reference has no golden vectors, so the oracle is checked against an
independent numpy restatement, brute-force leave-one-out, analytic known answers, finite differences
and a 50-digit mpmath twin.  Reference lines: gple/kernel.cpp:8-544, gple/kernel.h:285-332.
"""
import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn


def np_kernel(XL, XR, th, same):
    d = (XL[:, None, :] - XR[None, :, :]) / th[1:3]
    G = np.exp(-0.5 * (d * d).sum(-1))
    delta = np.eye(len(XL)) if same else (XL[:, None, :] == XR[None, :, :]).all(-1).astype(float)
    return th[0] ** 2 * (G + th[3] ** 2 * delta), d


@pytest.fixture(scope="module")
def small(oracle):
    X, y = syn.training_set(1, 0, 96)
    th = syn.theta_real()
    return X, y, th, oracle.TrainingKernel(th, X, y, True, True, True)


def test_kernel_and_derivatives_match_numpy(oracle):
    X, _ = syn.training_set(1, 0, 50)
    Xq, _ = syn.training_set(2, 0, 31)
    Xq[3] = X[7]  # exact coincidence exercises delta_kernel (kernel.cpp:16-29)
    th = np.array([1.3, 0.6, 0.9, 0.05])
    for XL, XR, same in ((X, X, True), (Xq, X, False)):
        K, dK = oracle.kernel_real(XL, XR, th, same, True)
        Kn, d = np_kernel(XL, XR, th, same)
        assert np.abs(K - Kn).max() <= 4e-16 * np.abs(Kn).max()
        G = Kn - (th[0] * th[3]) ** 2 * np.eye(len(XL)) if same else Kn
        assert np.abs(dK[0] - 2 * Kn / th[0]).max() <= 1e-15 * np.abs(Kn).max()
        for a in range(2):
            ref = G * d[..., a] ** 2 / th[1 + a]
            assert np.abs(dK[1 + a] - ref).max() <= 1e-15 * np.abs(ref).max()
        noise = 2 * th[0] ** 2 * th[3] * np.eye(len(XL)) if same else np.zeros_like(Kn)
        assert np.array_equal(dK[3], noise)
    assert K[3, 7] == pytest.approx(th[0] ** 2 * (1 + th[3] ** 2), rel=1e-15)


def test_inverse_and_solve(small):
    X, y, th, k = small
    Kn, _ = np_kernel(X, X, th, True)
    assert k.rescale == pytest.approx(10.0 / np.abs(y.real).max(), rel=1e-15)
    assert np.allclose(k.label, y.real * k.rescale, rtol=1e-15)
    inv = np.linalg.inv(Kn)
    assert np.abs(k.inverse - inv).max() <= 1e-9 * np.abs(inv).max()
    assert np.abs(k.inverse @ Kn - np.eye(len(X))).max() < 1e-9
    v = np.linalg.solve(Kn, k.label)
    assert np.abs(k.v - v).max() <= 1e-9 * np.abs(v).max()
    assert k.magnitude == pytest.approx(np.sqrt(abs(k.label @ v) / len(X)), rel=1e-9)


def test_loocv_error_is_brute_force_leave_one_out(oracle):
    X, y = syn.training_set(3, 0, 40)
    th = syn.theta_real(1.5)
    k = oracle.TrainingKernel(th, X, y, True, False, False)
    Kn, _ = np_kernel(X, X, th, True)
    lab = y.real * k.rescale
    tot = 0.0
    for i in range(len(X)):
        m = np.arange(len(X)) != i
        pred = Kn[i, m] @ np.linalg.solve(Kn[np.ix_(m, m)], lab[m])
        tot += (lab[i] - pred) ** 2
    assert k.error == pytest.approx(tot, rel=1e-7)


def test_analytic_known_answers(oracle):
    """Dense Gaussian label set => population -> 0.6, <r>/population -> centre, purity -> 0.36
    (kernel.cpp:286-335; a pure Gaussian Wigner function has purity population^2)."""
    X, y = syn.training_set(4, 0, 400)
    k = oracle.TrainingKernel(syn.theta_real(), X, y, True, True, False)
    assert k.population == pytest.approx(0.6, rel=2e-3)
    assert k.first_order / k.population == pytest.approx(np.array([0.0, syn.P0]), abs=2e-2)
    assert k.purity == pytest.approx(0.36, rel=5e-3)


def test_gradients_against_finite_differences(oracle):
    X, y = syn.training_set(5, 0, 60)
    th = np.array([1.0, 0.9 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 2e-2])
    k = oracle.TrainingKernel(th, X, y, True, True, True)
    for p in range(4):
        h = 1e-5 * th[p]
        tp, tm = th.copy(), th.copy()
        tp[p] += h
        tm[p] -= h
        kp, km = oracle.TrainingKernel(tp, X, y, True, True, False), oracle.TrainingKernel(tm, X, y, True, True, False)
        assert k.derror[p] == pytest.approx((kp.error - km.error) / (2 * h), rel=3e-5, abs=1e-4 * k.error)
        if p == 0:
            # quirk q4 (kernel.cpp:415,451): magnitude derivatives are hard-wired to zero
            assert k.dpopulation[0] == 0.0 and k.dpurity[0] == 0.0
            continue
        assert k.dpopulation[p] == pytest.approx((kp.population - km.population) / (2 * h), rel=3e-5)
        assert k.dpurity[p] == pytest.approx((kp.purity - km.purity) / (2 * h), rel=3e-5)
        assert np.allclose(k.dv(p), (kp.v - km.v) / (2 * h), rtol=1e-4, atol=1e-6 * np.abs(k.v).max())


def test_prediction_variance_cutoff(small, oracle):
    X, y, th, k = small
    Xq, yq = syn.extra_points(1, 0, X, 300)
    Xq[:40] += np.array([4.0, 0.0])  # far-away points: prediction ~ 0, gate closes
    Xq[40:80] += np.array([1.7, 0.0])  # tail points: transition band of the cubic gate
    yq = syn.labels(0, Xq, (0.0, syn.P0)).real
    r = k.predict(Xq, yq, False)
    Ks, _ = np_kernel(Xq, X, th, False)
    inv, v = k.inverse, k.v
    pred = Ks @ v
    var = th[0] ** 2 * (1 + th[3] ** 2) - np.einsum("mi,ij,mj->m", Ks, inv, Ks)
    assert np.abs(r["pred"] - pred).max() <= 1e-12 * np.abs(pred).max()
    assert np.abs(r["var"] - var).max() <= 1e-9 * th[0] ** 2  # k K^-1 k^T cancels against k**: ~1e-11 absolute noise between summation orders
    a = np.abs(pred) / np.sqrt(np.abs(var))
    gate = np.where(pred**2 >= 4 * var, 1.0, np.where(pred**2 <= var, 0.0, (5 - 2 * a) * (a - 1) ** 2))
    assert (gate == 0).any() and (gate == 1).any() and ((gate > 0) & (gate < 1)).any()
    assert np.abs(r["cutoff"] - pred * gate / k.rescale).max() <= 1e-7 * np.abs(pred).max() / k.rescale
    assert r["error"] == pytest.approx(((pred - yq * k.rescale) ** 2).sum(), rel=1e-10)


def test_validation_gradient(oracle):
    """With every gate open (cutoff == 1) the quirky gradient of kernel.cpp:524-541 is the true one."""
    X, y = syn.training_set(6, 0, 60)
    th = np.array([1.0, 0.9 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 2e-2])
    Xq, yq = syn.extra_points(6, 0, X, 80)
    Xq = X[np.arange(80) % 60] + 0.05 * (Xq - X[np.arange(80) % 60])
    yq = syn.labels(0, Xq, (0.0, syn.P0)).real
    k = oracle.TrainingKernel(th, X, y, True, False, True)
    r = k.predict(Xq, yq, True)
    keep = np.abs(r["pred"]) > 3.0 * np.sqrt(np.abs(r["var"]))  # gate comfortably open, also at theta +- h
    assert keep.sum() > 40
    Xq, yq = Xq[keep], yq[keep]
    r = k.predict(Xq, yq, True)
    assert np.allclose(r["cutoff"] * k.rescale, r["pred"], rtol=1e-14)
    for p in range(4):
        h = 1e-5 * th[p]
        tp, tm = th.copy(), th.copy()
        tp[p] += h
        tm[p] -= h
        ep = oracle.TrainingKernel(tp, X, y, True, False, False).predict(Xq, yq)["error"]
        em = oracle.TrainingKernel(tm, X, y, True, False, False).predict(Xq, yq)["error"]
        assert r["derror"][p] == pytest.approx((ep - em) / (2 * h), rel=5e-5, abs=1e-10)
    val, g = oracle.loose_function(th, X, y, Xq, yq.astype(complex), grad=True)
    assert val == pytest.approx(k.error + r["error"], rel=1e-14)
    assert np.allclose(g, k.derror + r["derror"], rtol=1e-14)


def test_mpmath_twin_bounds_oracle_error(oracle):
    """50-digit twin of K, K^-1, v, LOOCV error, population at N = 24: the FP64 oracle's own error."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    X, y = syn.training_set(7, 0, 24)
    th = syn.theta_real()
    k = oracle.TrainingKernel(th, X, y, True, True, False)
    n = len(X)
    Xm = [[mp.mpf(float(a)) for a in row] for row in X]
    t = [mp.mpf(float(a)) for a in th]
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            e = sum(((Xm[i][a] - Xm[j][a]) / t[1 + a]) ** 2 for a in range(2))
            K[i, j] = t[0] ** 2 * (mp.exp(-e / 2) + (t[3] ** 2 if i == j else 0))
    inv = K ** -1
    resc = mp.mpf(10) / max(abs(mp.mpf(float(a.real))) for a in y)
    lab = mp.matrix([mp.mpf(float(a.real)) * resc for a in y])
    v = inv * lab
    err = sum((v[i] / inv[i, i]) ** 2 for i in range(n))
    pop = 2 * mp.pi * t[0] ** 2 * t[1] * t[2] * sum(v) / resc
    inv_np = np.array([[float(inv[i, j]) for j in range(n)] for i in range(n)])
    assert np.abs(k.inverse - inv_np).max() <= 1e-10 * np.abs(inv_np).max()
    assert k.error == pytest.approx(float(err), rel=1e-9)
    assert k.population == pytest.approx(float(pop), rel=1e-10)


def test_best_effort_cpu_matches_the_oracle(oracle):
    """oracle/best_effort.py (the BLAS / LAPACK CPU baseline that bench.py reports next to the reference-shaped port) computes
    the same quantities as the oracle, for the real and the complex element."""
    from oracle import best_effort as be

    orc = oracle
    X, y = syn.training_set(5, 0, 200)
    th = np.array([1.0, 0.9 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 2e-2])
    o, m = orc.TrainingKernel(th, X, y), be.RealModel(th, X, y)
    Xq, _ = syn.extra_points(5, 0, X, 300)
    p = o.predict(Xq)
    f, var = m.predict(Xq)
    assert m.error == pytest.approx(o.error, rel=1e-9) and m.population == pytest.approx(o.population, rel=1e-11)
    assert np.abs(f - p["pred"]).max() <= 1e-11 * np.abs(p["pred"]).max() and np.abs(var - p["var"]).max() <= 1e-9
    thc = np.array([1.3, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])
    X, y = syn.training_set(4, 1, 150)
    o, m = orc.TrainingComplexKernel(thc, X, y), be.ComplexModel(thc, X, y)
    Xq, _ = syn.extra_points(4, 1, X, 200)
    p = o.predict(Xq)
    f, var = m.predict(Xq)
    assert m.error == pytest.approx(o.error, rel=1e-8)
    assert np.abs(f - p["pred"]).max() <= 1e-9 * np.abs(p["pred"]).max() and np.abs(var - p["var"]).max() <= 1e-9 * m.prior
