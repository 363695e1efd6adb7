"""Pins the CPU oracle's complex (widely-linear) kernel chain (oracle/gple_oracle_complex.hpp) against
an independent numpy restatement: augmented-matrix inverse, brute-force leave-one-out, the equivalent
real composite 2N x 2N Gaussian process (the form the CUDA path uses), and finite differences.
Reference lines: gple/complex_kernel.cpp:20-670.
"""
import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn


def gauss(XL, XR, mag, l):
    d = (XL[:, None, :] - XR[None, :, :]) / np.asarray(l)
    return mag**2 * np.exp(-0.5 * (d * d).sum(-1))


def sub_params(th):
    sR, lR, sI, lI = th[1], th[2:4], th[4], th[5:7]
    ss = lR**2 + lI**2
    sC = np.sqrt(sR * sI * np.prod(2 * lR * lI / ss))
    lC = np.sqrt(ss / 2)
    return (sR, lR), (sI, lI), (sC, lC)


def np_complex_kernel(XL, XR, th, same):
    (sR, lR), (sI, lI), (sC, lC) = sub_params(th)
    KR, KI, KC = gauss(XL, XR, sR, lR), gauss(XL, XR, sI, lI), gauss(XL, XR, sC, lC)
    delta = np.eye(len(XL)) if same else np.zeros((len(XL), len(XR)))
    K = th[0] ** 2 * (KR + KI + th[7] ** 2 * delta)
    Kt = th[0] ** 2 * (KR - KI + 2j * KC)
    return K, Kt, (KR, KI, KC)


THETA = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])


@pytest.fixture(scope="module")
def trained(oracle):
    X, y = syn.training_set(1, 1, 48)
    return X, y, oracle.TrainingComplexKernel(THETA, X, y, True, True, True)


def test_kernels_match_numpy(oracle):
    X, _ = syn.training_set(1, 1, 30)
    Xq, _ = syn.training_set(2, 1, 17)
    for XL, XR, same in ((X, X, True), (Xq, X, False)):
        K, Kt = oracle.kernel_complex(XL, XR, THETA, same, False)
        Kn, Ktn, _ = np_complex_kernel(XL, XR, THETA, same)
        assert np.abs(K - Kn).max() <= 1e-15 * np.abs(Kn).max()
        assert np.abs(Kt - Ktn).max() <= 1e-15 * np.abs(Ktn).max()


def test_kernel_derivatives_finite_difference(oracle):
    """d K / d theta and d Kt / d theta (complex_kernel.cpp:20-132) for the sub-kernel parameters.
    Magnitude (index 0) is exact; quirk q2 (no sigma^2 factor) is invisible because sigma == 1."""
    X, _ = syn.training_set(1, 1, 12)
    _, _, dK, dKt = oracle.kernel_complex(X, X, THETA, True, True)
    for p in range(8):
        h = 1e-6 * THETA[p]
        tp, tm = THETA.copy(), THETA.copy()
        tp[p] += h
        tm[p] -= h
        Kp, Ktp = oracle.kernel_complex(X, X, tp, True, False)
        Km, Ktm = oracle.kernel_complex(X, X, tm, True, False)
        assert np.abs(dK[p] - (Kp - Km) / (2 * h)).max() < 2e-8
        assert np.abs(dKt[p] - (Ktp - Ktm) / (2 * h)).max() < 2e-8


def test_augmented_inverse_blocks(trained):
    X, y, k = trained
    n = len(X)
    K, Kt, _ = np_complex_kernel(X, X, THETA, True)
    aug = np.block([[K.astype(complex), Kt], [Kt.conj(), K.astype(complex)]])
    inv = np.linalg.inv(aug)
    assert np.abs(k.P - inv[:n, :n]).max() <= 1e-9 * np.abs(inv).max()
    assert np.abs(k.Q - inv[n:, :n]).max() <= 1e-9 * np.abs(inv).max()
    lab = y * k.rescale
    v = (inv @ np.concatenate([lab, lab.conj()]))[:n]
    assert np.abs(k.v - v).max() <= 1e-9 * np.abs(v).max()
    assert k.rescale == pytest.approx(10.0 / np.abs(y).max(), rel=1e-15)
    assert k.magnitude == pytest.approx(np.sqrt(abs((lab.conj() @ v).real) / n), rel=1e-9)


def composite(X, th):
    """Real composite covariance of [Re f; Im f] (2N x 2N): the formulation used by the CUDA path."""
    K, Kt, (KR, KI, KC) = np_complex_kernel(X, X, th, True)
    n = len(X)
    s2, n2 = th[0] ** 2, th[7] ** 2
    return np.block([[s2 * (KR + 0.5 * n2 * np.eye(n)), s2 * KC], [s2 * KC, s2 * (KI + 0.5 * n2 * np.eye(n))]])


def test_equivalence_with_real_composite_gp(trained, oracle):
    """P = (Mrr + Mii + i (Mir - Mri)) / 4, Q = (Mrr - Mii - i (Mir + Mri)) / 4, v = (wr + i wi) / 2 with
    M = C^-1, w = M [Re y; Im y]; prediction and variance are those of the 2-output real GP."""
    X, y, k = trained
    n = len(X)
    Cc = composite(X, THETA)
    M = np.linalg.inv(Cc)
    Mrr, Mri, Mir, Mii = M[:n, :n], M[:n, n:], M[n:, :n], M[n:, n:]
    P = 0.25 * (Mrr + Mii + 1j * (Mir - Mri))
    Q = 0.25 * (Mrr - Mii - 1j * (Mir + Mri))
    sc = np.abs(k.P).max()
    assert np.abs(k.P - P).max() <= 1e-9 * sc
    assert np.abs(k.Q - Q).max() <= 1e-9 * sc
    lab = y * k.rescale
    w = M @ np.concatenate([lab.real, lab.imag])
    assert np.abs(k.v - 0.5 * (w[:n] + 1j * w[n:])).max() <= 1e-9 * np.abs(k.v).max()
    # prediction / variance
    Xq, _ = syn.extra_points(1, 1, X, 64)
    r = k.predict(Xq)
    (sR, lR), (sI, lI), (sC, lC) = sub_params(THETA)
    s2 = THETA[0] ** 2
    cr = s2 * np.hstack([gauss(Xq, X, sR, lR), gauss(Xq, X, sC, lC)])
    ci = s2 * np.hstack([gauss(Xq, X, sC, lC), gauss(Xq, X, sI, lI)])
    pred = cr @ w + 1j * (ci @ w)
    prior = s2 * (sR**2 + sI**2 + THETA[7] ** 2)
    var = prior - np.einsum("mi,ij,mj->m", cr, M, cr) - np.einsum("mi,ij,mj->m", ci, M, ci)
    assert np.abs(r["pred"] - pred).max() <= 1e-10 * np.abs(pred).max()
    assert np.abs(r["var"] - var).max() <= 1e-8 * prior


def test_complex_loocv_is_brute_force(oracle):
    X, y = syn.training_set(3, 1, 28)
    k = oracle.TrainingComplexKernel(THETA, X, y, True, False, False)
    Cc = composite(X, THETA)
    n = len(X)
    lab = y * k.rescale
    yy = np.concatenate([lab.real, lab.imag])
    tot = 0.0
    for i in range(n):
        out = np.array([i, n + i])
        keep = np.setdiff1d(np.arange(2 * n), out)
        pred = Cc[np.ix_(out, keep)] @ np.linalg.solve(Cc[np.ix_(keep, keep)], yy[keep])
        tot += ((yy[out] - pred) ** 2).sum()
    assert k.error == pytest.approx(tot, rel=1e-6)


def test_error_gradient_finite_difference(oracle):
    X, y = syn.training_set(4, 1, 36)
    k = oracle.TrainingComplexKernel(THETA, X, y, True, True, True)
    for p in range(8):
        h = 1e-5 * THETA[p]
        tp, tm = THETA.copy(), THETA.copy()
        tp[p] += h
        tm[p] -= h
        kp = oracle.TrainingComplexKernel(tp, X, y, True, True, False)
        km = oracle.TrainingComplexKernel(tm, X, y, True, True, False)
        assert k.derror[p] == pytest.approx((kp.error - km.error) / (2 * h), rel=1e-4, abs=1e-4 * k.error)
        assert np.allclose(k.dv(p), (kp.v - km.v) / (2 * h), rtol=1e-3, atol=1e-5 * np.abs(k.v).max() / THETA[p])
        if p in (1, 4):
            # sub-magnitude derivatives of the purity are free of quirk q10 (wrong sub-derivative index
            # for the length parameters, complex_kernel.cpp:534-564) and must match finite differences
            assert k.dpurity[p] == pytest.approx((kp.purity - km.purity) / (2 * h), rel=1e-4)


def test_validation_error_and_gradient(oracle):
    X, y = syn.training_set(5, 1, 36)
    Xq, yq = syn.extra_points(5, 1, X, 60)
    Xq = X[np.arange(60) % 36] + 0.05 * (Xq - X[np.arange(60) % 36])
    yq = syn.labels(1, Xq, (0.0, syn.P0))
    k = oracle.TrainingComplexKernel(THETA, X, y, True, False, True)
    r = k.predict(Xq, yq, True)
    keep = np.abs(r["pred"]) > 3.0 * np.sqrt(np.abs(r["var"]))
    assert keep.sum() > 30
    Xq, yq = Xq[keep], yq[keep]
    r = k.predict(Xq, yq, True)
    assert r["error"] == pytest.approx((np.abs(r["pred"] - yq * k.rescale) ** 2).sum(), rel=1e-12)
    for p in range(8):
        h = 1e-5 * THETA[p]
        tp, tm = THETA.copy(), THETA.copy()
        tp[p] += h
        tm[p] -= h
        ep = oracle.TrainingComplexKernel(tp, X, y, True, False, False).predict(Xq, yq)["error"]
        em = oracle.TrainingComplexKernel(tm, X, y, True, False, False).predict(Xq, yq)["error"]
        assert r["derror"][p] == pytest.approx((ep - em) / (2 * h), rel=2e-4, abs=1e-5 * r["error"])


def test_purity_formula_is_integral_of_squared_prediction(oracle):
    """complex_kernel.cpp:357-377: purity = (2 pi hbar) * integral |f(r)|^2 dGamma of the (uncut, unscaled)
    GP mean f.  Checked by brute-force quadrature of the oracle's own prediction on a grid."""
    X, y = syn.training_set(6, 1, 120)
    k = oracle.TrainingComplexKernel(THETA, X, y, True, True, False)
    gx = np.linspace(-6.0, 6.0, 141)
    gp = np.linspace(syn.P0 - 6.0, syn.P0 + 6.0, 141)
    G = np.stack(np.meshgrid(gx, gp, indexing="ij"), -1).reshape(-1, 2)
    f = k.predict(G)["pred"] / k.rescale
    quad = 2.0 * np.pi * (np.abs(f) ** 2).sum() * (gx[1] - gx[0]) * (gp[1] - gp[0])
    assert k.purity == pytest.approx(quad, rel=1e-2)
