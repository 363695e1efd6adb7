"""Pins the oracle's NLML / LLT objective (oracle/gple_oracle_nlml.hpp; formula of test/gpr.cpp:470-532) by an independent
numpy evaluation and by central finite differences of every gradient component."""
import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn
from oracle import oracle as orc

THETA_C = np.array([1.3, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])


def numpy_nlml(K, y):
    sign, logdet = np.linalg.slogdet(K)
    assert sign > 0
    return 0.5 * y @ np.linalg.solve(K, y) + 0.5 * logdet


def test_real_value_matches_numpy_and_gradient_matches_finite_differences():
    X, y = syn.training_set(41, 0, 90)
    th = syn.theta_real() * np.array([1.1, 1.0, 1.0, 3.0])
    k = orc.TrainingKernel(th, X, y, deriv=True)
    v, g = k.nlml(grad=True)
    assert v == pytest.approx(numpy_nlml(k.K, k.label), rel=1e-11)
    for p in range(4):
        h = 1e-6 * th[p]
        tp, tm = th.copy(), th.copy()
        tp[p] += h
        tm[p] -= h
        fd = (orc.TrainingKernel(tp, X, y).nlml() - orc.TrainingKernel(tm, X, y).nlml()) / (2 * h)
        assert g[p] == pytest.approx(fd, rel=2e-6, abs=1e-6 * np.abs(g).max()), p


def test_complex_value_matches_numpy_composite_and_gradient_matches_finite_differences():
    X, y = syn.training_set(42, 1, 60)
    k = orc.TrainingComplexKernel(THETA_C, X, y, deriv=True)
    v, g = k.nlml(grad=True)
    K, Kt = k.K, k.Kt
    Krr, Kii, Kri = 0.5 * (K + Kt.real), 0.5 * (K - Kt.real), 0.5 * Kt.imag
    Cov = np.block([[Krr, Kri], [Kri, Kii]])
    yy = np.concatenate([k.label.real, k.label.imag])
    assert v == pytest.approx(numpy_nlml(Cov, yy), rel=1e-10)
    for p in range(8):
        h = 1e-6 * THETA_C[p]
        tp, tm = THETA_C.copy(), THETA_C.copy()
        tp[p] += h
        tm[p] -= h
        fd = (orc.TrainingComplexKernel(tp, X, y).nlml() - orc.TrainingComplexKernel(tm, X, y).nlml()) / (2 * h)
        assert g[p] == pytest.approx(fd, rel=5e-6, abs=1e-6 * np.abs(g).max()), p


def test_not_positive_definite_is_nan():
    X, y = syn.training_set(43, 0, 20)
    X[1] = X[0]  # two identical points and no noise: singular covariance
    th = syn.theta_real() * np.array([1.0, 1.0, 1.0, 0.0])
    assert np.isnan(orc.TrainingKernel(th, X, y).nlml())
