"""Pins the CPU oracle's PES / evolve / observables restatement (oracle/gple_oracle_dynamics.hpp) with
identities the reference's own maths provides: eigen-decomposition of the diabatic
Hamiltonian, Hellmann-Feynman forces, the sign convention d_jk = F_jk / (E_j - E_k), free-streaming of
the Wigner function where forces and couplings vanish, and direct numpy sums for the observables.
Reference lines: gple/pes.cpp:42-189, gple/evolve.cpp:53-443, gple/predict.cpp:43-244, gple/mc.cpp:30-50.
"""
import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn


def diabatic(model, x):
    s = np.sign(x)
    if model == 0:
        v00 = s * 0.01 * (1 - np.exp(-s * 1.6 * x))
        return np.array([[v00, 0.005 * np.exp(-x * x)], [0.005 * np.exp(-x * x), -v00]])
    if model == 1:
        c = 0.015 * np.exp(-0.06 * x * x)
        return np.array([[0.0, c], [c, 0.05 - 0.10 * np.exp(-0.28 * x * x)]])
    c = 0.10 * (1 - s * (np.exp(-s * 0.90 * x) - 1))
    return np.array([[6e-4, c], [c, -6e-4]])


@pytest.mark.parametrize("model", [0, 1, 2])
def test_pes_identities(oracle, model):
    x = np.array([-7.3, -2.1, -0.4, 0.3, 1.9, 6.5])
    E, F, D = oracle.pes(model, x)
    h = 1e-5
    for i, xi in enumerate(x):
        w, U = np.linalg.eigh(diabatic(model, xi))
        assert E[i] == pytest.approx(w, abs=1e-15)
        wp = np.linalg.eigvalsh(diabatic(model, xi + h))
        wm = np.linalg.eigvalsh(diabatic(model, xi - h))
        assert F[i, [0, 2]] == pytest.approx(-(wp - wm) / (2 * h), rel=1e-6, abs=1e-11)
        dV = (diabatic(model, xi + h) - diabatic(model, xi - h)) / (2 * h)
        f10 = -(U[:, 1] @ dV @ U[:, 0])  # F = -dH/dx in the adiabatic basis, up to eigenvector signs
        assert abs(F[i, 1]) == pytest.approx(abs(f10), rel=1e-6, abs=1e-12)
        assert D[i] == pytest.approx(F[i, 1] / (E[i, 1] - E[i, 0]), rel=1e-14)


def test_free_streaming_conserves_density(oracle):
    """SAC at x ~ -10: forces and couplings are ~ e^{-16}, so the backward-propagated density at the
    moved point is the density at the old point (branch weights 1/4 + 1/2 + 1/4 = 1)."""
    g = syn.rng(9, 0)
    n = 64
    r = np.stack([syn.X0 + syn.SIGMA_X * g.standard_normal(n), syn.P0 + syn.SIGMA_P * g.standard_normal(n)], 1)
    an = np.array([syn.X0, syn.P0, syn.SIGMA_X, syn.SIGMA_P, 1.0, 0.0, 0.0, 0.0])
    rho = oracle.initial_distribution(an, r, 0, 0)
    assert rho.real == pytest.approx(syn.wigner_gaussian(r, (syn.X0, syn.P0)), rel=1e-14)
    pts = syn.points_aos(r, rho)
    out, _, _ = oracle.evolve(0, pts, None, None, syn.MASS, syn.DT, analytic=an)
    assert out[:, 0] == pytest.approx(r[:, 0] + syn.DT * r[:, 1] / syn.MASS, abs=1e-9)
    assert out[:, 1] == pytest.approx(r[:, 1], abs=1e-7)
    assert out[:, 2] == pytest.approx(rho.real, rel=1e-5)
    assert np.abs(out[:, 3]).max() < 1e-12
    q = oracle.backward_queries(0, out[0, 0], out[0, 1], syn.MASS, syn.DT, 0, 0)
    assert q[:, :, 0] == pytest.approx(r[0, 0], abs=1e-9)
    assert q[:, :, 1] == pytest.approx(r[0, 1], abs=1e-7)


def test_evolve_through_crossing_is_consistent_between_elements(oracle):
    """DAC near the first crossing with an analytic rho00-only distribution: population flows into rho11 and
    rho10; new_point_predict of an empty element equals the backward formula without the own density."""
    g = syn.rng(10, 0)
    n = 32
    c = (-1.5, syn.P0)
    r = np.stack([c[0] + syn.SIGMA_X * g.standard_normal(n), c[1] + syn.SIGMA_P * g.standard_normal(n)], 1)
    an = np.array([c[0], c[1], syn.SIGMA_X, syn.SIGMA_P, 1.0, 0.0, 0.0, 0.0])
    rho = oracle.initial_distribution(an, r, 0, 0)
    out, _, _ = oracle.evolve(1, syn.points_aos(r, rho), None, None, syn.MASS, 4.0, analytic=an)
    assert np.isfinite(out).all()
    assert (out[:, 2] <= rho.real.max() * 1.0001).all() and (out[:, 2] > 0).all()
    # branch geometry: the zero branch of the own element retraces the forward path exactly
    q = oracle.backward_queries(1, out[0, 0], out[0, 1], syn.MASS, 4.0, 0, 0)
    assert q[0, 1] == pytest.approx(r[0], abs=1e-10)
    # +-1 branches are displaced in momentum by -+ dt * F01 at x2
    assert q[0, 0, 1] != q[0, 2, 1]


def test_observable_sums(oracle):
    X, y = syn.training_set(11, 2, 500, centre=(0.3, syn.P0))
    pts = syn.points_aos(X, y)
    s = oracle.observable_sums(1, pts, syn.MASS, 1)
    E, _, _ = oracle.pes(1, X[:, 0])
    w = y.real
    ref = [w.sum(), (X[:, 0] * w).sum(), (X[:, 1] * w).sum(), X[:, 0].sum(), X[:, 1].sum(), (X[:, 0] ** 2).sum(), (X[:, 1] ** 2).sum(),
           ((X[:, 1] ** 2 / syn.MASS / 2 + E[:, 1]) * w).sum(), (np.abs(y) ** 2).sum()]
    assert s == pytest.approx(np.array(ref), rel=1e-12)
