"""The oracle restatement is pinned to the reference ITSELF: oracle/_ref holds the reference's own translation units
(gple/kernel.cpp, complex_kernel.cpp, pes.cpp, evolve.cpp, predict.cpp, mc.cpp), compiled unmodified against stand-in
headers (oracle/Makefile.ref, oracle/refstub/), and every oracle entry point that has a public counterpart in the reference
is compared with it here on the golden inputs and on a second, larger seeded set.

Tolerances.  Both sides run the same formulas in double precision; they differ in the summation order of the matrix
products (the reference's go through a BLAS, the oracle's are plain loops).  Well-conditioned outputs agree to ~1e-15;
everything that passes through K^-1 carries eps * cond(K) ~ 1e-16 * N / sigma_n^2, and the reference's gradient formulas of
population / purity / validation error square that condition number (kernel.cpp:337-364, 401-477, 524-541) -- those are
compared at the looser level the two CPU implementations themselves reach, which is also the level the CUDA path is held to.
"""
import os

import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.build(), reason="oracle/_ref is not built and /root/reference is not present")

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gple_golden_v1.npz"))
THETA_C2 = np.array([0.9, 1.1, 1.3 * syn.SIGMA_X, 0.8 * syn.SIGMA_P, 0.8, 0.9 * syn.SIGMA_X, 1.2 * syn.SIGMA_P, 3e-2])


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def elementwise(a, b, tol=1e-12):
    """north_star's kernel-matrix bar: every entry within 1e-12 relative (an entry exp(-a) carries ~a ulps of its argument)"""
    a, b = np.asarray(a), np.asarray(b)
    return bool((np.abs(a - b) <= tol * np.abs(b) + 1e-300).all())


def inputs(which):
    """(theta_r, theta_c, [(X, y)] * 3, (Xq, yq), (Xqc, yqc)) -- the golden set (N = 64) or a second seeded set (N = 150)."""
    if which == "golden":
        return G["theta_r"], G["theta_c"], [(G[f"X{e}"], G[f"y{e}"]) for e in range(3)], (G["Xq"], G["yq"]), (G["Xqc"], G["yqc"])
    centre = (0.3, syn.P0)
    sets = [syn.training_set(52, e, 150, centre) for e in range(3)]
    return (np.array([1.1, 1.2 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 3e-2]), THETA_C2, sets, syn.extra_points(52, 0, sets[0][0], 200, centre),
            syn.extra_points(52, 1, sets[1][0], 200, centre))


@pytest.mark.parametrize("which", ["golden", "second"])
def test_kernel_matrices_and_derivative_arrays(oracle, which):
    th, thc, sets, (Xq, _), _ = inputs(which)
    X = sets[0][0]
    for left, same in ((X, True), (Xq, False)):
        Ko, dKo = oracle.kernel_real(left, X, th, same, True)
        Kr, dKr = ref.kernel_real(left, X, th, same, True)
        assert elementwise(Kr, Ko) and elementwise(dKr, dKo)
        co, cr = oracle.kernel_complex(left, X, thc, same, True), ref.kernel_complex(left, X, thc, same, True)
        for a, b in zip(cr, co):
            # complex entries: compare against the modulus (real and imaginary parts are built from different sub-kernels)
            assert elementwise(a, b)
    # delta_kernel on equal columns of DIFFERENT buffers (kernel.cpp:8-31): a query that coincides with a training point
    Xq2 = Xq.copy()
    Xq2[3] = X[7]
    assert elementwise(ref.kernel_real(Xq2, X, th, False, False), oracle.kernel_real(Xq2, X, th, False, False))


@pytest.mark.parametrize("which", ["golden", "second"])
def test_training_kernel(oracle, which):
    th, _, sets, (Xq, yq), _ = inputs(which)
    o, r = oracle.TrainingKernel(th, *sets[0], True, True, True), ref.TrainingKernel(th, *sets[0], True, True, True)
    assert r.rescale == o.rescale
    assert rel(r.error, o.error) < 1e-9 and rel(r.population, o.population) < 1e-12 and rel(r.first_order, o.first_order) < 1e-12
    assert rel(r.purity, o.purity) < 1e-12 and rel(r.magnitude, o.magnitude) < 1e-12
    assert rel(r.v, o.v) < 1e-9 and rel(r.inverse, o.inverse) < 1e-9 and elementwise(r.K, o.K)
    assert rel(r.derror, o.derror) < 1e-8
    # cond(K)^2-limited gradients: the level at which the reference and its restatement agree with each other
    assert rel(r.dpopulation, o.dpopulation) < 1e-5 and rel(r.dpurity, o.dpurity) < 1e-3
    for p in range(4):
        assert rel(r.dv(p), o.dv(p)) < 1e-6 and elementwise(r.dK(p), o.dK(p))
    po, pr = o.predict(Xq, yq.real, True), r.predict(Xq, yq.real, True)
    prior = th[0] ** 2 * (1 + th[3] ** 2)
    assert rel(pr["pred"], po["pred"]) < 1e-12 and np.abs(pr["var"] - po["var"]).max() < 1e-9 * prior
    assert rel(pr["cutoff"], po["cutoff"]) < 1e-9 and rel(pr["error"], po["error"]) < 1e-10 and rel(pr["derror"], po["derror"]) < 1e-4
    # the flag combinations the callers use: (err, avg, deriv) = loose_function (T, F, grad), TrainingKernels (T, T, F), constraints (F, T, grad)
    for flags in ((True, False, False), (False, True, True), (False, False, False)):
        o2, r2 = oracle.TrainingKernel(th, *sets[0], *flags), ref.TrainingKernel(th, *sets[0], *flags)
        for f in ("error", "population", "purity"):
            a, b = getattr(r2, f), getattr(o2, f)
            assert (np.isnan(a) and np.isnan(b)) or rel(a, b) < 1e-9


@pytest.mark.parametrize("which", ["golden", "second"])
def test_training_complex_kernel(oracle, which):
    _, thc, sets, _, (Xqc, yqc) = inputs(which)
    o, r = oracle.TrainingComplexKernel(thc, *sets[1], True, True, True), ref.TrainingComplexKernel(thc, *sets[1], True, True, True)
    assert r.rescale == o.rescale
    assert rel(r.error, o.error) < 1e-9 and rel(r.purity, o.purity) < 1e-9 and rel(r.magnitude, o.magnitude) < 1e-10
    assert rel(r.P, o.P) < 1e-9 and rel(r.Q, o.Q) < 1e-9 and rel(r.v, o.v) < 1e-9
    assert rel(r.derror, o.derror) < 1e-8 and rel(r.dpurity, o.dpurity) < 1e-8
    for p in range(8):
        assert rel(r.dv(p), o.dv(p)) < 1e-7
    po, pr = o.predict(Xqc, yqc, True), r.predict(Xqc, yqc, True)
    prior = thc[0] ** 2 * (thc[1] ** 2 + thc[4] ** 2 + thc[7] ** 2)
    assert rel(pr["pred"], po["pred"]) < 1e-9 and np.abs(pr["var"] - po["var"]).max() < 1e-9 * prior
    assert rel(pr["cutoff"], po["cutoff"]) < 1e-9 and rel(pr["error"], po["error"]) < 1e-9 and rel(pr["derror"], po["derror"]) < 1e-7


@pytest.mark.parametrize("model", [0, 1, 2])
def test_pes(oracle, model):
    x = np.concatenate([np.linspace(-12, 12, 97), [1e-9, -1e-9, 0.37]])
    for a, b in zip(ref.pes(model, x), oracle.pes(model, x)):
        # adiabatic forces are differences of O(1e-2) diabatic terms (C^T F C): absolute floor at the scale of the array
        assert (np.abs(a - b) <= 1e-13 * np.abs(b) + 1e-15 * np.abs(b).max()).all()


@pytest.mark.parametrize("model", [0, 1, 2])
def test_evolve_new_point_predict_and_observables(oracle, model):
    th, thc, sets, _, _ = inputs("golden")
    o = [oracle.TrainingKernel(th, *sets[0]), oracle.TrainingComplexKernel(thc, *sets[1]), oracle.TrainingKernel(th, *sets[2])]
    r = [ref.TrainingKernel(th, *sets[0]), ref.TrainingComplexKernel(thc, *sets[1]), ref.TrainingKernel(th, *sets[2])]
    pts = [syn.points_aos(*s) for s in sets]
    for dt in (2.0, 0.5):
        a, b = oracle.evolve(model, pts[0], pts[1], pts[2], syn.MASS, dt, o[0], o[1], o[2]), ref.evolve(model, pts[0], pts[1], pts[2], syn.MASS, dt, r[0], r[1], r[2])
        for u, v in zip(a, b):
            assert np.array_equal(u[:, :2], v[:, :2])  # phase-space coordinates: bit for bit
            assert np.abs(u[:, 2:] - v[:, 2:]).max() < 1e-9 * np.abs(v[:, 2:]).max()
    # only rho00 populated (the t = 0 state): absent elements predict 0 (main.cpp:85-99)
    a, b = oracle.evolve(model, pts[0], None, None, syn.MASS, 2.0, o[0], None, None), ref.evolve(model, pts[0], None, None, syn.MASS, 2.0, r[0], None, None)
    assert np.array_equal(a[0][:, :2], b[0][:, :2]) and np.abs(a[0][:, 2:] - b[0][:, 2:]).max() < 1e-9 * np.abs(b[0][:, 2:]).max()
    an = np.array([-0.8, syn.P0, syn.SIGMA_X, syn.SIGMA_P, 0.8, 0.6, 0.1, -0.2])
    for u, v in zip(oracle.evolve(model, pts[0], pts[1], pts[2], syn.MASS, 2.0, analytic=an), ref.evolve(model, pts[0], pts[1], pts[2], syn.MASS, 2.0, analytic=an)):
        assert np.abs(u - v).max() < 1e-14 * np.abs(v).max()
    q = pts[0][:24, :2]
    for row, col in ((0, 0), (1, 0), (1, 1)):
        a, b = oracle.new_point_predict(model, q, syn.MASS, 2.0, row, col, *o), ref.new_point_predict(model, q, syn.MASS, 2.0, row, col, *r)
        assert rel(a, b) < 1e-9
        assert rel(oracle.initial_distribution(an, q, row, col), ref.initial_distribution(an, q, row, col)) < 1e-14
    for e, surface in ((0, 0), (2, 1)):
        s, ro = oracle.observable_sums(model, pts[e], syn.MASS, surface), ref.observables(model, pts[e], syn.MASS, surface)
        n = len(pts[e])
        mine = dict(x=s[1] / s[0], p=s[2] / s[0], std_x=np.sqrt(s[5] / n - (s[3] / n) ** 2), std_p=np.sqrt(s[6] / n - (s[4] / n) ** 2), energy=s[7] / s[0], purity_sum=s[8])
        for k, v in ro.items():
            assert abs(mine[k] / v - 1) < 1e-11, k


def test_training_kernels_aggregate_and_is_very_small(oracle):
    th, thc, sets, _, _ = inputs("golden")
    pts = [syn.points_aos(*s) for s in sets]
    o = [oracle.TrainingKernel(th, *sets[0]), oracle.TrainingComplexKernel(thc, *sets[1]), oracle.TrainingKernel(th, *sets[2])]
    energies = np.array([0.013, 0.041])
    agg = ref.training_kernels(np.concatenate([th, thc, th]), pts[0], pts[1], pts[2], energies)
    assert agg["population"] == pytest.approx(o[0].population + o[2].population, rel=1e-12)
    assert agg["energy"] == pytest.approx(o[0].population * energies[0] + o[2].population * energies[1], rel=1e-12)
    assert agg["purity"] == pytest.approx(o[0].purity + o[2].purity + 2.0 * o[1].purity, rel=1e-9)
    assert [agg["x"], agg["p"]] == pytest.approx(list(o[0].first_order + o[2].first_order), rel=1e-12)
    # an all-zero complex parameter vector switches the off-diagonal element off (predict.cpp:339-357)
    agg0 = ref.training_kernels(np.concatenate([th, np.zeros(8), th]), pts[0], pts[1], pts[2], energies)
    assert agg0["purity"] == pytest.approx(o[0].purity + o[2].purity, rel=1e-12)
    # is_very_small (evolve.cpp:444-478) with rho11 and rho10 empty: far from the crossing they stay small, at it they do not
    r0 = ref.TrainingKernel(th, *sets[0])
    for model, centre, expect in ((0, -8.0, True), (0, 0.0, False)):
        X, y = syn.training_set(61, 0, 48, (centre, syn.P0))
        k = ref.TrainingKernel(th, X, y)
        ko = oracle.TrainingKernel(th, X, y)
        flags = ref.is_very_small(model, syn.points_aos(X, y), None, None, syn.MASS, 1.0, k, None, None)
        assert flags[0] is False and flags[1] == expect and flags[2] == expect
        mine = [bool((np.abs(oracle.new_point_predict(model, X, syn.MASS, 1.0, row, col, ko, None, None)) ** 2 < 1e-10).all()) for row, col in ((1, 0), (1, 1))]
        assert mine == flags[1:]
    del r0
