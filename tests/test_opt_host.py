"""Host-side optimiser layer (predict.TrainingKernels, opt.Callbacks, opt.Optimization).

CPU: the layer is driven through the oracle-backed adapter (tests/oracle_backend.py) -- aggregation,
constraint Jacobians (finite differences) and a small end-to-end optimisation.
GPU: the same driver on the CUDA library must reach the same outcome as the oracle-backed run.
"""
import numpy as np
import pytest

import oracle_backend
from gaussian_process_liouville_equation_b200 import opt, predict
from gaussian_process_liouville_equation_b200 import synthetic as syn

THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 1e-2])


def make_sets(n, m, elements=(0, 1, 2), centre=(0.0, syn.P0)):
    ts, ets = [None] * 3, [None] * 3
    for e in elements:
        ts[e] = syn.training_set(60, e, n, centre)
        ets[e] = syn.extra_points(60, e, ts[e][0], m, centre)
    return ts, ets


def test_log_transforms_round_trip():
    for p in (syn.theta_real(), THETA_C):
        g = opt.local_parameter_to_global(p)
        assert np.allclose(opt.global_parameter_to_local(g), p, rtol=1e-15)
        grad = np.arange(1.0, len(p) + 1)
        gg = opt.local_gradient_to_global(p, grad)
        idx = [1, 4, 7] if len(p) == 8 else [3]
        assert np.allclose(gg[idx], grad[idx] * p[idx]) and np.allclose(np.delete(gg, idx), np.delete(grad, idx))
    lo, hi = opt.calculate_complex_kernel_bounds([0.1, 0.2], [1.0, 2.0])
    assert lo[0] == hi[0] == 1.0 and lo[7] == hi[7] == 1e-2 and lo[1] == 0.1 and hi[4] == 10.0 and (lo[[2, 5]] == 0.1).all() and (hi[[3, 6]] == 2.0).all()


def test_callbacks_aggregate_and_jacobians():
    ts, ets = make_sets(40, 60)
    cb = opt.Callbacks(ts, ets, Energies=np.array([0.05, 0.06]), TotalEnergy=0.055, Purity=1.0, backend=oracle_backend)
    x = np.concatenate([syn.theta_real(0.9), THETA_C, syn.theta_real(1.1)])
    v, g = cb.full_loose(x, True)
    parts = [oracle_backend.loose_function(x[sl], ts[e], ets[e], True) for e, sl in enumerate(predict.ELEMENT_SLICES)]
    assert v == pytest.approx(sum(p[0] for p in parts), rel=1e-14) and np.allclose(g, np.concatenate([p[1] for p in parts]), rtol=1e-14)
    xd = np.concatenate([x[0:4], x[12:16]])
    vd, gd = cb.diagonal_loose(xd, True)
    assert vd == pytest.approx(parts[0][0] + parts[2][0], rel=1e-14)
    res, jac = cb.full_constraints(x, True)
    assert jac.shape == (3, 16) and (jac[0, 4:12] == 0).all() and (jac[1, 4:12] == 0).all()
    for p in (1, 2, 5, 6, 9, 13, 14):  # length / sub-magnitude parameters (magnitude, noise are pinned by the bounds)
        if p in (6, 7, 9, 10):  # quirk q10: complex length derivatives of the purity are not true derivatives
            continue
        h = 1e-4 * x[p]  # the purity of the ill-conditioned complex element is noisy at the 1e-9 level
        xp, xm = x.copy(), x.copy()
        xp[p] += h
        xm[p] -= h
        fd = (cb.full_constraints(xp) - cb.full_constraints(xm)) / (2 * h)
        assert np.allclose(jac[:, p], fd, rtol=2e-3, atol=1e-7), p
    res2, jac2 = cb.diagonal_constraints(xd, 2, True)
    assert np.allclose(res2, res[:2], rtol=1e-12) and np.allclose(jac2, np.hstack([jac[:2, 0:4], jac[:2, 12:16]]), rtol=1e-12)


def run_optimisation(backend):
    """rho00-only C1-like case: Nelder-Mead per element, then SLSQP with population / energy / purity constraints."""
    n = 48
    X, y = syn.training_set(61, 0, n, (syn.X0, syn.P0))
    y = y / 0.6  # population 1, purity 1
    Xe, ye = syn.extra_points(61, 0, X, 5 * n, (syn.X0, syn.P0))
    density = [syn.points_aos(X, y), None, None]
    extra = [syn.points_aos(Xe, ye / 0.6), None, None]
    o = backend.observable_sums(0, density[0], syn.MASS, 0) if backend else None
    if o is None:
        from gaussian_process_liouville_equation_b200 import dynamics

        o = dynamics.observable_sums(0, density[0], syn.MASS, 0)
    optimizer = opt.Optimization((syn.SIGMA_X, syn.SIGMA_P), syn.MASS, 0, InitialTotalEnergy=o[7] / o[0], InitialPurity=1.0, backend=backend, max_global_evals=200)
    (err, steps, typ), check = optimizer.optimize(density, extra)
    k = predict.TrainingKernels(optimizer.get_parameters(), predict.construct_training_sets(density), True, True, False, backend)
    return optimizer.get_parameters()[0], err, k.calculate_population(), k.calculate_purity(), check, steps


def test_optimisation_on_oracle_backend():
    params, err, pop, pur, check, steps = run_optimisation(oracle_backend)
    lo = np.array([syn.SIGMA_X, syn.SIGMA_P]) / np.sqrt(48) * 0.5
    assert (params[1:3] > lo).all() and (params[1:3] < 4 * np.array([syn.SIGMA_X, syn.SIGMA_P])).all()
    assert abs(pop - 1.0) < 2 * opt.AverageTolerance and steps[0] > 10
    assert np.isfinite(err) and err < 1.0


@pytest.mark.gpu
def test_optimisation_gpu_matches_oracle_backed_run():
    a = run_optimisation(None)
    b = run_optimisation(oracle_backend)
    assert a[1] == pytest.approx(b[1], rel=1e-3)  # final loss
    assert a[2] == pytest.approx(b[2], rel=1e-4) and a[3] == pytest.approx(b[3], rel=1e-3)  # population, purity
    assert np.allclose(a[0][1:3], b[0][1:3], rtol=1e-2)  # optimised characteristic lengths


@pytest.mark.gpu
def test_training_kernels_and_extra_points_on_gpu(oracle):
    from gaussian_process_liouville_equation_b200 import mc

    ts, _ = make_sets(120, 10)
    pv = [syn.theta_real(), THETA_C, syn.theta_real()]
    k = predict.TrainingKernels(pv, ts, True, True, True)
    ko = predict.TrainingKernels(pv, ts, True, True, True, oracle_backend)
    assert k.calculate_population() == pytest.approx(ko.calculate_population(), rel=1e-9)
    assert k.calculate_purity() == pytest.approx(ko.calculate_purity(), rel=1e-8)
    assert np.abs(k.purity_derivative() - ko.purity_derivative()).max() <= 1e-6 * np.abs(ko.purity_derivative()).max()
    density = [syn.points_aos(*t) for t in ts]
    extra = mc.generate_extra_points(density, 300, k, syn.rng(62, 0), 1, syn.MASS)
    for e in range(3):
        assert extra[e].shape == (300, 4)
        o = ko[e].o.predict(extra[e][:, :2])["cutoff"]
        ref = o if e == 1 else o.astype(complex)
        got = extra[e][:, 2] + 1j * extra[e][:, 3]
        assert np.median(np.abs(got - ref)) <= 1e-9 * np.abs(ref).max()
    assert mc.is_very_small(density, syn.MASS, 1.0, k, 1) == [False, False, False]
    small = mc.is_very_small([density[0], None, None], syn.MASS, 1.0, [k[0], None, None], 0)
    assert small[0] is False and isinstance(small[1], bool)
