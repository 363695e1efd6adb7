"""GPU parity of the real-kernel chain against the CPU oracle, through the C-ABI (via the host mirror).

Tolerances are BASELINE.json's: kernel matrices 1e-12 relative, predictions / losses 1e-9 relative.
"Relative" for vectors means relative to the largest magnitude of the vector (the quantities are sums of
O(N) terms of that size).  The variance k** - k K^-1 k^T is a difference of nearly equal numbers in the
reference's own formulation (two summation orders of the ORACLE differ by ~1e-11 k**, see
tests/test_oracle_real.py), so it is compared relative to the prior variance k**.
"""
import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gk():
    from gaussian_process_liouville_equation_b200 import kernel

    return kernel


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("n,m", [(1, 1), (50, 31), (300, 257)])
def test_kernel_matrix_parity(gk, oracle, n, m):
    X, _ = syn.training_set(1, 0, n)
    Xq, _ = syn.training_set(2, 0, m)
    if n > 7 and m > 3:
        Xq[3] = X[7]
    th = np.array([1.3, 0.6, 0.9, 0.05])
    for XL, XR, same in ((X, X, True), (Xq, X, False)):
        K, dK = gk.kernel_matrix(XL, XR, th, same, True)
        Ko, dKo = oracle.kernel_real(XL, XR, th, same, True)
        assert np.abs(K - Ko).max() <= 1e-12 * np.abs(Ko).max()
        assert (np.abs(K - Ko) <= 1e-12 * np.abs(Ko) + 1e-300).all()  # elementwise relative
        for p in range(4):
            assert (np.abs(dK[p] - dKo[p]) <= 1e-12 * np.abs(dKo[p]) + 1e-300).all()


@pytest.mark.parametrize("n", [1, 96, 300, 700])
def test_training_scalars_parity(gk, oracle, n):
    X, y = syn.training_set(1, 0, n)
    th = syn.theta_real()
    k = gk.TrainingKernel(th, (X, y), True, True, False)
    o = oracle.TrainingKernel(th, X, y, True, True, False)
    assert k.status == 0
    assert k.get_rescale_factor() == pytest.approx(o.rescale, rel=1e-15)
    assert k.get_error() == pytest.approx(o.error, rel=1e-9)
    assert k.get_population() == pytest.approx(o.population, rel=1e-9)
    assert k.get_1st_order_average() == pytest.approx(o.first_order, rel=1e-9, abs=1e-9 * abs(o.first_order).max())
    assert k.get_purity() == pytest.approx(o.purity, rel=1e-9)
    assert k.get_magnitude() == pytest.approx(o.magnitude, rel=1e-9)
    assert rel(k.get_inverse_times_label(), o.v) <= 1e-8  # cond(K) ~ N / sigma_n^2 ~ 1e7: eps * cond
    assert rel(k.get_label(), o.label) <= 1e-15
    if n <= 300:
        assert rel(k.get_inverse(), o.inverse) <= 1e-9


def test_prediction_parity(gk, oracle):
    X, y = syn.training_set(1, 0, 300)
    th = syn.theta_real()
    Xq, _ = syn.extra_points(1, 0, X, 1500)
    Xq[:100] += np.array([4.0, 0.0])
    Xq[100:300] += np.array([1.7, 0.0])
    Xq[300] = X[5]  # exact coincidence: delta_kernel adds the noise term (kernel.cpp:16-29)
    yq = syn.labels(0, Xq, (0.0, syn.P0)).real
    k = gk.TrainingKernel(th, (X, y), True, True, False)
    o = oracle.TrainingKernel(th, X, y, True, True, False)
    p = gk.PredictiveKernel(Xq, k, False, yq)
    r = o.predict(Xq, yq, False)
    prior = th[0] ** 2 * (1 + th[3] ** 2)
    assert rel(p.get_prediction(), r["pred"]) <= 1e-9
    assert np.abs(p.get_variance() - r["var"]).max() <= 1e-9 * prior
    assert p.get_error() == pytest.approx(r["error"], rel=1e-9)
    # cutoff prediction: gate closed / open -> exact; inside the cubic band the gate inherits the variance noise
    gate_o = np.divide(r["cutoff"] * o.rescale, r["pred"], out=np.ones_like(r["pred"]), where=r["pred"] != 0)
    sharp = (gate_o == 0) | (gate_o == 1)
    assert sharp.any() and (~sharp).any()
    scale = np.abs(r["cutoff"]).max()
    near_edge = np.abs(r["pred"] ** 2 - r["var"]) < 1e-6 * r["var"]  # a gate flip here is rounding, not error
    ok = sharp & ~near_edge
    assert np.abs(p.get_cutoff_prediction() - r["cutoff"])[ok].max() <= 1e-9 * scale
    assert np.abs(p.get_cutoff_prediction() - r["cutoff"]).max() <= 1e-6 * scale


def test_single_point_prediction_like_predict_distribution(gk, oracle):
    """main.cpp:75-101: PredictiveKernel(r, kernel, false).get_cutoff_prediction().value() for one point."""
    X, y = syn.training_set(3, 0, 200)
    th = syn.theta_real()
    k = gk.TrainingKernel(th, (X, y))
    o = oracle.TrainingKernel(th, X, y)
    for r in (X[17] + 0.01, np.array([0.3, syn.P0 - 0.2])):
        a = gk.PredictiveKernel(r, k).get_cutoff_prediction()
        b = o.predict(r.reshape(1, 2))["cutoff"]
        assert a == pytest.approx(b, rel=1e-9)


def test_gradients_parity(gk, oracle):
    X, y = syn.training_set(5, 0, 200)
    th = np.array([1.0, 0.9 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 2e-2])
    k = gk.TrainingKernel(th, (X, y), True, True, True)
    o = oracle.TrainingKernel(th, X, y, True, True, True)
    sc = np.abs(o.derror).max()
    assert np.abs(k.get_error_derivative() - o.derror).max() <= 1e-8 * sc
    # the noise derivative is -2 sf^2 sn K^-2 y: squared condition number, ~1e-8 on either side
    assert np.abs(k.get_population_derivative() - o.dpopulation).max() <= 1e-7 * np.abs(o.dpopulation).max()
    assert np.abs(k.get_purity_derivative() - o.dpurity).max() <= 1e-5 * np.abs(o.dpurity).max()
    dv = k.get_inverse_times_label_derivative()
    for p in range(4):
        assert np.abs(dv[p] - o.dv(p)).max() <= 1e-8 * np.abs(o.dv(p)).max()
    # validation error gradient (quirk q1) and loose_function (opt.cpp:441-482)
    Xq, yq = syn.extra_points(5, 0, X, 1000)
    pq = gk.PredictiveKernel(Xq, k, True, yq.real)
    ro = o.predict(Xq, yq.real, True)
    assert pq.get_error() == pytest.approx(ro["error"], rel=1e-9)
    assert np.abs(pq.get_error_derivative() - ro["derror"]).max() <= 1e-7 * np.abs(ro["derror"]).max()
    from gaussian_process_liouville_equation_b200 import dynamics

    val, g = dynamics.loose_function(th, (X, y), (Xq, yq), grad=True)
    vo, go = oracle.loose_function(th, X, y, Xq, yq, grad=True)
    assert val == pytest.approx(vo, rel=1e-9)
    assert np.abs(g - go).max() <= 1e-7 * np.abs(go).max()
    assert dynamics.loose_function(th, (X, y), (Xq, yq)) == pytest.approx(vo, rel=1e-9)
    # the same split the way the host shares a trained model between objective and constraints (gple_validation_error)
    ve, vg = dynamics.validation_error(k, (Xq, yq), grad=True)
    assert k.get_error() + ve == pytest.approx(val, rel=1e-13) and np.abs(k.get_error_derivative() + vg - g).max() <= 1e-12 * np.abs(g).max()
    assert dynamics.validation_error(k, (Xq, yq)) == pytest.approx(ve, rel=1e-13)


def test_large_n_properties(gk):
    """N = 2048 (BASELINE config C2): size-independent properties instead of the (slow) oracle."""
    n = 2048
    X, y = syn.training_set(2, 0, n)
    th = syn.theta_real()
    k = gk.TrainingKernel(th, (X, y), True, True, False)
    assert k.status == 0
    v, lab = k.get_inverse_times_label(), k.get_label()
    K = gk.kernel_matrix(X, X, th, True)
    assert np.abs(K @ v - lab).max() <= 1e-8 * np.abs(lab).max()  # K (K^-1 y) = y
    assert k.get_population() == pytest.approx(0.6, rel=1e-3)
    assert k.get_purity() == pytest.approx(0.36, rel=2e-3)
    # linearity in the labels: v(2y) has the same rescaled solution, population doubles
    k2 = gk.TrainingKernel(th, (X, 2 * y), True, True, False)
    assert k2.get_population() == pytest.approx(2 * k.get_population(), rel=1e-10)
    # prediction at training points reproduces label - sigma_n^2-weighted residual: K* v with K* = K - sigma^2 I
    p = gk.PredictiveKernel(X[:512] + 1e-9, k)
    expect = (K[:512] - th[0] ** 2 * th[3] ** 2 * np.eye(n)[:512]) @ v
    assert np.abs(p.get_prediction() - expect).max() <= 1e-7 * np.abs(expect).max()
    assert (p.get_variance() > 0).all() and (p.get_variance() < th[0] ** 2 * (1 + th[3] ** 2)).all()


def test_not_spd_is_reported(gk):
    """Duplicate points with zero noise make K singular: status code, NaN scalars, no exception."""
    X, y = syn.training_set(1, 0, 64)
    X[1] = X[0]
    k = gk.TrainingKernel(np.array([1.0, 0.7, 0.7, 0.0]), (X, y), True, True, False)
    assert k.status == 2 and np.isnan(k.get_error())


def test_chunk_boundaries_and_batch_invariance(gk, oracle):
    """Predictions are independent of where a point falls in the batch: across the 18944-row chunk boundary, in the
    n-split tail chunk (partial row block), and as a single-point call."""
    X, y = syn.training_set(8, 0, 260)
    th = syn.theta_real()
    k = gk.TrainingKernel(th, (X, y))
    Q = 148 * 128 + 57
    Xq, _ = syn.extra_points(8, 0, X, Q)
    p = gk.PredictiveKernel(Xq, k)
    idx = np.array([0, 1, 127, 128, 9471, 18943, 18944, 18945, Q - 1])
    sub = gk.PredictiveKernel(Xq[idx], k)
    assert np.array_equal(p.get_prediction()[idx], sub.get_prediction())
    assert np.abs(p.get_variance()[idx] - sub.get_variance()).max() <= 1e-13 * th[0] ** 2
    for i in (18943, 18944, Q - 1):
        one = gk.PredictiveKernel(Xq[i], k)
        assert one.get_prediction()[0] == p.get_prediction()[i]
        assert abs(one.get_variance()[0] - p.get_variance()[i]) <= 1e-13
    o = oracle.TrainingKernel(th, X, y).predict(Xq[idx])
    assert np.abs(p.get_prediction()[idx] - o["pred"]).max() <= 1e-9 * np.abs(o["pred"]).max()
    assert np.abs(p.get_variance()[idx] - o["var"]).max() <= 1e-9 * th[0] ** 2


def test_argument_errors_are_status_codes(gk):
    from gaussian_process_liouville_equation_b200 import _lib as L

    ctx = L.default_context()
    import ctypes as C

    h, s = C.c_void_p(), L.RealScalars()
    th = syn.theta_real()
    X, y = syn.training_set(1, 0, 8)
    yv = np.ascontiguousarray(y).view(np.float64)
    assert ctx.lib.gple_train_real(ctx.h, L.addr(X), L.addr(yv), 0, L.addr(th), 3, C.byref(h), C.byref(s)) == L.ERR_ARG  # empty training set
    assert ctx.lib.gple_train_real(ctx.h, None, L.addr(yv), 8, L.addr(th), 3, C.byref(h), C.byref(s)) == L.ERR_ARG
    k = gk.TrainingKernel(th, (X, y))
    out = np.empty(4)
    assert ctx.lib.gple_predict_complex(ctx.h, k.h, L.addr(X), 2, None, L.addr(out), None, None, None, None) == L.ERR_ARG  # wrong kind
    assert ctx.lib.gple_predict_real(ctx.h, k.h, L.addr(X), 0, None, L.addr(out), None, None, None, None) == L.ERR_ARG  # no query
    assert b"model kind" in ctx.lib.gple_last_error(ctx.h) or b"query" in ctx.lib.gple_last_error(ctx.h)


def test_nlml_objective_parity(gk, oracle):
    """gple_model_nlml (the NLML / LLT objective of test/gpr.cpp:470-532) against the oracle: value 1e-9, gradient 1e-7."""
    X, y = syn.training_set(41, 0, 300)
    th = syn.theta_real() * np.array([1.1, 1.0, 1.0, 3.0])
    k = gk.TrainingKernel(th, (X, y))
    o = oracle.TrainingKernel(th, X, y, deriv=True)
    ov, og = o.nlml(grad=True)
    v, g = k.get_negative_log_marginal_likelihood(grad=True)
    assert v == pytest.approx(ov, rel=1e-9)
    assert np.abs(g - og).max() <= 1e-7 * np.abs(og).max()
    assert k.get_negative_log_marginal_likelihood() == v


@pytest.mark.parametrize("n", [1100, 1664, 2100, 4200])
def test_factorisation_schedules_on_ragged_block_counts(gk, n):
    """9, 13, 17 and 33 blocks of 128: odd block counts exercise the ragged pairs of the level-batched triangular inverse,
    n = 4200 the switch from the recursive form (above 4096 rows) to the right-looking sweep inside it."""
    X, y = syn.training_set(5, 0, n)
    th = syn.theta_real((2048.0 / max(n, 2048)) ** 0.5)
    k = gk.TrainingKernel(th, (X, y), True, False, False)
    assert k.status == 0
    K = gk.kernel_matrix(X, X, th, True)
    v, lab = k.get_inverse_times_label(), k.get_label()
    assert np.abs(K @ v - lab).max() <= 1e-8 * np.abs(lab).max()
    Kinv = k.get_inverse()
    assert np.abs(Kinv - Kinv.T).max() <= 1e-9 * np.abs(Kinv).max()
    probe = np.random.default_rng(n).standard_normal((n, 8))
    assert np.abs(K @ (Kinv @ probe) - probe).max() <= 1e-7 * np.abs(probe).max()
    assert k.get_error() == pytest.approx(float(np.sum((v / np.diag(Kinv)) ** 2)), rel=1e-9)


def test_factorisation_as_a_cuda_graph_changes_nothing(gk):
    """GPLE_OPT_FACTORISE_GRAPHS (default on): the factorisation schedule replayed from a captured CUDA graph launches the same
    kernels on the same buffers as the direct launches -- every scalar, v and a prediction bit for bit, over alternating sizes
    (one cached graph per size and buffer set) and repeated evaluations (the optimiser's pattern)."""
    from gaussian_process_liouville_equation_b200 import _lib as L

    ctx = L.default_context()
    out = {}
    for graphs in (True, False):
        ctx.set_factorise_graphs(graphs)
        try:
            res = []
            for n in (300, 700, 300, 1100, 700, 300):
                X, y = syn.training_set(5, 0, n)
                k = gk.TrainingKernel(syn.theta_real(), (X, y), True, True, False)
                p = gk.PredictiveKernel(X[:50] + 0.01, k)
                res.append((k.get_error(), k.get_population(), k.get_purity(), p.get_prediction().copy(), p.get_variance().copy()))
            out[graphs] = res
        finally:
            ctx.set_factorise_graphs(True)
    for a, b in zip(out[True], out[False]):
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
