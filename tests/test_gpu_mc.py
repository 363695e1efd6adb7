"""Metropolis sampling on the GPU (gple_markov_chains, gple_chain_autocorrelation) against the oracle's per-chain walk
(gple/mc.cpp:143-243) on the same Philox streams: every accept / reject decision, every chain state and the relabelled
densities must agree; the tuning logic of mc.py must then make the same choices on either backend."""
import numpy as np
import pytest

import oracle_backend
from gaussian_process_liouville_equation_b200 import complex_kernel, kernel, mc
from gaussian_process_liouville_equation_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
ANALYTIC = ((syn.X0, syn.P0), (syn.SIGMA_X, syn.SIGMA_P), (0.8, 0.6), (0.0, 0.4))
THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])


def start_points(config, n, centre):
    g = syn.rng(config, 0)
    pts = np.zeros((n, 4))
    pts[:, 0] = centre[0] + syn.SIGMA_X * g.standard_normal(n)
    pts[:, 1] = centre[1] + syn.SIGMA_P * g.standard_normal(n)
    return pts


def models(n, centre):
    sets = [syn.training_set(83, e, n, centre) for e in range(3)]
    th = [syn.theta_real(), THETA_C, syn.theta_real()]
    g = [kernel.TrainingKernel(th[0], sets[0]), complex_kernel.TrainingComplexKernel(th[1], sets[1]), kernel.TrainingKernel(th[2], sets[2])]
    o = [oracle_backend.TrainingKernel(th[0], sets[0]), oracle_backend.TrainingComplexKernel(th[1], sets[1]), oracle_backend.TrainingKernel(th[2], sets[2])]
    return g, o


@pytest.mark.parametrize("row,col", [(0, 0), (1, 0), (1, 1)])
def test_analytic_chains_parity(row, col):
    pts = start_points(84, 300, (syn.X0, syn.P0))
    a, acc_a, ch_a = mc.Sampler(9, analytic=ANALYTIC).chains(pts, 200, 0.5, row, col, want_chain=True)
    b, acc_b, ch_b = oracle_backend.Sampler(9, analytic=ANALYTIC).chains(pts, 200, 0.5, row, col, want_chain=True)
    assert np.array_equal(acc_a, acc_b)  # every decision identical
    assert np.abs(ch_a - ch_b).max() <= 1e-13 * np.abs(ch_b).max()
    assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()


def test_predicted_density_chains_parity():
    """predict_distribution target (main.cpp:75-101): real and complex element, plus an absent element (never moves)."""
    centre = (0.0, syn.P0)
    g, o = models(160, centre)
    pts = start_points(85, 64, centre)
    for row, col in ((0, 0), (1, 0), (1, 1)):
        a, acc_a, ch_a = mc.Sampler(3, kernels=g).chains(pts, 40, 0.3, row, col, want_chain=True)
        b, acc_b, ch_b = oracle_backend.Sampler(3, kernels=o).chains(pts, 40, 0.3, row, col, want_chain=True)
        same = np.all(np.abs(ch_a - ch_b) <= 1e-12 * np.abs(ch_b).max(), axis=(1, 2))
        assert same.mean() >= 0.95, (row, col, same.mean())  # a decision within rounding of the threshold may flip a chain
        assert np.abs(a[same, 2:] - b[same, 2:]).max() <= 1e-8 * np.abs(b[:, 2:]).max()
        assert np.array_equal(acc_a[same], acc_b[same]) and 0.0 < acc_a.mean() < 1.0
    a, acc_a, _ = mc.Sampler(3, kernels=[g[0], None, None]).chains(pts, 10, 0.3, 1, 1)
    assert np.array_equal(a[:, :2], pts[:, :2]) and np.all(acc_a == 0.0) and np.all(a[:, 2:] == 0.0)


def test_new_point_predict_chains_parity():
    centre = (-0.8, syn.P0)
    g, o = models(128, centre)
    pts = start_points(86, 48, centre)
    a, acc_a, ch_a = mc.Sampler(4, kernels=g, new_point=(1, syn.MASS, 1.0)).chains(pts, 25, 0.3, 1, 0, want_chain=True)
    b, acc_b, ch_b = oracle_backend.Sampler(4, kernels=o, new_point=(1, syn.MASS, 1.0)).chains(pts, 25, 0.3, 1, 0, want_chain=True)
    same = np.all(np.abs(ch_a - ch_b) <= 1e-12 * np.abs(ch_b).max(), axis=(1, 2))
    assert same.mean() >= 0.9, same.mean()
    assert np.abs(a[same, 2:] - b[same, 2:]).max() <= 1e-7 * np.abs(b[:, 2:]).max()


def test_autocorrelation_parity_and_tuning_logic():
    pts = start_points(87, 96, (syn.X0, syn.P0))
    sa, sb = mc.Sampler(5, analytic=ANALYTIC), oracle_backend.Sampler(5, analytic=ANALYTIC)
    _, _, ch = sb.chains(pts, 400, 0.5, 0, 0, want_chain=True)
    sa.calls = sb.calls
    assert np.allclose(sa.autocorrelation(ch), sb.autocorrelation(ch), rtol=1e-11, atol=1e-14)
    pa, pb = [mc.MCParameters() for _ in range(3)], [mc.MCParameters() for _ in range(3)]
    da = mc.monte_carlo_selection([pts, pts.copy(), None], pa, mc.Sampler(6, analytic=ANALYTIC))
    db = mc.monte_carlo_selection([pts, pts.copy(), None], pb, oracle_backend.Sampler(6, analytic=ANALYTIC))
    for e in range(2):
        assert pa[e].get_max_displacement() == pb[e].get_max_displacement() and pa[e].get_num_MC_steps() == pb[e].get_num_MC_steps()
        assert np.abs(da[e] - db[e]).max() <= 1e-12 * np.abs(db[e]).max()


def test_large_walk_keeps_the_target_distribution():
    """1e5 chains x 200 steps of the analytic target in one kernel: detailed balance at scale."""
    n = 100_000
    pts = start_points(88, n, (syn.X0, syn.P0))
    out, acc, _ = mc.Sampler(12, analytic=ANALYTIC).chains(pts, 200, 0.6, 0, 0)
    se = 5.0 / np.sqrt(n)
    assert abs(out[:, 0].mean() - syn.X0) < se * syn.SIGMA_X and abs(out[:, 1].mean() - syn.P0) < se * syn.SIGMA_P
    assert out[:, 0].std() == pytest.approx(syn.SIGMA_X, rel=0.01) and out[:, 1].std() == pytest.approx(syn.SIGMA_P, rel=0.01)
    assert 0.3 < acc.mean() < 0.9


def test_cpp_host_monte_carlo_selection_matches_oracle_sampler(tmp_path):
    """host/gple_mc.hpp drives the same walks as mc.py: identical tuned parameters and points as the oracle-backed run."""
    import subprocess

    from test_opt_cpp import compile_cpp

    exe = compile_cpp("mc_test", link=True)
    pts = start_points(89, 40, (syn.X0, syn.P0))
    path = tmp_path / "start.txt"
    path.write_text(f"{len(pts)}\n" + "".join(f"{float(x)!r} {float(p)!r}\n" for x, p in pts[:, :2]))
    out = subprocess.run([exe, str(path), "21"], capture_output=True, text=True, check=True).stdout
    got = {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}
    params = [mc.MCParameters() for _ in range(3)]
    ref = mc.monte_carlo_selection([pts, pts.copy(), None], params, oracle_backend.Sampler(21, analytic=ANALYTIC))
    for e in range(2):
        assert got[f"displacement{e}"] == params[e].get_max_displacement() and got[f"steps{e}"] == params[e].get_num_MC_steps()
        walked = np.array([[got[f"p{e}_{i}_{c}"] for c in ("x", "p", "re", "im")] for i in range(len(pts))])
        assert np.abs(walked - ref[e]).max() <= 1e-12 * np.abs(ref[e]).max()


def test_chain_blocks_reproduce_the_whole_set():
    """chain0: a block [lo, hi) of the points walked with chain0 = lo ends exactly where it ends inside the whole set
    (what a rank of a multi-GPU run does with its shard)."""
    pts = start_points(90, 101, (syn.X0, syn.P0))
    whole, acc, _ = mc.Sampler(13, analytic=ANALYTIC).chains(pts, 60, 0.5, 1, 0, stream=3)
    parts = [mc.Sampler(13, analytic=ANALYTIC).chains(pts[lo:hi], 60, 0.5, 1, 0, stream=3, chain0=lo) for lo, hi in ((0, 40), (40, 101))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), whole) and np.array_equal(np.concatenate([p[1] for p in parts]), acc)
    centre = (0.0, syn.P0)
    g, _ = models(96, centre)
    pts = start_points(91, 50, centre)
    whole, acc, _ = mc.Sampler(13, kernels=g).chains(pts, 15, 0.3, 0, 0, stream=4)
    parts = [mc.Sampler(13, kernels=g).chains(pts[lo:hi], 15, 0.3, 0, 0, stream=4, chain0=lo) for lo, hi in ((0, 17), (17, 50))]
    got = np.concatenate([p[0] for p in parts])
    assert np.array_equal(got[:, :2], whole[:, :2]) and np.abs(got[:, 2:] - whole[:, 2:]).max() <= 1e-12 * np.abs(whole[:, 2:]).max()
