"""Golden fixtures (tests/golden/gple_golden_v1.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle must keep reproducing them.  GPU: the CUDA path must match them without needing the oracle."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gple_golden_v1.npz"))


def close(a, b, tol):
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * np.abs(np.asarray(b)).max()


def test_oracle_reproduces_golden(oracle):
    k0 = oracle.TrainingKernel(G["theta_r"], G["X0"], G["y0"], True, True, True)
    k1 = oracle.TrainingComplexKernel(G["theta_c"], G["X1"], G["y1"], True, True, True)
    k2 = oracle.TrainingKernel(G["theta_r"], G["X2"], G["y2"], True, True, False)
    assert close([k0.rescale, k0.error, k0.population, *k0.first_order, k0.purity, k0.magnitude], G["r_scalars"], 1e-12)
    assert close(k0.derror, G["r_derror"], 1e-9) and close(k0.dpurity, G["r_dpurity"], 1e-9)
    assert close([k1.rescale, k1.error, k1.purity, k1.magnitude], G["c_scalars"], 1e-11)
    assert close(k1.derror, G["c_derror"], 1e-8) and close(k1.dpurity, G["c_dpurity"], 1e-8)
    p0 = k0.predict(G["Xq"], G["yq"].real, True)
    assert close(p0["pred"], G["r_pred"], 1e-12) and close(p0["cutoff"], G["r_cutoff"], 1e-9) and close(p0["derror"], G["r_vderr"], 1e-9)
    p1 = k1.predict(G["Xqc"], G["yqc"], True)
    assert close(p1["pred"], G["c_pred"], 1e-11) and close(p1["derror"], G["c_vderr"], 1e-8)
    pts = [np.column_stack([G[f"X{e}"], G[f"y{e}"].real, G[f"y{e}"].imag]) for e in range(3)]
    ev = oracle.evolve(1, pts[0], pts[1], pts[2], 2000.0, 2.0, k0, k1, k2)
    for e in range(3):
        assert close(ev[e], G[f"evolve_m1_e{e}"], 1e-9)


@pytest.mark.gpu
def test_gpu_matches_golden():
    from gaussian_process_liouville_equation_b200 import complex_kernel, dynamics, kernel

    k0 = kernel.TrainingKernel(G["theta_r"], (G["X0"], G["y0"]), True, True, True)
    k1 = complex_kernel.TrainingComplexKernel(G["theta_c"], (G["X1"], G["y1"]), True, True, False)
    k2 = kernel.TrainingKernel(G["theta_r"], (G["X2"], G["y2"]), True, True, False)
    got = [k0.get_rescale_factor(), k0.get_error(), k0.get_population(), *k0.get_1st_order_average(), k0.get_purity(), k0.get_magnitude()]
    assert np.allclose(got, G["r_scalars"], rtol=1e-9, atol=0)
    assert close(k0.get_error_derivative(), G["r_derror"], 1e-8)
    assert np.allclose([k1.get_rescale_factor(), k1.get_error(), k1.get_purity(), k1.get_magnitude()], G["c_scalars"], rtol=1e-8, atol=0)
    assert close(k0.get_inverse_times_label(), G["r_v"], 1e-8)
    assert close(k1.get_upper_part_of_augmented_inverse_times_label(), G["c_v"], 1e-8)
    p0 = kernel.PredictiveKernel(G["Xq"], k0, True, G["yq"].real)
    assert close(p0.get_prediction(), G["r_pred"], 1e-9) and p0.get_error() == pytest.approx(float(G["r_verr"]), rel=1e-9)
    assert close(p0.get_error_derivative(), G["r_vderr"], 1e-7)
    p1 = complex_kernel.PredictiveComplexKernel(G["Xqc"], k1, False, G["yqc"])
    assert close(p1.get_prediction(), G["c_pred"], 1e-8) and p1.get_error() == pytest.approx(float(G["c_verr"]), rel=1e-8)
    K, dK = kernel.kernel_matrix(G["X0"][:16], G["X0"][:16], G["theta_r"], True, True)
    assert (np.abs(K - G["K16"]) <= 1e-12 * np.abs(G["K16"])).all() and (np.abs(dK - G["dK16"]) <= 1e-12 * np.abs(G["dK16"]) + 1e-300).all()
    Kc, Ktc = complex_kernel.kernel_matrices(G["X1"][:16], G["X1"][:16], G["theta_c"], True)
    assert (np.abs(Kc - G["Kc16"]) <= 1e-12 * np.abs(G["Kc16"])).all() and (np.abs(Ktc - G["Ktc16"]) <= 1e-12 * np.abs(G["Ktc16"])).all()
    pts = [np.column_stack([G[f"X{e}"], G[f"y{e}"].real, G[f"y{e}"].imag]) for e in range(3)]
    for model in (0, 1, 2):
        ev = dynamics.evolve(model, pts, 2000.0, 2.0, [k0, k1, k2])
        for e in range(3):
            ref = G[f"evolve_m{model}_e{e}"]
            assert np.abs(ev[e][:, :2] - ref[:, :2]).max() <= 1e-13 * np.abs(ref[:, :2]).max()
            d = np.abs(ev[e][:, 2:] - ref[:, 2:])
            assert np.median(d) <= 1e-9 * np.abs(ref[:, 2:]).max() and d.max() <= 1e-5 * np.abs(ref[:, 2:]).max()
        E, F, D = dynamics.adiabatic_pes(model, np.linspace(-6, 6, 25))
        assert close(np.hstack([E, F, D[:, None]]), G[f"pes_m{model}"], 1e-12)
    assert close(dynamics.observable_sums(1, pts[2], 2000.0, 1), G["obs"], 1e-12)


G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gple_golden_v2.npz"))


def test_oracle_reproduces_golden_v2(oracle):
    """NLML objective and Metropolis chains (tests/golden/make_golden.py::main_v2)"""
    k0 = oracle.TrainingKernel(G["theta_r"], G["X0"], G["y0"], True, True, True)
    k1 = oracle.TrainingComplexKernel(G["theta_c"], G["X1"], G["y1"], True, True, True)
    k2 = oracle.TrainingKernel(G["theta_r"], G["X2"], G["y2"], True, True, False)
    v, g = k0.nlml(grad=True)
    assert v == pytest.approx(float(G2["r_nlml"]), rel=1e-12) and close(g, G2["r_dnlml"], 1e-9)
    v, g = k1.nlml(grad=True)
    assert v == pytest.approx(float(G2["c_nlml"]), rel=1e-11) and close(g, G2["c_dnlml"], 1e-8)
    pts, acc, chains = oracle.markov_chains(G2["mc_start"], 50, 0.5, 17, 5, 1, 0, analytic=list(G2["mc_analytic"]), want_chain=True)
    assert np.array_equal(acc, G2["mc_a_accept"]) and close(pts, G2["mc_a_pts"], 1e-14) and close(oracle.chain_autocorrelation(chains), G2["mc_a_autocor"], 1e-12)
    pts, acc, _ = oracle.markov_chains(G2["mc_p_start"], 20, 0.3, 17, 6, 0, 0, k00=k0, k10=k1, k11=k2)
    assert np.array_equal(acc, G2["mc_p_accept"]) and close(pts, G2["mc_p_pts"], 1e-12)
    assert np.array_equal(np.array([oracle.philox_draws(17, 5, c, s) for c in range(3) for s in range(3)]), G2["philox"])


@pytest.mark.gpu
def test_gpu_matches_golden_v2():
    from gaussian_process_liouville_equation_b200 import complex_kernel, kernel, mc

    k0 = kernel.TrainingKernel(G["theta_r"], (G["X0"], G["y0"]), True, True, True)
    k1 = complex_kernel.TrainingComplexKernel(G["theta_c"], (G["X1"], G["y1"]), True, True, True)
    k2 = kernel.TrainingKernel(G["theta_r"], (G["X2"], G["y2"]), True, True, False)
    v, g = k0.get_negative_log_marginal_likelihood(grad=True)
    assert v == pytest.approx(float(G2["r_nlml"]), rel=1e-9) and close(g, G2["r_dnlml"], 1e-7)
    v, g = k1.get_negative_log_marginal_likelihood(grad=True)
    assert v == pytest.approx(float(G2["c_nlml"]), rel=1e-9) and close(g, G2["c_dnlml"], 1e-7)
    a = G2["mc_analytic"]
    s = mc.Sampler(17, analytic=((a[0], a[1]), (a[2], a[3]), (a[4], a[5]), (a[6], a[7])))
    pts, acc, chains = s.chains(G2["mc_start"], 50, 0.5, 1, 0, want_chain=True, stream=5)
    assert np.array_equal(acc, G2["mc_a_accept"]) and close(pts, G2["mc_a_pts"], 1e-13) and close(s.autocorrelation(chains), G2["mc_a_autocor"], 1e-11)
    s = mc.Sampler(17, kernels=[k0, k1, k2])
    pts, acc, _ = s.chains(G2["mc_p_start"], 20, 0.3, 0, 0, stream=6)
    same = np.all(np.abs(pts[:, :2] - G2["mc_p_pts"][:, :2]) <= 1e-12 * np.abs(G2["mc_p_pts"][:, :2]).max(), axis=1)
    assert same.mean() >= 0.9 and np.array_equal(acc[same], G2["mc_p_accept"][same])
