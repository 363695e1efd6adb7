"""Generates tests/golden/gple_golden_v1.npz and gple_golden_v2.npz (NLML objective, Metropolis chains) from the CPU oracle on seeded
synthetic inputs.

The reference ships no golden vectors and cannot be built here, so these fixtures are
oracle outputs; they freeze the oracle (any later change to oracle/ must reproduce them) and give the GPU
tests a target that does not need the oracle library.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gaussian_process_liouville_equation_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402

THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])


def main():
    out = {}
    n, centre = 64, (-0.8, syn.P0)
    sets = [syn.training_set(40, e, n, centre) for e in range(3)]
    th = np.array([1.0, 0.9 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 2e-2])
    k0 = orc.TrainingKernel(th, *sets[0], True, True, True)
    k2 = orc.TrainingKernel(th, *sets[2], True, True, False)
    k1 = orc.TrainingComplexKernel(THETA_C, *sets[1], True, True, True)
    Xq, yq = syn.extra_points(40, 0, sets[0][0], 96, centre)
    p0 = k0.predict(Xq, yq.real, True)
    Xqc, yqc = syn.extra_points(40, 1, sets[1][0], 96, centre)
    p1 = k1.predict(Xqc, yqc, True)
    out.update(theta_r=th, theta_c=THETA_C, Xq=Xq, yq=yq, Xqc=Xqc, yqc=yqc)
    for e in range(3):
        out[f"X{e}"], out[f"y{e}"] = sets[e]
    out.update(r_scalars=np.array([k0.rescale, k0.error, k0.population, *k0.first_order, k0.purity, k0.magnitude]), r_derror=k0.derror,
               r_dpopulation=k0.dpopulation, r_dpurity=k0.dpurity, r_v=k0.v, r_pred=p0["pred"], r_var=p0["var"], r_cutoff=p0["cutoff"],
               r_verr=p0["error"], r_vderr=p0["derror"])
    out.update(c_scalars=np.array([k1.rescale, k1.error, k1.purity, k1.magnitude]), c_derror=k1.derror, c_dpurity=k1.dpurity, c_v=k1.v,
               c_pred=p1["pred"], c_var=p1["var"], c_cutoff=p1["cutoff"], c_verr=p1["error"], c_vderr=p1["derror"])
    K, dK = orc.kernel_real(sets[0][0][:16], sets[0][0][:16], th, True, True)
    Kc, Ktc = orc.kernel_complex(sets[1][0][:16], sets[1][0][:16], THETA_C, True, False)
    out.update(K16=K, dK16=dK, Kc16=Kc, Ktc16=Ktc)
    pts = [syn.points_aos(*s) for s in sets]
    for model in (0, 1, 2):
        ev = orc.evolve(model, pts[0], pts[1], pts[2], syn.MASS, 2.0, k0, k1, k2)
        for e in range(3):
            out[f"evolve_m{model}_e{e}"] = ev[e]
        E, F, D = orc.pes(model, np.linspace(-6, 6, 25))
        out[f"pes_m{model}"] = np.hstack([E, F, D[:, None]])
    out["obs"] = orc.observable_sums(1, pts[2], syn.MASS, 1)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gple_golden_v1.npz"), **out)
    print("wrote", len(out), "arrays")


def main_v2():
    """v2: the rows added after v1 -- NLML / LLT objective and Metropolis chains on the v1 element models."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gple_golden_v1.npz"))
    out = {}
    k0 = orc.TrainingKernel(g["theta_r"], g["X0"], g["y0"], True, True, True)
    k1 = orc.TrainingComplexKernel(g["theta_c"], g["X1"], g["y1"], True, True, True)
    k2 = orc.TrainingKernel(g["theta_r"], g["X2"], g["y2"], True, True, False)
    v, gr = k0.nlml(grad=True)
    out.update(r_nlml=np.array(v), r_dnlml=gr)
    v, gr = k1.nlml(grad=True)
    out.update(c_nlml=np.array(v), c_dnlml=gr)
    analytic = [syn.X0, syn.P0, syn.SIGMA_X, syn.SIGMA_P, 0.8, 0.6, 0.0, 0.4]
    rng = syn.rng(41, 0)
    start = np.zeros((32, 4))
    start[:, 0] = syn.X0 + syn.SIGMA_X * rng.standard_normal(32)
    start[:, 1] = syn.P0 + syn.SIGMA_P * rng.standard_normal(32)
    pts, acc, chains = orc.markov_chains(start, 50, 0.5, 17, 5, 1, 0, analytic=analytic, want_chain=True)
    out.update(mc_start=start, mc_analytic=np.array(analytic), mc_a_pts=pts, mc_a_accept=acc, mc_a_autocor=orc.chain_autocorrelation(chains))
    centre_pts = np.column_stack([g["X0"][:32], np.zeros((32, 2))])
    pts, acc, _ = orc.markov_chains(centre_pts, 20, 0.3, 17, 6, 0, 0, k00=k0, k10=k1, k11=k2)
    out.update(mc_p_start=centre_pts, mc_p_pts=pts, mc_p_accept=acc)
    out["philox"] = np.array([orc.philox_draws(17, 5, c, s) for c in range(3) for s in range(3)])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gple_golden_v2.npz"), **out)
    print("wrote", len(out), "arrays (v2)")


if __name__ == "__main__":
    if "--v2-only" not in sys.argv:
        main()
    main_v2()
