"""Generates tests/golden/gple_golden_v1.npz from the CPU oracle on seeded synthetic inputs.

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


if __name__ == "__main__":
    main()
