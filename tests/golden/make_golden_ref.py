"""Generates tests/golden/gple_golden_ref_v1.npz: outputs of the REFERENCE ITSELF (oracle/_ref = the unmodified gple/*.cpp
translation units compiled against oracle/refstub/) at BASELINE.json's own sizes, so that the GPU tests can be held to the
reference at N = 2048 / 4096 and over a 50-tick trajectory without paying its CPU cost on the GPU box (where /root/reference
does not exist).  Inputs are regenerated from seeds by the tests (gaussian_process_liouville_equation_b200/synthetic.py); only
hyper-parameters and outputs are stored.

  c2r_*  config C2, real element:    N = 2048, TrainingKernel (kernel.cpp:244-335) + PredictiveKernel on 512 queries (:481-522)
  c2c_*  config C2, complex element: N = 2048, TrainingComplexKernel (complex_kernel.cpp:221-377) + PredictiveComplexKernel (:594-646)
  c2e_*  one evolve step (evolve.cpp:377-423) of 128 points per element over the three N = 2048 models, SAC
  c5r_*  config C5 (smallest size), real element: N = 4096
  c1_*   config C1: 50 ticks of the deterministic tick loop (tests/trajectory.py; main.cpp:135-188 sequencing), SAC,
         N = 300 points per element, M = 1500 extra points per element, three elements populated, dt = 1

Run from the repo root in the container that has /root/reference (about half an hour on 8 cores):
    python tests/golden/make_golden_ref.py [c2r c2c c2e c5r c1]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajectory as tj  # noqa: E402
from gaussian_process_liouville_equation_b200 import synthetic as syn  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gple_golden_ref_v1.npz")
THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])
CENTRE = (0.0, syn.P0)  # the mid-crossing snapshot of SURVEY 8d


def theta_real(n):
    """characteristic lengths shrink with the point density so that cond(K) stays near the N = 2048 value"""
    return syn.theta_real((2048.0 / max(n, 2048)) ** 0.5)


def queries(config, element, X, q):
    """half jittered training points (gate mostly open), half pushed to the rim of the cloud (band / closed gate)"""
    Xq, _ = syn.extra_points(config, element, X, q, CENTRE)
    Xq[q // 2:, 0] += np.linspace(0.5, 3.5, q - q // 2) * syn.SIGMA_X
    return Xq


def c2r(out, n=2048, tag="c2r", config=2):
    X, y = syn.training_set(config, 0, n, CENTRE)
    th = theta_real(n)
    k = ref.TrainingKernel(th, X, y, True, True, False)
    p = k.predict(queries(config, 0, X, 512))
    out.update({f"{tag}_theta": th, f"{tag}_scalars": np.array([k.rescale, k.error, k.population, *k.first_order, k.purity, k.magnitude]), f"{tag}_v": k.v,
                f"{tag}_pred": p["pred"], f"{tag}_var": p["var"], f"{tag}_cutoff": p["cutoff"]})


def c5r(out):
    c2r(out, 4096, "c5r", 5)


def c2c(out):
    X, y = syn.training_set(2, 1, 2048, CENTRE)
    k = ref.TrainingComplexKernel(THETA_C, X, y, True, True, False)
    p = k.predict(queries(2, 1, X, 512))
    out.update(c2c_theta=THETA_C, c2c_scalars=np.array([k.rescale, k.error, k.purity, k.magnitude]), c2c_v=k.v, c2c_pred=p["pred"], c2c_var=p["var"], c2c_cutoff=p["cutoff"])


def c2e(out):
    sets = [syn.training_set(2, e, 2048, CENTRE) for e in range(3)]
    th = theta_real(2048)
    ks = [ref.TrainingKernel(th, *sets[0]), ref.TrainingComplexKernel(THETA_C, *sets[1]), ref.TrainingKernel(th, *sets[2])]
    pts = [syn.points_aos(*s)[:128] for s in sets]
    ev = ref.evolve(0, pts[0], pts[1], pts[2], syn.MASS, 1.0, *ks)
    for e in range(3):
        out[f"c2e_e{e}"] = ev[e]


def c1(out):
    thetas = [syn.theta_real(), THETA_C, syn.theta_real()]
    d0, e0 = tj.initial_state(1, 300, 1500, CENTRE)
    hist = {}

    def record(tick, density, kernels):
        if tick in (1, 10, 25):
            hist[tick] = [d.copy() for d in density]

    d, e, obs = tj.run(tj.CpuBackend(ref, "ref"), 0, thetas, d0, e0, syn.MASS, 1.0, 50, record)
    names, vals = tj.flatten(obs)
    out.update(c1_obs_names=np.array(names), c1_obs=vals, c1_theta_c=THETA_C)
    for i in range(3):
        out[f"c1_density_e{i}"], out[f"c1_extra_e{i}"] = d[i], e[i]
        for t, h in hist.items():
            out[f"c1_density_t{t}_e{i}"] = h[i]


def main():
    assert ref.build(), "needs /root/reference (oracle/_ref)"
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name in sys.argv[1:] or ["c2r", "c2c", "c2e", "c5r", "c1"]:
        t = time.time()
        globals()[name](out)
        print(f"{name}: {time.time() - t:.0f} s", flush=True)
        np.savez_compressed(OUT, **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
