"""High-precision arbiter (tests/golden/gple_arbiter_v1.npz, 40-digit mpmath evaluation of the reference's formulas, made
by tests/golden/make_arbiter.py): distance of every double-precision evaluation -- the compiled reference (oracle/_ref),
the oracle restatement, the CUDA path -- to the exact value of the reference formulation, for the quantities whose
double-precision evaluation is ill-conditioned: K^-1 y, LOOCV error, predictive variance, the cutoff prediction inside the
cubic band of the gate, and the whole complex chain (P, Q, v, error, prediction, variance, cutoff).

north_star's bar is 1e-9 relative for predictions and losses.  With the exact value at hand it is applied to ALL THREE
evaluations, including the complex element (whose GPU-vs-oracle tests at larger N use 1e-8 because there both sides carry
eps * cond(K) and no exact value is available).  The variance is compared relative to the prior variance k** (it is a
difference of two numbers of that size; at the coincident query it is exactly 0).
"""
import os

import numpy as np
import pytest

A = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gple_arbiter_v1.npz"))
PRIOR_R = A["theta_r"][0] ** 2 * (1 + A["theta_r"][3] ** 2)
PRIOR_C = A["theta_c"][0] ** 2 * (A["theta_c"][1] ** 2 + A["theta_c"][4] ** 2 + A["theta_c"][7] ** 2)


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max())


def distances_real(error, population, purity, v, pred, var, cutoff):
    return dict(error=rel(error, A["r_error"]), population=rel(population, A["r_population"]), purity=rel(purity, A["r_purity"]), v=rel(v, A["r_v"]),
                pred=rel(pred, A["r_pred"]), var=float(np.abs(var - A["r_var"]).max() / PRIOR_R), cutoff=rel(cutoff, A["r_cutoff"]))


def distances_complex(error, v, P, Q, pred, var, cutoff):
    d = dict(error=rel(error, A["c_error"]), v=rel(v, A["c_v"]), pred=rel(pred, A["c_pred"]), var=float(np.abs(var - A["c_var"]).max() / PRIOR_C), cutoff=rel(cutoff, A["c_cutoff"]))
    if P is not None:
        d.update(P=rel(P, A["c_P"]), Q=rel(Q, A["c_Q"]))
    return d


def test_fixture_exercises_the_cubic_band_and_a_coincidence():
    s = 10.0 / np.abs(A["y0"].real).max()
    gate = np.divide(A["r_cutoff"] * s, A["r_pred"], out=np.zeros_like(A["r_pred"]), where=A["r_pred"] != 0)
    assert ((gate > 1e-3) & (gate < 1 - 1e-3)).sum() >= 10 and (gate == 0).any() and (np.abs(gate - 1) < 1e-15).any()
    assert abs(A["r_var"][5]) < 1e-30 * PRIOR_R  # the coincident query: k** - k K^-1 k^T vanishes identically


@pytest.mark.parametrize("which", ["oracle", "ref"])
def test_cpu_evaluations_against_the_exact_values(oracle, which):
    from oracle import ref

    if which == "ref" and not ref.build():
        pytest.skip("oracle/_ref not available")
    m = oracle if which == "oracle" else ref
    k = m.TrainingKernel(A["theta_r"], A["X0"], A["y0"], True, True, False)
    p = k.predict(A["Xq"])
    d = distances_real(k.error, k.population, k.purity, k.v, p["pred"], p["var"], p["cutoff"])
    assert max(d.values()) < 1e-9, d
    assert d["population"] < 1e-13 and d["purity"] < 1e-13 and d["pred"] < 1e-13, d
    c = m.TrainingComplexKernel(A["theta_c"], A["X1"], A["y1"], True, True, False)
    pc = c.predict(A["Xqc"])
    dc = distances_complex(c.error, c.v, c.P, c.Q, pc["pred"], pc["var"], pc["cutoff"])
    assert max(dc.values()) < 1e-9, dc


@pytest.mark.gpu
def test_gpu_against_the_exact_values(oracle):
    """The CUDA path meets north_star's 1e-9 against the EXACT values, real and complex element alike, and is not further from
    them than the reference formulation evaluated in double precision (the oracle) by more than a factor that rounding explains."""
    from gaussian_process_liouville_equation_b200 import complex_kernel, kernel

    k = kernel.TrainingKernel(A["theta_r"], (A["X0"], A["y0"]), True, True, False)
    p = kernel.PredictiveKernel(A["Xq"], k)
    d = distances_real(k.get_error(), k.get_population(), k.get_purity(), k.get_inverse_times_label(), p.get_prediction(), p.get_variance(), p.get_cutoff_prediction())
    assert max(d.values()) < 1e-9, d
    c = complex_kernel.TrainingComplexKernel(A["theta_c"], (A["X1"], A["y1"]), True, True, False)
    pc = complex_kernel.PredictiveComplexKernel(A["Xqc"], c)
    dc = distances_complex(c.get_error(), c.get_upper_part_of_augmented_inverse_times_label(), c.get_upper_left_block_of_augmented_inverse(),
                           c.get_lower_left_block_of_augmented_inverse(), pc.get_prediction(), pc.get_variance(), pc.get_cutoff_prediction())
    assert max(dc.values()) < 1e-9, dc
    # the same distances for the oracle, for the record printed with -s and for the factor check
    ko = oracle.TrainingKernel(A["theta_r"], A["X0"], A["y0"], True, True, False)
    po = ko.predict(A["Xq"])
    do = distances_real(ko.error, ko.population, ko.purity, ko.v, po["pred"], po["var"], po["cutoff"])
    co = oracle.TrainingComplexKernel(A["theta_c"], A["X1"], A["y1"], True, True, False)
    pco = co.predict(A["Xqc"])
    dco = distances_complex(co.error, co.v, co.P, co.Q, pco["pred"], pco["var"], pco["cutoff"])
    print("distance to the exact values  (gpu | oracle)")
    for name, g, o in [("real " + n, d[n], do[n]) for n in d] + [("cplx " + n, dc[n], dco[n]) for n in dc]:
        print(f"  {name:16s} {g:9.2e} | {o:9.2e}")
        assert g <= max(50.0 * o, 1e-11), (name, g, o)


def test_compiled_reference_against_the_exact_values_at_baseline_sizes():
    """The reference's own outputs at N = 2048 / 4096 (gple_golden_ref_v1.npz, oracle/_ref) against the double-double arbiter
    (gple_arbiter_dd_v1.npz, tests/golden/arbiter_dd.cpp, itself pinned to the mpmath arbiter at 2e-16): how far a
    double-precision evaluation of the reference formulation is from its exact value at BASELINE.json's sizes.  The GPU
    counterpart is tests/test_gpu_baseline_sizes.py."""
    here = os.path.dirname(os.path.abspath(__file__))
    R = np.load(os.path.join(here, "golden", "gple_golden_ref_v1.npz"))
    D = np.load(os.path.join(here, "golden", "gple_arbiter_dd_v1.npz"))
    assert float(D["pin_distance_to_mpmath"]) < 5e-15
    for tag in ("c2r", "c5r", "c2c"):
        th = D[f"{tag}_theta"]
        prior = th[0] ** 2 * (1 + th[3] ** 2) if len(th) == 4 else th[0] ** 2 * (th[1] ** 2 + th[4] ** 2 + th[7] ** 2)
        d = dict(error=abs(R[f"{tag}_scalars"][1] / D[f"{tag}_error"] - 1), v=rel(R[f"{tag}_v"], D[f"{tag}_v"]), pred=rel(R[f"{tag}_pred"], D[f"{tag}_pred"]),
                 var=float(np.abs(R[f"{tag}_var"] - D[f"{tag}_var"]).max() / prior), cutoff=rel(R[f"{tag}_cutoff"], D[f"{tag}_cutoff"]))
        print(tag, d)
        assert max(d["error"], d["pred"], d["var"], d["cutoff"]) <= 1e-9 and d["v"] <= 5e-9, (tag, d)
