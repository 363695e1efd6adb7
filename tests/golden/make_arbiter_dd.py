"""Generates tests/golden/gple_arbiter_dd_v1.npz: the double-double (106-bit) evaluation of the reference's formulas
(tests/golden/arbiter_dd.cpp) at BASELINE.json's OWN sizes -- C2 real and complex element (N = 2048) and the smallest C5 size
(N = 4096, real) -- on exactly the inputs of gple_golden_ref_v1.npz (make_golden_ref.py), so that the three double-precision
evaluations (compiled reference, oracle, CUDA path) can each be measured against the true value where eps * cond(K) ~ 1e-8
makes them differ from one another (tests/test_arbiter.py, tests/test_gpu_baseline_sizes.py).

The program is first pinned against the 40-digit mpmath arbiter (gple_arbiter_v1.npz, N = 96, evaluated with the reference's
own P / Q formulas): agreement to double rounding is asserted before anything is written.
Run from the repo root (about ten minutes on 8 cores):  python tests/golden/make_arbiter_dd.py
"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from gaussian_process_liouville_equation_b200 import synthetic as syn  # noqa: E402
import make_golden_ref as mg  # noqa: E402

EXE = os.path.join(tempfile.gettempdir(), "gple_arbiter_dd")


def build():
    subprocess.check_call(["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-std=c++17", os.path.join(HERE, "arbiter_dd.cpp"), "-o", EXE])


def run(theta, X, y, Xq, exact_labels=False):
    """Returns dict(error, v, pred, var, cutoff); complex arrays for an 8-parameter theta."""
    cplx = len(theta) == 8
    N, Q = len(X), len(Xq)
    th = np.zeros(8)
    th[: len(theta)] = theta
    yy = np.zeros((N, 2))
    yy[:, 0], yy[:, 1] = np.real(y), np.imag(y) if cplx else 0.0
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            np.array([int(cplx) + 2 * int(exact_labels), N, Q], dtype=np.int64).tofile(f)
            th.tofile(f)
            np.ascontiguousarray(X, dtype=np.float64).tofile(f)
            yy.tofile(f)
            np.ascontiguousarray(Xq, dtype=np.float64).tofile(f)
        subprocess.check_call([EXE, fin, fout])
        o = np.fromfile(fout)
    err, o = o[0], o[1:]
    v, o = o[: 2 * N].reshape(N, 2), o[2 * N:]
    pred, o = o[: 2 * Q].reshape(Q, 2), o[2 * Q:]
    var, o = o[:Q], o[Q:]
    cut = o.reshape(Q, 2)
    if cplx:
        return dict(error=err, v=v[:, 0] + 1j * v[:, 1], pred=pred[:, 0] + 1j * pred[:, 1], var=var, cutoff=cut[:, 0] + 1j * cut[:, 1])
    return dict(error=err, v=v[:, 0].copy(), pred=pred[:, 0].copy(), var=var, cutoff=cut[:, 0].copy())


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max())


def pin_against_mpmath():
    A = np.load(os.path.join(HERE, "gple_arbiter_v1.npz"))
    r = run(A["theta_r"], A["X0"], A["y0"], A["Xq"], exact_labels=True)  # mpmath rescales the labels exactly
    c = run(A["theta_c"], A["X1"], A["y1"], A["Xqc"], exact_labels=True)
    worst = 0.0
    for got, want, scale in ((r["error"], A["r_error"], None), (r["v"], A["r_v"], None), (r["pred"], A["r_pred"], None), (r["var"], A["r_var"], A["theta_r"][0] ** 2), (r["cutoff"], A["r_cutoff"], None),
                             (c["error"], A["c_error"], None), (c["v"], A["c_v"], None), (c["pred"], A["c_pred"], None), (c["var"], A["c_var"], 2.0), (c["cutoff"], A["c_cutoff"], None)):
        d = float(np.abs(got - want).max() / (scale if scale is not None else np.abs(want).max()))
        worst = max(worst, d)
    print(f"double-double program vs 40-digit mpmath arbiter (N = 96, real + complex chain): max distance {worst:.2e}", flush=True)
    assert worst < 5e-15, worst  # both are rounded to double: agreement to rounding
    return worst


def main():
    build()
    out = dict(pin_distance_to_mpmath=pin_against_mpmath())
    for tag, config, n, element in (("c2r", 2, 2048, 0), ("c2c", 2, 2048, 1), ("c5r", 5, 4096, 0)):
        t = time.time()
        X, y = syn.training_set(config, element, n, mg.CENTRE)
        th = mg.THETA_C if element == 1 else mg.theta_real(n)
        r = run(th, X, y, mg.queries(config, element, X, 512))
        out.update({f"{tag}_theta": th, f"{tag}_error": r["error"], f"{tag}_v": r["v"], f"{tag}_pred": r["pred"], f"{tag}_var": r["var"], f"{tag}_cutoff": r["cutoff"]})
        print(f"{tag}: {time.time() - t:.0f} s", flush=True)
        np.savez_compressed(os.path.join(HERE, "gple_arbiter_dd_v1.npz"), **out)


if __name__ == "__main__":
    main()
