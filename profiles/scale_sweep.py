"""C5 scale sweep (BASELINE.json configs[4]): kernel build, Cholesky + inverse, prediction at N = 4096 / 8192 / 16384
training points on one B200, each phase timed with CUDA events inside the library (gple_profile_*), plus a whole
time step at reduced Q.  Usage (GPU box):  python profiles/scale_sweep.py [Q] > gpurun_out/scale_sweep.md"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L
from gaussian_process_liouville_equation_b200 import synthetic as syn

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
SIZES = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2048, 4096, 8192, 16384]
THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 1e-2])
ctx = L.Context(0)
lib = ctx.lib
dmma, dfma = ctx.fp64_peak()
print(f"# C5 scale sweep on 1 x B200 (Q = {Q} evolved points / element)\n")
print(f"Measured FP64 peaks: DMMA {dmma:.2f} TFLOP/s, DFMA {dfma:.2f} TFLOP/s; HBM copy peak 6537 GB/s (MEASURED_PEAKS.json).\n")
print("| N | element | train ms | potrf+trtri ms | TFLOP/s (2n^3/3) | K* build GB/s | variance GEMM TFLOP/s | % DMMA peak | predict ms (rows) |")
print("|---:|---|---:|---:|---:|---:|---:|---:|---:|")


def timed(fn):
    ctx.sync()
    t = time.perf_counter()
    r = fn()
    ctx.sync()
    return r, (time.perf_counter() - t) * 1e3


for N in SIZES:
    centre = (0.0, syn.P0)
    # lengths shrink with N so that the covariance keeps a comparable conditioning (bounds of opt.cpp:1036-1040 allow it)
    scale = max(0.25, (2048.0 / N) ** 0.5)
    for kind in ("real", "complex"):
        e = 0 if kind == "real" else 1
        X, y = syn.training_set(70, e, N, centre)
        th = syn.theta_real(scale) if kind == "real" else THETA_C * np.array([1, 1, scale, scale, 1, scale, scale, 1])
        Xq, _ = syn.extra_points(70, e, X, Q, centre)
        yv = np.ascontiguousarray(y).view(np.float64)
        h = C.c_void_p()
        ctx.profile_enable(True)
        for s in range(3):
            ctx.profile_read(s)

        def train():
            if kind == "real":
                sc = L.RealScalars()
                rc = lib.gple_train_real(ctx.h, L.addr(X), L.addr(yv), N, L.addr(th), 3, C.byref(h), C.byref(sc))
            else:
                sc = L.ComplexScalars()
                rc = lib.gple_train_complex(ctx.h, L.addr(X), L.addr(yv), N, L.addr(th), 3, C.byref(h), C.byref(sc))
            return rc, sc

        (rc, sc), _ = timed(train)  # warm-up (allocations)
        lib.gple_model_destroy(ctx.h, h)
        for s in range(3):
            ctx.profile_read(s)
        (rc, sc), train_ms = timed(train)
        fa_ms, _, fa_flops = ctx.profile_read(2)
        w = 2 if kind == "complex" else 1
        pred, var, cut = np.empty(w * Q), np.empty(Q), np.empty(w * Q)
        fn = lib.gple_predict_complex if kind == "complex" else lib.gple_predict_real
        _, pred_ms = timed(lambda: fn(ctx.h, h, L.addr(Xq), Q, None, L.addr(pred), L.addr(var), L.addr(cut), None, None))
        v_ms, v_n, v_flops = ctx.profile_read(0)
        k_ms, k_n, k_bytes = ctx.profile_read(1)
        lib.gple_model_destroy(ctx.h, h)
        ctx.profile_enable(False)
        status = "ok" if rc == 0 else f"rc={rc}"
        print(f"| {N} | {kind} ({status}) | {train_ms:.1f} | {fa_ms:.1f} | {fa_flops / fa_ms / 1e9:.2f} | {k_bytes / k_ms / 1e6:.0f} | {v_flops / v_ms / 1e9:.2f} | "
              f"{100 * v_flops / v_ms / 1e9 / dmma:.1f} | {pred_ms:.1f} ({w * Q}) |", flush=True)
