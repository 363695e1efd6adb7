"""C3 (BASELINE.json configs[2]): cost of the optimiser's inner loop on one B200 -- loose_function (opt.cpp:441-482:
TrainingKernel on N points + PredictiveKernel on M = 5N extra points) with and without gradient, real and complex
element, inputs passed as HOST arrays exactly as the NLopt callbacks pass them (so every evaluation pays its own H2D
copies and scalar read-backs).  Wall-clock per evaluation over `reps` back-to-back calls after warm-up.

Usage (GPU box):  python profiles/c3_loss_evaluations.py [reps] > gpurun_out/c3_loss_evaluations.md"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L
from gaussian_process_liouville_equation_b200 import synthetic as syn

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 20
SIZES = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [300, 1024, 2048, 4096]
THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 1e-2])
ctx = L.Context(0)
lib = ctx.lib
centre = (0.0, syn.P0)
print("# C3: loose_function evaluations on 1 x B200 (DAC-sized element models, M = 5N validation points, host arrays in)\n")
print("| N | element | value only: ms / eval | evals/s | value + gradient: ms / eval | evals/s | of which potrf+trtri ms | train only ms (err+avg) | train + all gradients ms |")
print("|---:|---|---:|---:|---:|---:|---:|---:|---:|")
for N in SIZES:
    for kind in ("real", "complex"):
        e = 0 if kind == "real" else 1
        X, y = syn.training_set(33, e, N, centre)
        Xe, ye = syn.extra_points(33, e, X, 5 * N, centre)
        yv = np.ascontiguousarray(y).view(np.float64)
        yev = np.ascontiguousarray(ye).view(np.float64)
        th = np.ascontiguousarray(syn.theta_real() if kind == "real" else THETA_C)
        npar = len(th)
        val = C.c_double()
        grad = np.empty(npar)

        def loose(with_grad):
            ctx.check(lib.gple_loose_function(ctx.h, L.addr(th), npar, L.addr(grad) if with_grad else None, L.addr(X), L.addr(yv), N, L.addr(Xe), L.addr(yev), 5 * N, C.byref(val)))

        def train(flags):
            h = C.c_void_p()
            if kind == "real":
                s = L.RealScalars()
                ctx.check(lib.gple_train_real(ctx.h, L.addr(X), L.addr(yv), N, L.addr(th), flags, C.byref(h), C.byref(s)))
            else:
                s = L.ComplexScalars()
                ctx.check(lib.gple_train_complex(ctx.h, L.addr(X), L.addr(yv), N, L.addr(th), flags, C.byref(h), C.byref(s)))
            lib.gple_model_destroy(ctx.h, h)

        def per_call(fn, reps=REPS):
            fn()
            fn()
            ctx.sync()
            t = time.perf_counter()
            for _ in range(reps):
                fn()
            ctx.sync()
            return (time.perf_counter() - t) * 1e3 / reps

        v_ms = per_call(lambda: loose(False))
        ctx.profile_enable(True)
        ctx.profile_read(2)
        g_ms = per_call(lambda: loose(True))
        fa_ms, _, _ = ctx.profile_read(2)
        ctx.profile_enable(False)
        t_ms = per_call(lambda: train(L.CALC_ERROR | L.CALC_AVERAGE))
        td_ms = per_call(lambda: train(L.CALC_ERROR | L.CALC_AVERAGE | L.CALC_DERIVATIVE))
        print(f"| {N} | {kind} | {v_ms:.2f} | {1e3 / v_ms:.1f} | {g_ms:.2f} | {1e3 / g_ms:.1f} | {fa_ms / (REPS + 2):.2f} | {t_ms:.2f} | {td_ms:.2f} |", flush=True)
