"""Factorisation as a CUDA graph (GPLE_OPT_FACTORISE_GRAPHS) against direct launches: potrf + trtri event time inside the library
and wall-clock per gple_train_real call (error + averages), one real element, 30 back-to-back calls after 3 warm-up calls.
Usage (GPU box):  python profiles/factorise_graphs.py > gpurun_out/factorise_graphs.md"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L
from gaussian_process_liouville_equation_b200 import synthetic as syn

SIZES = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [300, 1024, 2048, 4096]
ctx = L.Context(0)
lib = ctx.lib
print("| N | direct launches: potrf + trtri ms | train call ms (wall) | CUDA graph: potrf + trtri ms | train call ms (wall) | error identical |")
print("|---:|---:|---:|---:|---:|---|")
for N in SIZES:
    scale = max(0.25, (2048.0 / N) ** 0.5)
    X, y = syn.training_set(70, 0, N, (0.0, syn.P0))
    yv = np.ascontiguousarray(y).view(np.float64)
    th = syn.theta_real(scale)
    res = []
    for graphs in (False, True):
        ctx.set_factorise_graphs(graphs)
        ctx.profile_enable(True)
        best, wall, err = 1e30, [], None
        for it in range(33):
            h, sc = C.c_void_p(), L.RealScalars()
            ctx.profile_read(2)
            t0 = time.perf_counter()
            ctx.check(lib.gple_train_real(ctx.h, L.addr(X), L.addr(yv), N, L.addr(th), 3, C.byref(h), C.byref(sc)))
            t1 = time.perf_counter()
            ms, _, _ = ctx.profile_read(2)
            lib.gple_model_destroy(ctx.h, h)
            if it >= 3:
                best = min(best, ms)
                wall.append((t1 - t0) * 1e3)
            err = sc.error
        ctx.profile_enable(False)
        res.append((best, float(np.median(wall)), err))
    print(f"| {N} | {res[0][0]:.3f} | {res[0][1]:.3f} | {res[1][0]:.3f} | {res[1][1]:.3f} | {res[0][2] == res[1][2]} |", flush=True)
