"""Factorisation schedule sweep: potrf + trtri (2 n^3 / 3 flops, event-timed inside the library) of a real element for
several switch-over sizes between the right-looking 128-block sweep and the recursive form (gple_set_potrf_flat).
Usage (GPU box):  python profiles/tune_potrf.py > gpurun_out/tune_potrf.md"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L
from gaussian_process_liouville_equation_b200 import synthetic as syn

SIZES = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [384, 1024, 2048, 4096, 8192, 16384]
FLATS = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [128, 512, 1024, 2048, 4096, 8192, 16384]
ctx = L.Context(0)
lib = ctx.lib
print("# potrf + trtri of one real element on 1 x B200: ms (TFLOP/s at 2 n^3 / 3) per switch-over size; 128 = fully recursive\n")
print("| n | " + " | ".join(f"flat <= {f}" for f in FLATS) + " |")
print("|---:|" + "---:|" * len(FLATS))
ctx.profile_enable(True)
for N in SIZES:
    scale = max(0.25, (2048.0 / N) ** 0.5)
    X, y = syn.training_set(70, 0, N, (0.0, syn.P0))
    yv = np.ascontiguousarray(y).view(np.float64)
    th = syn.theta_real(scale)
    cells, prev = [], None
    for f in FLATS:
        if prev is not None and prev >= N:
            cells.append("=")
            continue
        prev = f
        lib.gple_set_potrf_flat(f)
        best = 1e30
        for it in range(4):
            h, sc = C.c_void_p(), L.RealScalars()
            ctx.profile_read(2)
            ctx.check(lib.gple_train_real(ctx.h, L.addr(X), L.addr(yv), N, L.addr(th), 1, C.byref(h), C.byref(sc)))
            ms, _, flops = ctx.profile_read(2)
            lib.gple_model_destroy(ctx.h, h)
            if it > 0:
                best = min(best, ms)
        cells.append(f"{best:.2f} ({flops / best / 1e9:.1f})")
    print(f"| {N} | " + " | ".join(cells) + " |", flush=True)
