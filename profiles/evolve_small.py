"""Latency of one evolve() call at main-loop sizes (N training points per element, Q evolved points per element, DAC, three
elements), gated and ungated variance.  Usage (GPU box):  python profiles/evolve_small.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L
from gaussian_process_liouville_equation_b200 import complex_kernel, dynamics, kernel
from gaussian_process_liouville_equation_b200 import synthetic as syn

THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])
ctx = L.default_context()
print("| N | Q per element | gated: ms / evolve | launches | every variance: ms / evolve | launches |")
print("|---:|---:|---:|---:|---:|---:|")
for N in (300, 1024):
    centre = (-0.3, syn.P0)
    sets = [syn.training_set(31, e, N, centre) for e in range(3)]
    g = [kernel.TrainingKernel(syn.theta_real(), sets[0]), complex_kernel.TrainingComplexKernel(THETA_C, sets[1]), kernel.TrainingKernel(syn.theta_real(), sets[2])]
    for Q in (N, 5 * N):
        pts = []
        for e in range(3):
            Xe, ye = syn.extra_points(31, e, sets[e][0], Q, centre)
            pts.append(syn.points_aos(Xe, ye))
        row = []
        for gated in (True, False):
            ctx.set_gated_variance(gated)
            for _ in range(3):
                dynamics.evolve(1, pts, syn.MASS, 1.0, g)
            l0 = ctx.launches
            t0 = time.perf_counter()
            for _ in range(20):
                dynamics.evolve(1, pts, syn.MASS, 1.0, g)
            row.append(((time.perf_counter() - t0) / 20 * 1e3, (ctx.launches - l0) // 20))
        ctx.set_gated_variance(True)
        print(f"| {N} | {Q} | {row[0][0]:.2f} | {row[0][1]} | {row[1][0]:.2f} | {row[1][1]} |", flush=True)
