"""C3 (BASELINE.json configs[2]): one full Optimization.optimize() on Tully's DAC with all three elements populated --
Nelder-Mead per element, SLSQP with the population / energy constraints on the diagonal, then all elements with the
purity constraint, plus the LocalInitial / Global restarts when the averages stay out of tolerance (opt.cpp:1019-1392).
Every loss, constraint and gradient evaluation runs on the GPU through the C-ABI; the driver itself is the scipy-based
mirror (NLopt is not installed).  Reports wall time, evaluation counts and evaluations / s.

Usage (GPU box):  python profiles/c3_optimize.py [N ...] > gpurun_out/c3_optimize.md"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L
from gaussian_process_liouville_equation_b200 import dynamics, opt
from gaussian_process_liouville_equation_b200 import synthetic as syn

SIZES = [int(a) for a in sys.argv[1:]] or [300, 1024]
DAC = 1
ctx = L.default_context()
print("# C3: full Optimization.optimize() on 1 x B200 (DAC, three elements, M = 5N extra points per element)\n")
print("| N | wall s | loose_function evaluations | of which with gradient | constraint evaluations | evaluations / s | kernel launches | result type | final loss | |population-1|, |E/E0-1|, |purity/p0-1| beyond tolerance |")
print("|---:|---:|---:|---:|---:|---:|---:|---|---:|---|")
for N in SIZES:
    centre = (0.0, syn.P0)
    density, extra = [], []
    for e in range(3):
        X, y = syn.training_set(35, e, N, centre)
        Xe, ye = syn.extra_points(35, e, X, 5 * N, centre)
        density.append(syn.points_aos(X, y))
        extra.append(syn.points_aos(Xe, ye))
    pops = [dynamics.observable_sums(DAC, density[e], syn.MASS, i) for i, e in enumerate((0, 2))]
    # the synthetic snapshot carries populations 0.6 / 0.4 by construction; total energy from the same labels
    e_tot = 0.6 * pops[0][7] / pops[0][0] + 0.4 * pops[1][7] / pops[1][0]
    counts = {"loose": 0, "grad": 0, "con": 0}
    lf, tk = dynamics.loose_function, opt.pr.TrainingKernels

    def counted_lf(x, ts, ets, grad=False):
        counts["loose"] += 1
        counts["grad"] += int(bool(grad))
        return lf(x, ts, ets, grad=grad)

    def counted_tk(*a, **k):
        counts["con"] += 1
        return tk(*a, **k)

    dynamics.loose_function = counted_lf
    opt.pr.TrainingKernels = counted_tk
    optimizer = opt.Optimization((syn.SIGMA_X, syn.SIGMA_P), syn.MASS, DAC, InitialTotalEnergy=e_tot, InitialPurity=syn.snapshot_purity(), max_global_evals=300)
    launches0 = ctx.launches
    ctx.sync()
    t = time.perf_counter()
    (err, steps, typ), check = optimizer.optimize(density, extra)
    ctx.sync()
    wall = time.perf_counter() - t
    dynamics.loose_function, opt.pr.TrainingKernels = lf, tk
    total = counts["loose"] + counts["con"]
    print(f"| {N} | {wall:.2f} | {counts['loose']} | {counts['grad']} | {counts['con']} | {total / wall:.1f} | {ctx.launches - launches0} | {typ} | {err:.4g} | {np.array2string(check, precision=3)} |", flush=True)
