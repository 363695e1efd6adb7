"""North-star measurement (BASELINE.json: "a full GPR-MQCLE time step at N = 16k training points and 10^6 MC points
... as a fraction of FP64 peak on 1 B200"): one time step = TrainingKernels rebuild + evolve of Q points per populated
element, through the C-ABI with device-resident inputs, timed with CUDA events on the library's stream.

Two states are measured:
  * "rho00": the t = 0 state of the reference (only rho00 populated, gple/main.cpp:57-64): train 1 real element,
    evolve Q points (2 queries per point into the rho00 model);
  * "3el":   all three elements populated (mid-crossing snapshot): train 2 real + 1 complex element, evolve Q3 points
    per element (8 queries per point and target element).
FP64 fraction = executed flops of the tensor-core kernels (factorise 2n^3/3, variance GEMM rows*n*(n+128), as counted
by gple_profile_*) / wall time of the whole step / measured DMMA peak (with the gate on this is the rate of the
EXECUTED flops: the gate removes work, it does not speed the kernels up) -- i.e. everything that is not a DMMA flop
(kernel build, mean, reductions, launch gaps) counts against it.  Reported with the bound-gated variance on (product
default) and off (every variance computed, the reference's amount of work).

Usage (GPU box):  python profiles/north_star_step.py [N] [Q] [Q3] > gpurun_out/north_star.md"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L
from gaussian_process_liouville_equation_b200 import synthetic as syn

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
Q3 = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000
THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 1e-2])
dev = torch.device("cuda", 0)
ctx = L.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
lib = ctx.lib
dmma, dfma = ctx.fp64_peak()
centre = (0.0, syn.P0)
scale = max(0.25, (2048.0 / N) ** 0.5)  # lengths shrink with N: comparable conditioning (allowed by opt.cpp:1036-1040)
thetas = [syn.theta_real(scale), THETA_C * np.array([1, 1, scale, scale, 1, scale, scale, 1]), syn.theta_real(scale)]

sets = [syn.training_set(90, e, N, centre) for e in range(3)]
d_X = [torch.from_numpy(s[0]).to(dev) for s in sets]
d_y = [torch.from_numpy(np.ascontiguousarray(s[1]).view(np.float64)).to(dev) for s in sets]


def points(q):
    out = []
    for e in range(3):
        Xe, ye = syn.extra_points(90, e, sets[e][0], q, centre)
        out.append(torch.from_numpy(syn.points_aos(Xe, ye)).to(dev))
    return out


def train(e):
    h = C.c_void_p()
    th = np.ascontiguousarray(thetas[e])
    if e == 1:
        s = L.ComplexScalars()
        ctx.check(lib.gple_train_complex(ctx.h, d_X[e].data_ptr(), d_y[e].data_ptr(), N, L.addr(th), L.CALC_ERROR | L.CALC_AVERAGE, C.byref(h), C.byref(s)))
    else:
        s = L.RealScalars()
        ctx.check(lib.gple_train_real(ctx.h, d_X[e].data_ptr(), d_y[e].data_ptr(), N, L.addr(th), L.CALC_ERROR | L.CALC_AVERAGE, C.byref(h), C.byref(s)))
    return h


def step(elements, pts0, q):
    pts = [p.clone() for p in pts0]
    torch.cuda.synchronize()
    for s in range(4):
        ctx.profile_read(s)
    ctx.gate_statistics()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    models = [train(e) if e in elements else None for e in range(3)]
    cnt = [q if e in elements else 0 for e in range(3)]
    ctx.check(lib.gple_evolve(ctx.h, 1, models[0], models[1], models[2], pts[0].data_ptr(), cnt[0], pts[1].data_ptr(), cnt[1], pts[2].data_ptr(), cnt[2], syn.MASS, syn.DT))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    for h in models:
        if h is not None:
            lib.gple_model_destroy(ctx.h, h)
    prof = [ctx.profile_read(s) for s in range(4)]
    return ms, prof, ctx.gate_statistics()


print(f"# North-star step on 1 x B200: N = {N} training points per element\n")
print(f"Measured FP64 peaks: DMMA {dmma:.2f} TFLOP/s, DFMA {dfma:.2f} TFLOP/s.  Times: CUDA events around the whole step (train + evolve), inputs resident in HBM.\n")
print("| state | evolved points / element | gated variance | step ms | steps/s | executed DMMA TFLOP | whole-step TFLOP/s | frac of DMMA peak | reference-formulation TFLOP/s | variance GEMM TFLOP/s (frac) | factorise ms (TFLOP/s) | K* build ms | mean ms | rows through stage A | rows needing the full variance |")
print("|---|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
ctx.profile_enable(True)
for name, elements, q in (("rho00", (0,), Q), ("3el", (0, 1, 2), Q3)):
    pts0 = points(q)
    # reference-formulation flops
    if name == "rho00":
        ref_flops = float(N) ** 3 + 4.0 * q * float(N) ** 2
    else:
        ref_flops = (26.0 + 1.0 / 3.0) * float(N) ** 3 + 160.0 * q * float(N) ** 2
    for gated in (True, False):
        ctx.set_gated_variance(gated)
        step(elements, pts0, q if gated else min(q, 20000))  # warm-up at the steady-state buffer sizes (the gated schedule owns the per-query buffers)
        ms, prof, gs = step(elements, pts0, q)
        flops = prof[0][2] + prof[2][2]
        frac_rows = gs[1] / gs[0] if gated and gs[0] else 1.0
        frac_full = gs[3] / gs[0] if gated and gs[0] else 1.0
        print(f"| {name} | {q} | {'on' if gated else 'off'} | {ms:.1f} | {1000.0 / ms:.4f} | {flops / 1e12:.1f} | {flops / ms / 1e9:.2f} | {flops / ms / 1e9 / dmma:.3f} | "
              f"{ref_flops / ms / 1e9:.1f} | {prof[0][2] / max(prof[0][0], 1e-9) / 1e9:.2f} ({prof[0][2] / max(prof[0][0], 1e-9) / 1e9 / dmma:.3f}) | "
              f"{prof[2][0]:.1f} ({prof[2][2] / max(prof[2][0], 1e-9) / 1e9:.2f}) | {prof[1][0]:.1f} | {prof[3][0]:.1f} | {frac_rows:.3f} | {frac_full:.3f} |", flush=True)
    del pts0
ctx.set_gated_variance(True)
ctx.profile_enable(False)
