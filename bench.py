#!/usr/bin/env python
"""bench.py -- one GPR-MQCLE time step per "step" on synthetic Tully-model inputs (BASELINE.json configs[1]).

Workload C2: Tully simple avoided crossing, N = 2048 training points per density-matrix element,
all three elements (rho00, rho10, rho11) populated (mid-crossing snapshot), Q = 1e5 evolved Monte-Carlo
phase-space points per element per step.  One step = what gple/main.cpp:135-188 does on a tick without
re-optimisation:
    TrainingKernels(params, density)   (kernel build + factorise + inverse + LOOCV error + averages, 3 elements)
  + evolve(points)                     (forward move, 9 back-propagated GPR predictions per point, recombination)
With 3 elements every element predictor answers 8 Q queries per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 (torchrun, one rank per GPU): the evolved points are block-partitioned over the ranks (strong scaling,
total work fixed), every rank rebuilds the three element models redundantly, and the only collective is the
NCCL all-gather of the evolved point sets at the end of the step (BASELINE.json north_star).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gaussian_process_liouville_equation_b200 import synthetic as syn  # noqa: E402

N_TRAIN = 2048
Q_POINTS = 100_000
PES_MODEL = 0  # SAC
CENTRE = (0.0, syn.P0)
THETA_R = syn.theta_real()
THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, syn.INITIAL_NOISE])
METRIC = "gpr_mqcle_time_steps_per_s"
UNIT = "steps/s"
WORKLOAD = ("C2: Tully SAC, N=2048 training points/element, 3 elements (rho00,rho10,rho11), Q=1e5 evolved MC points/element/step; "
            "step = TrainingKernels rebuild (3 elements) + evolve (8Q GPR predictions per element)")


def parse_schedule(text):
    """'1,5,10' or '1:0,5:1,10:2' -> [(re_end, im_end), ...] (cumulative 128-blocks)."""
    out = []
    for item in text.split(","):
        if item.strip():
            a, _, b = item.partition(":")
            out.append((int(a), int(b) if b else 0))
    return out


def make_inputs():
    sets = [syn.training_set(2, e, N_TRAIN, CENTRE) for e in range(3)]
    pts = []
    for e in range(3):
        Xe, ye = syn.extra_points(2, e, sets[e][0], Q_POINTS, CENTRE)
        pts.append(syn.points_aos(Xe, ye))
    return sets, pts


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation timed on the host cores, bounded sample
# ----------------------------------------------------------------------------------------------------------------
def cpu_backend():
    """oracle/_ref (the reference's unmodified translation units, compiled in the build container and shipped with the
    snapshot) when present, else the oracle restatement.  Returns (module, kind)."""
    try:
        from oracle import ref

        if ref.available():
            return ref, "reference"
    except Exception:
        pass
    from oracle import oracle as orc

    return orc, "port"


class CpuStep:
    """The bench step on the CPU, through the reference's own functions, as a bounded sample.

    Full-size and timed ONCE: the TrainingKernels rebuild of the step (TrainingKernel x 2 and TrainingComplexKernel at the
    workload's N; kernel.cpp:244-335, complex_kernel.cpp:221-377).  Timed per sample(): evolve() (evolve.cpp:377-423) of
    `points` phase-space points per element over those three models -- the same 9 back-propagated predictions per point the
    GPU step makes -- which is EXTRAPOLATED linearly to the Q points per element of the workload (points are independent:
    evolve.cpp:392-420 is a par_unseq loop over them)."""

    def __init__(self, sets, points=16):
        from oracle import oracle as orc

        self.mod, self.kind = cpu_backend()
        self.cores = orc.num_threads()
        self.points = points
        m = self.mod
        t = time.perf_counter()
        k0 = m.TrainingKernel(THETA_R, sets[0][0], sets[0][1], True, True, False)
        self.t_real = time.perf_counter() - t
        t = time.perf_counter()
        k1 = m.TrainingComplexKernel(THETA_C, sets[1][0], sets[1][1], True, True, False)
        self.t_cplx = time.perf_counter() - t
        k2 = m.TrainingKernel(THETA_R, sets[2][0], sets[2][1], True, True, False)
        self.models = [k0, k1, k2]
        self.pts = [syn.points_aos(*s) for s in sets]
        self.calls = 0

    def sample(self):
        """seconds of one full step, extrapolated from `points` evolved points per element"""
        lo = (self.calls * self.points) % (N_TRAIN - self.points)
        self.calls += 1
        pts = [p[lo:lo + self.points].copy() for p in self.pts]
        t = time.perf_counter()
        self.mod.evolve(PES_MODEL, pts[0], pts[1], pts[2], syn.MASS, syn.DT, *self.models)
        self.t_evolve = time.perf_counter() - t
        return 2 * self.t_real + self.t_cplx + self.t_evolve * (Q_POINTS / self.points)

    def describe(self):
        what = "oracle/_ref = the reference's own translation units compiled unmodified (stand-in Eigen/xtensor headers, host OpenBLAS for MKL)" if self.kind == "reference" else "oracle port of the reference (oracle/_ref not present)"
        return (f"{what}, {self.cores} threads: TrainingKernel N={N_TRAIN} measured once ({self.t_real:.2f} s, counted twice), TrainingComplexKernel N={N_TRAIN} measured once "
                f"({self.t_cplx:.1f} s), evolve() of {self.points} points/element over the three models measured per sample ({self.t_evolve:.2f} s) and "
                f"extrapolated x{Q_POINTS // self.points} to Q={Q_POINTS} points/element")


def cpu_best_effort_sample(sets):
    """The same C2 step on the host's BLAS / LAPACK (oracle/best_effort.py: potrf, triangular inverse, batched GEMM variance), bounded
    sample: the three trainings at full size, 4096 queries per predictor, extrapolated to the 8 Q queries each predictor serves."""
    from oracle import best_effort as be

    t = time.perf_counter()
    k0 = be.RealModel(THETA_R, sets[0][0], sets[0][1])
    t_train_real = time.perf_counter() - t
    t = time.perf_counter()
    k1 = be.ComplexModel(THETA_C, sets[1][0], sets[1][1])
    t_train_cplx = time.perf_counter() - t
    nq = 4096
    Xq, _ = syn.extra_points(3, 0, sets[0][0], nq, CENTRE)
    t = time.perf_counter()
    k0.predict(Xq)
    t_q_real = (time.perf_counter() - t) / nq
    Xq, _ = syn.extra_points(3, 1, sets[1][0], nq, CENTRE)
    t = time.perf_counter()
    k1.predict(Xq)
    t_q_cplx = (time.perf_counter() - t) / nq
    queries = 8 * Q_POINTS
    step_s = 2 * t_train_real + t_train_cplx + 2 * queries * t_q_real + queries * t_q_cplx
    desc = (f"numpy / scipy on the host BLAS: real train N={N_TRAIN} ({t_train_real:.2f}s), complex train ({t_train_cplx:.2f}s), {nq} real queries "
            f"({t_q_real * 1e6:.1f} us/query), {nq} complex queries ({t_q_cplx * 1e6:.1f} us/query), every variance computed; extrapolated to 8Q queries per predictor")
    return step_s, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sets = [syn.training_set(2, e, N_TRAIN, CENTRE) for e in range(3)]
    cpu = CpuStep(sets)
    times = []
    for i in range(args.warmup + args.steps):
        s = cpu.sample()
        if i >= args.warmup:
            times.append(s)
    step_s = float(np.mean(times))
    value = 1.0 / step_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "extrapolated": True,
        "config": {"workload": WORKLOAD, "note": "CPU arm: every step is a bounded sample of the workload (see cpu_baseline.sample); the full step would take hours"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.describe(), "extrapolated": True},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) == 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.rows[0][1]), "reasons": reasons}


def north_star_extra(ctx, dev, dmma_peak, q_points, steps):
    """BASELINE.json north_star on ONE GPU: a full time step at N = 16384 training points and 1e6 evolved MC points (configs[4]
    largest size).  State: the reference's t = 0 state (only rho00 populated, gple/main.cpp:57-64): the step trains one real
    element at N = 16384 (kernel build + factorise + inverse + LOOCV error + averages) and evolves 1e6 points (2 GPR queries
    per point into that model).  Timed like the headline: CUDA events around the whole step, inputs resident, bound-gated
    variance on (product default) and, once, off (every variance computed = the reference's amount of work).
    FP64 fraction = executed tensor-core flops (factorise 2n^3/3 + inverse, variance GEMM rows*n*(n+128)) / step time / DMMA peak:
    everything that is not a DMMA flop (kernel build, mean pass, reductions, launch gaps) counts against it."""
    import ctypes as C

    import torch

    from gaussian_process_liouville_equation_b200 import _lib as L
    from gaussian_process_liouville_equation_b200 import kernel as gk

    lib = ctx.lib
    N = 16384
    th = np.ascontiguousarray(syn.theta_real(max(0.25, (2048.0 / N) ** 0.5)))  # lengths shrink with N: cond(K) comparable to C2 (allowed by opt.cpp:1036-1040)
    X, y = syn.training_set(90, 0, N, CENTRE)
    Xe, ye = syn.extra_points(90, 0, X, q_points, CENTRE)
    d_X = torch.from_numpy(X).to(dev)
    d_y = torch.from_numpy(np.ascontiguousarray(y).view(np.float64)).to(dev)
    pts0 = torch.from_numpy(syn.points_aos(Xe, ye)).to(dev)
    scal = L.RealScalars()

    def step(q):
        pts = pts0[:q].clone()
        torch.cuda.synchronize()
        for slot in range(4):
            ctx.profile_read(slot)
        ctx.gate_statistics()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h = C.c_void_p()
        ctx.check(lib.gple_train_real(ctx.h, d_X.data_ptr(), d_y.data_ptr(), N, L.addr(th), L.CALC_ERROR | L.CALC_AVERAGE, C.byref(h), C.byref(scal)))
        ctx.check(lib.gple_evolve(ctx.h, PES_MODEL, h, None, None, pts.data_ptr(), q, None, 0, None, 0, syn.MASS, syn.DT))
        e1.record()
        torch.cuda.synchronize()
        lib.gple_model_destroy(ctx.h, h)
        return e0.elapsed_time(e1), [ctx.profile_read(slot) for slot in range(4)], ctx.gate_statistics()

    ctx.profile_enable(True)
    out = {"workload": f"north star / C5: Tully SAC t=0 state (rho00 populated), N={N} training points, Q={q_points} evolved MC points; step = TrainingKernel rebuild + evolve (2Q GPR predictions)",
           "n_train": N, "q_points": q_points, "dmma_peak_tflops": dmma_peak}
    for gated in (True, False):
        ctx.set_gated_variance(gated)
        step(q_points if gated else min(q_points, 20000))  # warm-up at the steady-state buffer sizes
        runs = [step(q_points) for _ in range(steps if gated else 1)]
        ms = float(np.mean([r[0] for r in runs]))
        prof, gs = runs[-1][1], runs[-1][2]
        flops = prof[0][2] + prof[2][2]
        ref_flops = float(N) ** 3 + 4.0 * q_points * float(N) ** 2  # reference formulation (LDLT + inverse ~ N^3, 2 Q queries x 2 N^2)
        out["gated" if gated else "every_variance_computed"] = {
            "ms_per_step": ms, "steps_per_s": 1000.0 / ms, "steps_timed": len(runs), "executed_dmma_tflop": flops / 1e12, "whole_step_tflops": flops / ms / 1e9,
            "frac_of_fp64_peak": flops / ms / 1e9 / dmma_peak, "reference_formulation_tflops_equiv": ref_flops / ms / 1e9,
            "variance_gemm_tflops": prof[0][2] / max(prof[0][0], 1e-9) / 1e9, "factorise_ms": prof[2][0], "factorise_tflops": prof[2][2] / max(prof[2][0], 1e-9) / 1e9,
            "kernel_build_ms": prof[1][0], "mean_pass_ms": prof[3][0], "rows_through_variance_gemm_fraction": (gs[1] / gs[0]) if gated and gs[0] else 1.0}
    ctx.set_gated_variance(True)
    ctx.profile_enable(False)
    del pts0
    # parity spot-check at this size through the reference-shaped API: K (K^-1 y) = y, and the LOOCV identity
    # error = sum_i (v_i / [K^-1]_ii)^2 (kernel.cpp:285-287) against the explicit inverse
    k = gk.TrainingKernel(th, (X, y), True, True, False, ctx=ctx)
    v, lab = k.get_inverse_times_label(), k.get_label()
    K = gk.kernel_matrix(X, X, th, True, ctx=ctx)
    Kinv = k.get_inverse()
    probe = np.random.default_rng(16384).standard_normal((N, 4))
    out["check"] = {"status": int(k.status), "population": k.get_population(), "loocv_error": k.get_error(),
                    "residual_K_v_minus_y_rel": float(np.abs(K @ v - lab).max() / np.abs(lab).max()),
                    "loocv_identity_rel": float(abs(k.get_error() / float(np.sum((v / np.diag(Kinv)) ** 2)) - 1.0)),
                    "K_Kinv_probe_rel": float(np.abs(K @ (Kinv @ probe) - probe).max() / np.abs(probe).max())}
    k.close()
    return out


def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from gaussian_process_liouville_equation_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = L.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1:
        # multi-GPU through the library: the context gets its own NCCL communicator (gple_ctx_comm_init); torch.distributed only
        # carries the 128-byte id to the other ranks and the barrier / max-over-ranks of the timing protocol
        box = [L.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(rank, world, box[0])
    if args.gate_stage_tiles is not None:
        ctx.set_gate_stage_tiles(args.gate_stage_tiles, args.gate_stage_tiles_im)
    if args.gate_stage2_tiles is not None:
        ctx.set_gate_stage2_tiles(args.gate_stage2_tiles)
    if args.gate_schedule_real is not None:
        ctx.set_gate_schedule(False, parse_schedule(args.gate_schedule_real))
    if args.gate_schedule_complex is not None:
        ctx.set_gate_schedule(True, parse_schedule(args.gate_schedule_complex))
    lib = ctx.lib
    # the three element models of a step are independent: they are factorised concurrently, one context (stream +
    # workspace) per element, from three host threads (the C-ABI is thread-safe across contexts)
    from concurrent.futures import ThreadPoolExecutor

    train_ctx = [L.Context(local), L.Context(local), ctx]
    if args.no_factorise_graphs:
        for c in train_ctx:
            c.set_factorise_graphs(False)
    pool = ThreadPoolExecutor(3)

    sets, pts_all = make_inputs()
    # block partition of the evolved points (strong scaling) inside gple_evolve_sharded: every rank holds the full point sets,
    # evolves its own block of each and the blocks are all-gathered in place by the library (ncclAllGather on the context's
    # stream); every rank keeps the full training sets
    lo, hi = L.partition(Q_POINTS, rank, world)
    nloc = hi - lo
    thetas = [THETA_R, THETA_C, THETA_R]
    d_X = [torch.from_numpy(s[0]).to(dev) for s in sets]
    d_y = [torch.from_numpy(np.ascontiguousarray(s[1]).view(np.float64)).to(dev) for s in sets]
    d_pts0 = [torch.from_numpy(p).to(dev) for p in pts_all]
    d_pts = [t.clone() for t in d_pts0]
    h_X = [torch.from_numpy(s[0]).pin_memory() for s in sets]
    h_y = [torch.from_numpy(np.ascontiguousarray(s[1]).view(np.float64)).pin_memory() for s in sets]
    h_pts0 = [torch.from_numpy(p).pin_memory() for p in pts_all]
    h_pts = [t.clone().pin_memory() for t in h_pts0]
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    last_scalars = {}

    def train_one(e, X, y):
        c = train_ctx[(e + 2) % 3]  # the complex element (largest) on the main context
        h = C.c_void_p()
        th = np.ascontiguousarray(thetas[e])
        if e == 1:
            s = L.ComplexScalars()
            c.check(lib.gple_train_complex(c.h, X[e].data_ptr(), y[e].data_ptr(), N_TRAIN, L.addr(th), L.CALC_ERROR | L.CALC_AVERAGE, C.byref(h), C.byref(s)))
            last_scalars["purity10"], last_scalars["error10"] = s.purity, s.error
        else:
            s = L.RealScalars()
            c.check(lib.gple_train_real(c.h, X[e].data_ptr(), y[e].data_ptr(), N_TRAIN, L.addr(th), L.CALC_ERROR | L.CALC_AVERAGE, C.byref(h), C.byref(s)))
            last_scalars[f"population{e}"] = s.population
        return h

    # N > 1: the three element models are independent (predict.cpp:390-393), so each is trained on ONE rank -- the complex
    # element (order 2N, ~8x the flops of a real one) on rank 0, the real ones on ranks 1 and 2 (both on rank 1 when there are
    # only two) -- and replicated with gple_model_bcast; a factorisation never spans GPUs (BASELINE.json north_star)
    owner = [1 % world, 0, 2 % world]

    def train_all(X, y):
        # gple_train_* returns after its stream is synchronised, so the models are complete when the threads join
        if world == 1:
            return list(pool.map(lambda e: train_one(e, X, y), range(3)))
        models = list(pool.map(lambda e: train_one(e, X, y) if owner[e] == rank else C.c_void_p(), range(3)))
        for e in range(3):
            ctx.check(lib.gple_model_bcast(ctx.h, C.byref(models[e]), owner[e]))
        return models

    def step(X, y, pts):
        """One time step through the C-ABI.  X, y, pts: device tensors (resident run) or pinned host tensors (e2e run)."""
        models = train_all(X, y)
        ctx.check(lib.gple_evolve_sharded(ctx.h, PES_MODEL, models[0], models[1], models[2], pts[0].data_ptr(), Q_POINTS, pts[1].data_ptr(), Q_POINTS, pts[2].data_ptr(), Q_POINTS, syn.MASS, syn.DT))
        for h in models:
            lib.gple_model_destroy(ctx.h, h)

    def reset():
        for a, b in zip(d_pts, d_pts0):
            a.copy_(b)
        for a, b in zip(h_pts, h_pts0):
            a.copy_(b)
        flush.fill_(1.0)  # evict L2 (126 MB) between timed steps
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        total = 0.0
        for _ in range(steps):
            reset()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            barrier()
            total += e0.elapsed_time(e1)
        t = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    resident = lambda: step(d_X, d_y, d_pts)  # noqa: E731

    def host_step():
        step(h_X, h_y, h_pts)  # host buffers in, full evolved sets back in the same host buffers (copies + exchange inside the call)

    for _ in range(max(args.warmup, 3)):
        reset()
        resident()
    torch.cuda.synchronize()
    dmma_peak, dfma_peak = ctx.fp64_peak()
    tile_peak = ctx.dmma_tile_peak()

    sampler = ClockSampler(local)
    sampler.start()
    ctx.profile_enable(True)
    for slot in range(4):
        ctx.profile_read(slot)
    for c in train_ctx[:2]:
        c.profile_enable(True)
        c.profile_read(2)
    launches0 = sum(c.launches for c in train_ctx)
    total_ms = timed(resident, args.steps)
    launches = sum(c.launches for c in train_ctx) - launches0
    prof = [ctx.profile_read(slot) for slot in range(4)]
    for c in train_ctx[:2]:  # factorisations of the two real elements ran concurrently on their own streams
        ms, n_, w_ = c.profile_read(2)
        prof[2] = (prof[2][0] + ms, prof[2][1] + n_, prof[2][2] + w_)
        c.profile_enable(False)
    ctx.profile_enable(False)
    clocks = sampler.stop()

    rows_total = rows_var = rows_zero = rows_b = 0
    for c in train_ctx:
        a_, b_, c_, d_ = c.gate_statistics()
        rows_total, rows_var, rows_zero, rows_b = rows_total + a_, rows_var + b_, rows_zero + c_, rows_b + d_
    # the same step with every variance computed (GPLE_OPT_GATED_VARIANCE = 0), for comparison
    ctx.set_gated_variance(False)
    reset()
    resident()
    full_ms = timed(resident, args.steps)
    ctx.set_gated_variance(True)

    reset()
    host_step()  # warm the host path
    e2e_ms = timed(host_step, args.steps)

    ms_per_step = total_ms / args.steps
    value = 1000.0 / ms_per_step
    e2e_value = 1000.0 / (e2e_ms / args.steps)
    h2d = sum(t.numel() * 8 for t in h_X + h_y + h_pts)
    d2h = sum(t.numel() * 8 for t in h_pts) + 3 * 160

    var_ms, var_n, var_flops = prof[0]
    kb_ms, kb_n, kb_bytes = prof[1]
    fa_ms, fa_n, fa_flops = prof[2]
    mean_ms, mean_n, mean_exps = prof[3]
    achieved = var_flops / (var_ms * 1e-3) / 1e12 if var_ms > 0 else 0.0
    # algorithmic flops of the reference formulation (2 Q N^2 per real query set, 16 Q N^2 complex)
    alg_flops_step = 8 * nloc * (2 + 2 + 16) * float(N_TRAIN) ** 2 + (1 + 1 + 24 + 1.0 / 3.0) * float(N_TRAIN) ** 3

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": f"evolved points block-partitioned over {world} GPU(s) inside gple_evolve_sharded; each element model trained on one rank and replicated (gple_model_bcast); one in-place NCCL all-gather of the evolved sets per element on the library's own communicator",
                   "l2": "512 MiB buffer written between timed steps; per-step working set (K* chunk 310-620 MB) exceeds L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "var_gemm_kernel (Z = K* W^T, fused row sum of squares; FP64 DMMA)", "achieved": achieved,
                     "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak if dmma_peak > 0 else None, "traffic": 2.338e9 * (var_flops / max(var_n, 1)) / 8.442e10,
                     "traffic_note": "ncu dram__bytes_read+write of a real-element launch (2.338 GB for 8.44e10 flops, profiles/r01_var_gemm_ncu_full.md), scaled by flops per launch; the launches of the multi-stage schedule measure 0.028-0.031 bytes per flop, early and late stages alike (profiles/r02_ncu_summary.md, addendum)",
                     "peak_source": "measured live by gple_measure_fp64_peak (register-resident DMMA.8x8x4 loop); MEASURED_PEAKS.json has no FP64 entry",
                     "launches": var_n, "avg_launch_ms": var_ms / max(var_n, 1), "share_of_step": var_ms / total_ms,
                     "flops_counted": "executed flops of every launch: 2*128^2*sum over its n-tile set of (tile+1) per row (= rows*n*(n+128) for the full triangular product; the reference formulation K* K^-1 k^T would be 2x that)"},
        "extra": {"fp64_dfma_peak_tflops": dfma_peak, "dmma_register_tile_ceiling_tflops": tile_peak,
                  "dmma_register_tile_ceiling_note": "same 8x4 DMMA register tile at 8 warps/SM with changing operands and no memory traffic: the ceiling of an mma.sync FP64 GEMM at this occupancy",
                  "kernel_build_gbs": kb_bytes / (kb_ms * 1e-3) / 1e9 if kb_ms > 0 else None, "kernel_build_share": kb_ms / total_ms,
                  "mean_pass_share": mean_ms / total_ms, "mean_pass_gexp_per_s": mean_exps / (mean_ms * 1e-3) / 1e9 if mean_ms > 0 else None,
                  "mean_pass_note": "kmean_kernel: K* v for all 8Q queries without storing K*; FP64-pipe bound (DFMA shares the pipe with DMMA, profiles/r02_kstar_fusion_decision.md): 19.5 FP64 instructions per kernel value with the table-based exp of csrc/gpr.cu",
                  "cholesky_inverse_tflops": fa_flops / (fa_ms * 1e-3) / 1e12 if fa_ms > 0 else None, "factorise_share": fa_ms / total_ms,
                  "gated_variance": {"enabled": True, "rows_total": rows_total, "rows_through_variance_gemm": rows_var, "rows_decided_zero": rows_zero, "rows_needing_the_full_variance": rows_b,
                                     "fraction": rows_var / max(rows_total, 1), "fraction_full": rows_b / max(rows_total, 1),
                                     "note": "exact short-cuts of the cutoff gate (kernel.h:301-332): gate == 1 when |f|^2 >= 4 k** (variance <= prior) or when |f|^2 >= 4 (k** - sum Z^2 over the columns of the first few training blocks) (variance <= variance given those points only); gate == 0 when |f|^2 <= noise/2 (variance >= noise); only the undecided queries get the full variance; outputs agree to rounding with the full computation (tests/test_gpu_dynamics.py), whose rate is reported next to it"},
                  "value_with_every_variance_computed": 1000.0 / (full_ms / args.steps), "ms_per_step_with_every_variance_computed": full_ms / args.steps,
                  "reference_formulation_tflops_equiv": alg_flops_step / ((full_ms / args.steps) * 1e-3) / 1e12,
                  "check": last_scalars},
    }
    if world == 1 and not args.no_north_star:
        pool.shutdown()
        del flush, d_pts, d_pts0
        torch.cuda.empty_cache()
        line["extra"]["north_star"] = north_star_extra(ctx, dev, dmma_peak, args.north_star_points, 2)
    if world == 1 and not args.no_cpu_baseline:
        cpu = CpuStep(sets)
        step_s = cpu.sample()
        line["cpu_baseline"] = {"value": 1.0 / step_s, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.describe(), "extrapolated": True}
        try:  # second CPU number of SURVEY.md 8d: best effort on the host's BLAS / LAPACK (reported, not the reference arm)
            be_s, be_desc = cpu_best_effort_sample(sets)
            line["cpu_baseline"]["best_effort_blas"] = {"value": 1.0 / be_s, "unit": UNIT, "sample": be_desc}
        except Exception as e:  # scipy missing on the box: the reference-shaped number stands alone
            line["cpu_baseline"]["best_effort_blas"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-north-star", action="store_true", help="skip extra.north_star (N=16384, 1e6 evolved points; about 40 s)")
    ap.add_argument("--north-star-points", type=int, default=1_000_000)
    ap.add_argument("--gate-stage-tiles", type=int, default=None, help="override GPLE_OPT_GATE_STAGE_TILES (tuning)")
    ap.add_argument("--gate-stage-tiles-im", type=int, default=-1, help="override GPLE_OPT_GATE_STAGE_TILES_IM (tuning)")
    ap.add_argument("--gate-stage2-tiles", type=int, default=None, help="override GPLE_OPT_GATE_STAGE2_TILES (tuning; 0 = stage B in one part)")
    ap.add_argument("--no-factorise-graphs", action="store_true", help="GPLE_OPT_FACTORISE_GRAPHS = 0: same kernels launched one by one (for the ncu launch list: ncu does not survive concurrent stream captures from several host threads)")
    ap.add_argument("--gate-schedule-real", default=None, help="explicit schedule of the staged bound, real elements: e.g. 1,5,10 (tuning)")
    ap.add_argument("--gate-schedule-complex", default=None, help="same for the complex element, re:im pairs: e.g. 1:0,5:1,10:2,16:6 (tuning)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"], help="c2 (default, BASELINE.json configs[1]); c4 = configs[3]: ECR, N=4096, 1e6 evolved points/element, meant for --gpus 8")
    args = ap.parse_args()
    if args.workload == "c4":
        global N_TRAIN, Q_POINTS, PES_MODEL, WORKLOAD, THETA_R, THETA_C
        N_TRAIN, Q_POINTS, PES_MODEL = 4096, 1_000_000, 2
        scale = (2048.0 / N_TRAIN) ** 0.5  # lengths shrink with N: comparable conditioning (allowed by the bounds of opt.cpp:1036-1040)
        THETA_R = syn.theta_real(scale)
        THETA_C = THETA_C * np.array([1, 1, scale, scale, 1, scale, scale, 1])
        WORKLOAD = ("C4: Tully extended coupling with reflection, N=4096 training points/element, 3 elements, Q=1e6 evolved MC points/element/step "
                    "sharded over the GPUs; step = TrainingKernels rebuild (3 elements) + evolve (8Q GPR predictions per element)")
        args.no_cpu_baseline = args.no_north_star = True  # the CPU sample is sized for C2
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
