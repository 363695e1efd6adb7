"""Host-side mirror of the path-feeding parts of the reference's gple/mc.h: the analytic initial Wigner
distribution, the extra ("validation") point generation and the Metropolis sampling of the elements
(mc.cpp:125-537): all chains of an element advance in lock-step on the GPU, one batched density evaluation
per step (gple_markov_chains), with the reference's displacement / chain-length auto-tuning on top.

The reference draws from a clock-seeded, thread-shared std::mt19937 (mc.cpp:17,87); here the caller passes a
numpy Generator (counter-based Philox streams from `synthetic.rng`) so that runs are reproducible.
"""
from __future__ import annotations

import numpy as np

from . import dynamics
from . import complex_kernel as ck
from . import kernel as rk


def initial_distribution(r0, SigmaR0, r, RowIndex, ColIndex, InitialPopulation=(1.0, 0.0), InitialPhaseFactor=(0.0, 0.0)):
    """mc.cpp:30-50 for points r (n, 2)"""
    r0, s, r = np.asarray(r0), np.asarray(SigmaR0), np.asarray(r, dtype=np.float64).reshape(-1, 2)
    gw = np.exp(-0.5 * (((r - r0) / s) ** 2).sum(1)) / (2.0 * np.pi * s.prod())
    sw = sum(p * p for p in InitialPopulation)
    return gw * InitialPopulation[RowIndex] * InitialPopulation[ColIndex] / sw * np.exp(1j * (InitialPhaseFactor[RowIndex] - InitialPhaseFactor[ColIndex]))


def predict_distribution(kernels, r, element):
    """main.cpp:75-101, batched: cutoff prediction of `element` at points r (n, 2); 0 where the element has no model."""
    r = np.asarray(r, dtype=np.float64).reshape(-1, 2)
    k = kernels[element]
    if k is None:
        return np.zeros(len(r), dtype=np.complex128)
    if element == 1:
        return ck.PredictiveComplexKernel(r, k).get_cutoff_prediction()
    return rk.PredictiveKernel(r, k).get_cutoff_prediction().astype(np.complex128)


def generate_extra_points(density, NumExtraPoints, kernels, rng, pes_model=0, mass=1.0, distribution=None):
    """mc.cpp:59-120: training point (cyclic) + N(0, sigma_element) jitter, labelled with the CURRENT prediction
    (one batched GPU prediction per element instead of NumExtraPoints single-point calls)."""
    out = []
    for e, pts in enumerate(density):
        if pts is None or len(pts) == 0:
            out.append(None)
            continue
        pts = np.asarray(pts, dtype=np.float64)
        sd = dynamics.calculate_standard_deviation_one_surface(pes_model, pts, mass)
        r = pts[np.arange(NumExtraPoints) % len(pts), :2] + sd * rng.standard_normal((NumExtraPoints, 2))
        rho = distribution(r, e) if distribution is not None else predict_distribution(kernels, r, e)
        out.append(np.column_stack([r, rho.real, rho.imag]))
    return out


def is_very_small(density, mass, dt, kernels, pes_model, epsilon=1e-10):
    """evolve.cpp:445-478: an EMPTY element is "small" iff new_point_predict of all test points (the rho00 points)
    has |rho|^2 < epsilon."""
    elements = ((0, 0), (1, 0), (1, 1))
    small = [False, False, False]
    for e, (row, col) in enumerate(elements):
        if density[e] is None or len(density[e]) == 0:
            r = np.asarray(density[0], dtype=np.float64)[:, :2]
            rho = dynamics.new_point_predict(pes_model, r, mass, dt, kernels, row, col)
            small[e] = bool(np.all(np.abs(rho) ** 2 < epsilon))
    return small


# ---- Metropolis sampling (mc.cpp:125-537) ---------------------------------------------------------------------
MaxAcceptRatio, MinAcceptRatio = 0.5, 0.15  # mc.cpp:19-20
PossibleDisplacement = (1e-4, 2e-4, 5e-4, 1e-3, 2e-3, 5e-3, 0.01, 0.02, 0.05, 0.1, 0.2, 0.5, 1.0, 2.0, 5.0, 10.0)  # mc.cpp:298
ELEMENTS = ((0, 0), (1, 0), (1, 1))


class MCParameters:
    """mc.h:45-121"""

    AboveMinFactor = 1.1

    def __init__(self, InitialSteps=200, InitialDisplacement=1.0):
        self.NOMC, self.displacement = InitialSteps, InitialDisplacement

    def set_num_MC_steps(self, n):
        self.NOMC = int(n)

    def set_displacement(self, d):
        self.displacement = float(d)

    def get_num_MC_steps(self):
        return self.NOMC

    def get_max_displacement(self):
        return self.displacement


class Sampler:
    """The DistributionFunction of the reference bound to the GPU chain runner.  Exactly one target:
    analytic = (r0, SigmaR0, InitialPopulation, InitialPhaseFactor)   initial_distribution (main.cpp:40-56)
    kernels                                                          predict_distribution (main.cpp:75-101)
    kernels + new_point = (pes_model, mass, dt)                      new_point_predict (main.cpp:153-156)
    Every call of `chains` advances a call counter that is folded into the Philox stream id, so successive walks of
    one sampler are independent while the whole sequence is a pure function of `seed`."""

    def __init__(self, seed, analytic=None, kernels=None, new_point=None, ctx=None):
        from . import _lib as L

        self.L, self.ctx, self.seed, self.calls = L, ctx or L.default_context(), int(seed), 0
        self.analytic, self.kernels, self.new_point = analytic, kernels, new_point

    def _source(self, row, col):
        L = self.L
        src = L.McSource()
        src.row, src.col = row, col
        if self.analytic is not None:
            r0, s0, pop, ph = self.analytic
            src.kind = L.MC_ANALYTIC
            src.analytic[:] = [r0[0], r0[1], s0[0], s0[1], pop[0], pop[1], ph[0], ph[1]]
            return src
        h = [k.h if k is not None else None for k in self.kernels]
        src.m00, src.m10, src.m11 = h
        if self.new_point is None:
            src.kind = L.MC_PREDICT
        else:
            src.kind = L.MC_NEW_POINT
            src.pes_model, src.mass, src.dt = int(self.new_point[0]), float(self.new_point[1]), float(self.new_point[2])
        return src

    def next_stream(self, element):
        self.calls += 1
        return self.calls * 4 + element

    def chains(self, pts, num_steps, max_displacement, row, col, want_chain=False, stream=None, chain0=0):
        """generate_markov_chain (mc.cpp:143-188) for all points: (pts (n, 4) at the last states, accept (n,), chains | None).
        Point k walks the Philox stream of chain `chain0 + k`: a rank that owns the block [lo, hi) of a sharded point set passes
        chain0 = lo and gets exactly the states it would have had in the whole set."""
        import ctypes as C

        L = self.L
        pts = np.array(L.f64(pts), copy=True)
        n = len(pts)
        stream = self.next_stream(row + col) if stream is None else stream
        accept = np.empty(n)
        chains = np.empty((n, num_steps + 1, 2)) if want_chain else None
        src = self._source(row, col)
        self.ctx.check(self.ctx.lib.gple_markov_chains(self.ctx.h, C.byref(src), L.addr(pts), n, int(num_steps), float(max_displacement), self.seed, int(stream), int(chain0), L.addr(accept),
                                                       L.addr(chains) if want_chain else None))
        return pts, accept, chains

    def autocorrelation(self, chains):
        """mc.cpp:230-243"""
        L = self.L
        chains = L.f64(chains)
        out = np.empty(chains.shape[1] // 2)
        self.ctx.check(self.ctx.lib.gple_chain_autocorrelation(self.ctx.h, L.addr(chains), chains.shape[0], chains.shape[1], L.addr(out)))
        return out

    def density(self, r, row, col):
        """distribution(r, row, col) for points r (n, 2)"""
        pts = np.zeros((len(r), 4))
        pts[:, :2] = r
        out, _, _ = self.chains(pts, 0, 1.0, row, col, stream=0)
        return out[:, 2] + 1j * out[:, 3]


def acceptance_optimize_displacement(MCParams, sampler, density, row, col):
    """mc.cpp:288-337: the largest displacement of the list whose mean acceptance ratio lies in (0.15, 0.5)"""
    MaxNOMC = 2 * 500
    for d in reversed(PossibleDisplacement):
        _, accept, _ = sampler.chains(density, MaxNOMC, d, row, col)
        ratio = float(accept.mean())
        if MinAcceptRatio < ratio < MaxAcceptRatio:
            MCParams.set_displacement(d)
            return


def autocorrelation_optimize_steps(MCParams, sampler, density, row, col):
    """mc.cpp:197-279: chain length = first lag whose |autocorrelation| is within 1.1x of the minimum (with the
    acceptance ratio of a chain of that length inside the window)"""
    MaxNOMC = 2 * 1000
    _, _, chains = sampler.chains(density, MaxNOMC, MCParams.get_max_displacement(), row, col, want_chain=True)
    autocor = sampler.autocorrelation(chains)
    length = len(autocor)
    min_start, min_step, min_ac = 0, 0, 0.0
    while True:
        min_start = min_step + 1
        if min_start >= length:  # nothing satisfies the window
            min_start = 1
            min_step = int(np.argmin(np.abs(autocor)))
            min_ac = float(np.abs(autocor)[min_step])
            break
        tail = np.abs(autocor[min_start:])
        min_step = int(np.argmin(tail)) + min_start
        min_ac = float(tail[min_step - min_start])
        _, acc, _ = sampler.chains(np.asarray(density)[:1], min_step, MCParams.get_max_displacement(), row, col)
        if MinAcceptRatio <= acc[0] <= MaxAcceptRatio:
            break
    for i in range(min_start, min_step):
        if abs(autocor[i]) <= MCParameters.AboveMinFactor * min_ac:
            min_step = i
            break
    MCParams.set_num_MC_steps(min_step)


def element_monte_carlo(density, MCParams, sampler, row, col):
    """mc.cpp:339-378: tune displacement and chain length, then walk every point and relabel it"""
    acceptance_optimize_displacement(MCParams, sampler, density, row, col)
    autocorrelation_optimize_steps(MCParams, sampler, density, row, col)
    out, _, _ = sampler.chains(density, MCParams.get_num_MC_steps(), MCParams.get_max_displacement(), row, col)
    return out


def monte_carlo_selection(density, MCParams, sampler):
    """mc.cpp:380-403"""
    out = []
    for e, (row, col) in enumerate(ELEMENTS):
        pts = density[e]
        out.append(element_monte_carlo(pts, MCParams[e], sampler, row, col) if pts is not None and len(pts) > 0 else pts)
    return out


def new_element_point_selection(density, extra_points, IsSmallOld, IsSmall, MCParams, sampler, rng):
    """mc.cpp:407-537: a newly populated element takes the N most important of all current coordinates (by |rho|^2 of
    `sampler`'s density, new_point_predict in main.cpp:147-157), replicated up to N, walks them, and gets extra points;
    a newly small element is emptied."""
    if list(IsSmallOld) == list(IsSmall):
        return density, extra_points
    density, extra_points = list(density), list(extra_points)
    NumPoints, NumExtraPoints = len(density[0]), len(extra_points[0])
    coords = np.concatenate([np.asarray(p, dtype=np.float64)[:, :2] for e in range(3) for p in (density[e], extra_points[e]) if p is not None and len(p) > 0])
    for e, (row, col) in enumerate(ELEMENTS):
        if IsSmallOld[e] and not IsSmall[e]:
            rho = sampler.density(coords, row, col)
            nonzero = int(np.count_nonzero(rho))
            keep = min(NumPoints, nonzero)
            order = np.argsort(-np.abs(rho) ** 2, kind="stable")[:keep]
            pts = np.column_stack([coords[order], rho[order].real, rho[order].imag])
            while NumPoints >= 2 * len(pts):
                pts = np.concatenate([pts, pts])
            if len(pts) < NumPoints:
                pts = np.concatenate([pts, pts[:NumPoints - len(pts)]])
            density[e] = element_monte_carlo(pts, MCParams[e], sampler, row, col)
            sd = np.asarray(density[e])[:, :2].std(axis=0)
            r = density[e][np.arange(NumExtraPoints) % len(density[e]), :2] + sd * rng.standard_normal((NumExtraPoints, 2))
            lab = sampler.density(r, row, col)
            extra_points[e] = np.column_stack([r, lab.real, lab.imag])
        elif not IsSmallOld[e] and IsSmall[e]:
            density[e], extra_points[e] = None, None
    return density, extra_points
