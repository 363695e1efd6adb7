"""Host-side mirror of the path-feeding parts of the reference's gple/mc.h: the analytic initial Wigner
distribution and the extra ("validation") point generation.  Metropolis sampling (mc.cpp:125-537) is out of the
hot-path scope.

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
