"""Host-side mirror of the reference's gple/predict.h: training-set construction and the aggregate of the (up to)
three element models with its analytic observables and their parameter gradients.

Elements are kept in the reference's lower-triangular order (rho00, rho10, rho11) (gple/storage.h:13-17,
gple/predict.cpp:512-559); the concatenated parameter vector has 4 + 8 + 4 = 16 entries (gple/predict.h:17).
"""
from __future__ import annotations

import numpy as np

from . import complex_kernel as ck
from . import kernel as rk

NumPES = 2
NumElements = 3
DIAGONAL = (0, 2)  # element index of rho00, rho11
NumTotalParameters = rk.NumTotalParameters * NumPES + ck.NumTotalParameters  # predict.h:17
ELEMENT_SLICES = (slice(0, 4), slice(4, 12), slice(12, 16))


def construct_training_sets(density):
    """construct_training_sets (predict.cpp:246-280): AoS points (n, 4) -> (feature (n, 2), complex label (n,))."""
    out = []
    for pts in density:
        if pts is None or len(pts) == 0:
            out.append(None)
        else:
            pts = np.asarray(pts, dtype=np.float64)
            out.append((np.ascontiguousarray(pts[:, :2]), pts[:, 2] + 1j * pts[:, 3]))
    return out


def _is_all_zero(params):
    return params is None or not np.any(np.asarray(params))


class TrainingKernels:
    """predict.h:89-143.  `backend` supplies TrainingKernel / TrainingComplexKernel (default: the CUDA mirror)."""

    def __init__(self, ParameterVectors, TrainingSets, IsToCalculateError=True, IsToCalculateAverage=True, IsToCalculateDerivative=False, backend=None):
        self.kernels = [None, None, None]
        real_cls = backend.TrainingKernel if backend else rk.TrainingKernel
        cplx_cls = backend.TrainingComplexKernel if backend else ck.TrainingComplexKernel
        for e in range(NumElements):
            ts = TrainingSets[e]
            if ts is None or len(ts[0]) == 0:
                continue
            if e == 1:
                if _is_all_zero(ParameterVectors[e]):  # predict.cpp:339-357
                    continue
                self.kernels[e] = cplx_cls(ParameterVectors[e], ts, IsToCalculateError, IsToCalculateAverage, IsToCalculateDerivative)
            else:
                self.kernels[e] = real_cls(ParameterVectors[e], ts, IsToCalculateError, IsToCalculateAverage, IsToCalculateDerivative)

    @classmethod
    def from_density(cls, ParameterVectors, density, backend=None):
        """predict.cpp:390-393"""
        return cls(ParameterVectors, construct_training_sets(density), True, True, False, backend)

    def __getitem__(self, e):
        return self.kernels[e]

    def calculate_population(self):  # predict.cpp:395-406
        return sum(self.kernels[e].get_population() for e in DIAGONAL if self.kernels[e] is not None)

    def calculate_1st_order_average(self):  # predict.cpp:408-419
        r = np.zeros(2)
        for e in DIAGONAL:
            if self.kernels[e] is not None:
                r += self.kernels[e].get_1st_order_average()
        return r

    def calculate_total_energy_average(self, Energies):  # predict.cpp:423-436
        return sum(self.kernels[e].get_population() * Energies[i] for i, e in enumerate(DIAGONAL) if self.kernels[e] is not None)

    def calculate_purity(self):  # predict.cpp:439-463
        r = sum(self.kernels[e].get_purity() for e in DIAGONAL if self.kernels[e] is not None)
        if self.kernels[1] is not None:
            r += 2.0 * self.kernels[1].get_purity()
        return r

    def population_derivative(self):  # predict.cpp:465-484 (diagonal parameters only: 2 x 4)
        r = np.zeros(NumPES * rk.NumTotalParameters)
        for i, e in enumerate(DIAGONAL):
            if self.kernels[e] is not None:
                r[4 * i:4 * i + 4] = self.kernels[e].get_population_derivative()
        return r

    def total_energy_derivative(self, Energies):  # predict.cpp:486-510
        r = np.zeros(NumPES * rk.NumTotalParameters)
        for i, e in enumerate(DIAGONAL):
            if self.kernels[e] is not None:
                r[4 * i:4 * i + 4] = self.kernels[e].get_population_derivative() * Energies[i]
        return r

    def purity_derivative(self):  # predict.cpp:512-559 (all 16 parameters, off-diagonal weighted by 2)
        r = np.zeros(NumTotalParameters)
        for e in range(NumElements):
            if self.kernels[e] is not None:
                r[ELEMENT_SLICES[e]] = self.kernels[e].get_purity_derivative() * (2.0 if e == 1 else 1.0)
        return r
