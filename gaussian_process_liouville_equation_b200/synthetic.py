"""Deterministic synthetic Tully-model inputs.

The reference seeds its RNG from the wall clock and shares it un-synchronised between threads
(gple/mc.cpp:17,87,168), so parity is only ever checked on fixed point sets.  This module is the one
place those point sets come from: counter-based Philox4x64 streams keyed by "gple", with the
(config id, element id) pair as the counter, so tests, bench.py and every rank of a multi-GPU run
regenerate bit-identical inputs without exchanging them.

Physical constants follow the reference's own test harness (test/continue_test.cpp:41,
test/stdafx.h:51): mass 2000, x0 = -10, p0 = 14.112, sigma_p = 0.7056, sigma_x = hbar / (2 sigma_p).
"""
from __future__ import annotations

import numpy as np

KEY = 0x67706C65  # "gple"
MASS = 2000.0
X0, P0 = -10.0, 14.112
SIGMA_P = 0.7056
SIGMA_X = 1.0 / (2.0 * SIGMA_P)
DT = 1.0
INITIAL_NOISE = 1e-2  # gple/opt.cpp:27
ELEMENTS = ((0, 0), (1, 0), (1, 1))  # lower-triangular order of gple/storage.h / evolve.cpp:196-199


def rng(config_id: int, element_id: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=KEY, counter=[stream, element_id, config_id, 0]))


def wigner_gaussian(r: np.ndarray, centre, sigma=(SIGMA_X, SIGMA_P)) -> np.ndarray:
    """gple/mc.cpp:41 initial Wigner Gaussian."""
    d = (r - np.asarray(centre)) / np.asarray(sigma)
    return np.exp(-0.5 * np.sum(d * d, axis=1)) / (2.0 * np.pi * sigma[0] * sigma[1])


def element_centre(element_id: int, centre):
    c = np.asarray(centre, dtype=np.float64)
    return c + np.array([0.5, -0.3]) if element_id == 2 else c


def labels(element_id: int, r: np.ndarray, centre) -> np.ndarray:
    """rho00 = 0.6 x Gaussian; rho11 = 0.4 x Gaussian shifted by (+0.5, -0.3); rho10 = sqrt(rho00 rho11) e^{i phase}."""
    c = np.asarray(centre, dtype=np.float64)
    g00 = 0.6 * wigner_gaussian(r, c)
    g11 = 0.4 * wigner_gaussian(r, c + np.array([0.5, -0.3]))
    if element_id == 0:
        return g00.astype(np.complex128)
    if element_id == 2:
        return g11.astype(np.complex128)
    phase = 0.7 * (r[:, 0] - c[0]) - 0.2 * (r[:, 1] - c[1])
    return np.sqrt(g00 * g11) * np.exp(1j * phase)


def snapshot_purity() -> float:
    """Purity of the three-element synthetic snapshot of `labels`: (2 pi hbar) int rho_ii^2 = pop_i^2 for the Wigner Gaussians
    (sigma_x sigma_p = hbar / 2) plus 2 (2 pi hbar) int |rho10|^2 = 2 * 0.24 * overlap of the two shifted Gaussians."""
    overlap = np.exp(-(0.5 ** 2) / (4.0 * SIGMA_X ** 2) - (0.3 ** 2) / (4.0 * SIGMA_P ** 2))
    return 0.6 ** 2 + 0.4 ** 2 + 2.0 * 0.24 * overlap


def training_set(config_id: int, element_id: int, n: int, centre=(0.0, P0)):
    """N i.i.d. features x ~ N(x_c, sigma_x^2), p ~ N(p_c, sigma_p^2) and their labels.

    Returns (X (n, 2) float64, y (n,) complex128).
    """
    g = rng(config_id, element_id, 0)
    c = element_centre(element_id, centre)
    X = np.empty((n, 2))
    X[:, 0] = c[0] + SIGMA_X * g.standard_normal(n)
    X[:, 1] = c[1] + SIGMA_P * g.standard_normal(n)
    return X, labels(element_id, X, centre)


def extra_points(config_id: int, element_id: int, X: np.ndarray, m: int, centre=(0.0, P0)):
    """gple/mc.cpp:59-94: training point (cyclic) + N(0, sigma_element) jitter; label = exact density."""
    g = rng(config_id, element_id, 1)
    sd = X.std(axis=0)  # gple/predict.cpp:109-126 (population standard deviation)
    idx = np.arange(m) % len(X)
    Xe = X[idx] + sd * g.standard_normal((m, 2))
    return Xe, labels(element_id, Xe, centre)


def theta_real(scale: float = 1.0) -> np.ndarray:
    """(sigma_f, l_x, l_p, sigma_n), gple/opt.cpp:25-27,286-305."""
    return np.array([1.0, scale * SIGMA_X, scale * SIGMA_P, INITIAL_NOISE])


def theta_complex(scale: float = 1.0) -> np.ndarray:
    """(sigma, sigma_R, l_Rx, l_Rp, sigma_I, l_Ix, l_Ip, sigma_n), gple/opt.cpp:306-332."""
    return np.array([1.0, 1.0, scale * SIGMA_X, scale * SIGMA_P, 1.0, scale * SIGMA_X, scale * SIGMA_P, INITIAL_NOISE])


def points_aos(X: np.ndarray, y: np.ndarray) -> np.ndarray:
    """(n, 4) float64 rows (x, p, Re rho, Im rho): the 32-byte PhaseSpacePoint of gple/storage.h:232-297."""
    out = np.empty((len(X), 4))
    out[:, :2] = X
    out[:, 2] = y.real
    out[:, 3] = y.imag
    return out
