"""Host-side mirror of the reference's gple/pes.h, gple/evolve.h and the observables of gple/predict.h.

Point sets are (n, 4) float64 arrays of (x, p, Re rho, Im rho): the 32-byte ``PhaseSpacePoint`` of
gple/storage.h:232-297.  Elements are addressed in lower-triangular order (rho00, rho10, rho11).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L


def _h(k):
    return None if k is None else k.h


def adiabatic_pes(model: int, x, ctx=None):
    """adiabatic_potential / force / coupling (pes.cpp:127-189): E (n, 2), F (n, 3: F00, F10, F11), d10 (n,)."""
    ctx = ctx or L.default_context()
    x = L.f64(x)
    n = len(x)
    E, F, D = np.empty((n, 2)), np.empty((n, 3)), np.empty(n)
    ctx.check(ctx.lib.gple_pes(ctx.h, int(model), L.addr(x), n, L.addr(E), L.addr(F), L.addr(D)))
    return E, F, D


def evolve(model: int, density, mass: float, dt: float, kernels, ctx=None):
    """evolve() (evolve.cpp:377-423) with the GPR-backed predict_distribution of main.cpp:75-101.

    density: [pts00, pts10, pts11] (None / empty for an unpopulated element); kernels: [k00, k10, k11]
    (TrainingKernel, TrainingComplexKernel, TrainingKernel or None).  Returns the evolved copies.
    """
    ctx = ctx or L.default_context()
    outs = [np.zeros((0, 4)) if a is None else L.f64(a).copy() for a in density]
    ctx.check(ctx.lib.gple_evolve(ctx.h, int(model), _h(kernels[0]), _h(kernels[1]), _h(kernels[2]),
                                  L.addr(outs[0]) if len(outs[0]) else None, len(outs[0]),
                                  L.addr(outs[1]) if len(outs[1]) else None, len(outs[1]),
                                  L.addr(outs[2]) if len(outs[2]) else None, len(outs[2]), float(mass), float(dt)))
    return outs


def evolve_sharded(model: int, density, mass: float, dt: float, kernels, ctx=None):
    """evolve() on G GPUs (one process each): `density` holds the FULL point sets on every rank; every rank evolves its own block
    and the blocks are all-gathered in place inside the library (gple_evolve_sharded).  Returns the full evolved sets."""
    ctx = ctx or L.default_context()
    outs = [np.zeros((0, 4)) if a is None else L.f64(a).copy() for a in density]
    ctx.check(ctx.lib.gple_evolve_sharded(ctx.h, int(model), _h(kernels[0]), _h(kernels[1]), _h(kernels[2]),
                                          L.addr(outs[0]) if len(outs[0]) else None, len(outs[0]),
                                          L.addr(outs[1]) if len(outs[1]) else None, len(outs[1]),
                                          L.addr(outs[2]) if len(outs[2]) else None, len(outs[2]), float(mass), float(dt)))
    return outs


def new_point_predict(model: int, r, mass: float, dt: float, kernels, row: int, col: int, ctx=None):
    """new_point_predict() (evolve.cpp:425-443) for n points r (n, 2); returns complex (n,)."""
    ctx = ctx or L.default_context()
    r = L.f64(r)
    out = np.empty(len(r), dtype=np.complex128)
    ctx.check(ctx.lib.gple_new_point_predict(ctx.h, int(model), _h(kernels[0]), _h(kernels[1]), _h(kernels[2]), L.addr(r), len(r), int(row), int(col), float(mass), float(dt), L.addr(out)))
    return out


def observable_sums(model: int, pts, mass: float, pes_index: int, ctx=None):
    """The nine running sums behind predict.cpp:65-244 (see include/gple_b200.h: gple_observables)."""
    ctx = ctx or L.default_context()
    pts = L.f64(pts)
    out = np.empty(9)
    ctx.check(ctx.lib.gple_observables(ctx.h, int(model), L.addr(pts), len(pts), float(mass), int(pes_index), L.addr(out)))
    return out


def calculate_population_each_surface(model, density, mass, ctx=None):
    """predict.cpp:65-87"""
    pop = np.array([observable_sums(model, density[i], mass, j, ctx)[0] if density[i] is not None and len(density[i]) else 0.0 for i, j in ((0, 0), (2, 1))])
    return pop / pop.sum()


def calculate_1st_order_average_one_surface(model, pts, mass, ctx=None):
    """predict.cpp:89-107"""
    s = observable_sums(model, pts, mass, 0, ctx)
    return s[1:3] / s[0]


def calculate_standard_deviation_one_surface(model, pts, mass, ctx=None):
    """predict.cpp:109-126"""
    s = observable_sums(model, pts, mass, 0, ctx)
    n = len(pts)
    return np.sqrt(s[5:7] / n - (s[3:5] / n) ** 2)


def calculate_total_energy_average_one_surface(model, pts, mass, pes_index, ctx=None):
    """predict.cpp:157-180"""
    s = observable_sums(model, pts, mass, pes_index, ctx)
    return s[7] / s[0]


def calculate_purity_one_element(model, pts, mass, ctx=None):
    """predict.cpp:224-244 (one element)"""
    return observable_sums(model, pts, mass, 0, ctx)[8]


def loose_function(x, TrainingSet, ExtraTrainingSet, grad: bool = False, ctx=None):
    """loose_function (opt.cpp:441-482) incl. make_normal (opt.cpp:420-431)."""
    import ctypes as C

    ctx = ctx or L.default_context()
    x = L.f64(x)
    X, y = L.f64(TrainingSet[0]), L.c128(TrainingSet[1])
    Xe, ye = L.f64(ExtraTrainingSet[0]), L.c128(ExtraTrainingSet[1])
    g = np.empty(len(x)) if grad else None
    val = np.empty(1)
    rc = ctx.check(ctx.lib.gple_loose_function(ctx.h, L.addr(x), len(x), L.addr(g), L.addr(X), L.addr(y), len(X), L.addr(Xe), L.addr(ye), len(Xe), L.addr(val)), allow=(L.ERR_NOT_SPD,))
    big = np.finfo(np.float64).max
    v = val[0] if (rc == L.OK and np.isfinite(val[0])) else big
    if grad:
        g = np.where(np.isfinite(g), g, big)
        return v, g
    return v


def validation_error(kernel, ExtraTrainingSet, grad: bool = False, ctx=None):
    """The validation half of loose_function (opt.cpp:455-470) on an already trained model: (error[, gradient])."""
    import ctypes as C

    ctx = ctx or L.default_context()
    Xe, ye = L.f64(ExtraTrainingSet[0]), L.c128(ExtraTrainingSet[1])
    err = C.c_double()
    g = np.empty(len(kernel.get_parameters())) if grad else None
    ctx.check(ctx.lib.gple_validation_error(ctx.h, kernel.h, L.addr(Xe), L.addr(ye), len(Xe), C.byref(err), L.addr(g)))
    return (err.value, g) if grad else err.value
