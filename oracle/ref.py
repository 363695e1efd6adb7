"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libgple_ref_{sac,dac,ecr}.so: the reference's OWN
hot-path translation units (gple/kernel.cpp, complex_kernel.cpp, pes.cpp, evolve.cpp, predict.cpp, mc.cpp), compiled
unmodified by oracle/Makefile.ref against the stand-in headers of oracle/refstub/ ("Oracle").

This is what pins the oracle restatement (oracle/oracle.py) to the reference: tests/test_ref_pins_oracle.py runs both on
the same inputs.  The wrappers are oracle.py's own classes and functions, re-bound to the `ref_*` entry points, so a test
written for one backend runs on the other.  Only tests/, smoke() and bench.py's CPU-baseline legs may import this module.

Entry points the reference has no public counterpart for (Metropolis walks on Philox streams, NLML, backward query
geometry, raw observable sums, the private label / dK^-1 members) are not provided and raise AttributeError.
"""
from __future__ import annotations

import ctypes as C
import glob
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SRC = "/root/reference/gaussian_process_liouville_equation"
_NAMES = {0: "sac", 1: "dac", 2: "ecr"}
SAC, DAC, ECR = 0, 1, 2


def lib_path(model: int) -> str:
    return os.path.join(_HERE, "_ref", f"libgple_ref_{_NAMES[int(model)]}.so")


def available() -> bool:
    return all(os.path.exists(lib_path(m)) for m in _NAMES)


def build() -> bool:
    """Compile oracle/_ref when the reference sources are present (this container); on the GPU box the prebuilt
    libraries travel with the snapshot and this is a no-op.  Returns available()."""
    if os.path.exists(os.path.join(_REF_SRC, "kernel.cpp")):
        subprocess.check_call(["make", "-f", "Makefile.ref", "-C", _HERE, "-j8"], stdout=subprocess.DEVNULL)
    return available()


def _blas_path() -> str:
    for pkg in ("scipy", "numpy"):
        spec = importlib.util.find_spec(pkg)
        if spec is None or not spec.submodule_search_locations:
            continue
        root = os.path.dirname(list(spec.submodule_search_locations)[0])
        hits = sorted(glob.glob(os.path.join(root, f"{pkg}.libs", "libscipy_openblas-*.so")))
        if hits:
            return hits[0]
    return ""


class _Proxy:
    """Maps oracle.py's `orc_name` calls onto `ref_name` of one model's library."""

    def __init__(self, cdll):
        self._l = cdll

    def __getattr__(self, name):
        if name.startswith("orc_"):
            fn = getattr(self._l, "ref_" + name[4:])
            if name in ("orc_train_real", "orc_train_complex"):
                fn.restype = C.c_void_p
            return fn
        raise AttributeError(name)


_backends = {}


def backend(model: int = DAC):
    """A private copy of the oracle.py module whose `lib()` is the compiled reference for `model`."""
    model = int(model)
    if model not in _backends:
        if not os.path.exists(lib_path(model)):
            raise RuntimeError(f"{lib_path(model)} is missing: run `make -f Makefile.ref -C oracle` where /root/reference exists")
        # the stand-in BLAS of the Eigen stub: scipy's bundled OpenBLAS (the reference links MKL)
        os.environ.setdefault("REFSTUB_BLAS", _blas_path())
        cdll = C.CDLL(lib_path(model))
        assert cdll.ref_model() == model
        spec = importlib.util.spec_from_file_location(f"oracle._ref_backend_{_NAMES[model]}", os.path.join(_HERE, "oracle.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        proxy = _Proxy(cdll)
        mod.lib = lambda: proxy
        mod.model_id = model
        mod.cdll = cdll
        _backends[model] = mod
    return _backends[model]


def have_blas(model: int = DAC) -> bool:
    return bool(backend(model).cdll.ref_have_blas())


# model-independent pieces (kernel.cpp, complex_kernel.cpp): any of the three libraries serves
def kernel_real(*a, **k):
    return backend(DAC).kernel_real(*a, **k)


def kernel_complex(*a, **k):
    return backend(DAC).kernel_complex(*a, **k)


def TrainingKernel(*a, **k):
    return backend(DAC).TrainingKernel(*a, **k)


def TrainingComplexKernel(*a, **k):
    return backend(DAC).TrainingComplexKernel(*a, **k)


def initial_distribution(*a, **k):
    return backend(DAC).initial_distribution(*a, **k)


# model-dependent pieces (pes.cpp, evolve.cpp, predict.cpp): dispatch on the model argument.  Element models trained by
# one library are plain heap objects of model-independent classes, so they may be handed to another model's evolve.
def pes(model, x):
    return backend(model).pes(model, x)


def evolve(model, *a, **k):
    return backend(model).evolve(model, *a, **k)


def new_point_predict(model, *a, **k):
    return backend(model).new_point_predict(model, *a, **k)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_dp = C.POINTER(C.c_double)


def observables(model, pts, mass, pes_index):
    """predict.cpp:65-244 on one diagonal element: dict(x, p, std_x, std_p, energy, purity_sum)."""
    pts = _f64(pts)
    out = np.empty(6)
    rc = backend(model).cdll.ref_observables(int(model), pts.ctypes.data_as(_dp), C.c_size_t(len(pts)), C.c_double(mass), int(pes_index), out.ctypes.data_as(_dp))
    assert rc == 0
    return dict(x=out[0], p=out[1], std_x=out[2], std_p=out[3], energy=out[4], purity_sum=out[5])


def training_kernels(theta16, pts00, pts10, pts11, energies):
    """TrainingKernels(params, density) (predict.cpp:362-463): dict(population, x, p, energy, purity)."""
    arrs = [np.zeros((0, 4)) if a is None else _f64(a) for a in (pts00, pts10, pts11)]
    theta16, energies = _f64(theta16), _f64(energies)
    out = np.empty(5)
    backend(DAC).cdll.ref_training_kernels(theta16.ctypes.data_as(_dp), arrs[0].ctypes.data_as(_dp), C.c_size_t(len(arrs[0])), arrs[1].ctypes.data_as(_dp), C.c_size_t(len(arrs[1])),
                                           arrs[2].ctypes.data_as(_dp), C.c_size_t(len(arrs[2])), energies.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
    return dict(population=out[0], x=out[1], p=out[2], energy=out[3], purity=out[4])


def is_very_small(model, pts00, pts10, pts11, mass, dt, k00=None, k10=None, k11=None):
    """evolve.cpp:444-478: flags of (0,0), (1,0), (1,1)."""
    arrs = [np.zeros((0, 4)) if a is None else _f64(a) for a in (pts00, pts10, pts11)]
    out = (C.c_int * 3)()
    h = [None if k is None else k.h for k in (k00, k10, k11)]
    rc = backend(model).cdll.ref_is_very_small(int(model), arrs[0].ctypes.data_as(_dp), C.c_size_t(len(arrs[0])), arrs[1].ctypes.data_as(_dp), C.c_size_t(len(arrs[1])),
                                               arrs[2].ctypes.data_as(_dp), C.c_size_t(len(arrs[2])), C.c_double(mass), C.c_double(dt), h[0], h[1], h[2], out)
    assert rc == 0
    return [bool(v) for v in out]
