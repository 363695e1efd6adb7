"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/gple_oracle_c.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The oracle is a CPU restatement of the reference
(kaigu1997/gaussian_process_liouville_equation, gaussian_process_liouville_equation/*.cpp).  The reference
holds no golden vectors; the oracle is pinned by the reference's own translation units compiled unmodified
into oracle/_ref (oracle/ref.py, tests/test_ref_pins_oracle.py) and by tests/test_oracle_*.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgple_oracle.so")

SAC, DAC, ECR = 0, 1, 2
_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp"))]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_train_real.restype = C.c_void_p
        L.orc_train_complex.restype = C.c_void_p
        L.orc_loose_function.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def num_threads() -> int:
    return lib().orc_num_threads()


def kernel_real(XL, XR, theta, same: bool, deriv: bool):
    """XL, XR: (n, 2) arrays of (x, p).  Returns K (nL, nR) [and dK (4, nL, nR)]."""
    XL, XR, theta = _f64(XL), _f64(XR), _f64(theta)
    nL, nR = len(XL), len(XR)
    K = np.empty((nR, nL))
    dK = np.empty((4, nR, nL)) if deriv else None
    lib().orc_kernel_real(_p(XL), C.c_size_t(nL), _p(XR), C.c_size_t(nR), _p(theta), int(same), int(deriv), _p(K), _p(dK))
    return (K.T, dK.transpose(0, 2, 1)) if deriv else K.T


def kernel_complex(XL, XR, theta, same: bool, deriv: bool):
    XL, XR, theta = _f64(XL), _f64(XR), _f64(theta)
    nL, nR = len(XL), len(XR)
    K = np.empty((nR, nL))
    Kt = np.empty((nR, nL), dtype=np.complex128)
    dK = np.empty((8, nR, nL)) if deriv else None
    dKt = np.empty((8, nR, nL), dtype=np.complex128) if deriv else None
    lib().orc_kernel_complex(_p(XL), C.c_size_t(nL), _p(XR), C.c_size_t(nR), _p(theta), int(same), int(deriv), _p(K), _p(Kt), _p(dK), _p(dKt))
    if deriv:
        return K.T, Kt.T, dK.transpose(0, 2, 1), dKt.transpose(0, 2, 1)
    return K.T, Kt.T


class TrainingKernel:
    """gple/kernel.cpp:244-479 (oracle).  X: (N, 2); y: complex (N,) (imaginary part ignored)."""

    def __init__(self, theta, X, y, err=True, avg=True, deriv=False):
        self.theta, self.X, self.y = _f64(theta), _f64(X), _c128(y)
        self.N = len(self.X)
        self.h = C.c_void_p(lib().orc_train_real(_p(self.theta), _p(self.X), self.y.ctypes.data_as(_dp), C.c_size_t(self.N), int(err), int(avg), int(deriv)))
        s = np.empty(19)
        lib().orc_train_real_scalars(self.h, _p(s))
        self.rescale, self.error, self.population = s[0], s[1], s[2]
        self.first_order = s[3:5].copy()
        self.purity, self.magnitude = s[5], s[6]
        self.derror, self.dpopulation, self.dpurity = s[7:11].copy(), s[11:15].copy(), s[15:19].copy()

    def _get(self, which, shape):
        out = np.empty(shape)
        if lib().orc_train_real_get(self.h, which, _p(out)) != 0:
            raise ValueError("quantity %d was not computed" % which)
        return out

    @property
    def K(self):
        return self._get(0, (self.N, self.N)).T

    @property
    def inverse(self):
        return self._get(1, (self.N, self.N)).T

    @property
    def v(self):
        return self._get(2, (self.N,))

    @property
    def label(self):
        return self._get(3, (self.N,))

    def dv(self, p):
        return self._get(4 + p, (self.N,))

    def dK(self, p):
        return self._get(8 + p, (self.N, self.N)).T

    def dinv(self, p):
        return self._get(12 + p, (self.N, self.N)).T

    def predict(self, Xq, yq=None, deriv=False):
        """gple/kernel.cpp:481-544.  Returns dict(pred, var, cutoff, error, derror)."""
        Xq = _f64(Xq)
        Q = len(Xq)
        pred, var, cut = np.empty(Q), np.empty(Q), np.empty(Q)
        err, derr = np.full(1, np.nan), np.full(4, np.nan)
        yq = None if yq is None else _f64(yq)
        lib().orc_predict_real(self.h, _p(Xq), C.c_size_t(Q), _p(yq), int(deriv), _p(pred), _p(var), _p(cut), _p(err), _p(derr))
        return dict(pred=pred, var=var, cutoff=cut, error=err[0], derror=derr)

    def nlml(self, grad=False):
        """test/gpr.cpp:470-532 on this model's kernel and rescaled labels: value, or (value, grad[4])."""
        lib().orc_nlml_real.restype = C.c_double
        g = np.empty(4) if grad else None
        v = lib().orc_nlml_real(self.h, _p(g))
        return (v, g) if grad else v

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_free_real(self.h)
            self.h = None


class TrainingComplexKernel:
    """gple/complex_kernel.cpp:221-592 (oracle)."""

    def __init__(self, theta, X, y, err=True, avg=True, deriv=False):
        self.theta, self.X, self.y = _f64(theta), _f64(X), _c128(y)
        self.N = len(self.X)
        self.h = C.c_void_p(lib().orc_train_complex(_p(self.theta), _p(self.X), self.y.ctypes.data_as(_dp), C.c_size_t(self.N), int(err), int(avg), int(deriv)))
        s = np.empty(20)
        lib().orc_train_complex_scalars(self.h, _p(s))
        self.rescale, self.error, self.purity, self.magnitude = s[0], s[1], s[2], s[3]
        self.derror, self.dpurity = s[4:12].copy(), s[12:20].copy()

    def _get(self, which, shape, cplx):
        out = np.empty(shape, dtype=np.complex128 if cplx else np.float64)
        if lib().orc_train_complex_get(self.h, which, out.ctypes.data_as(_dp)) != 0:
            raise ValueError("quantity %d was not computed" % which)
        return out

    @property
    def K(self):
        return self._get(0, (self.N, self.N), False).T

    @property
    def Kt(self):
        return self._get(1, (self.N, self.N), True).T

    @property
    def P(self):
        return self._get(2, (self.N, self.N), True).T

    @property
    def Q(self):
        return self._get(3, (self.N, self.N), True).T

    @property
    def v(self):
        return self._get(4, (self.N,), True)

    @property
    def label(self):
        return self._get(5, (self.N,), True)

    def dv(self, p):
        return self._get(8 + p, (self.N,), True)

    def predict(self, Xq, yq=None, deriv=False):
        Xq = _f64(Xq)
        Q = len(Xq)
        pred, cut = np.empty(Q, dtype=np.complex128), np.empty(Q, dtype=np.complex128)
        var = np.empty(Q)
        err, derr = np.full(1, np.nan), np.full(8, np.nan)
        yq = None if yq is None else _c128(yq)
        lib().orc_predict_complex(self.h, _p(Xq), C.c_size_t(Q), None if yq is None else yq.ctypes.data_as(_dp), int(deriv), pred.ctypes.data_as(_dp), _p(var), cut.ctypes.data_as(_dp), _p(err), _p(derr))
        return dict(pred=pred, var=var, cutoff=cut, error=err[0], derror=derr)

    def nlml(self, grad=False):
        """NLML of the composite [Re f; Im f] process (oracle/gple_oracle_nlml.hpp): value, or (value, grad[8])."""
        lib().orc_nlml_complex.restype = C.c_double
        g = np.empty(8) if grad else None
        v = lib().orc_nlml_complex(self.h, _p(g))
        return (v, g) if grad else v

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_free_complex(self.h)
            self.h = None


def loose_function(x, X, y, Xe, ye, grad=False):
    """gple/opt.cpp:441-482.  Returns loss or (loss, grad)."""
    x, X, y, Xe, ye = _f64(x), _f64(X), _c128(y), _f64(Xe), _c128(ye)
    g = np.empty(len(x)) if grad else None
    val = lib().orc_loose_function(_p(x), len(x), _p(g), _p(X), y.ctypes.data_as(_dp), C.c_size_t(len(X)), _p(Xe), ye.ctypes.data_as(_dp), C.c_size_t(len(Xe)))
    return (val, g) if grad else val


def pes(model, x):
    """Adiabatic E (n,2), F (n,3: F00,F10,F11), d10 (n,).  gple/pes.cpp:127-189."""
    x = _f64(x)
    n = len(x)
    E, F, D = np.empty((n, 2)), np.empty((n, 3)), np.empty(n)
    lib().orc_pes(int(model), _p(x), C.c_size_t(n), _p(E), _p(F), _p(D))
    return E, F, D


def _h(k):
    return None if k is None else k.h


def evolve(model, pts00, pts10, pts11, mass, dt, k00=None, k10=None, k11=None, analytic=None):
    """gple/evolve.cpp:377-423.  pts: (n, 4) arrays (x, p, Re rho, Im rho) or None; returns evolved copies."""
    outs = []
    for a in (pts00, pts10, pts11):
        outs.append(np.zeros((0, 4)) if a is None else _f64(a).copy())
    an = None if analytic is None else _f64(analytic)
    lib().orc_evolve(int(model), _p(outs[0]), C.c_size_t(len(outs[0])), _p(outs[1]), C.c_size_t(len(outs[1])), _p(outs[2]), C.c_size_t(len(outs[2])),
                     C.c_double(mass), C.c_double(dt), _h(k00), _h(k10), _h(k11), _p(an))
    return outs


def backward_queries(model, x, p, mass, dt, row, col):
    out = np.empty((3, 3, 2))
    lib().orc_backward_queries(int(model), C.c_double(x), C.c_double(p), C.c_double(mass), C.c_double(dt), int(row), int(col), _p(out))
    return out


def new_point_predict(model, r, mass, dt, row, col, k00=None, k10=None, k11=None):
    r = _f64(r)
    out = np.empty(len(r), dtype=np.complex128)
    lib().orc_new_point_predict(int(model), _p(r), C.c_size_t(len(r)), C.c_double(mass), C.c_double(dt), int(row), int(col), _h(k00), _h(k10), _h(k11), out.ctypes.data_as(_dp))
    return out


def observable_sums(model, pts, mass, pes_index):
    pts = _f64(pts)
    out = np.empty(9)
    lib().orc_observable_sums(int(model), _p(pts), C.c_size_t(len(pts)), C.c_double(mass), int(pes_index), _p(out))
    return out


def initial_distribution(analytic, r, row, col):
    analytic, r = _f64(analytic), _f64(r)
    out = np.empty(len(r), dtype=np.complex128)
    lib().orc_initial_distribution(_p(analytic), _p(r), C.c_size_t(len(r)), int(row), int(col), out.ctypes.data_as(_dp))
    return out


def markov_chains(pts, num_steps, max_displacement, seed, stream, row, col, analytic=None, k00=None, k10=None, k11=None, new_point=None, want_chain=False, chain0=0):
    """gple/mc.cpp:143-188 for every point of pts (n, 4) at once, chain i on Philox stream (seed, stream, i).
    Distribution: analytic[8] (initial_distribution), else predict_distribution of the models, else -- with
    new_point = (model, mass, dt) -- new_point_predict.  Returns (pts_out (n, 4), accept (n,), chains (n, steps + 1, 2) | None)."""
    pts = np.array(_f64(pts), copy=True)
    n = len(pts)
    kind = 0 if analytic is not None else (2 if new_point is not None else 1)
    model, mass, dt = new_point if new_point is not None else (0, 1.0, 0.0)
    accept = np.empty(n)
    chains = np.empty((n, num_steps + 1, 2)) if want_chain else None
    lib().orc_markov_chains(kind, _p(None if analytic is None else _f64(analytic)), _h(k00), _h(k10), _h(k11), int(model), C.c_double(mass), C.c_double(dt), int(row), int(col), _p(pts),
                            C.c_size_t(n), C.c_size_t(num_steps), C.c_double(max_displacement), C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(chain0), _p(accept), _p(chains))
    return pts, accept, chains


def chain_autocorrelation(chains):
    """gple/mc.cpp:230-243: mean autocorrelation over the chains, chains (n, len, 2) -> (len // 2,)"""
    chains = _f64(chains)
    n, length = chains.shape[0], chains.shape[1]
    out = np.empty(length // 2)
    lib().orc_chain_autocorrelation(_p(chains), C.c_size_t(n), C.c_size_t(length), _p(out))
    return out


def philox_draws(seed, stream, chain, step):
    out = np.empty(3)
    lib().orc_philox_draws(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(chain), C.c_uint32(step), _p(out))
    return out
