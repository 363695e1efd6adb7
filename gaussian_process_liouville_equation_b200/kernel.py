"""Host-side mirror of the reference's gple/kernel.h over the C-ABI (same names and getter semantics).

Reference classes: KernelBase (kernel.h:29-106), TrainingKernel (kernel.h:111-280), PredictiveKernel
(kernel.h:336-403).  Features are (n, 2) arrays of (x, p) -- the memory layout of the reference's
column-major ``PhasePoints`` (stdafx.h:153).  All arithmetic happens in libgple_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

NumTotalParameters = 4  # kernel.h:33
RescaleMaximum = 10.0  # kernel.h:37


def kernel_matrix(left_feature, right_feature, parameter, same_set: bool, derivative: bool = False, ctx=None):
    """KernelBase::KernelBase (kernel.cpp:217-242): returns K (nL, nR) [and dK (4, nL, nR)]."""
    ctx = ctx or L.default_context()
    XL, XR, th = L.f64(left_feature), L.f64(right_feature), L.f64(parameter)
    nL, nR = len(XL), len(XR)
    K = np.empty((nR, nL))
    dK = np.empty((4, nR, nL)) if derivative else None
    ctx.check(ctx.lib.gple_kernel_real(ctx.h, L.addr(XL), nL, L.addr(XR), nR, L.addr(th), int(same_set), L.addr(K), L.addr(dK)))
    return (K.T, dK.transpose(0, 2, 1)) if derivative else K.T


class TrainingKernel:
    """kernel.h:111-280.  `TrainingSet` = (feature (N, 2), label complex (N,))."""

    def __init__(self, Parameter, TrainingSet, IsToCalculateError=True, IsToCalculateAverage=True, IsToCalculateDerivative=False, ctx=None):
        self.ctx = ctx or L.default_context()
        feature, label = TrainingSet
        self._params = L.f64(Parameter)
        assert self._params.shape == (NumTotalParameters,)
        self._X, self._y = L.f64(feature), L.c128(label)
        self.N = len(self._X)
        flags = (L.CALC_ERROR if IsToCalculateError else 0) | (L.CALC_AVERAGE if IsToCalculateAverage else 0) | (L.CALC_DERIVATIVE if IsToCalculateDerivative else 0)
        self._flags = flags
        h, s = C.c_void_p(), L.RealScalars()
        self.status = self.ctx.check(self.ctx.lib.gple_train_real(self.ctx.h, L.addr(self._X), L.addr(self._y), self.N, L.addr(self._params), flags, C.byref(h), C.byref(s)), allow=(L.ERR_NOT_SPD,))
        self.h, self._s = h, s

    # --- getters, kernel.h:136-243 ---
    def get_parameters(self):
        return self._params.copy()

    def get_left_feature(self):
        return self._X

    def get_rescale_factor(self):
        return self._s.rescale

    def get_magnitude(self):
        return self._s.magnitude

    def get_negative_log_marginal_likelihood(self, grad: bool = False):
        """NLML / LLT objective of test/gpr.cpp:470-532 on this model (value, or (value, gradient[4]))."""
        v = C.c_double()
        g = np.empty(4) if grad else None
        self.ctx.check(self.ctx.lib.gple_model_nlml(self.ctx.h, self.h, C.byref(v), L.addr(g) if grad else None))
        return (v.value, g) if grad else v.value

    def _need(self, flag, what):
        assert self._flags & flag, f"{what} was not requested at construction (reference asserts has_value())"

    def get_error(self):
        self._need(L.CALC_ERROR, "error")
        return self._s.error

    def get_population(self):
        self._need(L.CALC_AVERAGE, "population")
        return self._s.population

    def get_1st_order_average(self):
        self._need(L.CALC_AVERAGE, "first order average")
        return np.array(self._s.first_order[:])

    def get_purity(self):
        self._need(L.CALC_AVERAGE, "purity")
        return self._s.purity

    def get_error_derivative(self):
        self._need(L.CALC_ERROR | L.CALC_DERIVATIVE, "error derivative")
        return np.array(self._s.d_error[:])

    def get_population_derivative(self):
        self._need(L.CALC_AVERAGE | L.CALC_DERIVATIVE, "population derivative")
        return np.array(self._s.d_population[:])

    def get_purity_derivative(self):
        self._need(L.CALC_AVERAGE | L.CALC_DERIVATIVE, "purity derivative")
        return np.array(self._s.d_purity[:])

    def _field(self, which, shape):
        out = np.empty(shape)
        self.ctx.check(self.ctx.lib.gple_model_get(self.ctx.h, self.h, which, L.addr(out)))
        return out

    def get_inverse(self):
        return self._field(L.FIELD_INVERSE, (self.N, self.N)).T

    def get_inverse_times_label(self):
        return self._field(L.FIELD_INV_LABEL, (self.N,))

    def get_label(self):
        return self._field(L.FIELD_LABEL, (self.N,))

    def get_inverse_times_label_derivative(self):
        self._need(L.CALC_DERIVATIVE, "derivative of inverse times label")
        return [self._field(L.FIELD_INV_LABEL_DERIV + p, (self.N,)) for p in range(NumTotalParameters)]

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.ctx.lib.gple_model_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PredictiveKernel:
    """kernel.h:336-403: batched prediction of `TestFeature` (M, 2) with a TrainingKernel."""

    def __init__(self, TestFeature, kernel: TrainingKernel, IsToCalculateDerivative=False, TestLabel=None):
        ctx = kernel.ctx
        Xq = L.f64(TestFeature).reshape(-1, 2)
        M = len(Xq)
        yq = None if TestLabel is None else L.f64(TestLabel)
        self._pred, self._var, self._cut = np.empty(M), np.empty(M), np.empty(M)
        err = np.full(1, np.nan)
        derr = np.full(NumTotalParameters, np.nan)
        want_grad = IsToCalculateDerivative and yq is not None
        ctx.check(ctx.lib.gple_predict_real(ctx.h, kernel.h, L.addr(Xq), M, L.addr(yq), L.addr(self._pred), L.addr(self._var), L.addr(self._cut),
                                            L.addr(err) if yq is not None else None, L.addr(derr) if want_grad else None))
        self._err, self._derr, self._has_label, self._has_grad = err[0], derr, yq is not None, want_grad

    def get_prediction(self):
        return self._pred

    def get_variance(self):
        return self._var

    def get_cutoff_prediction(self):
        return self._cut

    def get_error(self):
        assert self._has_label
        return self._err

    def get_error_derivative(self):
        assert self._has_grad
        return self._derr
