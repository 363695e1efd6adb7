"""Host-side mirror of the reference's gple/complex_kernel.h over the C-ABI.

Reference classes: ComplexKernelBase (complex_kernel.h:14-145), TrainingComplexKernel (:150-318),
PredictiveComplexKernel (:323-391).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

NumTotalParameters = 8  # complex_kernel.h:22


def kernel_matrices(left_feature, right_feature, parameter, same_set: bool, ctx=None):
    """ComplexKernelBase::ComplexKernelBase (complex_kernel.cpp:134-200): K (real) and Kt (complex), (nL, nR)."""
    ctx = ctx or L.default_context()
    XL, XR, th = L.f64(left_feature), L.f64(right_feature), L.f64(parameter)
    nL, nR = len(XL), len(XR)
    K = np.empty((nR, nL))
    Kt = np.empty((nR, nL), dtype=np.complex128)
    ctx.check(ctx.lib.gple_kernel_complex(ctx.h, L.addr(XL), nL, L.addr(XR), nR, L.addr(th), int(same_set), L.addr(K), L.addr(Kt)))
    return K.T, Kt.T


def kernel_derivatives(left_feature, right_feature, parameter, same_set: bool, ctx=None):
    """calculate_derivative / calculate_pseudo_derivative (complex_kernel.cpp:20-59, 74-132): dK (8, nL, nR) real, dKt (8, nL, nR) complex."""
    ctx = ctx or L.default_context()
    XL, XR, th = L.f64(left_feature), L.f64(right_feature), L.f64(parameter)
    nL, nR = len(XL), len(XR)
    dK = np.empty((8, nR, nL))
    dKt = np.empty((8, nR, nL), dtype=np.complex128)
    ctx.check(ctx.lib.gple_kernel_complex_derivatives(ctx.h, L.addr(XL), nL, L.addr(XR), nR, L.addr(th), int(same_set), L.addr(dK), L.addr(dKt)))
    return dK.transpose(0, 2, 1), dKt.transpose(0, 2, 1)


class TrainingComplexKernel:
    def __init__(self, Parameter, TrainingSet, IsToCalculateError=True, IsToCalculateAverage=True, IsToCalculateDerivative=False, ctx=None):
        self.ctx = ctx or L.default_context()
        feature, label = TrainingSet
        self._params = L.f64(Parameter)
        assert self._params.shape == (NumTotalParameters,)
        self._X, self._y = L.f64(feature), L.c128(label)
        self.N = len(self._X)
        flags = (L.CALC_ERROR if IsToCalculateError else 0) | (L.CALC_AVERAGE if IsToCalculateAverage else 0) | (L.CALC_DERIVATIVE if IsToCalculateDerivative else 0)
        self._flags = flags
        h, s = C.c_void_p(), L.ComplexScalars()
        self.status = self.ctx.check(self.ctx.lib.gple_train_complex(self.ctx.h, L.addr(self._X), L.addr(self._y), self.N, L.addr(self._params), flags, C.byref(h), C.byref(s)), allow=(L.ERR_NOT_SPD,))
        self.h, self._s = h, s

    def get_parameters(self):
        return self._params.copy()

    def get_rescale_factor(self):
        return self._s.rescale

    def get_magnitude(self):
        return self._s.magnitude

    def get_negative_log_marginal_likelihood(self, grad: bool = False):
        """NLML / LLT objective of test/gpr.cpp:470-532 on this model (value, or (value, gradient[8]))."""
        v = C.c_double()
        g = np.empty(8) if grad else None
        self.ctx.check(self.ctx.lib.gple_model_nlml(self.ctx.h, self.h, C.byref(v), L.addr(g) if grad else None))
        return (v.value, g) if grad else v.value

    def get_error(self):
        assert self._flags & L.CALC_ERROR
        return self._s.error

    def get_purity(self):
        assert self._flags & L.CALC_AVERAGE
        return self._s.purity

    def get_error_derivative(self):
        assert self._flags & L.CALC_ERROR and self._flags & L.CALC_DERIVATIVE
        return np.array(self._s.d_error[:])

    def get_purity_derivative(self):
        assert self._flags & L.CALC_AVERAGE and self._flags & L.CALC_DERIVATIVE
        return np.array(self._s.d_purity[:])

    def _field(self, which, shape):
        out = np.empty(shape, dtype=np.complex128)
        self.ctx.check(self.ctx.lib.gple_model_get(self.ctx.h, self.h, which, L.addr(out)))
        return out

    def get_upper_left_block_of_augmented_inverse(self):
        return self._field(L.FIELD_UPPER_LEFT, (self.N, self.N)).T

    def get_lower_left_block_of_augmented_inverse(self):
        return self._field(L.FIELD_LOWER_LEFT, (self.N, self.N)).T

    def get_upper_part_of_augmented_inverse_times_label(self):
        return self._field(L.FIELD_INV_LABEL, (self.N,))

    def get_label(self):
        return self._field(L.FIELD_LABEL, (self.N,))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.ctx.lib.gple_model_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PredictiveComplexKernel:
    def __init__(self, TestFeature, kernel: TrainingComplexKernel, IsToCalculateDerivative=False, TestLabel=None):
        ctx = kernel.ctx
        Xq = L.f64(TestFeature).reshape(-1, 2)
        M = len(Xq)
        yq = None if TestLabel is None else L.c128(TestLabel)
        self._pred, self._cut = np.empty(M, dtype=np.complex128), np.empty(M, dtype=np.complex128)
        self._var = np.empty(M)
        err = np.full(1, np.nan)
        derr = np.full(NumTotalParameters, np.nan)
        want_grad = IsToCalculateDerivative and yq is not None
        ctx.check(ctx.lib.gple_predict_complex(ctx.h, kernel.h, L.addr(Xq), M, L.addr(yq), L.addr(self._pred), L.addr(self._var), L.addr(self._cut),
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
