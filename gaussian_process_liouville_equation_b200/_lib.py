"""ctypes binding of libgple_b200.so (the C-ABI declared in include/gple_b200.h).

There is deliberately no CPU fallback: if the CUDA library cannot be loaded, or no CUDA device is
present, every compute entry point raises.  torch is NOT required here; arguments may be numpy arrays
(host pointers, copied in/out by the library) or anything exposing ``data_ptr()`` (torch CUDA tensors,
passed through as device pointers).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgple_b200.so")

OK, ERR_ARG, ERR_NOT_SPD, ERR_CUDA, ERR_STATE, ERR_COMM = 0, 1, 2, 3, 4, 5
COMM_ID_BYTES = 128
CALC_ERROR, CALC_AVERAGE, CALC_DERIVATIVE = 1, 2, 4
SAC, DAC, ECR = 0, 1, 2
FIELD_INVERSE, FIELD_INV_LABEL, FIELD_LABEL, FIELD_UPPER_LEFT, FIELD_LOWER_LEFT, FIELD_INV_LABEL_DERIV = 1, 2, 3, 4, 5, 8

_vp, _dp, _sz = C.c_void_p, C.c_void_p, C.c_size_t  # double* passed as raw addresses (host or device)


class RealScalars(C.Structure):
    _fields_ = [("rescale", C.c_double), ("error", C.c_double), ("population", C.c_double), ("first_order", C.c_double * 2),
                ("purity", C.c_double), ("magnitude", C.c_double), ("d_error", C.c_double * 4), ("d_population", C.c_double * 4),
                ("d_purity", C.c_double * 4)]


class ComplexScalars(C.Structure):
    _fields_ = [("rescale", C.c_double), ("error", C.c_double), ("purity", C.c_double), ("magnitude", C.c_double),
                ("d_error", C.c_double * 8), ("d_purity", C.c_double * 8)]


class McSource(C.Structure):
    """gple_mc_source (include/gple_b200.h)"""
    _fields_ = [("kind", C.c_int), ("row", C.c_int), ("col", C.c_int), ("analytic", C.c_double * 8), ("m00", C.c_void_p), ("m10", C.c_void_p),
                ("m11", C.c_void_p), ("pes_model", C.c_int), ("mass", C.c_double), ("dt", C.c_double)]


MC_ANALYTIC, MC_PREDICT, MC_NEW_POINT = 0, 1, 2

# name -> (restype, argtypes); must list every symbol of include/gple_b200.h (tests/test_abi.py checks this)
SIGNATURES = {
    "gple_version": (C.c_char_p, []),
    "gple_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "gple_ctx_destroy": (C.c_int, [_vp]),
    "gple_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "gple_ctx_sync": (C.c_int, [_vp]),
    "gple_ctx_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "gple_ctx_set_gate_schedule": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "gple_gate_schedule_automatic": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int]),
    "gple_gate_statistics": (C.c_int, [_vp, C.POINTER(C.c_ulonglong)]),
    "gple_last_error": (C.c_char_p, [_vp]),
    "gple_launch_count": (C.c_ulonglong, [_vp]),
    "gple_kernel_real": (C.c_int, [_vp, _dp, _sz, _dp, _sz, _dp, C.c_int, _dp, _dp]),
    "gple_kernel_complex_derivatives": (C.c_int, [_vp, _dp, _sz, _dp, _sz, _dp, C.c_int, _dp, _dp]),
    "gple_kernel_complex": (C.c_int, [_vp, _dp, _sz, _dp, _sz, _dp, C.c_int, _dp, _dp]),
    "gple_train_real": (C.c_int, [_vp, _dp, _dp, _sz, _dp, C.c_uint, C.POINTER(_vp), C.POINTER(RealScalars)]),
    "gple_train_complex": (C.c_int, [_vp, _dp, _dp, _sz, _dp, C.c_uint, C.POINTER(_vp), C.POINTER(ComplexScalars)]),
    "gple_model_get": (C.c_int, [_vp, _vp, C.c_int, _dp]),
    "gple_markov_chains": (C.c_int, [_vp, C.POINTER(McSource), _dp, _sz, _sz, C.c_double, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, _dp, _dp]),
    "gple_chain_autocorrelation": (C.c_int, [_vp, _dp, _sz, _sz, _dp]),
    "gple_model_nlml": (C.c_int, [_vp, _vp, _dp, _dp]),
    "gple_model_is_complex": (C.c_int, [_vp]),
    "gple_model_size": (_sz, [_vp]),
    "gple_model_destroy": (C.c_int, [_vp, _vp]),
    "gple_predict_real": (C.c_int, [_vp, _vp, _dp, _sz, _dp, _dp, _dp, _dp, _dp, _dp]),
    "gple_predict_complex": (C.c_int, [_vp, _vp, _dp, _sz, _dp, _dp, _dp, _dp, _dp, _dp]),
    "gple_validation_error": (C.c_int, [_vp, _vp, _dp, _dp, _sz, _dp, _dp]),
    "gple_loose_function": (C.c_int, [_vp, _dp, C.c_int, _dp, _dp, _dp, _sz, _dp, _dp, _sz, _dp]),
    "gple_pes": (C.c_int, [_vp, C.c_int, _dp, _sz, _dp, _dp, _dp]),
    "gple_evolve": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _dp, _sz, _dp, _sz, _dp, _sz, C.c_double, C.c_double]),
    "gple_comm_unique_id": (C.c_int, [C.c_char_p]),
    "gple_ctx_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p]),
    "gple_ctx_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gple_partition": (C.c_int, [_sz, C.c_int, C.c_int, C.POINTER(_sz), C.POINTER(_sz)]),
    "gple_allgather_points": (C.c_int, [_vp, _dp, _sz]),
    "gple_model_bcast": (C.c_int, [_vp, C.POINTER(C.c_void_p), C.c_int]),
    "gple_allreduce_sum": (C.c_int, [_vp, _dp, _sz]),
    "gple_evolve_sharded": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _dp, _sz, _dp, _sz, _dp, _sz, C.c_double, C.c_double]),
    "gple_new_point_predict": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _dp, _sz, C.c_int, C.c_int, C.c_double, C.c_double, _dp]),
    "gple_observables": (C.c_int, [_vp, C.c_int, _dp, _sz, C.c_double, C.c_int, _dp]),
    "gple_tune_variance_gemm": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "gple_set_potrf_flat": (C.c_int, [C.c_int]),
    "gple_set_variance_gemm_variant": (C.c_int, [C.c_int]),
    "gple_profile_enable": (C.c_int, [_vp, C.c_int]),
    "gple_profile_read": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong), C.POINTER(C.c_double)]),
    "gple_measure_dmma_tile_peak": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "gple_measure_fp64_peak": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class GpleError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"gple status {status}: {message}")
        self.status = status


def addr(a):
    """Raw address of a numpy array (host) or of an object with data_ptr() (device tensor); None -> NULL."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "pass C-contiguous numpy arrays"
    return a.ctypes.data


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


class Context:
    """One CUDA device + stream (gple_ctx).  Fails loudly without a GPU."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _vp()
        rc = self.lib.gple_ctx_create(int(device), C.byref(h))
        if rc != OK:
            raise GpleError(rc, "gple_ctx_create failed: no usable CUDA device (this library has no CPU fallback)")
        self.h = h

    def check(self, rc, allow=()):
        if rc != OK and rc not in allow:
            raise GpleError(rc, self.lib.gple_last_error(self.h).decode())
        return rc

    def set_stream(self, stream_ptr):
        self.check(self.lib.gple_ctx_set_stream(self.h, stream_ptr))

    def set_gated_variance(self, on: bool):
        self.check(self.lib.gple_ctx_set_option(self.h, 1, int(on)))

    def set_refine_solution(self, on: bool):
        self.check(self.lib.gple_ctx_set_option(self.h, 4, int(on)))

    def set_factorise_graphs(self, on: bool):
        self.check(self.lib.gple_ctx_set_option(self.h, 5, int(on)))

    def set_gate_stage2_tiles(self, tiles: int):
        self.check(self.lib.gple_ctx_set_option(self.h, 6, int(tiles)))

    def set_gate_schedule(self, complex_element: bool, stages):
        """stages: sequence of cumulative (re_end, im_end) block boundaries (128 training points each); empty = automatic."""
        st = [(int(s), 0) if np.isscalar(s) else (int(s[0]), int(s[1])) for s in stages]
        re = (C.c_int * max(1, len(st)))(*[s[0] for s in st])
        im = (C.c_int * max(1, len(st)))(*[s[1] for s in st])
        self.check(self.lib.gple_ctx_set_gate_schedule(self.h, int(bool(complex_element)), len(st), C.cast(re, C.c_void_p), C.cast(im, C.c_void_p)))

    def set_gate_stage_tiles(self, tiles: int, tiles_im: int = -1):
        self.check(self.lib.gple_ctx_set_option(self.h, 2, int(tiles)))
        self.check(self.lib.gple_ctx_set_option(self.h, 3, int(tiles_im)))

    def gate_statistics(self):
        out = (C.c_ulonglong * 4)()
        self.check(self.lib.gple_gate_statistics(self.h, out))
        return int(out[0]), int(out[1]), int(out[2]), int(out[3])

    def sync(self):
        self.check(self.lib.gple_ctx_sync(self.h))

    # ---- multi-GPU (one process per GPU): NCCL communicator of this context ------------------------------
    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        """unique_id: the 128 bytes rank 0 got from comm_unique_id(), handed over by the host's own channel."""
        assert len(unique_id) == COMM_ID_BYTES
        self.check(self.lib.gple_ctx_comm_init(self.h, int(rank), int(nranks), unique_id))

    def comm_info(self):
        r, n = C.c_int(), C.c_int()
        self.check(self.lib.gple_ctx_comm_info(self.h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def allgather_points(self, pts, total: int):
        """in place: pts = full (total, 4) array (numpy or device tensor) with this rank's block valid"""
        self.check(self.lib.gple_allgather_points(self.h, addr(pts), int(total)))

    def allreduce_sum(self, values):
        self.check(self.lib.gple_allreduce_sum(self.h, addr(values), int(values.size if isinstance(values, np.ndarray) else values.numel())))

    @property
    def launches(self) -> int:
        return int(self.lib.gple_launch_count(self.h))

    def profile_enable(self, on: bool):
        self.check(self.lib.gple_profile_enable(self.h, int(on)))

    def profile_read(self, slot: int):
        """(total_ms, launches, work) of the profiled kernel since the last read."""
        ms, n, w = C.c_double(), C.c_ulonglong(), C.c_double()
        self.check(self.lib.gple_profile_read(self.h, int(slot), C.byref(ms), C.byref(n), C.byref(w)))
        return ms.value, int(n.value), w.value

    def tune_variance_gemm(self, variant: int, rows: int, n: int, iters: int = 5) -> float:
        ms = C.c_double()
        self.check(self.lib.gple_tune_variance_gemm(self.h, variant, rows, n, iters, C.byref(ms)))
        return ms.value

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self.check(self.lib.gple_measure_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def dmma_tile_peak(self) -> float:
        a = C.c_double()
        self.check(self.lib.gple_measure_dmma_tile_peak(self.h, C.byref(a)))
        return a.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.gple_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = load().gple_comm_unique_id(buf)
    if rc != OK:
        raise GpleError(rc, "gple_comm_unique_id failed (NCCL not loadable?)")
    return buf.raw


def gate_schedule_automatic(complex_element: bool, blocks: int):
    """The automatic schedule of the staged bound for a model of `blocks` 128-blocks: [(re_end, im_end), ...] (host only)."""
    re, im = (C.c_int * 16)(), (C.c_int * 16)()
    n = load().gple_gate_schedule_automatic(int(bool(complex_element)), int(blocks), C.cast(re, C.c_void_p), C.cast(im, C.c_void_p), 16)
    if n < 0:
        raise GpleError(-n, "gple_gate_schedule_automatic: bad argument")
    return [(int(re[k]), int(im[k])) for k in range(min(n, 16))]


def partition(total: int, rank: int, nranks: int):
    lo, hi = _sz(), _sz()
    rc = load().gple_partition(int(total), int(rank), int(nranks), C.byref(lo), C.byref(hi))
    if rc != OK:
        raise GpleError(rc, "gple_partition: bad rank / size")
    return int(lo.value), int(hi.value)


_default = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context(int(os.environ.get("LOCAL_RANK", "0")) if "GPLE_DEVICE" not in os.environ else int(os.environ["GPLE_DEVICE"]))
    return _default
