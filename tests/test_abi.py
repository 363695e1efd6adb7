"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
declared in include/gple_b200.h, and refuses (loudly, no CPU fallback) to create a context without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gple_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gple_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for must in ("gple_ctx_create", "gple_kernel_real", "gple_train_real", "gple_train_complex", "gple_predict_real",
                 "gple_predict_complex", "gple_evolve", "gple_pes", "gple_observables", "gple_loose_function"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from gaussian_process_liouville_equation_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "libgple_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in include/gple_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == header_symbols(), "python binding table out of sync with the header"
    assert b"sm_100a" in _lib.load().gple_version()


def test_no_cpu_fallback_without_a_gpu():
    import torch

    from gaussian_process_liouville_equation_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GpleError):
        _lib.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gaussian_process_liouville_equation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
