"""The C++ host mirror (host/gple_host.hpp) compiles against the C-ABI (CPU check) and, on the GPU box,
reproduces the oracle for a TrainingKernels -> PredictiveKernel -> evolve sequence written like gple/main.cpp."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "host_mirror_test")
LIBDIR = os.path.join(ROOT, "gaussian_process_liouville_equation_b200")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", SRC, "-o", EXE, f"-L{LIBDIR}", "-lgple_b200", f"-Wl,-rpath,{LIBDIR}"])


def test_host_mirror_compiles_and_links():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_mirror_matches_oracle(oracle, tmp_path):
    build()
    out = subprocess.run([EXE, str(tmp_path)], capture_output=True, text=True, check=True).stdout
    got = {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}
    n, sx, sp, p0 = 200, 1.0 / (2.0 * 0.7056), 0.7056, 14.112
    i = np.arange(1, n + 1)
    u, v = np.fmod(0.5 + 0.6180339887498949 * i, 1.0), np.fmod(0.5 + 0.7548776662466927 * i, 1.0)
    X = np.stack([-0.8 + sx * 3.0 * (2 * u - 1), p0 + sp * 3.0 * (2 * v - 1)], 1)
    g = np.exp(-0.5 * (((X[:, 0] + 0.8) / sx) ** 2 + ((X[:, 1] - p0) / sp) ** 2)) / (2 * np.pi * sx * sp)
    ys = [0.6 * g + 0j, np.sqrt(0.24) * g * np.exp(1j * (0.7 * (X[:, 0] + 0.8) - 0.2 * (X[:, 1] - p0))), 0.4 * g + 0j]
    tr = np.array([1.0, sx, sp, 1e-2])
    tc = np.array([1.0, 1.2, 0.8 * sx, 1.1 * sp, 0.7, 1.1 * sx, 0.9 * sp, 2e-2])
    k0, k1, k2 = oracle.TrainingKernel(tr, X, ys[0]), oracle.TrainingComplexKernel(tc, X, ys[1]), oracle.TrainingKernel(tr, X, ys[2])
    Kb, dKb = oracle.kernel_real(X[:12], X[:12], tr, True, True)
    _, Ktb, dKcb, dKtb = oracle.kernel_complex(X[:12], X[:12], tc, True, True)
    assert got["kb_K_3_5"] == pytest.approx(Kb[3, 5], rel=1e-12) and got["kb_K_4_4"] == pytest.approx(Kb[4, 4], rel=1e-12)
    assert got["kb_dK1_3_5"] == pytest.approx(dKb[1][3, 5], rel=1e-11) and got["kb_dK3_4_4"] == pytest.approx(dKb[3][4, 4], rel=1e-12)
    assert got["ckb_Kt_3_5_im"] == pytest.approx(Ktb[3, 5].imag, rel=1e-12) and got["ckb_dKt2_3_5_im"] == pytest.approx(dKtb[2][3, 5].imag, rel=1e-11)
    assert got["ckb_dK7_4_4"] == pytest.approx(dKcb[7][4, 4], rel=1e-12)
    assert got["population"] == pytest.approx(k0.population + k2.population, rel=1e-9)
    assert got["purity"] == pytest.approx(k0.purity + k2.purity + 2 * k1.purity, rel=1e-8)
    assert got["error00"] == pytest.approx(k0.error, rel=1e-8) and got["error10"] == pytest.approx(k1.error, rel=1e-7)
    p = k0.predict(np.array([[-0.7, p0 + 0.1], [3.0, p0]]))
    assert got["cutoff0"] == pytest.approx(p["cutoff"][0], rel=1e-8) and got["cutoff1"] == pytest.approx(p["cutoff"][1], rel=1e-8, abs=1e-12)
    assert got["var0"] == pytest.approx(p["var"][0], abs=1e-9)
    pts = [np.column_stack([X, y.real, y.imag]) for y in ys]
    ev = oracle.evolve(1, pts[0], pts[1], pts[2], 2000.0, 2.0, k0, k1, k2)
    assert got["x00_0"] == pytest.approx(ev[0][0, 0], rel=1e-13)
    sc = np.abs(ev[0][:, 2]).max()
    assert got["rho00_0"] == pytest.approx(ev[0][0, 2], abs=1e-6 * sc)
    assert got["rho10_5_re"] == pytest.approx(ev[1][5, 2], abs=1e-6 * sc) and got["rho10_5_im"] == pytest.approx(ev[1][5, 3], abs=1e-6 * sc)
    E, _, _ = oracle.pes(1, np.array([0.3]))
    assert got["E1_at_0.3"] == pytest.approx(E[0, 1], rel=1e-14)
    # output_phase (gple/output.cpp:181-233): 3 elements x (Re, Im) cutoff-prediction lines + 3 variance lines, blank-line terminated
    grid = np.array([[-0.8 + sx * (ix - 2.0), p0 + sp * (ip - 1.5) * 1.5] for ix in range(5) for ip in range(4)])
    phase = [np.array(line.split(), dtype=float) for line in (tmp_path / "phase.txt").read_text().split("\n")[:6]]
    var = [np.array(line.split(), dtype=float) for line in (tmp_path / "var.txt").read_text().split("\n")[:3]]
    assert (tmp_path / "phase.txt").read_text().endswith("\n\n") and all(len(v) == 20 for v in phase + var)
    for e, k in enumerate((k0, k1, k2)):
        p = k.predict(grid)
        scale = np.abs(p["cutoff"]).max()
        assert np.abs(phase[2 * e] - p["cutoff"].real).max() <= 1e-6 * scale and np.abs(phase[2 * e + 1] - p["cutoff"].imag).max() <= 1e-6 * scale
        assert np.abs(var[e] - p["var"]).max() <= 1e-9 * (tr[0] ** 2 if e != 1 else tc[0] ** 2 * (tc[1] ** 2 + tc[4] ** 2))
