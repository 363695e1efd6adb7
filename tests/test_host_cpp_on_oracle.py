"""Runs the host-side C++ (host/gple_host.hpp, gple_opt.hpp, gple_mc.hpp, examples/mqcle_run.cpp) on the CPU: the test binaries are
linked against tests/cpp/mock_abi.cpp, a stand-in for libgple_b200.so that answers the C-ABI with the oracle.  This executes the
host LOGIC (aggregators, callbacks, optimiser, sampler tuning, main loop) in the CPU tier; the numerics of the product are only
ever tested on the GPU (tests/test_gpu_*.py, test_opt_cpp.py, test_example_run.py run the same binaries on the real library)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_backend
from gaussian_process_liouville_equation_b200 import mc, opt, predict
from gaussian_process_liouville_equation_b200 import synthetic as syn
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_DIR = os.path.join(ROOT, "tests", "cpp", "_build", "mock")


_mock_built = False


def build_mock():
    global _mock_built
    if _mock_built:
        return
    _mock_built = True
    os.makedirs(MOCK_DIR, exist_ok=True)
    orc.build()
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-pthread", "-shared", os.path.join(ROOT, "tests", "cpp", "mock_abi.cpp"), "-o", os.path.join(MOCK_DIR, "libgple_b200.so")])


def compile_on_mock(source, name):
    build_mock()
    exe = os.path.join(MOCK_DIR, name)
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-pthread", source, "-o", exe, f"-L{MOCK_DIR}", "-lgple_b200", f"-Wl,-rpath,{MOCK_DIR}"])
    return exe


def parse(out):
    return {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}


def test_host_mirror_logic(tmp_path):
    exe = compile_on_mock(os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "host_mirror_test")
    got = parse(subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, check=True).stdout)
    n, sx, sp, p0 = 200, 1.0 / (2.0 * 0.7056), 0.7056, 14.112
    i = np.arange(1, n + 1)
    u, v = np.fmod(0.5 + 0.6180339887498949 * i, 1.0), np.fmod(0.5 + 0.7548776662466927 * i, 1.0)
    X = np.stack([-0.8 + sx * 3.0 * (2 * u - 1), p0 + sp * 3.0 * (2 * v - 1)], 1)
    g = np.exp(-0.5 * (((X[:, 0] + 0.8) / sx) ** 2 + ((X[:, 1] - p0) / sp) ** 2)) / (2 * np.pi * sx * sp)
    ys = [0.6 * g + 0j, np.sqrt(0.24) * g * np.exp(1j * (0.7 * (X[:, 0] + 0.8) - 0.2 * (X[:, 1] - p0))), 0.4 * g + 0j]
    tr = np.array([1.0, sx, sp, 1e-2])
    tc = np.array([1.0, 1.2, 0.8 * sx, 1.1 * sp, 0.7, 1.1 * sx, 0.9 * sp, 2e-2])
    k0, k1, k2 = orc.TrainingKernel(tr, X, ys[0]), orc.TrainingComplexKernel(tc, X, ys[1]), orc.TrainingKernel(tr, X, ys[2])
    # the aggregators of TrainingKernels (predict.cpp:395-463) on top of the same oracle numbers; the C++ side computes exp / pow
    # of the synthetic labels itself, so agreement is to rounding of those inputs, not bitwise
    Kb, dKb = orc.kernel_real(X[:12], X[:12], tr, True, True)
    _, Ktb, dKcb, dKtb = orc.kernel_complex(X[:12], X[:12], tc, True, True)
    assert got["kb_K_3_5"] == pytest.approx(Kb[3, 5], rel=1e-12) and got["kb_dK1_3_5"] == pytest.approx(dKb[1][3, 5], rel=1e-11) and got["kb_dK3_4_4"] == pytest.approx(dKb[3][4, 4], rel=1e-12)
    assert got["ckb_Kt_3_5_im"] == pytest.approx(Ktb[3, 5].imag, rel=1e-12) and got["ckb_dKt2_3_5_im"] == pytest.approx(dKtb[2][3, 5].imag, rel=1e-11) and got["ckb_dK7_4_4"] == pytest.approx(dKcb[7][4, 4], rel=1e-12)
    assert got["population"] == pytest.approx(k0.population + k2.population, rel=1e-10)
    assert got["purity"] == pytest.approx(k0.purity + k2.purity + 2 * k1.purity, rel=1e-10)
    assert got["error00"] == pytest.approx(k0.error, rel=1e-8) and got["error10"] == pytest.approx(k1.error, rel=1e-8)
    phase = [np.array(line.split(), dtype=float) for line in (tmp_path / "phase.txt").read_text().split("\n")[:6]]
    assert all(len(p) == 20 for p in phase) and np.all(phase[1] == 0.0) and np.all(phase[5] == 0.0) and np.any(phase[3] != 0.0)


def test_cpp_optimiser_logic_against_the_python_driver(tmp_path):
    """Optimization::optimize (gple_opt.hpp on nlopt_lite.hpp) and the scipy-based opt.py, both on the oracle's callbacks."""
    from test_opt_cpp import write_points

    exe = compile_on_mock(os.path.join(ROOT, "tests", "cpp", "opt_test.cpp"), "opt_test")
    n = 48
    X, y = syn.training_set(61, 0, n, (syn.X0, syn.P0))
    y = y / 0.6
    Xe, ye = syn.extra_points(61, 0, X, 5 * n, (syn.X0, syn.P0))
    density, extra = [syn.points_aos(X, y), None, None], [syn.points_aos(Xe, ye / 0.6), None, None]
    o = oracle_backend.observable_sums(0, density[0], syn.MASS, 0)
    e0 = o[7] / o[0]
    path = os.path.join(tmp_path, "points.txt")
    write_points(path, density, extra)
    got = parse(subprocess.run([exe, path, "0", repr(syn.MASS), repr(e0), "1.0"], capture_output=True, text=True, check=True).stdout)
    ref = opt.Optimization((syn.SIGMA_X, syn.SIGMA_P), syn.MASS, 0, InitialTotalEnergy=e0, InitialPurity=1.0, backend=oracle_backend, max_global_evals=200)
    (err, steps, typ), check = ref.optimize(density, extra)
    k = predict.TrainingKernels(ref.get_parameters(), predict.construct_training_sets(density), True, True, False, oracle_backend)
    assert np.isfinite(got["error"]) and got["error"] <= 1.5 * err + 1e-12
    assert abs(got["population"] - 1.0) <= max(2 * opt.AverageTolerance, 1.5 * abs(k.calculate_population() - 1.0))
    assert abs(got["purity"] - 1.0) <= max(2 * opt.AverageTolerance, 1.5 * abs(k.calculate_purity() - 1.0))
    assert got["lb0_1"] <= got["theta0_1"] <= got["ub0_1"] and got["theta0_3"] == opt.InitialNoise
    assert got["steps0"] > 10 and got["evaluations"] > got["steps0"]


def test_cpp_sampler_logic_is_identical_to_the_python_mirror(tmp_path):
    """monte_carlo_selection of gple_mc.hpp and of mc.py make the same calls: on the same (oracle) chains the results agree to rounding."""
    exe = compile_on_mock(os.path.join(ROOT, "tests", "cpp", "mc_test.cpp"), "mc_test")
    g = syn.rng(89, 0)
    pts = np.zeros((24, 4))
    pts[:, 0] = syn.X0 + syn.SIGMA_X * g.standard_normal(24)
    pts[:, 1] = syn.P0 + syn.SIGMA_P * g.standard_normal(24)
    path = tmp_path / "start.txt"
    path.write_text(f"{len(pts)}\n" + "".join(f"{float(x)!r} {float(p)!r}\n" for x, p in pts[:, :2]))
    got = parse(subprocess.run([exe, str(path), "21"], capture_output=True, text=True, check=True).stdout)
    params = [mc.MCParameters() for _ in range(3)]
    analytic = ((syn.X0, syn.P0), (syn.SIGMA_X, syn.SIGMA_P), (0.8, 0.6), (0.0, 0.4))
    ref = mc.monte_carlo_selection([pts, pts.copy(), None], params, oracle_backend.Sampler(21, analytic=analytic))
    for e in range(2):
        assert got[f"displacement{e}"] == params[e].get_max_displacement() and got[f"steps{e}"] == params[e].get_num_MC_steps()
        walked = np.array([[got[f"p{e}_{i}_{c}"] for c in ("x", "p", "re", "im")] for i in range(len(pts))])
        assert np.abs(walked - ref[e]).max() <= 1e-13 * np.abs(ref[e]).max()  # the mock and the oracle library are separate builds of the same code


def test_main_loop_logic():
    """examples/mqcle_run.cpp on the oracle: far from the crossing nothing but rho00 is populated and the averages are conserved."""
    exe = compile_on_mock(os.path.join(ROOT, "examples", "mqcle_run.cpp"), "mqcle_run")
    out = subprocess.run([exe, "32", "3", "0", "1", "7"], capture_output=True, text=True, check=True, timeout=900).stdout
    ticks = [[float(v) for v in line.split()[2:]] for line in out.splitlines() if line.startswith("tick ")]
    info = {line.split()[0]: line.split()[1:] for line in out.splitlines() if not line.startswith("tick ")}
    assert len(ticks) == 4 and info["elements"] == ["32", "0", "0"]
    pop0, e0, pur0 = ticks[0]
    assert abs(pop0 - 1.0) < 0.15
    for pop, e, pur in ticks:
        assert abs(pop - pop0) < 0.05 and abs(e / e0 - 1.0) < 0.03 and abs(pur - pur0) < 0.15


def test_cpp_optimiser_three_elements_logic(tmp_path):
    """Diagonal and full constrained stages (diagonal_loose / full_loose, diagonal_constraints / full_constraints, the ModelCache
    shared between them) with all three elements populated."""
    from test_opt_cpp import write_points

    exe = compile_on_mock(os.path.join(ROOT, "tests", "cpp", "opt_test.cpp"), "opt_test")
    n, centre = 16, (0.0, syn.P0)
    density, extra = [], []
    for e in range(3):
        X, y = syn.training_set(63, e, n, centre)
        Xe, ye = syn.extra_points(63, e, X, 5 * n, centre)
        density.append(syn.points_aos(X, y))
        extra.append(syn.points_aos(Xe, ye))
    o = [oracle_backend.observable_sums(1, density[e], syn.MASS, i) for i, e in enumerate((0, 2))]
    e0 = 0.6 * o[0][7] / o[0][0] + 0.4 * o[1][7] / o[1][0]
    path = os.path.join(tmp_path, "points.txt")
    write_points(path, density, extra)
    got = parse(subprocess.run([exe, path, "1", repr(syn.MASS), repr(e0), repr(syn.snapshot_purity()), "40", "80"], capture_output=True, text=True, check=True, timeout=900).stdout)
    assert np.isfinite(got["error"]) and got["error"] >= 0.0
    assert abs(got["population"] - 1.0) < 0.2 and abs(got["energy"] / e0 - 1.0) < 0.2
    for e, npar in enumerate((4, 8, 4)):
        assert all(np.isfinite(got[f"theta{e}_{p}"]) for p in range(npar))
    assert got["steps0"] > 0 and got["steps1"] > 0 and got["steps2"] > 0 and got["steps3"] > 0 and got["steps4"] > 0


def test_speculative_restart_stages_change_nothing(tmp_path):
    """Optimization::set_speculative_restarts: stages 2 and 3 of optimize() run ahead of time on their own threads, contexts and
    copies of the minimisers; the outcome -- every parameter, the error, the result type, the evaluation counts -- is the one of the
    sequential run, whichever stage is finally kept."""
    from test_opt_cpp import write_points

    exe = compile_on_mock(os.path.join(ROOT, "tests", "cpp", "opt_test.cpp"), "opt_test")
    n, centre = 16, (0.0, syn.P0)
    density, extra = [], []
    for e in range(3):
        X, y = syn.training_set(63, e, n, centre)
        Xe, ye = syn.extra_points(63, e, X, 5 * n, centre)
        density.append(syn.points_aos(X, y))
        extra.append(syn.points_aos(Xe, ye))
    o = [oracle_backend.observable_sums(1, density[e], syn.MASS, i) for i, e in enumerate((0, 2))]
    e0 = 0.6 * o[0][7] / o[0][0] + 0.4 * o[1][7] / o[1][0]
    path = os.path.join(tmp_path, "points.txt")
    write_points(path, density, extra)
    # an unreachable purity target makes every stage miss the averages, so all three are needed and compared (the early-exit path,
    # where the stages running ahead are stopped and discarded, runs in test_cpp_optimiser_logic_against_the_python_driver)
    runs = []
    for spec in ("1", "0"):
        env = dict(os.environ, GPLE_SPECULATIVE_RESTARTS=spec)
        out = subprocess.run([exe, path, "1", repr(syn.MASS), repr(e0), "3.0", "40", "80"], capture_output=True, text=True, check=True, timeout=900, env=env).stdout
        runs.append({k: v for k, v in (line.split() for line in out.strip().splitlines()) if k != "wall_s"})
    assert runs[0] == runs[1]  # text-identical: bitwise equal doubles at 17 digits
    assert int(runs[0]["evaluations"]) > 300


def test_main_loop_logic_at_the_crossing():
    """Started at the crossing the run must populate rho10 and rho11: is_very_small, new_element_point_selection (Metropolis
    tuning on the new_point_predict target, extra points) and the element-change re-optimisation of main.cpp:145-162."""
    exe = compile_on_mock(os.path.join(ROOT, "examples", "mqcle_run.cpp"), "mqcle_run")
    out = subprocess.run([exe, "12", "2", "0", "1", "7", "-0.3"], capture_output=True, text=True, check=True, timeout=900).stdout
    ticks = [[float(v) for v in line.split()[2:]] for line in out.splitlines() if line.startswith("tick ")]
    info = {line.split()[0]: line.split()[1:] for line in out.splitlines() if not line.startswith("tick ")}
    assert len(ticks) == 3 and info["elements"] == ["12", "12", "12"] and int(info["optimisations"][0]) >= 2
    for pop, e, pur in ticks:
        assert 0.8 < pop < 1.2 and 0.8 < pur < 1.2
