"""The C++ optimiser mirror (host/gple_opt.hpp + host/nlopt_lite.hpp).

CPU: nlopt_lite reaches the known optima of analytic problems; gple_opt.hpp compiles and links against the C-ABI.
GPU: Optimization::optimize on a C1-like case reaches an outcome as good as the scipy-based Python driver run on the
oracle's callbacks (parity on outcomes), and on a three-element case stays within its bounds with all
elements optimised concurrently."""
import os
import subprocess

import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
LIBDIR = os.path.join(ROOT, "gaussian_process_liouville_equation_b200")


def compile_cpp(name, link):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    cmd = ["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", "-pthread", os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe]
    if link:
        cmd += [f"-L{LIBDIR}", "-lgple_b200", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.check_call(cmd)
    return exe


def test_nlopt_lite_on_analytic_problems():
    exe = compile_cpp("nlopt_lite_test", link=False)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "FAIL" not in r.stdout, r.stdout


def test_optimiser_mirror_compiles_and_links():
    assert os.path.exists(compile_cpp("opt_test", link=True))


def test_metropolis_mirror_compiles_and_links():
    assert os.path.exists(compile_cpp("mc_test", link=True))


def write_points(path, density, extra):
    with open(path, "w") as f:
        for group in (density, extra):
            for pts in group:
                if pts is None:
                    f.write("0\n")
                    continue
                f.write(f"{len(pts)}\n")
                for row in pts:
                    f.write(" ".join(repr(float(v)) for v in row) + "\n")


def run_cpp(density, extra, pes_model, e0, purity, tmp_path, *caps):
    exe = compile_cpp("opt_test", link=True)
    path = os.path.join(tmp_path, "points.txt")
    write_points(path, density, extra)
    out = subprocess.run([exe, path, str(pes_model), repr(syn.MASS), repr(e0), repr(purity), *map(str, caps)], capture_output=True, text=True, check=True).stdout
    return {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}


@pytest.mark.gpu
def test_cpp_optimiser_matches_python_driver_outcome(tmp_path):
    import oracle_backend
    from gaussian_process_liouville_equation_b200 import opt, predict

    n = 48
    X, y = syn.training_set(61, 0, n, (syn.X0, syn.P0))
    y = y / 0.6  # population 1, purity 1
    Xe, ye = syn.extra_points(61, 0, X, 5 * n, (syn.X0, syn.P0))
    density, extra = [syn.points_aos(X, y), None, None], [syn.points_aos(Xe, ye / 0.6), None, None]
    o = oracle_backend.observable_sums(0, density[0], syn.MASS, 0)
    e0 = o[7] / o[0]
    got = run_cpp(density, extra, 0, e0, 1.0, str(tmp_path))
    ref = opt.Optimization((syn.SIGMA_X, syn.SIGMA_P), syn.MASS, 0, InitialTotalEnergy=e0, InitialPurity=1.0, backend=oracle_backend, max_global_evals=200)
    (err, steps, typ), check = ref.optimize(density, extra)
    k = predict.TrainingKernels(ref.get_parameters(), predict.construct_training_sets(density), True, True, False, oracle_backend)
    # outcomes: the C++ driver ends at a loss no worse than 1.5x the scipy driver's, with the averages as close to their targets
    assert np.isfinite(got["error"]) and got["error"] <= 1.5 * err + 1e-12
    assert abs(got["population"] - 1.0) <= max(2 * opt.AverageTolerance, 1.5 * abs(k.calculate_population() - 1.0))
    assert abs(got["purity"] - 1.0) <= max(2 * opt.AverageTolerance, 1.5 * abs(k.calculate_purity() - 1.0))
    assert got["lb0_1"] <= got["theta0_1"] <= got["ub0_1"] and got["theta0_3"] == opt.InitialNoise
    assert got["steps0"] > 10 and got["evaluations"] > got["steps0"]


@pytest.mark.gpu
def test_cpp_optimiser_three_elements(tmp_path):
    n, centre = 64, (0.0, syn.P0)
    density, extra = [], []
    for e in range(3):
        X, y = syn.training_set(63, e, n, centre)
        Xe, ye = syn.extra_points(63, e, X, 5 * n, centre)
        density.append(syn.points_aos(X, y))
        extra.append(syn.points_aos(Xe, ye))
    import oracle_backend

    o = [oracle_backend.observable_sums(1, density[e], syn.MASS, i) for i, e in enumerate((0, 2))]
    e0 = 0.6 * o[0][7] / o[0][0] + 0.4 * o[1][7] / o[1][0]
    got = run_cpp(density, extra, 1, e0, 1.0, str(tmp_path), 100, 300)
    assert np.isfinite(got["error"]) and got["error"] >= 0.0
    assert abs(got["population"] - 1.0) < 0.2 and abs(got["energy"] / e0 - 1.0) < 0.2
    for e, npar in enumerate((4, 8, 4)):
        assert all(np.isfinite(got[f"theta{e}_{p}"]) for p in range(npar))
    assert got["steps0"] > 0 and got["steps1"] > 0 and got["steps2"] > 0 and got["steps4"] > 0


@pytest.mark.gpu
def test_speculative_restart_stages_change_nothing_on_the_gpu(tmp_path, monkeypatch):
    """Optimization::set_speculative_restarts on the real library: three restart stages on nine contexts at once against the same
    stages one after the other -- identical parameters, error, result type and evaluation counts (the library's results do not
    depend on what else runs on the device)."""
    n, centre = 96, (0.0, syn.P0)
    density, extra = [], []
    for e in range(3):
        X, y = syn.training_set(63, e, n, centre)
        Xe, ye = syn.extra_points(63, e, X, 5 * n, centre)
        density.append(syn.points_aos(X, y))
        extra.append(syn.points_aos(Xe, ye))
    import oracle_backend

    o = [oracle_backend.observable_sums(1, density[e], syn.MASS, i) for i, e in enumerate((0, 2))]
    e0 = 0.6 * o[0][7] / o[0][0] + 0.4 * o[1][7] / o[1][0]
    runs = []
    for mode in ("1", "0"):
        monkeypatch.setenv("GPLE_SPECULATIVE_RESTARTS", mode)
        got = run_cpp(density, extra, 1, e0, 3.0, str(tmp_path), 100, 300)  # purity 3 is unreachable: every stage is needed
        runs.append({k: v for k, v in got.items() if k != "wall_s"})
    assert runs[0] == runs[1] and runs[0]["evaluations"] > 500


@pytest.mark.gpu
def test_cpp_optimiser_outcome_at_n300_is_verified_by_the_oracle_and_locally_optimal(tmp_path):
    """C3-shaped run (DAC, three populated elements, N = 300, M = 5 N) through the C++ host on the GPU, checked three ways
    (the 20 % asserts of the small test say nothing):
      1. the CPU oracle, evaluated at the returned parameters, reproduces the reported averages (1e-7 relative) and loss (1e-4):
         the optimiser did not converge on a numerical artefact of the CUDA path;
      2. the constraints hold far inside AverageTolerance (opt.h:13): population, energy and purity within 1e-3 of their targets;
      3. a different constrained optimiser (scipy SLSQP on the Python mirror of the callbacks) started from the returned point
         inside a +-25 % box finds no feasible point whose loss is more than 1 % lower: the point is a constrained local minimum."""
    import oracle_backend
    from scipy.optimize import minimize

    from gaussian_process_liouville_equation_b200 import opt, predict

    n, centre, dac = 300, (0.0, syn.P0), 1
    density, extra = [], []
    for e in range(3):
        X, y = syn.training_set(35, e, n, centre)
        Xe, ye = syn.extra_points(35, e, X, 5 * n, centre)
        density.append(syn.points_aos(X, y))
        extra.append(syn.points_aos(Xe, ye))
    o = [oracle_backend.observable_sums(dac, density[e], syn.MASS, i) for i, e in enumerate((0, 2))]
    energies = np.array([o[0][7] / o[0][0], o[1][7] / o[1][0]])
    e0, purity0 = 0.6 * energies[0] + 0.4 * energies[1], syn.snapshot_purity()
    got = run_cpp(density, extra, dac, e0, purity0, str(tmp_path), 300, 2000)
    x = np.array([got[f"theta{e}_{p}"] for e, npar in enumerate((4, 8, 4)) for p in range(npar)])
    assert np.all(np.isfinite(x)) and np.isfinite(got["error"])
    # 2. constraints
    assert abs(got["population"] - 1.0) < 1e-3 and abs(got["energy"] / e0 - 1.0) < 1e-3 and abs(got["purity"] / purity0 - 1.0) < 1e-3
    # 1. the oracle at the returned point (the magnitudes were re-derived after the optimisation, opt.cpp:1179-1195: the averages
    # are reported with them; the loss was minimised at InitialMagnitude)
    ts, ets = predict.construct_training_sets(density), predict.construct_training_sets(extra)
    k = predict.TrainingKernels([x[sl] for sl in predict.ELEMENT_SLICES], ts, False, True, False, oracle_backend)
    assert abs(k.calculate_population() / got["population"] - 1.0) < 1e-7
    assert abs(k.calculate_total_energy_average(energies) / got["energy"] - 1.0) < 1e-7
    assert abs(k.calculate_purity() / got["purity"] - 1.0) < 1e-7
    x_opt = x.copy()
    for sl in predict.ELEMENT_SLICES:
        x_opt[sl.start] = opt.InitialMagnitude
    cb_oracle = opt.Callbacks(ts, ets, energies, e0, purity0, oracle_backend)
    loss_oracle = cb_oracle.full_loose(x_opt)
    # (1e-4: the reported error is the minimum the constrained stage saw, the parameters are its final iterate -- 2.7e-5 apart here)
    assert abs(loss_oracle / got["error"] - 1.0) < 1e-4, (loss_oracle, got["error"])
    # 3. SLSQP polish on the GPU-backed Python callbacks: magnitudes and noises stay fixed, the rest may move by 25 %
    cb = opt.Callbacks(ts, ets, energies, e0, purity0)
    fixed = np.zeros(16, dtype=bool)
    for sl, npar in zip(predict.ELEMENT_SLICES, (4, 8, 4)):
        fixed[sl.start] = fixed[sl.start + npar - 1] = True
    free = ~fixed

    def embed(z):
        v = x_opt.copy()
        v[free] = z
        return v

    def f(z):
        v, g = cb.full_loose(embed(z), True)
        return v, g[free]

    cons = dict(type="eq", fun=lambda z: cb.full_constraints(embed(z), False), jac=lambda z: cb.full_constraints(embed(z), True)[1][:, free])
    z0 = x_opt[free]
    res = minimize(f, z0, jac=True, method="SLSQP", bounds=list(zip(0.8 * z0, 1.25 * z0)), constraints=[cons], options=dict(ftol=1e-12, maxiter=60))
    feasible = np.abs(cb.full_constraints(embed(res.x), False)).max() < 1e-3
    assert not (feasible and res.fun < 0.99 * got["error"]), (res.fun, got["error"], res.x / z0)
