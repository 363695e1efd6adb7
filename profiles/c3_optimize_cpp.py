"""C3 through the C++ host (tests/cpp/opt_test.cpp -> host/gple_opt.hpp -> C-ABI): one full Optimization::optimize() on
Tully's DAC with three populated elements, same synthetic inputs as profiles/c3_optimize.py.
Usage (GPU box):  python profiles/c3_optimize_cpp.py [N ...] > gpurun_out/c3_optimize_cpp.md"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gaussian_process_liouville_equation_b200 import dynamics
from gaussian_process_liouville_equation_b200 import synthetic as syn
from test_opt_cpp import run_cpp

SIZES = [int(a) for a in sys.argv[1:]] or [300, 1024]
DAC = 1
print("# C3: full Optimization::optimize() through the C++ host on 1 x B200 (DAC, three elements, M = 5N extra points per element)\n")
print("| N | restart stages | wall s | evaluations (NM x3, diagonal, full) | evaluations / s | result type | final loss | population | energy / E0 | purity |")
print("|---:|---|---:|---|---:|---|---:|---:|---:|---:|")
for N in SIZES:
    centre = (0.0, syn.P0)
    density, extra = [], []
    for e in range(3):
        X, y = syn.training_set(35, e, N, centre)
        Xe, ye = syn.extra_points(35, e, X, 5 * N, centre)
        density.append(syn.points_aos(X, y))
        extra.append(syn.points_aos(Xe, ye))
    pops = [dynamics.observable_sums(DAC, density[e], syn.MASS, i) for i, e in enumerate((0, 2))]
    e_tot = 0.6 * pops[0][7] / pops[0][0] + 0.4 * pops[1][7] / pops[1][0]
    purity0 = syn.snapshot_purity()
    for mode in ("1", "0"):  # Optimization::set_speculative_restarts: stages 2 and 3 ahead of time / one after the other
        os.environ["GPLE_SPECULATIVE_RESTARTS"] = mode
        with tempfile.TemporaryDirectory() as tmp:
            got = run_cpp(density, extra, DAC, e_tot, purity0, tmp, 300, 2000)
        steps = [int(got[f"steps{i}"]) for i in range(5)]
        print(f"| {N} | {'concurrent' if mode == '1' else 'sequential'} | {got['wall_s']:.2f} | {steps} | {got['evaluations'] / got['wall_s']:.1f} | {int(got['type'])} | {got['error']:.4g} | {got['population']:.4f} | {got['energy'] / e_tot:.4f} | {got['purity']:.4f} |", flush=True)
