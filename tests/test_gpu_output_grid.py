"""output_phase at the reference's own size (gple/output.cpp:181-233; grid of gple/input.cpp:39-71: 200 x 200 = 40 000 phase-space
points): the batched PredictiveKernel / PredictiveComplexKernel calls of output.cpp:204,218 -- cutoff prediction AND variance of
all three elements on the whole grid -- through the C-ABI, against the CPU oracle on every grid point (only a
20-point grid had been checked).  The grid spans +-6 sigma around the training cloud, so it covers the point cloud (gate 1), its
rim (the cubic band of cutoff_factor, kernel.h:301-332) and empty phase space (gate 0, variance = prior)."""
import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])


def test_output_grid_of_40000_points_three_elements(oracle):
    from gaussian_process_liouville_equation_b200 import complex_kernel, kernel

    n, centre, side = 300, (-0.5, syn.P0), 200
    x = centre[0] + syn.SIGMA_X * np.linspace(-6.0, 6.0, side)
    p = centre[1] + syn.SIGMA_P * np.linspace(-6.0, 6.0, side)
    grid = np.ascontiguousarray(np.stack(np.meshgrid(x, p, indexing="ij"), axis=-1).reshape(-1, 2))
    assert len(grid) == 40000
    sets = [syn.training_set(71, e, n, centre) for e in range(3)]
    seen_band = 0
    for e in range(3):
        if e == 1:
            g = complex_kernel.PredictiveComplexKernel(grid, complex_kernel.TrainingComplexKernel(THETA_C, sets[e]))
            o = oracle.TrainingComplexKernel(THETA_C, *sets[e]).predict(grid)
            prior = THETA_C[0] ** 2 * (THETA_C[1] ** 2 + THETA_C[4] ** 2)
        else:
            g = kernel.PredictiveKernel(grid, kernel.TrainingKernel(syn.theta_real(), sets[e]))
            o = oracle.TrainingKernel(syn.theta_real(), *sets[e]).predict(grid)
            prior = syn.theta_real()[0] ** 2
        scale = np.abs(o["pred"]).max()
        assert np.abs(g.get_prediction() - o["pred"]).max() <= 1e-9 * scale
        assert np.abs(g.get_variance() - o["var"]).max() <= 1e-9 * prior
        # the cutoff prediction: 1e-9 where the gate is decided; inside the cubic band it inherits the rounding of the variance on
        # BOTH sides (a difference of nearly equal numbers in the reference formulation: profiles/r02_parity_distances.md)
        d = np.abs(g.get_cutoff_prediction() - o["cutoff"])
        gate = np.abs(o["cutoff"]) / np.maximum(np.abs(o["pred"]), 1e-300)
        decided = (gate < 1e-12) | (gate > 1.0 - 1e-12)
        seen_band += int((~decided).sum())
        assert d[decided].max() <= 1e-9 * scale and d.max() <= 1e-6 * scale
        # far from the cloud the element is empty: gate exactly 0, variance back at (almost) the prior
        far = np.abs(grid[:, 0] - centre[0]) > 5.5 * syn.SIGMA_X
        assert np.all(g.get_cutoff_prediction()[far] == 0.0) and g.get_variance()[far].min() >= 0.5 * prior
    assert seen_band > 100  # the grid does sample the cubic band
