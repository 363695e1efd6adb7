"""TEST INFRASTRUCTURE: the deterministic part of the reference's tick loop (gple/main.cpp:135-188) on three
interchangeable backends -- the CUDA path (through the C-ABI), the oracle restatement, and the compiled reference
(oracle/_ref) -- so that a whole TRAJECTORY can be compared, not only single steps (BASELINE config C1).

One tick = evolve(density); evolve(extra_points) with the GPR-backed predict_distribution (main.cpp:75-101, 140-141);
all_kernels = TrainingKernels(parameters, density) (main.cpp:176, predict.cpp:390-393).  Monte Carlo re-selection and
hyper-parameter re-optimisation are left out on purpose: they draw random numbers / take data-dependent branches whose
outcome is not a function of the inputs alone (fixed points, fixed parameters, as VERDICT r1 asks).
"""
import numpy as np

from gaussian_process_liouville_equation_b200 import synthetic as syn


class GpuBackend:
    name = "gpu"

    def __init__(self):
        from gaussian_process_liouville_equation_b200 import complex_kernel, dynamics, kernel

        self.k, self.ck, self.dyn = kernel, complex_kernel, dynamics

    def train(self, thetas, density):
        out = [None, None, None]
        for e in range(3):
            if density[e] is not None and len(density[e]):
                X, y = np.ascontiguousarray(density[e][:, :2]), density[e][:, 2] + 1j * density[e][:, 3]
                out[e] = (self.ck.TrainingComplexKernel if e == 1 else self.k.TrainingKernel)(thetas[e], (X, y), True, True, False)
        return out

    def evolve(self, model, pts, mass, dt, kernels):
        return self.dyn.evolve(model, pts, mass, dt, kernels)

    def scalars(self, kernels):
        pop = [kernels[e].get_population() if kernels[e] is not None else 0.0 for e in (0, 2)]
        r = sum(kernels[e].get_1st_order_average() for e in (0, 2) if kernels[e] is not None)
        pur = sum(kernels[e].get_purity() * (2.0 if e == 1 else 1.0) for e in range(3) if kernels[e] is not None)
        return dict(population=np.array(pop), first_order=np.asarray(r), purity=pur)

    def sums(self, model, pts, mass, surface):
        return self.dyn.observable_sums(model, pts, mass, surface)


class CpuBackend:
    """oracle (module oracle.oracle) or compiled reference (module oracle.ref)"""

    def __init__(self, module, name):
        self.m, self.name = module, name

    def train(self, thetas, density):
        out = [None, None, None]
        for e in range(3):
            if density[e] is not None and len(density[e]):
                X, y = np.ascontiguousarray(density[e][:, :2]), density[e][:, 2] + 1j * density[e][:, 3]
                out[e] = (self.m.TrainingComplexKernel if e == 1 else self.m.TrainingKernel)(thetas[e], X, y, True, True, False)
        return out

    def evolve(self, model, pts, mass, dt, kernels):
        return self.m.evolve(model, pts[0], pts[1], pts[2], mass, dt, kernels[0], kernels[1], kernels[2])

    def scalars(self, kernels):
        pop = [kernels[e].population if kernels[e] is not None else 0.0 for e in (0, 2)]
        r = sum(kernels[e].first_order for e in (0, 2) if kernels[e] is not None)
        pur = sum(kernels[e].purity * (2.0 if e == 1 else 1.0) for e in range(3) if kernels[e] is not None)
        return dict(population=np.array(pop), first_order=np.asarray(r), purity=pur)

    def sums(self, model, pts, mass, surface):
        from oracle import oracle as orc

        return orc.observable_sums(model, pts, mass, surface)  # plain sums over the points; the reference exposes only ratios


def initial_state(config_id, n, m, centre, populated=(0, 1, 2)):
    """density (n points / element) and extra points (m / element) of the synthetic snapshot of SURVEY 8d"""
    density, extra = [None, None, None], [None, None, None]
    for e in populated:
        X, y = syn.training_set(config_id, e, n, centre)
        Xe, ye = syn.extra_points(config_id, e, X, m, centre)
        density[e], extra[e] = syn.points_aos(X, y), syn.points_aos(Xe, ye)
    return density, extra


def run(backend, model, thetas, density, extra, mass, dt, ticks, record=None):
    """Returns (density, extra, observables after the last tick).  record(tick, density, kernels) is called after every tick."""
    kernels = backend.train(thetas, density)
    for tick in range(1, ticks + 1):
        density = backend.evolve(model, density, mass, dt, kernels)
        extra = backend.evolve(model, extra, mass, dt, kernels)
        density = [d if len(d) else None for d in density]
        extra = [d if len(d) else None for d in extra]
        kernels = backend.train(thetas, density)
        if record is not None:
            record(tick, density, kernels)
    return density, extra, observables(backend, model, density, kernels, mass)


def observables(backend, model, density, kernels, mass):
    """The numbers the reference prints every output step (output.cpp:48-132): per-surface populations, energies, <x>, <p>
    from the MC integrals over the points (predict.cpp:65-244), and population / <r> / purity from the element models."""
    out = dict(backend.scalars(kernels))
    mc = {}
    for e, surface in ((0, 0), (2, 1)):
        if density[e] is not None:
            s = backend.sums(model, density[e], mass, surface)
            mc[surface] = dict(weight=s[0], x=s[1] / s[0], p=s[2] / s[0], energy=s[7] / s[0], purity_sum=s[8])
    if density[1] is not None:
        mc["offdiag_purity_sum"] = backend.sums(model, density[1], mass, 0)[8]
    out["mc"] = mc
    return out


def flatten(obs):
    """observables -> (names, values) for a tolerance check"""
    names, vals = [], []
    for i, v in enumerate(obs["population"]):
        names.append(f"population[{i}]")
        vals.append(v)
    for i, v in enumerate(obs["first_order"]):
        names.append(f"first_order[{i}]")
        vals.append(v)
    names.append("purity")
    vals.append(obs["purity"])
    for k, d in obs["mc"].items():
        if isinstance(d, dict):
            for kk, v in d.items():
                names.append(f"mc[{k}].{kk}")
                vals.append(v)
        else:
            names.append(f"mc.{k}")
            vals.append(d)
    return names, np.array(vals, dtype=float)
