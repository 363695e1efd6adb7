"""TEST INFRASTRUCTURE: adapter that lets the host-side drivers (predict.TrainingKernels, opt.Callbacks,
opt.Optimization) run on top of the CPU oracle instead of the CUDA library, so that their logic can be tested
without a GPU and their GPU results can be compared with an oracle-backed run."""
import numpy as np

from oracle import oracle as orc


class TrainingKernel:
    def __init__(self, Parameter, TrainingSet, err=True, avg=True, deriv=False):
        self.o = orc.TrainingKernel(Parameter, TrainingSet[0], TrainingSet[1], err, avg, deriv)

    def get_error(self):
        return self.o.error

    def get_population(self):
        return self.o.population

    def get_1st_order_average(self):
        return self.o.first_order

    def get_purity(self):
        return self.o.purity

    def get_magnitude(self):
        return self.o.magnitude

    def get_error_derivative(self):
        return self.o.derror

    def get_population_derivative(self):
        return self.o.dpopulation

    def get_purity_derivative(self):
        return self.o.dpurity


class TrainingComplexKernel:
    def __init__(self, Parameter, TrainingSet, err=True, avg=True, deriv=False):
        self.o = orc.TrainingComplexKernel(Parameter, TrainingSet[0], TrainingSet[1], err, avg, deriv)

    def get_error(self):
        return self.o.error

    def get_purity(self):
        return self.o.purity

    def get_magnitude(self):
        return self.o.magnitude

    def get_error_derivative(self):
        return self.o.derror

    def get_purity_derivative(self):
        return self.o.dpurity


def loose_function(x, TrainingSet, ExtraTrainingSet, grad=False):
    return orc.loose_function(x, TrainingSet[0], TrainingSet[1], ExtraTrainingSet[0], ExtraTrainingSet[1], grad)


def observable_sums(model, pts, mass, pes_index):
    return orc.observable_sums(model, np.asarray(pts), mass, pes_index)


class Sampler:
    """Oracle twin of gaussian_process_liouville_equation_b200.mc.Sampler (same Philox streams, same call counter)."""

    def __init__(self, seed, analytic=None, kernels=None, new_point=None):
        self.seed, self.calls = int(seed), 0
        self.analytic, self.kernels, self.new_point = analytic, kernels, new_point

    def next_stream(self, element):
        self.calls += 1
        return self.calls * 4 + element

    def chains(self, pts, num_steps, max_displacement, row, col, want_chain=False, stream=None, chain0=0):
        stream = self.next_stream(row + col) if stream is None else stream
        analytic = None
        if self.analytic is not None:
            r0, s0, pop, ph = self.analytic
            analytic = [r0[0], r0[1], s0[0], s0[1], pop[0], pop[1], ph[0], ph[1]]
        k = [None, None, None] if self.kernels is None else [getattr(x, "o", x) for x in self.kernels]
        return orc.markov_chains(pts, num_steps, max_displacement, self.seed, stream, row, col, analytic=analytic, k00=k[0], k10=k[1], k11=k[2], new_point=self.new_point, want_chain=want_chain, chain0=chain0)

    def autocorrelation(self, chains):
        return orc.chain_autocorrelation(chains)

    def density(self, r, row, col):
        pts = np.zeros((len(r), 4))
        pts[:, :2] = r
        out, _, _ = self.chains(pts, 0, 1.0, row, col, stream=0)
        return out[:, 2] + 1j * out[:, 3]
