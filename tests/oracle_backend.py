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
