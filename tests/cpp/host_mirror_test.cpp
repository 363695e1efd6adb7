// Exercises the C++ host mirror (gple_host.hpp) the way gple/main.cpp:160-176 uses the reference classes:
// build TrainingKernels from the density, evolve one step, read the analytic averages.  Prints key numbers
// as "name value" lines; tests/test_host_mirror.py compares them with the oracle.
#include "../../gaussian_process_liouville_equation_b200/host/gple_host.hpp"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>

using namespace gple_host;

int main(int argc, char** argv)
{
	const std::size_t n = 200;
	const double sx = 1.0 / (2.0 * 0.7056), sp = 0.7056, p0 = 14.112;
	AllPoints density;
	// deterministic low-discrepancy points (no RNG dependency between C++ and Python)
	for (std::size_t e = 0; e < NumElements; e++)
	{
		for (std::size_t i = 0; i < n; i++)
		{
			const double u = std::fmod(0.5 + 0.6180339887498949 * double(i + 1), 1.0), v = std::fmod(0.5 + 0.7548776662466927 * double(i + 1), 1.0);
			const double x = -0.8 + sx * 3.0 * (2.0 * u - 1.0), p = p0 + sp * 3.0 * (2.0 * v - 1.0);
			const double g = std::exp(-0.5 * (std::pow((x + 0.8) / sx, 2) + std::pow((p - p0) / sp, 2))) / (2.0 * M_PI * sx * sp);
			std::complex<double> rho = (e == 0 ? 0.6 : 0.4) * g;
			if (e == 1)
			{
				rho = std::sqrt(0.24) * g * std::exp(std::complex<double>(0.0, 0.7 * (x + 0.8) - 0.2 * (p - p0)));
			}
			density[e].push_back(PhaseSpacePoint{{x, p}, rho});
		}
	}
	const ParameterVector tr{1.0, sx, sp, 1e-2}, tc{1.0, 1.2, 0.8 * sx, 1.1 * sp, 0.7, 1.1 * sx, 0.9 * sp, 2e-2};
	{
		// gple/kernel.h:29-106, complex_kernel.h:14-145: kernel matrices and their derivative arrays of the first 12 points
		PhasePoints f(12);
		for (std::size_t i = 0; i < 12; i++)
		{
			f(0, i) = density[0][i].r[0];
			f(1, i) = density[0][i].r[1];
		}
		const KernelBase kb({1.0, {sx, sp}, 1e-2}, f, f, true);
		std::printf("kb_K_3_5 %.17g\nkb_K_4_4 %.17g\nkb_dK1_3_5 %.17g\nkb_dK3_4_4 %.17g\n", kb.get_kernel()(3, 5), kb.get_kernel()(4, 4), kb.get_derivative()[1](3, 5), kb.get_derivative()[3](4, 4));
		const ComplexKernelBase ckb({1.0, 1.2, 0.8 * sx, 1.1 * sp, 0.7, 1.1 * sx, 0.9 * sp, 2e-2}, f, f, true);
		std::printf("ckb_Kt_3_5_im %.17g\nckb_dKt2_3_5_im %.17g\nckb_dK7_4_4 %.17g\n", ckb.get_pseudo_kernel()(3, 5).imag(), ckb.get_pseudo_derivatives()[2](3, 5).imag(), ckb.get_derivatives()[7](4, 4));
	}
	const TrainingKernels kernels({tr, tc, tr}, density);
	std::printf("population %.17g\n", kernels.calculate_population());
	std::printf("purity %.17g\n", kernels.calculate_purity());
	std::printf("error00 %.17g\n", kernels.Diagonal[0]->get_error());
	std::printf("error10 %.17g\n", kernels.OffDiagonal->get_error());
	PhasePoints q(2);
	q(0, 0) = -0.7;
	q(1, 0) = p0 + 0.1;
	q(0, 1) = 3.0;
	q(1, 1) = p0;
	const PredictiveKernel pk(q, *kernels.Diagonal[0], false);
	std::printf("cutoff0 %.17g\ncutoff1 %.17g\nvar0 %.17g\n", pk.get_cutoff_prediction()[0], pk.get_cutoff_prediction()[1], pk.get_variance()[0]);
	if (argc > 1)
	{
		// gple/main.cpp:118 output_phase on a small grid (gple/input.cpp:39-71 builds it the same way: x-major)
		PhasePoints grid(5 * 4);
		for (std::size_t ix = 0; ix < 5; ix++)
		{
			for (std::size_t ip = 0; ip < 4; ip++)
			{
				grid(0, ix * 4 + ip) = -0.8 + sx * (double(ix) - 2.0);
				grid(1, ix * 4 + ip) = p0 + sp * (double(ip) - 1.5) * 1.5;
			}
		}
		std::ofstream phase(std::string(argv[1]) + "/phase.txt"), variance(std::string(argv[1]) + "/var.txt");
		phase << std::setprecision(17);
		variance << std::setprecision(17);
		output_phase(phase, variance, kernels, grid);
	}
	evolve(density, 2000.0, 2.0, kernels, GPLE_DAC);
	std::printf("x00_0 %.17g\nrho00_0 %.17g\nrho10_5_re %.17g\nrho10_5_im %.17g\n", density[0][0].r[0], density[0][0].rho.real(), density[1][5].rho.real(), density[1][5].rho.imag());
	const auto pop = calculate_population_each_surface(density, 2000.0, GPLE_DAC);
	std::printf("pop0 %.17g\n", pop[0]);
	std::printf("E1_at_0.3 %.17g\n", adiabatic_potential(0.3, GPLE_DAC)[1]);
	return 0;
}
