// Runs the C++ optimiser mirror (host/gple_opt.hpp) the way gple/main.cpp:103-131 drives the reference's Optimization:
// reads the density and the extra points written by tests/test_opt_cpp.py (text: count, then x p re im per line, for
// each of the three elements, density first, then extra points), optimises, and prints "name value" lines.
#include "../../gaussian_process_liouville_equation_b200/host/gple_opt.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>

using namespace gple_host;

static AllPoints read_points(std::ifstream& in)
{
	AllPoints pts;
	for (auto& e : pts)
	{
		std::size_t n = 0;
		in >> n;
		e.resize(n);
		for (auto& p : e)
		{
			double re = 0.0, im = 0.0;
			in >> p.r[0] >> p.r[1] >> re >> im;
			p.rho = {re, im};
		}
	}
	return pts;
}

int main(int argc, char** argv)
{
	if (argc < 6)
	{
		std::fprintf(stderr, "usage: opt_test points.txt pes_model mass total_energy purity [max_global] [max_constrained]\n");
		return 2;
	}
	std::ifstream in(argv[1]);
	const AllPoints density = read_points(in), extra = read_points(in);
	const int pes_model = std::atoi(argv[2]);
	const double mass = std::atof(argv[3]), e0 = std::atof(argv[4]), purity = std::atof(argv[5]);
	const double sp = 0.7056, sx = 1.0 / (2.0 * sp);
	Optimization optimizer({sx, sp}, {20.0, 40.0}, mass, pes_model, e0, purity);
	optimizer.set_maximum_evaluations(argc > 6 ? std::atoi(argv[6]) : 200, argc > 7 ? std::atoi(argv[7]) : 1000);
	if (const char* spec = std::getenv("GPLE_SPECULATIVE_RESTARTS")) // "0": the restart stages one after the other
	{
		optimizer.set_speculative_restarts(std::atoi(spec) != 0);
	}
	const auto t0 = std::chrono::steady_clock::now();
	const auto [err, steps, type] = optimizer.optimize(density, extra);
	const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	const TrainingKernels k(optimizer.get_parameters(), construct_training_sets(density), true, true, false);
	std::printf("error %.17g\ntype %d\nwall_s %.6f\n", err, int(type), wall);
	std::size_t total = 0;
	for (std::size_t i = 0; i < steps.size(); i++)
	{
		std::printf("steps%zu %zu\n", i, steps[i]);
		total += steps[i];
	}
	std::printf("evaluations %zu\n", total);
	std::printf("population %.17g\npurity %.17g\n", k.calculate_population(), k.calculate_purity());
	const QuantumVectorD E = calculate_total_energy_average_each_surface(density, mass, pes_model);
	std::printf("energy %.17g\n", k.calculate_total_energy_average(E));
	const auto& pv = optimizer.get_parameters();
	for (std::size_t e = 0; e < NumElements; e++)
	{
		for (std::size_t p = 0; p < pv[e].size(); p++)
		{
			std::printf("theta%zu_%zu %.17g\n", e, p, pv[e][p]);
		}
	}
	const auto lb = optimizer.get_lower_bounds(), ub = optimizer.get_upper_bounds();
	std::printf("lb0_1 %.17g\nub0_1 %.17g\n", lb[0][1], ub[0][1]);
	return 0;
}
