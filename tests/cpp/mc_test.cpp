// monte_carlo_selection through the C++ host (host/gple_mc.hpp) on the analytic initial distribution, as gple/main.cpp:35-57
// does at start-up.  Reads start points (count, then x p per line); prints the tuned parameters and the walked points.
#include "../../gaussian_process_liouville_equation_b200/host/gple_mc.hpp"

#include <cstdio>
#include <fstream>

using namespace gple_host;

int main(int argc, char** argv)
{
	if (argc < 3)
	{
		return 2;
	}
	std::ifstream in(argv[1]);
	const unsigned long long seed = std::strtoull(argv[2], nullptr, 10);
	std::size_t n = 0;
	in >> n;
	AllPoints density;
	density[0].resize(n);
	for (auto& p : density[0])
	{
		in >> p.r[0] >> p.r[1];
	}
	density[1] = density[0];
	const double sp = 0.7056, sx = 1.0 / (2.0 * sp);
	Sampler sampler(seed, {-10.0, 14.112}, {sx, sp}, {0.8, 0.6}, {0.0, 0.4});
	std::array<MCParameters, NumElements> params;
	monte_carlo_selection(density, params, sampler);
	for (std::size_t e = 0; e < 2; e++)
	{
		std::printf("displacement%zu %.17g\nsteps%zu %zu\n", e, params[e].get_max_displacement(), e, params[e].get_num_MC_steps());
		for (std::size_t i = 0; i < n; i++)
		{
			std::printf("p%zu_%zu_x %.17g\np%zu_%zu_p %.17g\np%zu_%zu_re %.17g\np%zu_%zu_im %.17g\n", e, i, density[e][i].r[0], e, i, density[e][i].r[1], e, i, density[e][i].rho.real(), e, i, density[e][i].rho.imag());
		}
	}
	return 0;
}
