// CPU-only checks of host/nlopt_lite.hpp on analytic problems (run by tests/test_nlopt_lite.py).
#include "../../gaussian_process_liouville_equation_b200/host/nlopt_lite.hpp"

#include <atomic>
#include <cstdio>
#include <cstdlib>

namespace
{
int failures = 0;
void expect(const bool ok, const char* what)
{
	std::printf("%s %s\n", ok ? "ok  " : "FAIL", what);
	failures += ok ? 0 : 1;
}
double rosenbrock(const std::vector<double>& x, std::vector<double>& g, void*)
{
	const double a = 1.0 - x[0], b = x[1] - x[0] * x[0];
	if (!g.empty())
	{
		g[0] = -2.0 * a - 400.0 * x[0] * b;
		g[1] = 200.0 * b;
	}
	return a * a + 100.0 * b * b;
}
/// minimise x^2 + 2 y^2 + z^2 subject to x + y + z = 1, x - y = 0.25; z is pinned by its bounds in one test
double quad(const std::vector<double>& x, std::vector<double>& g, void*)
{
	if (!g.empty())
	{
		g[0] = 2.0 * x[0];
		g[1] = 4.0 * x[1];
		g[2] = 2.0 * x[2];
	}
	return x[0] * x[0] + 2.0 * x[1] * x[1] + x[2] * x[2];
}
void quad_con(unsigned m, double* r, unsigned n, const double* x, double* grad, void*)
{
	r[0] = x[0] + x[1] + x[2] - 1.0;
	if (m > 1)
	{
		r[1] = x[0] - x[1] - 0.25;
	}
	if (grad != nullptr)
	{
		const double J[2][3] = {{1.0, 1.0, 1.0}, {1.0, -1.0, 0.0}};
		for (unsigned i = 0; i < m; i++)
		{
			for (unsigned j = 0; j < n; j++)
			{
				grad[i * n + j] = J[i][j];
			}
		}
	}
}
/// six-hump camel: global minima -1.0316 at (0.0898, -0.7126) and (-0.0898, 0.7126), several local ones
double camel(const std::vector<double>& x, std::vector<double>&, void*)
{
	const double a = x[0], b = x[1];
	return (4.0 - 2.1 * a * a + a * a * a * a / 3.0) * a * a + a * b + (-4.0 + 4.0 * b * b) * b * b;
}
void tolerances(nlopt_lite::opt& o)
{
	o.set_xtol_rel(1e-5); // gple/opt.cpp:342-355
	o.set_ftol_rel(1e-5);
	o.set_xtol_abs(1e-15);
	o.set_ftol_abs(1e-15);
}
} // namespace

int main()
{
	using namespace nlopt_lite;
	{
		opt o(LN_NELDERMEAD, 2);
		tolerances(o);
		o.set_initial_step(0.5);
		o.set_lower_bounds({-2.0, -2.0});
		o.set_upper_bounds({2.0, 2.0});
		o.set_min_objective(rosenbrock, nullptr);
		o.set_maxeval(4000);
		std::vector<double> x{-1.2, 1.0};
		double f = 0.0;
		o.optimize(x, f);
		expect(std::abs(x[0] - 1.0) < 1e-2 && std::abs(x[1] - 1.0) < 2e-2 && f < 1e-4, "Nelder-Mead reaches the Rosenbrock minimum");
		expect(o.get_numevals() > 10 && o.get_numevals() <= 4000, "Nelder-Mead counts evaluations");
	}
	{
		// the minimum (1, 1) is outside the box: the bound-constrained minimum is on the face x = 0.5
		opt o(LN_NELDERMEAD, 2);
		tolerances(o);
		o.set_initial_step(0.5);
		o.set_lower_bounds({-2.0, -2.0});
		o.set_upper_bounds({0.5, 2.0});
		o.set_min_objective(rosenbrock, nullptr);
		o.set_maxeval(4000);
		std::vector<double> x{-1.2, 1.0};
		double f = 0.0;
		o.optimize(x, f);
		expect(std::abs(x[0] - 0.5) < 1e-3 && std::abs(x[1] - 0.25) < 1e-2 && std::abs(f - 0.25) < 1e-4, "Nelder-Mead respects bounds");
	}
	{
		opt o(LD_SLSQP, 2);
		tolerances(o);
		o.set_lower_bounds({-2.0, -2.0});
		o.set_upper_bounds({2.0, 2.0});
		o.set_min_objective(rosenbrock, nullptr);
		o.set_maxeval(2000);
		std::vector<double> x{-1.2, 1.0};
		double f = 0.0;
		o.optimize(x, f);
		expect(std::abs(x[0] - 1.0) < 1e-3 && std::abs(x[1] - 1.0) < 2e-3, "projected BFGS reaches the Rosenbrock minimum");
	}
	{
		opt o(AUGLAG_EQ, 3);
		tolerances(o);
		opt sub(LD_SLSQP, 3);
		tolerances(sub);
		o.set_local_optimizer(sub);
		o.set_lower_bounds({-5.0, -5.0, -5.0});
		o.set_upper_bounds({5.0, 5.0, 5.0});
		o.set_min_objective(quad, nullptr);
		o.add_equality_mconstraint(quad_con, nullptr, {1e-8, 1e-8});
		o.set_maxeval(5000);
		std::vector<double> x{0.0, 0.0, 0.0};
		double f = 0.0;
		o.optimize(x, f);
		// x = y + 1/4, z = 3/4 - 2y: d/dy [(y + 1/4)^2 + 2 y^2 + (3/4 - 2y)^2] = 14 y - 5/2 = 0 -> y = 5/28; f is flat along the
		// feasible line (f'' = 14), so ftol_rel = 1e-5 bounds the position error by ~1e-3
		expect(std::abs(x[1] - 5.0 / 28.0) < 1e-3 && std::abs(x[0] - x[1] - 0.25) < 2e-6 && std::abs(x[0] + x[1] + x[2] - 1.0) < 2e-6 && std::abs(f - 45.0 / 112.0) < 1e-5, "augmented Lagrangian solves the equality-constrained quadratic");
	}
	{
		// z pinned at 0.5 by lb == ub (the reference pins magnitude and noise this way, opt.cpp:33-60)
		opt o(AUGLAG_EQ, 3);
		tolerances(o);
		opt sub(LD_SLSQP, 3);
		tolerances(sub);
		o.set_local_optimizer(sub);
		o.set_lower_bounds({-5.0, -5.0, 0.5});
		o.set_upper_bounds({5.0, 5.0, 0.5});
		o.set_min_objective(quad, nullptr);
		o.add_equality_mconstraint(quad_con, nullptr, {1e-8});
		o.set_maxeval(5000);
		std::vector<double> x{0.0, 0.0, 0.0};
		double f = 0.0;
		o.optimize(x, f);
		// x + y = 0.5, minimise x^2 + 2 y^2 -> x = 1/3, y = 1/6
		expect(x[2] == 0.5 && std::abs(x[0] - 1.0 / 3.0) < 1e-3 && std::abs(x[0] + x[1] - 0.5) < 2e-6 && std::abs(f - (1.0 / 9.0 + 2.0 / 36.0 + 0.25)) < 1e-5, "pinned parameters are eliminated");
	}
	{
		opt o(GN_DIRECT_L, 2);
		tolerances(o);
		o.set_lower_bounds({-3.0, -2.0});
		o.set_upper_bounds({3.0, 2.0});
		o.set_min_objective(camel, nullptr);
		o.set_maxeval(3000);
		std::vector<double> x{2.5, 1.5};
		double f = 0.0;
		o.optimize(x, f);
		expect(std::abs(f + 1.0316) < 1e-3 && std::abs(std::abs(x[0]) - 0.0898) < 2e-2 && std::abs(std::abs(x[1]) - 0.7126) < 2e-2, "DIRECT-L finds a global minimum of the six-hump camel function");
	}
	{
		// the caller's stop flag (nlopt's force_stop): raised from inside the objective after 25 evaluations, every algorithm
		// returns at its next iteration; a copy of the optimiser carries the flag along, the subsidiary optimiser too
		struct Counter
		{
			int calls = 0;
			std::atomic<bool> stop{false};
		};
		auto counted = [](const std::vector<double>& x, std::vector<double>& g, void* data)
		{
			Counter* c = static_cast<Counter*>(data);
			if (++c->calls >= 25)
			{
				c->stop.store(true);
			}
			return rosenbrock(x, g, nullptr);
		};
		bool all = true;
		for (const algorithm a : {LN_NELDERMEAD, LD_SLSQP, AUGLAG_EQ, GN_DIRECT_L})
		{
			Counter c;
			opt o(a, 2);
			o.set_xtol_rel(1e-14);
			o.set_ftol_rel(1e-14);
			if (a == AUGLAG_EQ)
			{
				opt sub(LD_SLSQP, 2);
				sub.set_xtol_rel(1e-14);
				sub.set_ftol_rel(1e-14);
				o.set_local_optimizer(sub);
			}
			o.set_lower_bounds({-5.0, -5.0});
			o.set_upper_bounds({5.0, 5.0});
			o.set_min_objective(counted, &c);
			o.set_maxeval(100000);
			o.set_stop_flag(&c.stop);
			opt running(o); // what Optimization::optimize hands to a stage that runs ahead of time
			std::vector<double> x{-1.2, 1.0};
			double f = 0.0;
			running.optimize(x, f);
			all = all && c.calls >= 25 && c.calls < 120 && running.get_numevals() == c.calls;
		}
		expect(all, "a raised stop flag ends every algorithm within one iteration");
	}
	return failures == 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
