// nlopt_lite.hpp -- the four NLopt algorithms the reference's optimiser uses (gple/opt.h:51-55, gple/opt.cpp:333-336),
// written from their published descriptions behind NLopt's own C++ call shapes, so that host code written against
// <nlopt.hpp> (gple/opt.cpp) compiles against this header instead.  NLopt itself is an un-vendored, un-pinned
// dependency of the reference (gple/stdafx.h:52, makefile:4) and is not available here.
//
//   LN_NELDERMEAD  Nelder & Mead simplex (Comput. J. 7, 308 (1965)) with bound constraints by clipping (Box 1965),
//                  initial simplex x0 + step e_i as in NLopt's nldrmd
//   LD_SLSQP       in the reference it only ever runs as the subsidiary optimiser of AUGLAG_EQ, where it sees a smooth,
//                  bound-constrained problem; here that role is filled by a projected BFGS with Armijo backtracking
//   AUGLAG_EQ      augmented Lagrangian for the equality constraints (Birgin & Martinez, Optim. Methods Softw. 23, 177
//                  (2008); Conn, Gould & Toint 1991): L = f + sum lambda_i h_i + rho/2 sum h_i^2, lambda += rho h,
//                  rho *= 10 whenever the infeasibility did not halve
//   GN_DIRECT_L    DIviding RECTangles, locally biased variant (Jones et al. 1993; Gablonsky & Kelley 2001): size =
//                  longest side, at most one potentially optimal rectangle per size
// Parity with NLopt is on outcomes, not on iterates; the deterministic contract of the hot path is the callbacks.
// Parameters whose lower and upper bounds coincide are eliminated before the algorithm runs (NLopt does the same).
// No GPU dependency: tests/cpp/nlopt_lite_test.cpp exercises this header on analytic problems.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace nlopt_lite
{
enum algorithm
{
	LN_NELDERMEAD,
	LD_SLSQP,
	AUGLAG_EQ,
	GN_DIRECT_L
};
enum result
{
	SUCCESS = 1,
	FTOL_REACHED = 3,
	XTOL_REACHED = 4,
	MAXEVAL_REACHED = 5
};
/// nlopt::vfunc: grad.empty() means "no gradient wanted"
using vfunc = double (*)(const std::vector<double>& x, std::vector<double>& grad, void* data);
/// nlopt::mfunc: grad == nullptr means "no gradient wanted"; grad[i * n + j] = d result_i / d x_j
using mfunc = void (*)(unsigned m, double* result, unsigned n, const double* x, double* grad, void* data);

class opt
{
public:
	opt() = default;
	opt(const algorithm a, const unsigned n): Algo(a), Dim(n), Lower(n, -HUGE_VAL), Upper(n, HUGE_VAL) {}
	algorithm get_algorithm() const { return Algo; }
	unsigned get_dimension() const { return Dim; }
	const char* get_algorithm_name() const
	{
		switch (Algo)
		{
		case LN_NELDERMEAD:
			return "Nelder-Mead simplex algorithm (local, no-derivative)";
		case LD_SLSQP:
			return "projected BFGS in the role of SLSQP under AUGLAG (local, derivative)";
		case AUGLAG_EQ:
			return "Augmented Lagrangian method for equality constraints (needs sub-algorithm)";
		default:
			return "DIRECT-L (global, no-derivative)";
		}
	}
	void set_min_objective(const vfunc f, void* data)
	{
		Objective = f;
		ObjectiveData = data;
	}
	void add_equality_mconstraint(const mfunc c, void* data, const std::vector<double>& tol)
	{
		Constraint = c;
		ConstraintData = data;
		ConstraintTolerance = tol;
	}
	void remove_equality_constraints()
	{
		Constraint = nullptr;
		ConstraintTolerance.clear();
	}
	void set_lower_bounds(const std::vector<double>& lb) { Lower = checked(lb); }
	void set_upper_bounds(const std::vector<double>& ub) { Upper = checked(ub); }
	const std::vector<double>& get_lower_bounds() const { return Lower; }
	const std::vector<double>& get_upper_bounds() const { return Upper; }
	void set_xtol_rel(const double t) { XTolRel = t; }
	void set_ftol_rel(const double t) { FTolRel = t; }
	void set_xtol_abs(const double t) { XTolAbs = t; }
	void set_ftol_abs(const double t) { FTolAbs = t; }
	void set_maxeval(const int n) { MaxEval = n; }
	void set_initial_step(const double s) { InitialStep = s; }
	void set_local_optimizer(const opt& local) { LocalStore.assign(1, local); }
	/// nlopt::opt::force_stop, as a flag the caller owns: once it reads true the running optimisation stops at its next
	/// iteration (the result is then whatever was reached; a caller that raises the flag discards it)
	void set_stop_flag(const std::atomic<bool>* flag)
	{
		StopFlag = flag;
		for (opt& l : LocalStore)
		{
			l.set_stop_flag(flag);
		}
	}
	int get_numevals() const { return NumEvals; }

	/// Minimise from x (in place); returns the reason for stopping, the minimum in opt_f.
	result optimize(std::vector<double>& x, double& opt_f)
	{
		if (Objective == nullptr || x.size() != Dim)
		{
			throw std::invalid_argument("nlopt_lite::opt::optimize: objective missing or wrong dimension");
		}
		NumEvals = 0;
		for (unsigned i = 0; i < Dim; i++)
		{
			if (Lower[i] > Upper[i])
			{
				throw std::invalid_argument("nlopt_lite::opt::optimize: lower bound above upper bound");
			}
			x[i] = std::min(std::max(x[i], Lower[i]), Upper[i]);
		}
		// eliminate the pinned dimensions
		Free.clear();
		for (unsigned i = 0; i < Dim; i++)
		{
			if (Upper[i] > Lower[i])
			{
				Free.push_back(i);
			}
		}
		Full = x;
		std::vector<double> z(Free.size()), lo(Free.size()), hi(Free.size());
		for (std::size_t k = 0; k < Free.size(); k++)
		{
			z[k] = x[Free[k]];
			lo[k] = Lower[Free[k]];
			hi[k] = Upper[Free[k]];
		}
		result r = SUCCESS;
		if (Free.empty())
		{
			opt_f = value(z);
		}
		else
		{
			switch (Algo)
			{
			case LN_NELDERMEAD:
				r = nelder_mead(z, lo, hi, opt_f);
				break;
			case LD_SLSQP:
				r = projected_bfgs(z, lo, hi, opt_f, [this](const std::vector<double>& p, std::vector<double>* g) { return value(p, g); });
				break;
			case AUGLAG_EQ:
				r = auglag(z, lo, hi, opt_f);
				break;
			default:
				r = direct_l(z, lo, hi, opt_f);
				break;
			}
		}
		for (std::size_t k = 0; k < Free.size(); k++)
		{
			x[Free[k]] = z[k];
		}
		return r;
	}

private:
	algorithm Algo = LN_NELDERMEAD;
	unsigned Dim = 0;
	vfunc Objective = nullptr;
	void* ObjectiveData = nullptr;
	mfunc Constraint = nullptr;
	void* ConstraintData = nullptr;
	std::vector<double> ConstraintTolerance;
	std::vector<double> Lower, Upper;
	double XTolRel = 0.0, FTolRel = 0.0, XTolAbs = 0.0, FTolAbs = 0.0, InitialStep = 0.0;
	int MaxEval = 0, NumEvals = 0;
	std::vector<opt> LocalStore; // at most one element: the subsidiary optimiser (by value; a vector allows the recursive member)
	// scratch of the running optimisation
	std::vector<unsigned> Free;
	std::vector<double> Full;

	static constexpr double FeasibilityFloor = 1e-6;
	static const std::vector<double>& checked(const std::vector<double>& v) { return v; }
	const std::atomic<bool>* StopFlag = nullptr;
	bool budget_left() const { return (MaxEval <= 0 || NumEvals < MaxEval) && !(StopFlag != nullptr && StopFlag->load(std::memory_order_relaxed)); }

	/// objective in the reduced coordinates
	double value(const std::vector<double>& z, std::vector<double>* grad = nullptr)
	{
		for (std::size_t k = 0; k < Free.size(); k++)
		{
			Full[Free[k]] = z[k];
		}
		std::vector<double> g(grad != nullptr ? Dim : 0);
		NumEvals++;
		const double f = Objective(Full, g, ObjectiveData);
		if (grad != nullptr)
		{
			grad->resize(Free.size());
			for (std::size_t k = 0; k < Free.size(); k++)
			{
				(*grad)[k] = g[Free[k]];
			}
		}
		return f;
	}
	/// equality constraints in the reduced coordinates; jac (m x free, row-major) optional
	void constraints(const std::vector<double>& z, std::vector<double>& h, std::vector<double>* jac)
	{
		const unsigned m = unsigned(ConstraintTolerance.size());
		for (std::size_t k = 0; k < Free.size(); k++)
		{
			Full[Free[k]] = z[k];
		}
		h.assign(m, 0.0);
		std::vector<double> J(jac != nullptr ? std::size_t(m) * Dim : 0);
		Constraint(m, h.data(), Dim, Full.data(), jac != nullptr ? J.data() : nullptr, ConstraintData);
		if (jac != nullptr)
		{
			jac->assign(std::size_t(m) * Free.size(), 0.0);
			for (unsigned i = 0; i < m; i++)
			{
				for (std::size_t k = 0; k < Free.size(); k++)
				{
					(*jac)[i * Free.size() + k] = J[std::size_t(i) * Dim + Free[k]];
				}
			}
		}
	}
	bool x_converged(const std::vector<double>& a, const std::vector<double>& b) const
	{
		for (std::size_t i = 0; i < a.size(); i++)
		{
			if (std::abs(a[i] - b[i]) > XTolAbs && std::abs(a[i] - b[i]) > XTolRel * 0.5 * (std::abs(a[i]) + std::abs(b[i])))
			{
				return false;
			}
		}
		return XTolAbs > 0.0 || XTolRel > 0.0;
	}
	bool f_converged(const double a, const double b) const
	{
		const double d = std::abs(a - b);
		return (FTolAbs > 0.0 && d <= FTolAbs) || (FTolRel > 0.0 && d <= FTolRel * 0.5 * (std::abs(a) + std::abs(b))) || (a == b && (FTolAbs > 0.0 || FTolRel > 0.0));
	}

	// ---------------------------------------------------------------------------------------- Nelder-Mead
	result nelder_mead(std::vector<double>& z, const std::vector<double>& lo, const std::vector<double>& hi, double& fmin)
	{
		const std::size_t n = z.size();
		const double step = InitialStep > 0.0 ? InitialStep : 1.0;
		auto clip = [&](std::vector<double>& p)
		{
			for (std::size_t i = 0; i < n; i++)
			{
				p[i] = std::min(std::max(p[i], lo[i]), hi[i]);
			}
		};
		std::vector<std::vector<double>> pts(n + 1, z);
		std::vector<double> f(n + 1);
		f[0] = value(z);
		for (std::size_t i = 0; i < n; i++)
		{
			double& c = pts[i + 1][i];
			c = z[i] + step;
			if (c > hi[i])
			{
				c = (hi[i] - z[i] > 0.1 * step) ? hi[i] : z[i] - step;
			}
			if (c < lo[i])
			{
				c = (z[i] - lo[i] > 0.1 * step) ? lo[i] : 0.5 * (lo[i] + hi[i]);
			}
			if (c == z[i])
			{
				c = 0.5 * (lo[i] + hi[i]);
			}
			f[i + 1] = value(pts[i + 1]);
		}
		std::vector<std::size_t> order(n + 1);
		std::vector<double> centroid(n), xr(n), xe(n), xc(n);
		result r = MAXEVAL_REACHED;
		while (budget_left())
		{
			std::iota(order.begin(), order.end(), 0);
			std::sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return f[a] < f[b]; });
			const std::size_t best = order[0], worst = order[n], second = order[n - 1];
			if (f_converged(f[best], f[worst]))
			{
				r = FTOL_REACHED;
				break;
			}
			bool xconv = XTolAbs > 0.0 || XTolRel > 0.0;
			for (std::size_t k = 1; k <= n && xconv; k++)
			{
				xconv = x_converged(pts[best], pts[order[k]]);
			}
			if (xconv)
			{
				r = XTOL_REACHED;
				break;
			}
			for (std::size_t i = 0; i < n; i++)
			{
				double s = 0.0;
				for (std::size_t k = 0; k < n; k++)
				{
					s += pts[order[k]][i];
				}
				centroid[i] = s / double(n);
			}
			auto along = [&](const double t, std::vector<double>& out)
			{
				for (std::size_t i = 0; i < n; i++)
				{
					out[i] = centroid[i] + t * (pts[worst][i] - centroid[i]);
				}
				clip(out);
			};
			along(-1.0, xr);
			const double fr = value(xr);
			if (fr < f[best])
			{
				along(-2.0, xe);
				const double fe = value(xe);
				if (fe < fr)
				{
					pts[worst] = xe;
					f[worst] = fe;
				}
				else
				{
					pts[worst] = xr;
					f[worst] = fr;
				}
			}
			else if (fr < f[second])
			{
				pts[worst] = xr;
				f[worst] = fr;
			}
			else
			{
				const bool outside = fr < f[worst];
				along(outside ? -0.5 : 0.5, xc);
				const double fc = value(xc);
				if (fc < (outside ? fr : f[worst]))
				{
					pts[worst] = xc;
					f[worst] = fc;
				}
				else
				{
					for (std::size_t k = 1; k <= n; k++)
					{
						auto& p = pts[order[k]];
						for (std::size_t i = 0; i < n; i++)
						{
							p[i] = pts[best][i] + 0.5 * (p[i] - pts[best][i]);
						}
						f[order[k]] = value(p);
					}
				}
			}
		}
		const std::size_t b = std::size_t(std::min_element(f.begin(), f.end()) - f.begin());
		z = pts[b];
		fmin = f[b];
		return r;
	}

	// ---------------------------------------------------------------------------------------- projected BFGS
	template <typename F>
	result projected_bfgs(std::vector<double>& z, const std::vector<double>& lo, const std::vector<double>& hi, double& fmin, F&& fg)
	{
		const std::size_t n = z.size();
		std::vector<double> g(n), gn(n), d(n), zn(n), s(n), y(n), Hy(n);
		std::vector<double> H(n * n, 0.0);
		auto reset = [&]()
		{
			std::fill(H.begin(), H.end(), 0.0);
			for (std::size_t i = 0; i < n; i++)
			{
				H[i * n + i] = 1.0;
			}
		};
		reset();
		double f = fg(z, &g);
		result r = MAXEVAL_REACHED;
		while (budget_left())
		{
			// variables held at a bound by the gradient
			std::vector<char> active(n, 0);
			double pg = 0.0;
			for (std::size_t i = 0; i < n; i++)
			{
				active[i] = (z[i] <= lo[i] && g[i] > 0.0) || (z[i] >= hi[i] && g[i] < 0.0);
				pg += active[i] ? 0.0 : g[i] * g[i];
			}
			if (pg == 0.0)
			{
				r = SUCCESS;
				break;
			}
			double slope = 0.0;
			for (std::size_t i = 0; i < n; i++)
			{
				double v = 0.0;
				if (!active[i])
				{
					for (std::size_t j = 0; j < n; j++)
					{
						v -= active[j] ? 0.0 : H[i * n + j] * g[j];
					}
				}
				d[i] = v;
				slope += v * g[i];
			}
			if (!(slope < 0.0))
			{
				reset();
				slope = 0.0;
				for (std::size_t i = 0; i < n; i++)
				{
					d[i] = active[i] ? 0.0 : -g[i];
					slope += d[i] * g[i];
				}
			}
			// Armijo backtracking along the projected path
			double alpha = 1.0, fn = f;
			bool accepted = false;
			for (int it = 0; it < 40 && budget_left(); it++)
			{
				double decrease = 0.0;
				for (std::size_t i = 0; i < n; i++)
				{
					zn[i] = std::min(std::max(z[i] + alpha * d[i], lo[i]), hi[i]);
					decrease += g[i] * (zn[i] - z[i]);
				}
				fn = fg(zn, &gn);
				if (std::isfinite(fn) && fn <= f + 1e-4 * decrease)
				{
					accepted = true;
					break;
				}
				alpha *= 0.5;
			}
			if (!accepted)
			{
				r = FTOL_REACHED; // no further progress possible along a descent direction
				break;
			}
			double sy = 0.0, ss = 0.0, yy = 0.0;
			for (std::size_t i = 0; i < n; i++)
			{
				s[i] = zn[i] - z[i];
				y[i] = gn[i] - g[i];
				sy += s[i] * y[i];
				ss += s[i] * s[i];
				yy += y[i] * y[i];
			}
			const bool xc = x_converged(z, zn), fc = f_converged(f, fn);
			z = zn;
			g = gn;
			f = fn;
			if (xc)
			{
				r = XTOL_REACHED;
				break;
			}
			if (fc)
			{
				r = FTOL_REACHED;
				break;
			}
			if (sy > 1e-10 * std::sqrt(ss * yy))
			{
				// H <- (I - s y^T / sy) H (I - y s^T / sy) + s s^T / sy
				double yHy = 0.0;
				for (std::size_t i = 0; i < n; i++)
				{
					double v = 0.0;
					for (std::size_t j = 0; j < n; j++)
					{
						v += H[i * n + j] * y[j];
					}
					Hy[i] = v;
					yHy += v * y[i];
				}
				for (std::size_t i = 0; i < n; i++)
				{
					for (std::size_t j = 0; j < n; j++)
					{
						H[i * n + j] += (1.0 + yHy / sy) * s[i] * s[j] / sy - (Hy[i] * s[j] + s[i] * Hy[j]) / sy;
					}
				}
			}
		}
		fmin = f;
		return r;
	}

	// ---------------------------------------------------------------------------------------- augmented Lagrangian
	result auglag(std::vector<double>& z, const std::vector<double>& lo, const std::vector<double>& hi, double& fmin)
	{
		const std::size_t n = z.size(), m = ConstraintTolerance.size();
		const opt inner = LocalStore.empty() ? opt(LD_SLSQP, Dim) : LocalStore.front();
		if (Constraint == nullptr || m == 0)
		{
			return projected_bfgs(z, lo, hi, fmin, [this](const std::vector<double>& p, std::vector<double>* g) { return value(p, g); });
		}
		std::vector<double> lambda(m, 0.0), h(m), jac, gf;
		double rho = 1.0;
		auto lagrangian = [&](const std::vector<double>& p, std::vector<double>* g)
		{
			const double f = value(p, g != nullptr ? &gf : nullptr);
			constraints(p, h, g != nullptr ? &jac : nullptr);
			double L = f;
			for (std::size_t i = 0; i < m; i++)
			{
				L += lambda[i] * h[i] + 0.5 * rho * h[i] * h[i];
			}
			if (g != nullptr)
			{
				*g = gf;
				for (std::size_t i = 0; i < m; i++)
				{
					const double c = lambda[i] + rho * h[i];
					for (std::size_t k = 0; k < n; k++)
					{
						(*g)[k] += c * jac[i * n + k];
					}
				}
			}
			return L;
		};
		double f = value(z);
		constraints(z, h, nullptr);
		double h2 = 0.0, icm = 0.0;
		for (std::size_t i = 0; i < m; i++)
		{
			h2 += h[i] * h[i];
			icm = std::max(icm, std::abs(h[i]));
		}
		rho = std::max(1e-6, std::min(10.0, h2 > 0.0 ? 2.0 * std::abs(f) / h2 : 10.0));
		std::vector<double> zprev = z;
		double fprev = f;
		int stagnant = 0;
		result r = MAXEVAL_REACHED;
		// the inner solver runs with the subsidiary optimiser's tolerances
		const double sub_xr = inner.XTolRel > 0.0 ? inner.XTolRel : XTolRel, sub_fr = inner.FTolRel > 0.0 ? inner.FTolRel : FTolRel;
		const double sub_xa = inner.XTolAbs, sub_fa = inner.FTolAbs;
		for (int outer = 0; outer < 200 && budget_left(); outer++)
		{
			double L = 0.0;
			{
				const double xr = XTolRel, fr = FTolRel, xa = XTolAbs, fa = FTolAbs;
				XTolRel = sub_xr;
				FTolRel = sub_fr;
				XTolAbs = sub_xa;
				FTolAbs = sub_fa;
				projected_bfgs(z, lo, hi, L, lagrangian);
				XTolRel = xr;
				FTolRel = fr;
				XTolAbs = xa;
				FTolAbs = fa;
			}
			f = value(z);
			constraints(z, h, nullptr);
			double icm_new = 0.0;
			bool feasible = true;
			for (std::size_t i = 0; i < m; i++)
			{
				icm_new = std::max(icm_new, std::abs(h[i]));
				// the reference passes tolerance 0 (opt.cpp:756, 968), with which NLopt's AUGLAG only ever stops on an
				// exact zero or an exception of the inner solver; a floor keeps the stopping rule meaningful
				feasible = feasible && std::abs(h[i]) <= std::max(ConstraintTolerance[i], FeasibilityFloor);
				lambda[i] = std::min(std::max(lambda[i] + rho * h[i], -1e20), 1e20);
			}
			const bool improved = icm_new <= 0.5 * icm;
			if (!improved)
			{
				rho *= 10.0;
			}
			stagnant = improved ? 0 : stagnant + 1;
			icm = icm_new;
			const bool xc = x_converged(zprev, z), fc = f_converged(fprev, f);
			if (feasible && (xc || fc))
			{
				r = xc ? XTOL_REACHED : FTOL_REACHED;
				break;
			}
			if (icm == 0.0 || (xc && stagnant >= 4) || rho > 1e12)
			{
				r = SUCCESS; // exactly feasible, or the infeasibility no longer responds to the penalty
				break;
			}
			zprev = z;
			fprev = f;
		}
		fmin = f;
		return r;
	}

	// ---------------------------------------------------------------------------------------- DIRECT-L
	result direct_l(std::vector<double>& z, const std::vector<double>& lo, const std::vector<double>& hi, double& fmin)
	{
		const std::size_t n = z.size();
		struct Rect
		{
			std::vector<double> c; // centre in the unit cube
			std::vector<int> level; // number of trisections per dimension
			double f;
			int minlevel;
		};
		auto eval = [&](const std::vector<double>& u)
		{
			std::vector<double> p(n);
			for (std::size_t i = 0; i < n; i++)
			{
				p[i] = lo[i] + u[i] * (hi[i] - lo[i]);
			}
			return value(p);
		};
		std::vector<Rect> rects;
		rects.push_back(Rect{std::vector<double>(n, 0.5), std::vector<int>(n, 0), 0.0, 0});
		rects[0].f = eval(rects[0].c);
		double best = rects[0].f;
		std::vector<double> bestc = rects[0].c;
		const double eps = 1e-4;
		const int maxeval = MaxEval > 0 ? MaxEval : 2000;
		result r = MAXEVAL_REACHED;
		int stall = 0;
		while (NumEvals < maxeval && budget_left())
		{
			// best rectangle of every size class (size = longest side = 3^-minlevel)
			int maxl = 0;
			for (const auto& q : rects)
			{
				maxl = std::max(maxl, q.minlevel);
			}
			std::vector<int> rep(maxl + 1, -1);
			for (std::size_t k = 0; k < rects.size(); k++)
			{
				int& slot = rep[rects[k].minlevel];
				if (slot < 0 || rects[k].f < rects[slot].f)
				{
					slot = int(k);
				}
			}
			// lower-right convex hull over (size, f) of the representatives, largest size first
			std::vector<int> cand;
			for (int l = 0; l <= maxl; l++)
			{
				if (rep[l] >= 0)
				{
					cand.push_back(rep[l]);
				}
			}
			auto size_of = [&](const int k) { return std::pow(3.0, -rects[k].minlevel); };
			std::vector<int> hull;
			for (const int k : cand) // one representative per size, sizes strictly decreasing
			{
				hull.push_back(k);
			}
			std::vector<int> selected;
			for (std::size_t a = 0; a < hull.size(); a++)
			{
				const int j = hull[a];
				const double dj = size_of(j), fj = rects[j].f;
				// potentially optimal: exists K >= 0 with fj - K dj <= fi - K di for all i, and fj - K dj <= best - eps |best|
				double Klow = 0.0, Khigh = HUGE_VAL;
				bool ok = true;
				for (const int i : hull)
				{
					if (i == j)
					{
						continue;
					}
					const double di = size_of(i), fi = rects[i].f;
					if (di < dj)
					{
						Klow = std::max(Klow, (fj - fi) / (dj - di));
					}
					else
					{
						Khigh = std::min(Khigh, (fi - fj) / (di - dj));
					}
				}
				if (Klow > Khigh)
				{
					ok = false;
				}
				if (ok && Khigh < HUGE_VAL && fj - Khigh * dj > best - eps * std::abs(best))
				{
					ok = false; // cannot improve on the incumbent by more than eps |f_min|
				}
				if (ok)
				{
					selected.push_back(j);
				}
			}
			if (selected.empty())
			{
				selected.push_back(cand.front());
			}
			const double before = best;
			for (const int j : selected)
			{
				if (NumEvals >= maxeval)
				{
					break;
				}
				// sample along the longest sides
				const int lvl = rects[j].minlevel;
				const double delta = std::pow(3.0, -(lvl + 1));
				struct Probe
				{
					std::size_t dim;
					double fp, fm, w;
				};
				std::vector<Probe> probes;
				for (std::size_t i = 0; i < n && NumEvals + 2 <= maxeval + 1; i++)
				{
					if (rects[j].level[i] != lvl)
					{
						continue;
					}
					std::vector<double> cp = rects[j].c, cm = rects[j].c;
					cp[i] += delta;
					cm[i] -= delta;
					const double fp = eval(cp), fm = eval(cm);
					probes.push_back(Probe{i, fp, fm, std::min(fp, fm)});
					if (fp < best)
					{
						best = fp;
						bestc = cp;
					}
					if (fm < best)
					{
						best = fm;
						bestc = cm;
					}
				}
				std::sort(probes.begin(), probes.end(), [](const Probe& a, const Probe& b) { return a.w < b.w; });
				for (const Probe& pr : probes)
				{
					rects[j].level[pr.dim]++;
					Rect rp = rects[j], rm = rects[j];
					rp.c[pr.dim] += delta;
					rm.c[pr.dim] -= delta;
					rp.f = pr.fp;
					rm.f = pr.fm;
					rp.minlevel = *std::min_element(rp.level.begin(), rp.level.end());
					rm.minlevel = rp.minlevel;
					rects.push_back(std::move(rp));
					rects.push_back(std::move(rm));
				}
				rects[j].minlevel = *std::min_element(rects[j].level.begin(), rects[j].level.end());
			}
			// convergence: the best value has not moved by more than the tolerances for 5 sweeps and the best box is small
			stall = f_converged(before, best) ? stall + 1 : 0;
			if (stall >= 5 && (FTolRel > 0.0 || FTolAbs > 0.0))
			{
				r = FTOL_REACHED;
				break;
			}
		}
		for (std::size_t i = 0; i < n; i++)
		{
			z[i] = lo[i] + bestc[i] * (hi[i] - lo[i]);
		}
		fmin = best;
		return r;
	}
};
} // namespace nlopt_lite
