// gple_opt.hpp -- C++ host mirror of the reference's optimiser (gple/opt.h, gple/opt.cpp) over the C-ABI.
//
// Same surface as the reference: class Optimization with optimize(density, extra_points) -> (error, steps, type),
// get_parameters(), get_lower_bounds(), get_upper_bounds(); the NLopt callbacks keep NLopt's shapes
//   objective : double f(const std::vector<double>& x, std::vector<double>& grad, void* params)      (opt.cpp:441, 594, 844)
//   constraint: void c(unsigned m, double* result, unsigned n, const double* x, double* grad, void*)  (opt.cpp:644, 879)
// with `grad.empty()` / `grad == nullptr` meaning "no gradient", and `params` pointing to tuples of references to the
// caller-owned training sets (opt.cpp:538-545, 752-756).  All numerics behind the callbacks run on the GPU
// (gple_loose_function, gple_train_*); the independent elements of one evaluation run concurrently, one library
// context (stream) per element -- the single-GPU form of the element x candidate sharding of SURVEY.md section 8e.
// NLopt is not available: host/nlopt_lite.hpp supplies the four algorithms behind NLopt's call shapes.
// Header-only; link with libgple_b200.so.
#pragma once
#include "gple_host.hpp"
#include "nlopt_lite.hpp"

#include <cfloat>
#include <numeric>

namespace gple_host
{
namespace nlopt = nlopt_lite;

constexpr double AverageTolerance = 0.05; // gple/opt.h:13
constexpr double InitialMagnitude = 1.0;  // gple/opt.cpp:25
constexpr double InitialNoise = 1e-2;	  // gple/opt.cpp:27
using Bounds = std::array<ParameterVector, 2>;
using QuantumVectorD = std::array<double, NumPES>;
using AllParameters = std::array<ParameterVector, NumElements>; // QuantumStorage<ParameterVector> in lower-triangular order

/// gple/opt.cpp:33-60: magnitude and noise pinned, characteristic lengths in [lb, ub]
inline Bounds calculate_kernel_bounds(const ClassicalPhaseVector& lb, const ClassicalPhaseVector& ub)
{
	return {ParameterVector{InitialMagnitude, lb[0], lb[1], InitialNoise}, ParameterVector{InitialMagnitude, ub[0], ub[1], InitialNoise}};
}
/// gple/opt.cpp:66-104
inline Bounds calculate_complex_kernel_bounds(const ClassicalPhaseVector& lb, const ClassicalPhaseVector& ub)
{
	return {
		ParameterVector{InitialMagnitude, InitialMagnitude / 10.0, lb[0], lb[1], InitialMagnitude / 10.0, lb[0], lb[1], InitialNoise},
		ParameterVector{InitialMagnitude, InitialMagnitude * 10.0, ub[0], ub[1], InitialMagnitude * 10.0, ub[0], ub[1], InitialNoise}};
}
/// Positions handled in log space by the global optimiser (gple/opt.cpp:109-145)
inline std::vector<std::size_t> log_positions(const std::size_t n)
{
	return n == 8 ? std::vector<std::size_t>{1, 4, 7} : std::vector<std::size_t>{3};
}
/// gple/opt.cpp:109-145
inline ParameterVector local_parameter_to_global(ParameterVector p)
{
	for (const std::size_t i : log_positions(p.size()))
	{
		p[i] = std::log(p[i]);
	}
	return p;
}
/// gple/opt.cpp:197-232
inline ParameterVector global_parameter_to_local(ParameterVector p)
{
	for (const std::size_t i : log_positions(p.size()))
	{
		p[i] = std::exp(p[i]);
	}
	return p;
}
/// gple/opt.cpp:155-191: d/d(ln x) = x d/dx
inline ParameterVector local_gradient_to_global(const ParameterVector& local, ParameterVector grad)
{
	for (const std::size_t i : log_positions(local.size()))
	{
		grad[i] *= local[i];
	}
	return grad;
}

using ElementTrainingParameters = std::tuple<const ElementTrainingSet&, const ElementTrainingSet&>;					   // opt.cpp:16
using AnalyticalLooseFunctionParameters = std::tuple<const AllTrainingSets&, const AllTrainingSets&>;					   // opt.cpp:19
using AnalyticalConstraintParameters = std::tuple<const AllTrainingSets&, const QuantumVectorD&, const double&, const double&>; // opt.cpp:22

/// loose_function (gple/opt.cpp:441-482): LOOCV error of the training set + squared error on the extra set.  Written like the
/// reference's body -- a TrainingKernel on the training set, then the prediction error on the extra set -- with the trained model
/// taken from the ModelCache (the averages are computed along, so that the constraint callback evaluated at the same x reuses it).
inline double loose_function(const ParameterVector& x, ParameterVector& grad, void* params)
{
	const auto& [ts, ets] = *static_cast<ElementTrainingParameters*>(params);
	const bool deriv = !grad.empty();
	const std::size_t M = std::get<0>(ets).cols();
	double value = std::numeric_limits<double>::quiet_NaN(), val = 0.0;
	ParameterVector vg(grad.size(), 0.0);
	auto validate = [&](const gple_model* m)
	{
		Context::check(gple_validation_error(Context::get(), m, std::get<0>(ets).data(), reinterpret_cast<const double*>(std::get<1>(ets).data()), M, &val, deriv ? vg.data() : nullptr), "loose_function");
	};
	if (x.size() == TrainingKernel::NumTotalParameters)
	{
		const auto k = ModelCache::instance().real(ts, x, true, true, deriv);
		if (k->status() == GPLE_OK)
		{
			validate(k->handle());
			value = k->get_error() + val;
			const auto d = k->get_error_derivative();
			for (std::size_t p = 0; p < grad.size(); p++)
			{
				grad[p] = d[p] + vg[p];
			}
		}
	}
	else
	{
		const auto k = ModelCache::instance().complex(ts, x, true, true, deriv);
		if (k->status() == GPLE_OK)
		{
			validate(k->handle());
			value = k->get_error() + val;
			const auto d = k->get_error_derivative();
			for (std::size_t p = 0; p < grad.size(); p++)
			{
				grad[p] = d[p] + vg[p];
			}
		}
	}
	if (!(value == value))
	{
		std::fill(grad.begin(), grad.end(), std::numeric_limits<double>::quiet_NaN()); // not positive definite: opt.cpp:476-480
	}
	make_normal(value);
	for (double& g : grad)
	{
		make_normal(g);
	}
	return value;
}
/// gple/opt.cpp:489-497
inline double loose_function_global_wrapper(const ParameterVector& x, ParameterVector& grad, void* params)
{
	const ParameterVector local = global_parameter_to_local(x);
	ParameterVector g(grad.size());
	const double r = loose_function(local, g, params);
	if (!grad.empty())
	{
		grad = local_gradient_to_global(local, g);
	}
	return r;
}

/// Sum of the elements' loose functions with the gradient scattered into `grad`; elements run concurrently
inline double sum_loose(const ParameterVector& x, ParameterVector& grad, const AllTrainingSets& ts, const AllTrainingSets& ets, const std::array<std::size_t, NumElements>& offset, const std::array<bool, NumElements>& use)
{
	std::array<double, NumElements> value{0.0, 0.0, 0.0};
	std::array<ParameterVector, NumElements> g;
	// multi-GPU: every element's loss on the rank that owns the element, values and gradients summed over the ranks
	const int ranks = Context::num_ranks(), me = ranks > 1 ? Context::rank() : 0;
	for_each_element_concurrently(
		[&](const std::size_t e)
		{
			if (!use[e] || std::get<0>(ts[e]).cols() == 0 || element_owner(e, ranks) != me)
			{
				return;
			}
			const std::size_t np = e == 1 ? 8 : 4;
			const ParameterVector xe(x.cbegin() + offset[e], x.cbegin() + offset[e] + np);
			g[e].assign(grad.empty() ? 0 : np, 0.0);
			ElementTrainingParameters etp = std::tie(ts[e], ets[e]);
			value[e] = loose_function(xe, g[e], static_cast<void*>(&etp));
		}
	);
	if (ranks > 1)
	{
		std::vector<double> buf(NumElements * 9, 0.0);
		for (std::size_t e = 0; e < NumElements; e++)
		{
			buf[e * 9] = value[e];
			std::copy(g[e].cbegin(), g[e].cend(), buf.begin() + e * 9 + 1);
		}
		all_reduce_sum(buf);
		for (std::size_t e = 0; e < NumElements; e++)
		{
			value[e] = buf[e * 9];
			if (use[e] && std::get<0>(ts[e]).cols() != 0)
			{
				g[e].assign(buf.cbegin() + e * 9 + 1, buf.cbegin() + e * 9 + 1 + (grad.empty() ? 0 : (e == 1 ? 8 : 4)));
			}
		}
	}
	double err = 0.0;
	for (std::size_t e = 0; e < NumElements; e++)
	{
		err += value[e];
		if (!grad.empty() && !g[e].empty())
		{
			std::copy(g[e].cbegin(), g[e].cend(), grad.begin() + offset[e]);
		}
	}
	make_normal(err);
	for (double& d : grad)
	{
		make_normal(d);
	}
	return err;
}
/// diagonal_loose (gple/opt.cpp:594-617): x = [theta_00, theta_11]
inline double diagonal_loose(const ParameterVector& x, ParameterVector& grad, void* params)
{
	const auto& [ts, ets] = *static_cast<AnalyticalLooseFunctionParameters*>(params);
	std::fill(grad.begin(), grad.end(), 0.0);
	return sum_loose(x, grad, ts, ets, {0, 0, 4}, {true, false, true});
}
/// full_loose (gple/opt.cpp:844-870): x = [theta_00, theta_10, theta_11]
inline double full_loose(const ParameterVector& x, ParameterVector& grad, void* params)
{
	const auto& [ts, ets] = *static_cast<AnalyticalLooseFunctionParameters*>(params);
	std::fill(grad.begin(), grad.end(), 0.0);
	return sum_loose(x, grad, ts, ets, {0, 4, 12}, {true, true, true});
}

/// diagonal_constraints (gple/opt.cpp:644-719): [population - 1, energy - E0, (purity - purity0)] and the m x 8 Jacobian
inline void diagonal_constraints(const unsigned m, double* result, const unsigned n, const double* x, double* grad, void* params)
{
	const auto& [ts, Energies, TotalEnergy, Purity] = *static_cast<AnalyticalConstraintParameters*>(params);
	assert(n == 8 && (m == 2 || m == 3));
	// construct_all_parameters_from_diagonal (opt.cpp:622-636): the off-diagonal element is absent (all-zero parameters)
	const AllParameters all{ParameterVector(x, x + 4), ParameterVector(8, 0.0), ParameterVector(x + 4, x + 8)};
	const TrainingKernels k(all, ts, true, true, grad != nullptr, &ModelCache::instance(), true);
	result[0] = k.calculate_population() - 1.0;
	result[1] = k.calculate_total_energy_average(Energies) - TotalEnergy;
	if (m == 3)
	{
		result[2] = k.calculate_purity() - Purity;
	}
	for (unsigned i = 0; i < m; i++)
	{
		make_normal(result[i]);
	}
	if (grad != nullptr)
	{
		const ParameterVector dp = k.population_derivative(), de = k.total_energy_derivative(Energies);
		std::copy(dp.cbegin(), dp.cend(), grad);
		std::copy(de.cbegin(), de.cend(), grad + n);
		if (m == 3)
		{
			const ParameterVector du = k.purity_derivative();
			std::copy(du.cbegin(), du.cbegin() + 4, grad + 2 * n);
			std::copy(du.cbegin() + 12, du.cend(), grad + 2 * n + 4);
		}
		for (unsigned i = 0; i < m * n; i++)
		{
			make_normal(grad[i]);
		}
	}
}
/// full_constraints (gple/opt.cpp:879-929): [population - 1, energy - E0, purity - purity0] and the 3 x 16 Jacobian
inline void full_constraints(const unsigned m, double* result, const unsigned n, const double* x, double* grad, void* params)
{
	const auto& [ts, Energies, TotalEnergy, Purity] = *static_cast<AnalyticalConstraintParameters*>(params);
	assert(n == 16 && m == 3);
	const AllParameters all{ParameterVector(x, x + 4), ParameterVector(x + 4, x + 12), ParameterVector(x + 12, x + 16)};
	const TrainingKernels k(all, ts, true, true, grad != nullptr, &ModelCache::instance(), true);
	result[0] = k.calculate_population() - 1.0;
	result[1] = k.calculate_total_energy_average(Energies) - TotalEnergy;
	result[2] = k.calculate_purity() - Purity;
	for (unsigned i = 0; i < m; i++)
	{
		make_normal(result[i]);
	}
	if (grad != nullptr)
	{
		std::fill(grad, grad + m * n, 0.0);
		const ParameterVector dp = k.population_derivative(), de = k.total_energy_derivative(Energies), du = k.purity_derivative();
		std::copy(dp.cbegin(), dp.cbegin() + 4, grad);
		std::copy(dp.cbegin() + 4, dp.cend(), grad + 12);
		std::copy(de.cbegin(), de.cbegin() + 4, grad + n);
		std::copy(de.cbegin() + 4, de.cend(), grad + n + 12);
		std::copy(du.cbegin(), du.cend(), grad + 2 * n);
		for (unsigned i = 0; i < m * n; i++)
		{
			make_normal(grad[i]);
		}
	}
}

/// Nine MC sums of one element's points (gple_observables)
inline std::array<double, 9> observable_sums(const ElementPoints& pts, const double mass, const int pes_model, const int pes_index)
{
	std::array<double, 9> o{};
	Context::check(gple_observables(Context::get(), pes_model, reinterpret_cast<const double*>(pts.data()), pts.size(), mass, pes_index, o.data()), "observables");
	return o;
}
/// gple/predict.cpp:109-126 (population standard deviation of the positions and momenta of one element's points)
inline ClassicalPhaseVector calculate_standard_deviation_one_surface(const ElementPoints& pts, const double mass, const int pes_model)
{
	const auto o = observable_sums(pts, mass, pes_model, 0);
	const double n = double(pts.size());
	return {std::sqrt(o[5] / n - (o[3] / n) * (o[3] / n)), std::sqrt(o[6] / n - (o[4] / n) * (o[4] / n))};
}
/// gple/predict.cpp:182-190
inline QuantumVectorD calculate_total_energy_average_each_surface(const AllPoints& density, const double mass, const int pes_model)
{
	QuantumVectorD r{0.0, 0.0};
	for (std::size_t s = 0; s < NumPES; s++)
	{
		if (!density[2 * s].empty())
		{
			const auto o = observable_sums(density[2 * s], mass, pes_model, int(s));
			r[s] = o[7] / o[0];
		}
	}
	return r;
}

/// gple/opt.h:17-105
class Optimization final
{
public:
	enum OptimizationType
	{
		Default,
		LocalPrevious,
		LocalInitial,
		Global
	};
	using Result = std::tuple<double, std::vector<std::size_t>, OptimizationType>;

	/// gple/opt.cpp:273-416.  SigmaR0 = InitParams.get_sigma_r0(), rSize = get_rmax() - get_rmin() (gple/input.h)
	Optimization(
		const ClassicalPhaseVector& SigmaR0,
		const ClassicalPhaseVector& rSize,
		const double Mass,
		const int PESModel,
		const double InitialTotalEnergy,
		const double InitialPurity,
		const nlopt::algorithm LocalDiagonalAlgorithm = nlopt::LN_NELDERMEAD,
		const nlopt::algorithm LocalOffDiagonalAlgorithm = nlopt::LN_NELDERMEAD,
		const nlopt::algorithm ConstraintAlgorithm = nlopt::LD_SLSQP,
		const nlopt::algorithm GlobalAlgorithm = nlopt::GN_DIRECT_L
	):
		TotalEnergy(InitialTotalEnergy), Purity(InitialPurity), mass(Mass), pes_model(PESModel),
		InitialKernelParameter{InitialMagnitude, SigmaR0[0], SigmaR0[1], InitialNoise},
		InitialComplexKernelParameter{InitialMagnitude, InitialMagnitude, SigmaR0[0], SigmaR0[1], InitialMagnitude, SigmaR0[0], SigmaR0[1], InitialNoise},
		LocalMinimizers{nlopt::opt(LocalDiagonalAlgorithm, 4), nlopt::opt(LocalOffDiagonalAlgorithm, 8), nlopt::opt(LocalDiagonalAlgorithm, 4)},
		DiagonalMinimizer(nlopt::AUGLAG_EQ, 8), FullMinimizer(nlopt::AUGLAG_EQ, 16),
		GlobalMinimizers{nlopt::opt(GlobalAlgorithm, 4), nlopt::opt(GlobalAlgorithm, 8), nlopt::opt(GlobalAlgorithm, 4)},
		ParameterVectors{InitialKernelParameter, InitialComplexKernelParameter, InitialKernelParameter}
	{
		auto set_optimizer = [](nlopt::opt& o) // opt.cpp:340-355
		{
			o.set_xtol_rel(1e-5);
			o.set_ftol_rel(1e-5);
			o.set_xtol_abs(1e-15);
			o.set_ftol_abs(1e-15);
			if (o.get_algorithm() == nlopt::LN_NELDERMEAD)
			{
				o.set_initial_step(0.5);
			}
		};
		for (std::size_t e = 0; e < NumElements; e++)
		{
			set_optimizer(LocalMinimizers[e]);
			set_optimizer(GlobalMinimizers[e]);
			GlobalMinimizers[e].set_maxeval(MaximumGlobalEvaluations);
		}
		for (nlopt::opt* o : {&DiagonalMinimizer, &FullMinimizer}) // opt.cpp:384-389
		{
			set_optimizer(*o);
			nlopt::opt sub(ConstraintAlgorithm, o->get_dimension());
			set_optimizer(sub);
			o->set_local_optimizer(sub);
			o->set_maxeval(MaximumConstrainedEvaluations);
		}
		// opt.cpp:391-414: minimal characteristic length 1/100, maximal = the size of the phase-space box
		const ClassicalPhaseVector lmin{0.01, 0.01};
		const Bounds kb = calculate_kernel_bounds(lmin, rSize), cb = calculate_complex_kernel_bounds(lmin, rSize);
		set_optimizer_bounds({kb, cb, kb});
	}

	/// Evaluation caps: the reference gives the global optimiser 100000 evaluations (opt.cpp:339) and AUGLAG none; both are
	/// settable here because a run-away optimisation is the only unbounded host loop of a time step
	void set_maximum_evaluations(const int global, const int constrained)
	{
		for (auto& g : GlobalMinimizers)
		{
			g.set_maxeval(global);
		}
		DiagonalMinimizer.set_maxeval(constrained);
		FullMinimizer.set_maxeval(constrained);
	}
	const AllParameters& get_parameters() const { return ParameterVectors; }
	AllParameters get_lower_bounds() const { return {LocalMinimizers[0].get_lower_bounds(), LocalMinimizers[1].get_lower_bounds(), LocalMinimizers[2].get_lower_bounds()}; }
	AllParameters get_upper_bounds() const { return {LocalMinimizers[0].get_upper_bounds(), LocalMinimizers[1].get_upper_bounds(), LocalMinimizers[2].get_upper_bounds()}; }

	/// Restart stages ahead of time (default on; single-GPU runs only).  optimize() tries up to three starts one after the other
	/// (opt.cpp:1320-1383): the previous parameters, the initial parameters, a global search -- each a full local sequence of
	/// thousands of millisecond-sized, latency-bound evaluations that leave most of the GPU idle, and the later ones only run when
	/// the earlier result misses the averages.  With this option stages 2 and 3 start at once on threads and contexts of their own
	/// (slot groups 1 and 2) next to stage 1; the acceptance rules are then applied in the reference's order to the finished
	/// results, and a stage whose result turns out not to be needed is stopped at its next iteration and discarded.  Evaluations are
	/// deterministic, so the outcome (parameters, error, evaluation counts, result type) is the sequential one.
	void set_speculative_restarts(const bool on) { SpeculativeRestarts = on; }

	/// gple/opt.cpp:1019-1392
	Result optimize(const AllPoints& density, const AllPoints& extra_points)
	{
		const AllTrainingSets TrainingSets = construct_training_sets(density), ExtraTrainingSets = construct_training_sets(extra_points);
		// the cache is keyed by the addresses of these training sets: nothing cached may outlive them
		ModelCache::instance().clear();
		struct CacheGuard
		{
			~CacheGuard() { ModelCache::instance().clear(); }
		} cache_guard;
		const QuantumVectorD Energies = calculate_total_energy_average_each_surface(density, mass, pes_model);
		// opt.cpp:1026-1052: bounds of the characteristic lengths from the spread of the current points
		std::array<Bounds, NumElements> ParameterBounds;
		for (std::size_t e = 0; e < NumElements; e++)
		{
			if (!density[e].empty())
			{
				const ClassicalPhaseVector sd = calculate_standard_deviation_one_surface(density[e], mass, pes_model);
				const double rn = std::sqrt(double(density[e].size()));
				const ClassicalPhaseVector lb{sd[0] / rn, sd[1] / rn}, ub{2.0 * sd[0], 2.0 * sd[1]};
				ParameterBounds[e] = e == 1 ? calculate_complex_kernel_bounds(lb, ub) : calculate_kernel_bounds(lb, ub);
			}
			else
			{
				ParameterBounds[e] = {LocalMinimizers[e].get_lower_bounds(), LocalMinimizers[e].get_upper_bounds()};
			}
		}
		set_optimizer_bounds(ParameterBounds);
		const StageInputs in{TrainingSets, ExtraTrainingSets, Energies, ParameterBounds, !density[1].empty()};
		auto any = [](const std::array<double, 3>& c) { return c[0] != 0.0 || c[1] != 0.0 || c[2] != 0.0; };
		// opt.cpp:1272-1318
		auto compare_and_overwrite = [this](Result& result, std::array<double, 3>& check, const Result& result_new, const std::array<double, 3>& check_new, const AllParameters& pv_new)
		{
			int better = 0, worse = 0;
			for (std::size_t i = 0; i < 3; i++)
			{
				better += (check_new[i] < check[i] && check[i] > 2.0 * AverageTolerance) ? 1 : 0;
				worse += (check_new[i] > check[i] && check_new[i] > 2.0 * AverageTolerance) ? 1 : 0;
			}
			const double sum_old = check[0] + check[1] + check[2], sum_new = check_new[0] + check_new[1] + check_new[2];
			if (better > worse || (better == worse && (sum_new < sum_old || std::get<0>(result_new) < std::get<0>(result))))
			{
				ParameterVectors = pv_new;
				std::get<0>(result) = std::get<0>(result_new);
				auto& steps = std::get<1>(result);
				const auto& more = std::get<1>(result_new);
				for (std::size_t i = 0; i < steps.size() && i < more.size(); i++)
				{
					steps[i] += more[i];
				}
				std::get<2>(result) = std::get<2>(result_new);
				check = check_new;
			}
		};
		// stages 2 and 3 ahead of time: a copy of this optimiser (same bounds, tolerances, targets) per stage, on its own slot group
		const bool speculate = SpeculativeRestarts && Context::num_ranks() == 1 && Context::group() == 0;
		struct Speculation
		{
			std::unique_ptr<Optimization> worker;
			std::atomic<bool> stop{false};
			std::thread th;
			StageOutcome out;
			std::exception_ptr error;
			void finish()
			{
				if (th.joinable())
				{
					th.join();
				}
			}
			void cancel()
			{
				stop.store(true);
				finish();
			}
			~Speculation() { cancel(); }
		};
		std::array<Speculation, 2> ahead; // destroyed (stopped and joined) before the cache guard and the training sets
		if (speculate)
		{
			for (int k = 0; k < 2; k++)
			{
				Speculation& sp = ahead[k];
				sp.worker.reset(new Optimization(*this));
				sp.worker->set_stop_flag(&sp.stop);
				sp.th = std::thread(
					[&sp, &in, k]()
					{
						const ContextGroup group(k + 1);
						try
						{
							sp.out = k == 0 ? sp.worker->stage_from_initial(in) : sp.worker->stage_from_global(in);
						}
						catch (...)
						{
							sp.error = std::current_exception();
						}
					}
				);
			}
		}
		auto stage = [&](const int k) -> StageOutcome
		{
			if (!speculate)
			{
				return k == 0 ? stage_from_initial(in) : stage_from_global(in);
			}
			ahead[k].finish();
			if (ahead[k].error)
			{
				std::rethrow_exception(ahead[k].error);
			}
			return ahead[k].out;
		};

		// 1. from the previous parameters (opt.cpp:1320-1333)
		Result result = do_optimize(in, ParameterVectors, LocalPrevious);
		std::array<double, 3> check = check_averages(in, ParameterVectors);
		if (!any(check))
		{
			return result;
		}
		// 2. from the initial parameters (opt.cpp:1335-1352)
		{
			const StageOutcome s2 = stage(0);
			compare_and_overwrite(result, check, s2.result, s2.check, s2.pv);
			if (!any(check))
			{
				return result;
			}
		}
		// 3. global search per element in log space, then the local sequence from there (opt.cpp:1354-1383)
		{
			const StageOutcome s3 = stage(1);
			compare_and_overwrite(result, check, s3.result, s3.check, s3.pv);
		}
		return result;
	}

private:
	/// what every restart stage of one optimize() call works from
	struct StageInputs
	{
		const AllTrainingSets& TrainingSets;
		const AllTrainingSets& ExtraTrainingSets;
		const QuantumVectorD& Energies;
		const std::array<Bounds, NumElements>& ParameterBounds;
		bool OffDiagonalPopulated;
	};
	struct StageOutcome
	{
		AllParameters pv;
		Result result;
		std::array<double, 3> check{0.0, 0.0, 0.0};
	};
	bool SpeculativeRestarts = true;
	void set_stop_flag(const std::atomic<bool>* flag)
	{
		for (std::size_t e = 0; e < NumElements; e++)
		{
			LocalMinimizers[e].set_stop_flag(flag);
			GlobalMinimizers[e].set_stop_flag(flag);
		}
		DiagonalMinimizer.set_stop_flag(flag);
		FullMinimizer.set_stop_flag(flag);
	}
	static void move_into_bounds(const StageInputs& in, AllParameters& pv)
	{
		for (std::size_t e = 0; e < NumElements; e++)
		{
			for (std::size_t p = 0; p < pv[e].size(); p++)
			{
				pv[e][p] = std::clamp(pv[e][p], in.ParameterBounds[e][0][p], in.ParameterBounds[e][1][p]);
			}
		}
	}
	/// opt.cpp:1101-1198: the local sequence (per element, diagonal with constraints, all elements with constraints) from pv
	Result do_optimize(const StageInputs& in, AllParameters& pv, const OptimizationType type)
	{
		for (auto& p : pv)
		{
			p[0] = InitialMagnitude;
		}
		move_into_bounds(in, pv);
		auto [err, steps] = optimize_elementwise(in.TrainingSets, in.ExtraTrainingSets, LocalMinimizers, pv, false);
		if (in.OffDiagonalPopulated)
		{
			const auto [derr, dsteps] = optimize_diagonal(in.TrainingSets, in.ExtraTrainingSets, in.Energies, pv, false);
			(void)derr;
			const auto [ferr, fsteps] = optimize_full(in.TrainingSets, in.ExtraTrainingSets, in.Energies, pv);
			err = ferr;
			steps.push_back(dsteps);
			steps.push_back(fsteps);
		}
		else
		{
			const auto [derr, dsteps] = optimize_diagonal(in.TrainingSets, in.ExtraTrainingSets, in.Energies, pv, true);
			err = derr;
			steps.push_back(dsteps);
			steps.push_back(0);
		}
		// opt.cpp:1179-1195: the magnitude follows from the optimised kernel
		const TrainingKernels k(pv, in.TrainingSets, false, false, false);
		for (std::size_t i = 0; i < NumPES; i++)
		{
			if (k.Diagonal[i])
			{
				pv[2 * i][0] = k.Diagonal[i]->get_magnitude();
			}
		}
		if (k.OffDiagonal)
		{
			pv[1][0] = k.OffDiagonal->get_magnitude();
		}
		return {err, steps, type};
	}
	/// opt.cpp:1200-1270: relative deviations of the averages, 0 when within tolerance
	std::array<double, 3> check_averages(const StageInputs& in, const AllParameters& pv) const
	{
		const TrainingKernels k(pv, in.TrainingSets, false, true, false);
		auto beyond = [](const double calc, const double ref)
		{
			const double err = std::abs(calc / ref - 1.0);
			return err < AverageTolerance ? 0.0 : err;
		};
		return {beyond(k.calculate_population(), 1.0), beyond(k.calculate_total_energy_average(in.Energies), TotalEnergy), beyond(k.calculate_purity(), Purity)};
	}
	/// stage 2 (opt.cpp:1335-1352): the local sequence from the initial parameters
	StageOutcome stage_from_initial(const StageInputs& in)
	{
		StageOutcome o;
		o.pv = AllParameters{InitialKernelParameter, InitialComplexKernelParameter, InitialKernelParameter};
		o.result = do_optimize(in, o.pv, LocalInitial);
		o.check = check_averages(in, o.pv);
		return o;
	}
	/// stage 3 (opt.cpp:1354-1383): global search per element in log space, then the local sequence from there
	StageOutcome stage_from_global(const StageInputs& in)
	{
		StageOutcome o;
		o.pv = AllParameters{InitialKernelParameter, InitialComplexKernelParameter, InitialKernelParameter};
		move_into_bounds(in, o.pv);
		const auto [gerr, gsteps] = optimize_elementwise(in.TrainingSets, in.ExtraTrainingSets, GlobalMinimizers, o.pv, true);
		(void)gerr;
		o.result = do_optimize(in, o.pv, Global);
		auto& steps = std::get<1>(o.result);
		for (std::size_t i = 0; i < gsteps.size() && i < steps.size(); i++)
		{
			steps[i] += gsteps[i];
		}
		o.check = check_averages(in, o.pv);
		return o;
	}
	static constexpr int MaximumGlobalEvaluations = 100000;		// opt.cpp:339
	static constexpr int MaximumConstrainedEvaluations = 2000;	// none in the reference (see set_maximum_evaluations)
	const double TotalEnergy;
	const double Purity;
	const double mass;
	const int pes_model;
	const ParameterVector InitialKernelParameter;
	const ParameterVector InitialComplexKernelParameter;
	std::array<nlopt::opt, NumElements> LocalMinimizers;
	nlopt::opt DiagonalMinimizer;
	nlopt::opt FullMinimizer;
	std::array<nlopt::opt, NumElements> GlobalMinimizers;
	AllParameters ParameterVectors;

	/// set_optimizer_bounds (gple/opt.cpp:240-271); the global optimisers work in log space on some parameters
	void set_optimizer_bounds(const std::array<Bounds, NumElements>& b)
	{
		ParameterVector dl, du, fl, fu;
		for (std::size_t e = 0; e < NumElements; e++)
		{
			LocalMinimizers[e].set_lower_bounds(b[e][0]);
			LocalMinimizers[e].set_upper_bounds(b[e][1]);
			GlobalMinimizers[e].set_lower_bounds(local_parameter_to_global(b[e][0]));
			GlobalMinimizers[e].set_upper_bounds(local_parameter_to_global(b[e][1]));
			fl.insert(fl.end(), b[e][0].cbegin(), b[e][0].cend());
			fu.insert(fu.end(), b[e][1].cbegin(), b[e][1].cend());
			if (e != 1)
			{
				dl.insert(dl.end(), b[e][0].cbegin(), b[e][0].cend());
				du.insert(du.end(), b[e][1].cbegin(), b[e][1].cend());
			}
		}
		DiagonalMinimizer.set_lower_bounds(dl);
		DiagonalMinimizer.set_upper_bounds(du);
		FullMinimizer.set_lower_bounds(fl);
		FullMinimizer.set_upper_bounds(fu);
	}

	/// optimize_elementwise (gple/opt.cpp:517-587): every populated element on its own, the three optimisations concurrently
	static std::tuple<double, std::vector<std::size_t>> optimize_elementwise(const AllTrainingSets& ts, const AllTrainingSets& ets, std::array<nlopt::opt, NumElements>& minimizers, AllParameters& pv, const bool is_global)
	{
		std::array<double, NumElements> err{0.0, 0.0, 0.0};
		std::vector<std::size_t> steps(NumElements, 0);
		// multi-GPU: an element's optimisation runs on the rank that owns the element; parameters, error and evaluation count
		// are then summed over the ranks (zeros elsewhere), so every rank continues from identical parameters
		const int ranks = Context::num_ranks(), me = ranks > 1 ? Context::rank() : 0;
		for_each_element_concurrently(
			[&](const std::size_t e)
			{
				if (std::get<0>(ts[e]).cols() == 0 || element_owner(e, ranks) != me)
				{
					return;
				}
				ElementTrainingParameters etp = std::tie(ts[e], ets[e]);
				nlopt::opt& o = minimizers[e];
				ParameterVector x = pv[e];
				if (is_global)
				{
					o.set_min_objective(loose_function_global_wrapper, static_cast<void*>(&etp));
					x = local_parameter_to_global(x);
				}
				else
				{
					o.set_min_objective(loose_function, static_cast<void*>(&etp));
				}
				o.optimize(x, err[e]);
				pv[e] = is_global ? global_parameter_to_local(x) : x;
				steps[e] = std::size_t(o.get_numevals());
			}
		);
		if (ranks > 1)
		{
			std::vector<double> buf(NumElements * 10, 0.0);
			for (std::size_t e = 0; e < NumElements; e++)
			{
				if (std::get<0>(ts[e]).cols() != 0 && element_owner(e, ranks) == me)
				{
					std::copy(pv[e].cbegin(), pv[e].cend(), buf.begin() + e * 10);
					buf[e * 10 + 8] = err[e];
					buf[e * 10 + 9] = double(steps[e]);
				}
			}
			all_reduce_sum(buf);
			for (std::size_t e = 0; e < NumElements; e++)
			{
				if (std::get<0>(ts[e]).cols() != 0)
				{
					std::copy(buf.cbegin() + e * 10, buf.cbegin() + e * 10 + pv[e].size(), pv[e].begin());
					err[e] = buf[e * 10 + 8];
					steps[e] = std::size_t(buf[e * 10 + 9]);
				}
			}
		}
		return {err[0] + err[1] + err[2], steps};
	}
	/// optimize_diagonal (gple/opt.cpp:730-800)
	std::tuple<double, std::size_t> optimize_diagonal(const AllTrainingSets& ts, const AllTrainingSets& ets, const QuantumVectorD& Energies, AllParameters& pv, const bool with_purity)
	{
		ParameterVector x = pv[0];
		x.insert(x.end(), pv[2].cbegin(), pv[2].cend());
		AnalyticalLooseFunctionParameters alfp = std::tie(ts, ets);
		DiagonalMinimizer.set_min_objective(diagonal_loose, static_cast<void*>(&alfp));
		AnalyticalConstraintParameters acp = std::tie(ts, Energies, TotalEnergy, Purity);
		DiagonalMinimizer.add_equality_mconstraint(diagonal_constraints, static_cast<void*>(&acp), ParameterVector(with_purity ? 3 : 2, 0.0));
		double err = 0.0;
		DiagonalMinimizer.optimize(x, err);
		DiagonalMinimizer.remove_equality_constraints();
		pv[0].assign(x.cbegin(), x.cbegin() + 4);
		pv[2].assign(x.cbegin() + 4, x.cend());
		return {err, std::size_t(DiagonalMinimizer.get_numevals())};
	}
	/// optimize_full (gple/opt.cpp:940-1015)
	std::tuple<double, std::size_t> optimize_full(const AllTrainingSets& ts, const AllTrainingSets& ets, const QuantumVectorD& Energies, AllParameters& pv)
	{
		ParameterVector x = pv[0];
		x.insert(x.end(), pv[1].cbegin(), pv[1].cend());
		x.insert(x.end(), pv[2].cbegin(), pv[2].cend());
		AnalyticalLooseFunctionParameters alfp = std::tie(ts, ets);
		FullMinimizer.set_min_objective(full_loose, static_cast<void*>(&alfp));
		AnalyticalConstraintParameters acp = std::tie(ts, Energies, TotalEnergy, Purity);
		FullMinimizer.add_equality_mconstraint(full_constraints, static_cast<void*>(&acp), ParameterVector(3, 0.0));
		double err = 0.0;
		FullMinimizer.optimize(x, err);
		FullMinimizer.remove_equality_constraints();
		pv[0].assign(x.cbegin(), x.cbegin() + 4);
		pv[1].assign(x.cbegin() + 4, x.cbegin() + 12);
		pv[2].assign(x.cbegin() + 12, x.cend());
		return {err, std::size_t(FullMinimizer.get_numevals())};
	}
};
} // namespace gple_host
