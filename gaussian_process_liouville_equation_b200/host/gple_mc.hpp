// gple_mc.hpp -- C++ host mirror of the reference's Metropolis sampling (gple/mc.h, gple/mc.cpp:125-403) over the C-ABI.
//
// Same names and roles as the reference (MCParameters, monte_carlo_selection, element_monte_carlo and the two tuning
// passes); the DistributionFunction argument becomes a Sampler that names the target density (analytic initial
// distribution, the element models' prediction, or new_point_predict) so that all chains of an element can advance in
// lock-step on the GPU (gple_markov_chains).  The reference's clock-seeded shared mt19937 (mc.cpp:17) is replaced by one
// Philox stream per chain; a Sampler folds a call counter into the stream id, so successive walks are independent and the
// whole sequence is a pure function of the seed.  Header-only; link with libgple_b200.so.
#pragma once
#include "gple_host.hpp"

#include <cmath>

namespace gple_host
{
constexpr double MaxAcceptRatio = 0.5, MinAcceptRatio = 0.15; // gple/mc.cpp:18-19

/// gple/mc.h:45-121
class MCParameters final
{
	std::size_t NOMC;
	double displacement;

public:
	static constexpr double AboveMinFactor = 1.1;
	MCParameters(const std::size_t InitialSteps = 200, const double InitialDisplacement = 1.0): NOMC(InitialSteps), displacement(InitialDisplacement) {}
	void set_num_MC_steps(const std::size_t NOMC_) { NOMC = NOMC_; }
	void set_displacement(const double displacement_) { displacement = displacement_; }
	std::size_t get_num_MC_steps() const { return NOMC; }
	double get_max_displacement() const { return displacement; }
};

/// The target density of a walk + the Philox seed
class Sampler
{
public:
	/// initial_distribution (gple/mc.cpp:30-50, bound in main.cpp:40-56)
	Sampler(const unsigned long long Seed, const ClassicalPhaseVector& r0, const ClassicalPhaseVector& SigmaR0, const std::array<double, NumPES>& InitialPopulation = {1.0, 0.0}, const std::array<double, NumPES>& InitialPhaseFactor = {0.0, 0.0}):
		seed(Seed)
	{
		src.kind = GPLE_MC_ANALYTIC;
		const double a[8] = {r0[0], r0[1], SigmaR0[0], SigmaR0[1], InitialPopulation[0], InitialPopulation[1], InitialPhaseFactor[0], InitialPhaseFactor[1]};
		std::copy(a, a + 8, src.analytic);
	}
	/// predict_distribution (gple/main.cpp:75-101)
	Sampler(const unsigned long long Seed, const TrainingKernels& kernels): seed(Seed)
	{
		src.kind = GPLE_MC_PREDICT;
		src.m00 = kernels.handle(0);
		src.m10 = kernels.handle(1);
		src.m11 = kernels.handle(2);
	}
	/// new_point_predict (gple/evolve.cpp:425-443, bound in main.cpp:153-156)
	Sampler(const unsigned long long Seed, const TrainingKernels& kernels, const int pes_model, const double mass, const double dt): Sampler(Seed, kernels)
	{
		src.kind = GPLE_MC_NEW_POINT;
		src.pes_model = pes_model;
		src.mass = mass;
		src.dt = dt;
	}
	/// generate_markov_chain (gple/mc.cpp:125-160) for every point: points end at the last state of their chain, relabelled
	/// with the density there; returns the acceptance ratios; `chains` (optional) receives every state, n x (steps + 1) x 2
	std::vector<double> chains(ElementPoints& pts, const std::size_t NumSteps, const double MaxDisplacement, const std::size_t RowIndex, const std::size_t ColIndex, std::vector<double>* chains_out = nullptr)
	{
		calls++;
		gple_mc_source s = src;
		s.row = int(RowIndex);
		s.col = int(ColIndex);
		std::vector<double> accept(pts.size());
		if (chains_out != nullptr)
		{
			chains_out->assign(2 * pts.size() * (NumSteps + 1), 0.0);
		}
		Context::check(
			gple_markov_chains(Context::get(), &s, reinterpret_cast<double*>(pts.data()), pts.size(), NumSteps, MaxDisplacement, seed, calls * 4 + RowIndex + ColIndex, 0, accept.data(), chains_out != nullptr ? chains_out->data() : nullptr),
			"markov chains"
		);
		return accept;
	}
	/// gple/mc.cpp:187-203
	std::vector<double> autocorrelation(const std::vector<double>& chain_states, const std::size_t NumChains, const std::size_t Length) const
	{
		std::vector<double> out(Length / 2);
		Context::check(gple_chain_autocorrelation(Context::get(), chain_states.data(), NumChains, Length, out.data()), "autocorrelation");
		return out;
	}

private:
	gple_mc_source src{};
	unsigned long long seed;
	unsigned long long calls = 0;
};

/// gple/mc.cpp:286-335: the largest displacement of the list whose mean acceptance ratio lies inside (0.15, 0.5)
inline void acceptance_optimize_displacement(MCParameters& MCParams, Sampler& sampler, const ElementPoints& density, const std::size_t RowIndex, const std::size_t ColIndex)
{
	static constexpr std::size_t MaxNOMC = PhaseDim * 500;
	static constexpr std::array PossibleDisplacement{1e-4, 2e-4, 5e-4, 1e-3, 2e-3, 5e-3, 0.01, 0.02, 0.05, 0.1, 0.2, 0.5, 1.0, 2.0, 5.0, 10.0};
	for (std::size_t i = PossibleDisplacement.size(); i-- > 0;)
	{
		ElementPoints walk = density;
		const std::vector<double> acc = sampler.chains(walk, MaxNOMC, PossibleDisplacement[i], RowIndex, ColIndex);
		double ratio = 0.0;
		for (const double a : acc) // numpy's pairwise mean is not reproduced bit for bit; the window test does not need it
		{
			ratio += a;
		}
		ratio /= double(acc.size());
		if (ratio < MaxAcceptRatio && ratio > MinAcceptRatio)
		{
			MCParams.set_displacement(PossibleDisplacement[i]);
			return;
		}
	}
}

/// gple/mc.cpp:162-260: chain length = first lag whose |autocorrelation| is within 1.1x of the minimum
inline void autocorrelation_optimize_steps(MCParameters& MCParams, Sampler& sampler, const ElementPoints& density, const std::size_t RowIndex, const std::size_t ColIndex)
{
	static constexpr std::size_t MaxNOMC = PhaseDim * 1000;
	ElementPoints walk = density;
	std::vector<double> states;
	sampler.chains(walk, MaxNOMC, MCParams.get_max_displacement(), RowIndex, ColIndex, &states);
	const std::vector<double> AutoCors = sampler.autocorrelation(states, density.size(), MaxNOMC + 1);
	const std::size_t len = AutoCors.size();
	std::size_t min_start_step = 0, min_autocor_step = 0;
	double min_auto_cor = 0.0;
	auto argmin_abs = [&AutoCors](const std::size_t from)
	{
		std::size_t best = from;
		for (std::size_t i = from; i < AutoCors.size(); i++)
		{
			if (std::abs(AutoCors[i]) < std::abs(AutoCors[best]))
			{
				best = i;
			}
		}
		return best;
	};
	while (true)
	{
		min_start_step = min_autocor_step + 1;
		if (min_start_step >= len)
		{
			min_start_step = 1;
			min_autocor_step = argmin_abs(0);
			min_auto_cor = std::abs(AutoCors[min_autocor_step]);
			break;
		}
		min_autocor_step = argmin_abs(min_start_step);
		min_auto_cor = std::abs(AutoCors[min_autocor_step]);
		ElementPoints one(density.cbegin(), density.cbegin() + 1);
		const double acc = sampler.chains(one, min_autocor_step, MCParams.get_max_displacement(), RowIndex, ColIndex)[0];
		if (acc <= MaxAcceptRatio && acc >= MinAcceptRatio)
		{
			break;
		}
	}
	for (std::size_t i = min_start_step; i < min_autocor_step; i++)
	{
		if (std::abs(AutoCors[i]) <= MCParameters::AboveMinFactor * min_auto_cor)
		{
			min_autocor_step = i;
			break;
		}
	}
	MCParams.set_num_MC_steps(min_autocor_step);
}

/// gple/mc.cpp:337-378
inline void element_monte_carlo(ElementPoints& density, MCParameters& MCParams, Sampler& sampler, const std::size_t RowIndex, const std::size_t ColIndex)
{
	acceptance_optimize_displacement(MCParams, sampler, density, RowIndex, ColIndex);
	autocorrelation_optimize_steps(MCParams, sampler, density, RowIndex, ColIndex);
	sampler.chains(density, MCParams.get_num_MC_steps(), MCParams.get_max_displacement(), RowIndex, ColIndex);
}

/// gple/mc.cpp:380-403
inline void monte_carlo_selection(AllPoints& density, std::array<MCParameters, NumElements>& MCParams, Sampler& sampler)
{
	static constexpr std::size_t Row[NumElements] = {0, 1, 1}, Col[NumElements] = {0, 0, 1};
	for (std::size_t e = 0; e < NumElements; e++)
	{
		if (!density[e].empty())
		{
			element_monte_carlo(density[e], MCParams[e], sampler, Row[e], Col[e]);
		}
	}
}
} // namespace gple_host
