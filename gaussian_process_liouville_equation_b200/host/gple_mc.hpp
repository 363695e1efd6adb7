// gple_mc.hpp -- C++ host mirror of the reference's Metropolis sampling (gple/mc.h, gple/mc.cpp:125-403) over the C-ABI.
//
// Same names and roles as the reference (MCParameters, monte_carlo_selection, element_monte_carlo and the two tuning
// passes); the DistributionFunction argument becomes a Sampler that names the target density (analytic initial
// distribution, the element models' prediction, or new_point_predict) so that all chains of an element can advance in
// lock-step on the GPU (gple_markov_chains).  The reference's clock-seeded shared mt19937 (mc.cpp:17) is replaced by one
// Philox stream per chain; a Sampler folds a call counter into the stream id, so successive walks are independent and the
// whole sequence is a pure function of the seed.  Header-only; link with libgple_b200.so.
#pragma once
#include "gple_opt.hpp"

#include <cmath>
#include <random>

namespace gple_host
{
constexpr double MaxAcceptRatio = 0.5, MinAcceptRatio = 0.15; // gple/mc.cpp:19-20

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
	/// generate_markov_chain (gple/mc.cpp:143-188) for every point: points end at the last state of their chain, relabelled
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
	/// distribution(r, RowIndex, ColIndex) at the coordinates of `pts` (their densities are overwritten); no walk, no draw
	void relabel(ElementPoints& pts, const std::size_t RowIndex, const std::size_t ColIndex) const
	{
		if (pts.empty())
		{
			return;
		}
		gple_mc_source s = src;
		s.row = int(RowIndex);
		s.col = int(ColIndex);
		Context::check(gple_markov_chains(Context::get(), &s, reinterpret_cast<double*>(pts.data()), pts.size(), 0, 1.0, seed, 0, 0, nullptr, nullptr), "relabel");
	}
	/// gple/mc.cpp:230-243
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

/// gple/mc.cpp:288-337: the largest displacement of the list whose mean acceptance ratio lies inside (0.15, 0.5)
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

/// gple/mc.cpp:197-279: chain length = first lag whose |autocorrelation| is within 1.1x of the minimum
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

/// gple/mc.cpp:339-378
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
/// generate_extra_points (gple/mc.cpp:59-120): training point (cyclic) + N(0, sigma_element) jitter per dimension, labelled
/// with the current distribution -- one batched evaluation per element instead of NumExtraPoints single-point calls.
/// Host-side randomness (the jitter) comes from the caller's engine, as the reference draws it from its global one.
template <typename Engine>
inline AllPoints generate_extra_points(const AllPoints& density, const std::size_t NumExtraPoints, const Sampler& sampler, Engine& engine, const double mass, const int pes_model)
{
	static constexpr std::size_t Row[NumElements] = {0, 1, 1}, Col[NumElements] = {0, 0, 1};
	AllPoints result;
	for (std::size_t e = 0; e < NumElements; e++)
	{
		if (density[e].empty())
		{
			continue;
		}
		const ClassicalPhaseVector sd = calculate_standard_deviation_one_surface(density[e], mass, pes_model);
		std::array<std::normal_distribution<double>, PhaseDim> normdists{std::normal_distribution<double>(0.0, sd[0]), std::normal_distribution<double>(0.0, sd[1])};
		result[e].resize(NumExtraPoints);
		for (std::size_t i = 0; i < NumExtraPoints; i++)
		{
			result[e][i].r = density[e][i % density[e].size()].r;
			for (std::size_t d = 0; d < PhaseDim; d++)
			{
				result[e][i].r[d] += normdists[d](engine);
			}
		}
		sampler.relabel(result[e], Row[e], Col[e]);
	}
	return result;
}

/// is_very_small (gple/evolve.cpp:445-478): an EMPTY element is small iff new_point_predict at all test points (the rho00
/// points) has |rho|^2 below (1e-5)^2; a populated element is never small.
inline std::array<bool, NumElements> is_very_small(const AllPoints& density, const double mass, const double dt, const TrainingKernels& kernels, const int pes_model)
{
	static constexpr int Row[NumElements] = {0, 1, 1}, Col[NumElements] = {0, 0, 1};
	static constexpr double epsilon = 1e-5 * 1e-5;
	std::array<bool, NumElements> result{false, false, false};
	for (std::size_t e = 0; e < NumElements; e++)
	{
		if (!density[e].empty() || density[0].empty())
		{
			continue;
		}
		const std::size_t n = density[0].size();
		std::vector<double> r(2 * n), rho(2 * n);
		for (std::size_t i = 0; i < n; i++)
		{
			r[2 * i] = density[0][i].r[0];
			r[2 * i + 1] = density[0][i].r[1];
		}
		Context::check(gple_new_point_predict(Context::get(), pes_model, kernels.handle(0), kernels.handle(1), kernels.handle(2), r.data(), n, Row[e], Col[e], mass, dt, rho.data()), "is_very_small");
		bool small = true;
		for (std::size_t i = 0; i < n && small; i++)
		{
			small = rho[2 * i] * rho[2 * i] + rho[2 * i + 1] * rho[2 * i + 1] < epsilon;
		}
		result[e] = small;
	}
	return result;
}

/// new_element_point_selection (gple/mc.cpp:407-537): a newly populated element takes the N most important of all current
/// coordinates (|rho|^2 of `sampler`'s density: new_point_predict in main.cpp:147-157), replicated up to N, walks them and gets
/// extra points; a newly small element is emptied.
template <typename Engine>
inline void new_element_point_selection(AllPoints& density, AllPoints& extra_points, const std::array<bool, NumElements>& IsSmallOld, const std::array<bool, NumElements>& IsSmall, std::array<MCParameters, NumElements>& MCParams, Sampler& sampler, Engine& engine, const double mass, const int pes_model)
{
	static constexpr std::size_t Row[NumElements] = {0, 1, 1}, Col[NumElements] = {0, 0, 1};
	if (IsSmallOld == IsSmall)
	{
		return;
	}
	const std::size_t NumPoints = density[0].size(), NumExtraPoints = extra_points[0].size();
	ElementPoints PossibleCoordinates;
	for (std::size_t e = 0; e < NumElements; e++)
	{
		PossibleCoordinates.insert(PossibleCoordinates.end(), density[e].cbegin(), density[e].cend());
		PossibleCoordinates.insert(PossibleCoordinates.end(), extra_points[e].cbegin(), extra_points[e].cend());
	}
	for (std::size_t e = 0; e < NumElements; e++)
	{
		if (IsSmallOld[e] && !IsSmall[e])
		{
			ElementPoints element_density = PossibleCoordinates;
			sampler.relabel(element_density, Row[e], Col[e]);
			const std::size_t NumNonZeroPoints = std::size_t(std::count_if(element_density.cbegin(), element_density.cend(), [](const PhaseSpacePoint& psp) { return psp.rho != 0.0; }));
			const std::size_t keep = std::min(NumPoints, NumNonZeroPoints);
			if (keep == 0)
			{
				continue;
			}
			std::nth_element(element_density.begin(), element_density.begin() + keep, element_density.end(), [](const PhaseSpacePoint& a, const PhaseSpacePoint& b) { return std::norm(a.rho) > std::norm(b.rho); });
			element_density.resize(keep);
			while (NumPoints >= 2 * element_density.size())
			{
				const ElementPoints copy = element_density;
				element_density.insert(element_density.end(), copy.cbegin(), copy.cend());
			}
			if (element_density.size() < NumPoints)
			{
				const ElementPoints copy(element_density.cbegin(), element_density.cbegin() + (NumPoints - element_density.size()));
				element_density.insert(element_density.end(), copy.cbegin(), copy.cend());
			}
			element_monte_carlo(element_density, MCParams[e], sampler, Row[e], Col[e]);
			density[e] = std::move(element_density);
			AllPoints only;
			only[e] = density[e];
			extra_points[e] = generate_extra_points(only, NumExtraPoints, sampler, engine, mass, pes_model)[e];
		}
		else if (!IsSmallOld[e] && IsSmall[e])
		{
			density[e].clear();
			extra_points[e].clear();
		}
	}
}
} // namespace gple_host
