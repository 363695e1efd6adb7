// The reference's main loop (gple/main.cpp:19-202) written against the C++ host headers of this repo: initial Metropolis
// selection from the analytic Wigner distribution, hyper-parameter optimisation, then per tick
//   evolve(density), evolve(extra_points) -> is_very_small -> [new element selection + optimisation] ->
//   [routine / average-triggered re-optimisation] -> TrainingKernels rebuild
// with every numerical step on the GPU through the C-ABI.  Output: one "tick population energy purity" line per tick
// (the quantities of output_average / spdlog in main.cpp:117-126) and a summary.
//
//   mqcle_run N ticks reopt_freq pes_model(0 SAC, 1 DAC, 2 ECR) [seed] [x0]
//
// Multi-GPU (one process per GPU): start G copies with GPLE_RANK = 0 .. G-1, GPLE_WORLD_SIZE = G,
// GPLE_COMM_FILE = a path all of them see (rank 0 publishes the NCCL id there) and GPLE_LOCAL_DEVICE = the GPU of the copy
// (torchrun's RANK / WORLD_SIZE / LOCAL_RANK are understood as well).  Every rank runs the same loop on the same seeds;
// evolve() moves only the rank's block of each point set and all-gathers the evolved sets through the library
// (gple_evolve_sharded), so every rank rebuilds its models from the full sets exactly like main.cpp:140-141, 176.  The
// "hash" lines (FNV-1a over the bytes of all points after the tick) are identical for every G.
#include "../gaussian_process_liouville_equation_b200/host/gple_mc.hpp"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

using namespace gple_host;

int main(int argc, char** argv)
{
	if (argc < 5)
	{
		std::fprintf(stderr, "usage: mqcle_run N ticks reopt_freq pes_model [seed] [x0]\n");
		return 2;
	}
	const std::size_t NumPoints = std::strtoull(argv[1], nullptr, 10), TotalTicks = std::strtoull(argv[2], nullptr, 10), ReoptFreq = std::strtoull(argv[3], nullptr, 10);
	const int pes_model = std::atoi(argv[4]);
	const unsigned long long seed = argc > 5 ? std::strtoull(argv[5], nullptr, 10) : 20240229ull;
	const std::size_t NumExtraPoints = NumPoints * 5; // main.cpp:35
	// test/continue_test.cpp:41, test/stdafx.h:51
	const double mass = 2000.0, dt = 1.0, sp = 0.7056, sx = 1.0 / (2.0 * sp);
	const ClassicalPhaseVector r0{argc > 6 ? std::atof(argv[6]) : -10.0, 14.112}, SigmaR0{sx, sp};
	std::mt19937_64 engine(seed);
	auto env = [](const char* a, const char* b) -> const char*
	{
		const char* v = std::getenv(a);
		return v != nullptr ? v : std::getenv(b);
	};
	const int rank = env("GPLE_RANK", "RANK") != nullptr ? std::atoi(env("GPLE_RANK", "RANK")) : 0;
	const int world = env("GPLE_WORLD_SIZE", "WORLD_SIZE") != nullptr ? std::atoi(env("GPLE_WORLD_SIZE", "WORLD_SIZE")) : 1;
	const int local = env("GPLE_LOCAL_DEVICE", "LOCAL_RANK") != nullptr ? std::atoi(env("GPLE_LOCAL_DEVICE", "LOCAL_RANK")) : 0;
	Context::device() = local;
	if (world > 1)
	{
		const char* file = std::getenv("GPLE_COMM_FILE");
		if (file == nullptr)
		{
			std::fprintf(stderr, "mqcle_run: GPLE_COMM_FILE is not set\n");
			return 2;
		}
		Context::init_distributed(rank, world, file, local);
	}
	if (rank != 0)
	{
		if (std::freopen("/dev/null", "w", stdout) == nullptr) // rank 0 reports; the others compute the same numbers
		{
			return 2;
		}
	}
	const auto begin = std::chrono::steady_clock::now();

	// main.cpp:36-57: N copies of r0, then the Metropolis walk in the initial distribution
	std::array<MCParameters, NumElements> MCParams;
	Sampler initdist(seed, r0, SigmaR0, {1.0, 0.0}, {0.0, 0.0});
	AllPoints density;
	density[0].assign(NumPoints, PhaseSpacePoint{r0, {0.0, 0.0}});
	std::array<bool, NumElements> IsSmall{false, true, true};
	monte_carlo_selection(density, MCParams, initdist);
	// main.cpp:58-69
	const double TotalEnergy = calculate_total_energy_average_each_surface(density, mass, pes_model)[0];
	const double Purity = 1.0;
	AllPoints extra_points = generate_extra_points(density, NumExtraPoints, initdist, engine, mass, pes_model);
	// main.cpp:70-74
	Optimization optimizer(SigmaR0, {20.0, 40.0}, mass, pes_model, TotalEnergy, Purity);
	optimizer.set_maximum_evaluations(300, 600);
	Optimization::Result opt_result = optimizer.optimize(density, extra_points);
	std::unique_ptr<TrainingKernels> all_kernels = std::make_unique<TrainingKernels>(optimizer.get_parameters(), density);
	std::size_t optimisations = 1;
	double t_opt = 0.0, t_evolve = 0.0, t_rebuild = 0.0, t_select = 0.0, t_evolve_late = 0.0;
	std::size_t late_ticks = 0;
	auto now = []() { return std::chrono::steady_clock::now(); };
	auto since = [](const std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
	auto output = [&](const std::size_t tick)
	{
		const QuantumVectorD E = calculate_total_energy_average_each_surface(density, mass, pes_model);
		std::printf("tick %zu %.15e %.15e %.15e\n", tick, all_kernels->calculate_population(), all_kernels->calculate_total_energy_average(E), all_kernels->calculate_purity());
	};
	auto hash_points = [&](const std::size_t tick)
	{
		std::uint64_t h = 1469598103934665603ull;
		for (const AllPoints* set : {&density, &extra_points})
		{
			for (const ElementPoints& pts : *set)
			{
				const unsigned char* b = reinterpret_cast<const unsigned char*>(pts.data());
				for (std::size_t i = 0; i < pts.size() * sizeof(PhaseSpacePoint); i++)
				{
					h = (h ^ b[i]) * 1099511628211ull;
				}
			}
		}
		std::printf("hash%zu %u %u\n", tick, unsigned(h & 0xffffffffu), unsigned(h >> 32));
	};
	output(0);
	std::printf("displacement %.17g\nmc_steps %zu\n", MCParams[0].get_max_displacement(), MCParams[0].get_num_MC_steps());

	// main.cpp:135-202
	for (std::size_t iTick = 1; iTick <= TotalTicks; iTick++)
	{
		const std::array<bool, NumElements> IsSmallOld = IsSmall;
		auto t0 = now();
		evolve(density, mass, dt, *all_kernels, pes_model);
		evolve(extra_points, mass, dt, *all_kernels, pes_model);
		IsSmall = is_very_small(density, mass, dt, *all_kernels, pes_model);
		t_evolve += since(t0);
		if (2 * iTick > TotalTicks) // second half of the run: kernels loaded, workspaces grown, all elements settled
		{
			t_evolve_late += since(t0);
			late_ticks++;
		}
		bool IsOptimized = false;
		auto reoptimize = [&]()
		{
			const auto t1 = now();
			opt_result = optimizer.optimize(density, extra_points);
			t_opt += since(t1);
			all_kernels = std::make_unique<TrainingKernels>(optimizer.get_parameters(), density);
			Sampler predict(seed + iTick, *all_kernels);
			extra_points = generate_extra_points(density, NumExtraPoints, predict, engine, mass, pes_model);
			IsOptimized = true;
			optimisations++;
		};
		if (IsSmallOld != IsSmall)
		{
			t0 = now();
			Sampler new_point(seed + iTick, *all_kernels, pes_model, mass, dt); // main.cpp:153-156
			new_element_point_selection(density, extra_points, IsSmallOld, IsSmall, MCParams, new_point, engine, mass, pes_model);
			t_select += since(t0);
			reoptimize();
		}
		if (ReoptFreq > 0 && iTick % ReoptFreq == 0 && !IsOptimized)
		{
			reoptimize();
		}
		if (!IsOptimized)
		{
			t0 = now();
			all_kernels = std::make_unique<TrainingKernels>(optimizer.get_parameters(), density);
			t_rebuild += since(t0);
			const double pop = all_kernels->calculate_population(), pur = all_kernels->calculate_purity();
			if (pur > (1.0 + 2.0 * AverageTolerance) * Purity || pop > 1.0 + 2.0 * AverageTolerance || pop < 1.0 - 2.0 * AverageTolerance)
			{
				reoptimize(); // main.cpp:176-186
			}
		}
		output(iTick);
		hash_points(iTick);
	}
	const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count();
	std::printf("ranks %d\nelements %zu %zu %zu\noptimisations %zu\nwall_s %.3f\n", Context::num_ranks(), density[0].size(), density[1].size(), density[2].size(), optimisations, wall);
	std::printf("seconds_evolve_second_half %.4f\nticks_second_half %zu\n", t_evolve_late, late_ticks);
	std::printf("seconds_evolve %.4f\nseconds_rebuild %.4f\nseconds_optimise_in_loop %.4f\nseconds_new_element_selection %.4f\n", t_evolve, t_rebuild, t_opt, t_select);
	return 0;
}
