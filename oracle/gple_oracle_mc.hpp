// TEST INFRASTRUCTURE ONLY (see gple_oracle.hpp).  Metropolis sampling of the reference, gple/mc.cpp:125-403, restated
// per chain exactly as generate_markov_chain (:125-160) runs it, plus the chain autocorrelation of
// autocorrelation_optimize_steps (:176-203).
//
// The reference draws from one clock-seeded std::mt19937 shared un-synchronised between the threads of par_unseq
// (mc.cpp:17, 168, 173), so its chains are not reproducible even against itself.  To make the sampler testable, chain
// i draws from its own counter-based stream: Philox4x32-10 (Salmon et al., SC'11; the published constants below), key =
// the 64-bit seed, counter = (chain index low, chain index high, step, stream << 1 | block), block 0 -> the two
// displacement uniforms, block 1 -> the acceptance uniform.  A uniform double takes 53 bits of two output words.
#pragma once
#include "gple_oracle_dynamics.hpp"

#include <cstdint>

namespace orc
{
struct Philox
{
	static void mulhilo(const std::uint32_t a, const std::uint32_t b, std::uint32_t& hi, std::uint32_t& lo)
	{
		const std::uint64_t p = std::uint64_t(a) * b;
		hi = std::uint32_t(p >> 32);
		lo = std::uint32_t(p);
	}
	/// Philox4x32-10: ctr, key -> 4 words
	static std::array<std::uint32_t, 4> block(std::array<std::uint32_t, 4> c, std::array<std::uint32_t, 2> k)
	{
		for (int round = 0; round < 10; round++)
		{
			std::uint32_t h0, l0, h1, l1;
			mulhilo(0xD2511F53u, c[0], h0, l0);
			mulhilo(0xCD9E8D57u, c[2], h1, l1);
			c = {h1 ^ c[1] ^ k[0], l1, h0 ^ c[3] ^ k[1], l0};
			k[0] += 0x9E3779B9u;
			k[1] += 0xBB67AE85u;
		}
		return c;
	}
	/// [0, 1) with 53 random bits
	static double uniform(const std::uint32_t hi, const std::uint32_t lo)
	{
		return double((std::uint64_t(hi >> 5) << 26) | std::uint64_t(lo >> 6)) * (1.0 / 9007199254740992.0);
	}
	/// the three uniforms of (chain, step): displacement x, displacement p, acceptance
	static std::array<double, 3> draws(const std::uint64_t seed, const std::uint64_t stream, const std::uint64_t chain, const std::uint32_t step)
	{
		const std::array<std::uint32_t, 2> key{std::uint32_t(seed), std::uint32_t(seed >> 32)};
		const auto a = block({std::uint32_t(chain), std::uint32_t(chain >> 32), step, std::uint32_t(stream << 1)}, key);
		const auto b = block({std::uint32_t(chain), std::uint32_t(chain >> 32), step, std::uint32_t(stream << 1) | 1u}, key);
		return {uniform(a[0], a[1]), uniform(a[2], a[3]), uniform(b[0], b[1])};
	}
};

/// generate_markov_chain (gple/mc.cpp:143-188) for chain `chain` starting at (x, p).
/// chain_out (optional): 2 * (num_steps + 1) doubles.  Returns the acceptance ratio; (x, p) and rho end at the last state.
inline double markov_chain(const Distribution& distribution, const std::size_t row, const std::size_t col, double& x, double& p, cplx& rho, const std::size_t num_steps, const double max_displacement, const std::uint64_t seed, const std::uint64_t stream, const std::uint64_t chain, double* chain_out)
{
	rho = distribution(x, p, row, col);
	double weight_old = std::abs(rho);
	if (chain_out != nullptr)
	{
		chain_out[0] = x;
		chain_out[1] = p;
	}
	std::size_t acc = 0;
	for (std::size_t it = 0; it < num_steps; it++)
	{
		const auto u = Philox::draws(seed, stream, chain, std::uint32_t(it));
		// std::uniform_real_distribution(-d, d) (mc.cpp:125-133)
		const double xn = x + (2.0 * u[0] - 1.0) * max_displacement, pn = p + (2.0 * u[1] - 1.0) * max_displacement;
		const cplx rho_new = distribution(xn, pn, row, col);
		const double weight_new = std::abs(rho_new);
		if (weight_new > weight_old || weight_new / weight_old > u[2]) // mc.cpp:173
		{
			x = xn;
			p = pn;
			rho = rho_new;
			weight_old = weight_new;
			acc++;
		}
		if (chain_out != nullptr)
		{
			chain_out[2 * (it + 1)] = x;
			chain_out[2 * (it + 1) + 1] = p;
		}
	}
	return num_steps > 0 ? double(acc) / double(num_steps) : 0.0;
}

/// Autocorrelation of one chain (gple/mc.cpp:230-243): out[j] = sum_i (r_i - avg).(r_{i+j} - avg) / (len - j), j < len / 2
inline void chain_autocorrelation(const double* chain, const std::size_t len, double* out)
{
	double ax = 0.0, ap = 0.0;
	for (std::size_t i = 0; i < len; i++)
	{
		ax += chain[2 * i];
		ap += chain[2 * i + 1];
	}
	ax /= double(len);
	ap /= double(len);
	for (std::size_t j = 0; j < len / 2; j++)
	{
		double s = 0.0;
		for (std::size_t i = 0; i + j < len; i++)
		{
			s += (chain[2 * i] - ax) * (chain[2 * (i + j)] - ax) + (chain[2 * i + 1] - ap) * (chain[2 * (i + j) + 1] - ap);
		}
		out[j] = s / double(len - j);
	}
}
} // namespace orc
