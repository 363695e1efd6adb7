// Metropolis sampling of the density-matrix elements on the GPU (gple/mc.cpp:125-403).
//
// The reference walks one Markov chain per phase-space point (generate_markov_chain, mc.cpp:143-188) under par_unseq and
// evaluates the target density -- the analytic initial Wigner function (mc.cpp:30-50), the GPR prediction
// (main.cpp:75-101) or new_point_predict (evolve.cpp:425-443) -- once per step and chain.  Here all chains of an element
// advance in LOCK-STEP: one step = proposal kernel -> ONE batched density evaluation of all n proposals (the batched
// prediction path of gpr.cu / evolve.cu) -> accept kernel.  The analytic target needs no model, so its whole walk runs
// inside a single kernel, one chain per thread.
//
// Randomness: the reference shares one clock-seeded std::mt19937 between its threads (mc.cpp:17, 168, 173), which is
// not reproducible; here chain i owns the counter-based stream Philox4x32-10(key = seed, counter = (i, step, stream,
// block)) (Salmon et al., SC'11), so a run is a pure function of (seed, stream) however the chains are scheduled or
// sharded over GPUs.  block 0 -> the two displacement uniforms of std::uniform_real_distribution(-d, d)
// (mc.cpp:125-133), block 1 -> the acceptance uniform (mc.cpp:153, 173).
#include "mc.cuh"

#include <math_constants.h>

#include "evolve.cuh"
#include "gpr.cuh"

namespace gple
{
namespace
{
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
	for (int round = 0; round < 10; round++)
	{
		const uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
		const uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
		const uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
		c[0] = n0;
		c[1] = l1;
		c[2] = n2;
		c[3] = l0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
}
/// [0, 1) with 53 random bits
__device__ __forceinline__ double uniform53(const uint32_t hi, const uint32_t lo)
{
	return double((uint64_t(hi >> 5) << 26) | uint64_t(lo >> 6)) * (1.0 / 9007199254740992.0);
}
/// the three uniforms of (chain, step): displacement x, displacement p, acceptance
__device__ __forceinline__ void chain_draws(const unsigned long long seed, const unsigned long long stream, const unsigned long long chain, const uint32_t step, double (&u)[3])
{
	const uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
	uint32_t a[4] = {uint32_t(chain), uint32_t(chain >> 32), step, uint32_t(stream << 1)};
	uint32_t b[4] = {uint32_t(chain), uint32_t(chain >> 32), step, uint32_t(stream << 1) | 1u};
	philox4x32_10(a, k0, k1);
	philox4x32_10(b, k0, k1);
	u[0] = uniform53(a[0], a[1]);
	u[1] = uniform53(a[2], a[3]);
	u[2] = uniform53(b[0], b[1]);
}

struct Analytic
{
	double x0, p0, sx, sp, pop[2], phase[2];
};
/// initial_distribution (gple/mc.cpp:30-50)
__device__ __forceinline__ double2 initial_distribution(const Analytic& a, const double x, const double p, const int row, const int col)
{
	const double dx = (x - a.x0) / a.sx, dp = (p - a.p0) / a.sp;
	const double gw = exp(-(dx * dx + dp * dp) / 2.0) / (2.0 * CUDART_PI * (a.sx * a.sp));
	const double sw = 0.0 + a.pop[0] * a.pop[0] + a.pop[1] * a.pop[1];
	const double mag = gw * a.pop[row] * a.pop[col] / sw;
	double s, c;
	sincos(a.phase[row] - a.phase[col], &s, &c);
	return make_double2(mag * c, mag * s);
}

/// The whole walk of the analytic target, one chain per thread.  pts: (x, p, re, im) in/out.
__global__ void __launch_bounds__(256) analytic_chains_kernel(const Analytic a, const int row, const int col, double4* __restrict__ pts, const long long n, const unsigned num_steps, const double disp, const unsigned long long seed, const unsigned long long stream, const unsigned long long chain0, double* __restrict__ accept, double2* __restrict__ chain_out)
{
	const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k >= n)
	{
		return;
	}
	double x = pts[k].x, p = pts[k].y;
	double2 rho = initial_distribution(a, x, p, row, col);
	double w_old = hypot(rho.x, rho.y);
	double2* out = chain_out != nullptr ? chain_out + size_t(k) * (num_steps + 1) : nullptr;
	if (out != nullptr)
	{
		out[0] = make_double2(x, p);
	}
	unsigned acc = 0;
	for (unsigned it = 0; it < num_steps; it++)
	{
		double u[3];
		chain_draws(seed, stream, chain0 + k, it, u);
		const double xn = x + (2.0 * u[0] - 1.0) * disp, pn = p + (2.0 * u[1] - 1.0) * disp;
		const double2 rn = initial_distribution(a, xn, pn, row, col);
		const double w_new = hypot(rn.x, rn.y);
		if (w_new > w_old || w_new / w_old > u[2]) // mc.cpp:173
		{
			x = xn;
			p = pn;
			rho = rn;
			w_old = w_new;
			acc++;
		}
		if (out != nullptr)
		{
			out[it + 1] = make_double2(x, p);
		}
	}
	pts[k] = make_double4(x, p, rho.x, rho.y);
	if (accept != nullptr)
	{
		accept[k] = num_steps > 0 ? double(acc) / double(num_steps) : 0.0;
	}
}

/// proposals of one lock-step move: r_new = r + U(-d, d)^2
__global__ void __launch_bounds__(256) propose_kernel(const double4* __restrict__ pts, const long long n, const unsigned step, const double disp, const unsigned long long seed, const unsigned long long stream, const unsigned long long chain0, double2* __restrict__ r_new)
{
	const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k >= n)
	{
		return;
	}
	double u[3];
	chain_draws(seed, stream, chain0 + k, step, u);
	const double4 pt = pts[k];
	r_new[k] = make_double2(pt.x + (2.0 * u[0] - 1.0) * disp, pt.y + (2.0 * u[1] - 1.0) * disp);
}

/// rho: nb doubles per chain (1: real element, 2: complex).  First call (step == ~0u) only installs the start density.
__global__ void __launch_bounds__(256) accept_kernel(double4* __restrict__ pts, const long long n, const unsigned step, const unsigned num_steps, const unsigned long long seed, const unsigned long long stream, const unsigned long long chain0, const double2* __restrict__ r_new, const double* __restrict__ rho_new, const int nb, unsigned* __restrict__ acc, double2* __restrict__ chain_out)
{
	const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k >= n)
	{
		return;
	}
	double4 pt = pts[k];
	const double re = rho_new != nullptr ? rho_new[k * nb] : 0.0, im = (rho_new != nullptr && nb == 2) ? rho_new[k * nb + 1] : 0.0;
	double2* out = chain_out != nullptr ? chain_out + size_t(k) * (num_steps + 1) : nullptr;
	if (step == 0xffffffffu)
	{
		pt.z = re;
		pt.w = im;
		pts[k] = pt;
		if (out != nullptr)
		{
			out[0] = make_double2(pt.x, pt.y);
		}
		return;
	}
	double u[3];
	chain_draws(seed, stream, chain0 + k, step, u);
	const double w_old = hypot(pt.z, pt.w), w_new = hypot(re, im);
	if (w_new > w_old || w_new / w_old > u[2]) // mc.cpp:173
	{
		const double2 r = r_new[k];
		pt = make_double4(r.x, r.y, re, im);
		pts[k] = pt;
		acc[k]++;
	}
	if (out != nullptr)
	{
		out[step + 1] = make_double2(pt.x, pt.y);
	}
}

__global__ void accept_ratio_kernel(const unsigned* __restrict__ acc, const long long n, const unsigned num_steps, double* __restrict__ out)
{
	const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k < n)
	{
		out[k] = num_steps > 0 ? double(acc[k]) / double(num_steps) : 0.0;
	}
}

__global__ void coords_kernel(const double4* __restrict__ pts, const long long n, double2* __restrict__ r)
{
	const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k < n)
	{
		r[k] = make_double2(pts[k].x, pts[k].y);
	}
}

/// Autocorrelation of every chain (mc.cpp:230-243), one CTA per chain, accumulated into part[chain][len / 2]:
///   part[j] = sum_i (r_i - avg) . (r_{i+j} - avg) / (len - j)
__global__ void __launch_bounds__(256) autocorrelation_kernel(const double2* __restrict__ chains, const int len, double* __restrict__ part)
{
	extern __shared__ double2 sm_chain[]; // len entries, centred
	__shared__ double red[2 * 8];
	const double2* c = chains + size_t(blockIdx.x) * len;
	double sx = 0.0, sp = 0.0;
	for (int i = threadIdx.x; i < len; i += 256)
	{
		const double2 v = c[i];
		sm_chain[i] = v;
		sx += v.x;
		sp += v.y;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		sx += __shfl_xor_sync(0xffffffffu, sx, o);
		sp += __shfl_xor_sync(0xffffffffu, sp, o);
	}
	if ((threadIdx.x & 31) == 0)
	{
		red[threadIdx.x >> 5] = sx;
		red[8 + (threadIdx.x >> 5)] = sp;
	}
	__syncthreads();
	double ax = 0.0, ap = 0.0;
	for (int w = 0; w < 8; w++)
	{
		ax += red[w];
		ap += red[8 + w];
	}
	ax /= double(len);
	ap /= double(len);
	__syncthreads();
	for (int i = threadIdx.x; i < len; i += 256)
	{
		sm_chain[i] = make_double2(sm_chain[i].x - ax, sm_chain[i].y - ap);
	}
	__syncthreads();
	for (int j = threadIdx.x; j < len / 2; j += 256)
	{
		double s = 0.0;
		for (int i = 0; i + j < len; i++)
		{
			const double2 a = sm_chain[i], b = sm_chain[i + j];
			s += a.x * b.x + a.y * b.y;
		}
		part[size_t(blockIdx.x) * (len / 2) + j] = s / double(len - j);
	}
}

/// out[j] = mean over the chains, summed in chain order (deterministic)
__global__ void autocorrelation_mean_kernel(const double* __restrict__ part, const long long n, const int half, double* __restrict__ out)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= half)
	{
		return;
	}
	double s = 0.0;
	for (long long k = 0; k < n; k++)
	{
		s += part[size_t(k) * half + j] / double(n);
	}
	out[j] = s;
}
} // namespace

void markov_chains_device(gple_ctx* ctx, const gple_mc_source& src, double* d_pts, const size_t n, const size_t num_steps, const double max_displacement, const unsigned long long seed, const unsigned long long stream, const unsigned long long chain0, double* d_accept, double* d_chain)
{
	const unsigned grid = unsigned((n + 255) / 256);
	double4* pts = reinterpret_cast<double4*>(d_pts);
	double2* chain = reinterpret_cast<double2*>(d_chain);
	if (src.kind == GPLE_MC_ANALYTIC)
	{
		const Analytic a{src.analytic[0], src.analytic[1], src.analytic[2], src.analytic[3], {src.analytic[4], src.analytic[5]}, {src.analytic[6], src.analytic[7]}};
		GPLE_LAUNCH(ctx, analytic_chains_kernel, grid, 256, 0, a, src.row, src.col, pts, (long long)n, unsigned(num_steps), max_displacement, seed, stream, chain0, d_accept, chain);
		return;
	}
	const gple_model* models[3] = {src.m00, src.m10, src.m11};
	const int element = src.row + src.col; // (0,0) -> 0, (1,0) -> 1, (1,1) -> 2
	const gple_model* own = models[element];
	const int nb = src.kind == GPLE_MC_NEW_POINT ? 2 : (element == 1 ? 2 : 1);
	double2* r_new = ctx->ws.get<double2>("mc.r_new", n);
	double* rho_new = ctx->ws.get<double>("mc.rho_new", 2 * n);
	unsigned* acc = ctx->ws.get<unsigned>("mc.acc", n);
	GPLE_CUDA(cudaMemsetAsync(acc, 0, n * sizeof(unsigned), ctx->stream));
	// density at `r` into rho_new; returns false when the target is identically zero (no model)
	auto density = [&](const double2* r) -> bool
	{
		if (src.kind == GPLE_MC_NEW_POINT)
		{
			new_point_predict_device(ctx, src.pes_model, models, reinterpret_cast<const double*>(r), n, src.row, src.col, src.mass, src.dt, rho_new);
			return true;
		}
		if (own == nullptr)
		{
			return false;
		}
		predict_device(ctx, own, reinterpret_cast<const double*>(r), n, nullptr, nullptr, nullptr, rho_new, nullptr);
		return true;
	};
	GPLE_LAUNCH(ctx, coords_kernel, grid, 256, 0, pts, (long long)n, r_new);
	const bool has0 = density(r_new);
	GPLE_LAUNCH(ctx, accept_kernel, grid, 256, 0, pts, (long long)n, 0xffffffffu, unsigned(num_steps), seed, stream, chain0, r_new, has0 ? rho_new : nullptr, nb, acc, chain);
	for (unsigned it = 0; it < unsigned(num_steps); it++)
	{
		GPLE_LAUNCH(ctx, propose_kernel, grid, 256, 0, pts, (long long)n, it, max_displacement, seed, stream, chain0, r_new);
		const bool has = density(r_new);
		GPLE_LAUNCH(ctx, accept_kernel, grid, 256, 0, pts, (long long)n, it, unsigned(num_steps), seed, stream, chain0, r_new, has ? rho_new : nullptr, nb, acc, chain);
	}
	if (d_accept != nullptr)
	{
		GPLE_LAUNCH(ctx, accept_ratio_kernel, grid, 256, 0, acc, (long long)n, unsigned(num_steps), d_accept);
	}
}

void mc_setup_attributes()
{
	GPLE_CUDA(cudaFuncSetAttribute(autocorrelation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

void chain_autocorrelation_device(gple_ctx* ctx, const double* d_chains, const size_t n, const size_t len, double* d_out)
{
	if (len * sizeof(double2) > 200 * 1024)
	{
		throw ArgError{"gple_chain_autocorrelation: chains longer than 12800 states do not fit in shared memory"};
	}
	const int half = int(len / 2);
	double* part = ctx->ws.get<double>("mc.autocor", n * size_t(half));
	GPLE_LAUNCH(ctx, autocorrelation_kernel, unsigned(n), 256, len * sizeof(double2), reinterpret_cast<const double2*>(d_chains), int(len), part);
	GPLE_LAUNCH(ctx, autocorrelation_mean_kernel, (half + 255) / 256, 256, 0, part, (long long)n, half, d_out);
}

} // namespace gple
