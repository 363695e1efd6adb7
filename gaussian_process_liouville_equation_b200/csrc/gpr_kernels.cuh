// Device kernels of the GPR path shared by the real and the complex (composite) element models.
//
// A "block spec" describes a covariance made of up to 2 x 2 Gaussian-ARD blocks:
//   real element    : 1 x 1, sigma_f^2 (exp(-r^2/2) + sigma_n^2 delta)                 (gple/kernel.h:25-28)
//   complex element : the widely-linear process of gple/complex_kernel.h:12-13 written as the covariance
//                     of [Re f; Im f]:  [[s^2 (K_R + sn^2/2 d), s^2 K_C], [s^2 K_C, s^2 (K_I + sn^2/2 d)]]
//                     (K = K_rr + K_ii, Kt = K_rr - K_ii + 2 i K_ri).
#pragma once
#include "common.cuh"

namespace gple
{
struct GaussBlock
{
	double mag2;	 // prefactor of the exponential
	double inv_lx;	 // 1 / l_x
	double inv_lp;	 // 1 / l_p
	double diag_add; // added where the two points are the same point (noise term)
};

struct BlockSpec
{
	int nb; // 1 (real) or 2 (composite)
	GaussBlock b[2][2];
};

__device__ __forceinline__ double gauss_value(const GaussBlock& s, const double2 a, const double2 c)
{
	const double dx = (a.x - c.x) * s.inv_lx, dp = (a.y - c.y) * s.inv_lp;
	return s.mag2 * exp(-0.5 * (dx * dx + dp * dp));
}

/// Deterministic block reduction of `NV` running sums per thread; result valid in thread 0.
template <int NV, int THREADS>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* scratch /* [NV * THREADS / 32] */)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int i = 0; i < NV; i++)
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
		}
	}
	if (lane == 0)
	{
#pragma unroll
		for (int i = 0; i < NV; i++)
		{
			scratch[i * (THREADS / 32) + warp] = v[i];
		}
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
#pragma unroll
		for (int i = 0; i < NV; i++)
		{
			double s = 0.0;
			for (int w = 0; w < THREADS / 32; w++)
			{
				s += scratch[i * (THREADS / 32) + w];
			}
			v[i] = s;
		}
	}
	__syncthreads();
}

} // namespace gple
