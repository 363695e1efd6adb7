// Latency of the 8 x 8 diagonal-block Cholesky inside the factorisation leaf, single warp, three codings
// (profiles/r02_leaf_latency.md).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_diag8 microbench_diag8.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int FW = 8, LP = 132;

__device__ __forceinline__ double fast_rsqrt_checked(const double d)
{
	if (!(d > 1e-290 && d < 1e290))
	{
		return rsqrt(d);
	}
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
	const double h = 0.5 * d;
	y = y * fma(-h * y, y, 1.5);
	y = y * fma(-h * y, y, 1.5);
	return y;
}
__device__ __forceinline__ double fast_rsqrt_plain(const double d)
{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
	const double h = 0.5 * d;
	y = y * fma(-h * y, y, 1.5);
	y = y * fma(-h * y, y, 1.5);
	return y;
}

// V0: as in the leaf today (checked rsqrt, pivot test with a branch)
__device__ __noinline__ void v0(const double* S, double* Ld, double* rsd, int* info)
{
	double l[FW][FW], rs[FW];
#pragma unroll
	for (int i = 0; i < FW; i++)
#pragma unroll
		for (int k = 0; k <= i; k++)
			l[i][k] = S[i * LP + k];
	bool bad = false;
#pragma unroll
	for (int j = 0; j < FW; j++)
	{
		double d = l[j][j];
		if (!(d > 0.0))
		{
			if (!bad && (threadIdx.x & 31) == 0)
				atomicCAS(info, 0, j + 1);
			bad = true;
			d = 1.0;
		}
		rs[j] = fast_rsqrt_checked(d);
		l[j][j] = d * rs[j];
#pragma unroll
		for (int i = j + 1; i < FW; i++)
			l[i][j] *= rs[j];
#pragma unroll
		for (int k = j + 1; k < FW; k++)
#pragma unroll
			for (int i = k; i < FW; i++)
				l[i][k] = fma(-l[i][j], l[k][j], l[i][k]);
	}
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int i = 0; i < FW; i++)
		if (lane == i)
		{
#pragma unroll
			for (int k = 0; k < FW; k++)
				Ld[i * FW + k] = k <= i ? l[i][k] : 0.0;
			rsd[i] = rs[i];
		}
}

// V1: no branch inside the chain: pivots outside (1e-290, 1e290) only raise a flag (the caller redoes the block the slow way)
__device__ __noinline__ void v1(const double* S, double* Ld, double* rsd, int* info)
{
	double l[FW][FW], rs[FW];
#pragma unroll
	for (int i = 0; i < FW; i++)
#pragma unroll
		for (int k = 0; k <= i; k++)
			l[i][k] = S[i * LP + k];
	bool ok = true;
#pragma unroll
	for (int j = 0; j < FW; j++)
	{
		const double d = l[j][j];
		ok = ok && d > 1e-290 && d < 1e290;
		rs[j] = fast_rsqrt_plain(d);
		l[j][j] = d * rs[j];
#pragma unroll
		for (int i = j + 1; i < FW; i++)
			l[i][j] *= rs[j];
#pragma unroll
		for (int k = j + 1; k < FW; k++)
#pragma unroll
			for (int i = k; i < FW; i++)
				l[i][k] = fma(-l[i][j], l[k][j], l[i][k]);
	}
	const int lane = threadIdx.x & 31;
	if (!ok && lane == 0)
		atomicCAS(info, 0, 1);
#pragma unroll
	for (int i = 0; i < FW; i++)
		if (lane == i)
		{
#pragma unroll
			for (int k = 0; k < FW; k++)
				Ld[i * FW + k] = k <= i ? l[i][k] : 0.0;
			rsd[i] = rs[i];
		}
}

// V2: V1 with the next pivot's rsqrt chain shortened: d_{j+1} is updated FIRST in each step, and its update uses
// rs^2 = 1/d_j computed off the Newton result in parallel with the column scaling
__device__ __noinline__ void v2(const double* S, double* Ld, double* rsd, int* info)
{
	double l[FW][FW], rs[FW];
#pragma unroll
	for (int i = 0; i < FW; i++)
#pragma unroll
		for (int k = 0; k <= i; k++)
			l[i][k] = S[i * LP + k];
	bool ok = true;
#pragma unroll
	for (int j = 0; j < FW; j++)
	{
		const double d = l[j][j];
		ok = ok && d > 1e-290 && d < 1e290;
		rs[j] = fast_rsqrt_plain(d);
		const double inv = rs[j] * rs[j]; // 1 / d_j
		// trailing update with the UNSCALED column: a_ik -= a_ij a_kj / d_j  (the scaled column is formed on the side)
#pragma unroll
		for (int k = j + 1; k < FW; k++)
		{
			const double f = l[k][j] * inv;
#pragma unroll
			for (int i = k; i < FW; i++)
				l[i][k] = fma(-l[i][j], f, l[i][k]);
		}
		l[j][j] = d * rs[j];
#pragma unroll
		for (int i = j + 1; i < FW; i++)
			l[i][j] *= rs[j];
	}
	const int lane = threadIdx.x & 31;
	if (!ok && lane == 0)
		atomicCAS(info, 0, 1);
#pragma unroll
	for (int i = 0; i < FW; i++)
		if (lane == i)
		{
#pragma unroll
			for (int k = 0; k < FW; k++)
				Ld[i * FW + k] = k <= i ? l[i][k] : 0.0;
			rsd[i] = rs[i];
		}
}

__global__ void bench(long long* ticks, int* info, double* out)
{
	__shared__ double S[8 * LP], Ld[64], rsd[8];
	for (int e = threadIdx.x; e < 8 * LP; e += 32)
	{
		const int i = e / LP, k = e % LP;
		S[e] = k < 8 ? exp(-0.02 * (i - k) * (i - k)) + (i == k ? 0.5 : 0.0) : 0.0;
	}
	__syncwarp();
	for (int rep = 0; rep < 3; rep++)
	{
		long long t0 = clock64();
		v0(S, Ld, rsd, info);
		__syncwarp();
		long long t1 = clock64();
		v1(S, Ld, rsd, info);
		__syncwarp();
		long long t2 = clock64();
		v2(S, Ld, rsd, info);
		__syncwarp();
		long long t3 = clock64();
		if (threadIdx.x == 0)
		{
			ticks[0] = t1 - t0;
			ticks[1] = t2 - t1;
			ticks[2] = t3 - t2;
		}
	}
	out[threadIdx.x] = Ld[threadIdx.x] + rsd[threadIdx.x & 7];
}

int main()
{
	long long* ticks;
	int* info;
	double* out;
	cudaMalloc(&ticks, 64);
	cudaMalloc(&info, 4);
	cudaMalloc(&out, 32 * 8);
	cudaMemset(info, 0, 4);
	bench<<<1, 32>>>(ticks, info, out);
	long long h[3];
	cudaMemcpy(h, ticks, sizeof(h), cudaMemcpyDeviceToHost);
	double ho[32];
	cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
	std::printf("8 x 8 block factor, one warp, third repetition:\n  V0 checked rsqrt + pivot branch   %6lld cycles\n  V1 branch-free chain             %6lld cycles\n  V2 branch-free, update before scale %6lld cycles\n  (out[0] = %.15g)\n", h[0], h[1], h[2], ho[0]);
	return cudaGetLastError() != cudaSuccess;
}
