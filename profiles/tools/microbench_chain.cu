// Where do the cycles of the 8 x 8 diagonal-block factor go?  Single warp, clock64 between the stages.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int FW = 8, LP = 132;
__device__ __forceinline__ double fast_rsqrt_plain(const double d)
{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
	const double h = 0.5 * d;
	y = y * fma(-h * y, y, 1.5);
	y = y * fma(-h * y, y, 1.5);
	return y;
}
__device__ __forceinline__ long long tick()
{
	long long t;
	asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
	return t;
}
__global__ void bench(long long* ticks, double* out, double seed)
{
	__shared__ double S[8 * LP], Ld[64], rsd[8];
	for (int e = threadIdx.x; e < 8 * LP; e += 32)
	{
		const int i = e / LP, k = e % LP;
		S[e] = k < 8 ? exp(-0.02 * (i - k) * (i - k)) + (i == k ? 0.5 : 0.0) : 0.0;
	}
	__syncwarp();
	const int lane = threadIdx.x & 31;
	for (int rep = 0; rep < 3; rep++)
	{
		const long long t0 = tick();
		double l[FW][FW], rs[FW];
#pragma unroll
		for (int i = 0; i < FW; i++)
#pragma unroll
			for (int k = 0; k <= i; k++)
				l[i][k] = S[i * LP + k];
		// force the loads to complete
		double sum = 0.0;
#pragma unroll
		for (int i = 0; i < FW; i++)
#pragma unroll
			for (int k = 0; k <= i; k++)
				sum += l[i][k];
		if (sum == 1234.5)
			out[33] = sum;
		const long long t1 = tick();
#pragma unroll
		for (int j = 0; j < FW; j++)
		{
			const double d = l[j][j];
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
		double sum2 = 0.0;
#pragma unroll
		for (int i = 0; i < FW; i++)
#pragma unroll
			for (int k = 0; k <= i; k++)
				sum2 += l[i][k];
		if (sum2 == 1234.5)
			out[34] = sum2;
		const long long t2 = tick();
#pragma unroll
		for (int i = 0; i < FW; i++)
			if (lane == i)
			{
#pragma unroll
				for (int k = 0; k < FW; k++)
					Ld[i * FW + k] = k <= i ? l[i][k] : 0.0;
				rsd[i] = rs[i];
			}
		__syncwarp();
		const long long t3 = tick();
		// synthetic: 8 x (rsqrt seed + 8 dependent FP64 operations)
		double x = seed + Ld[0];
#pragma unroll
		for (int j = 0; j < 8; j++)
		{
			double r = fast_rsqrt_plain(x);
			r = r * x;
			x = fma(-r, r, x + 1.0);
		}
		if (x == 1234.5)
			out[35] = x;
		const long long t4 = tick();
		// synthetic: 64 dependent alternating DMUL / DFMA
		double y = seed;
#pragma unroll
		for (int j = 0; j < 32; j++)
		{
			y = y * 1.0000001;
			y = fma(y, 0.999999, 1e-9);
		}
		if (y == 1234.5)
			out[36] = y;
		const long long t5 = tick();
		// synthetic: 16 dependent rsqrt.approx.ftz.f64 alone
		double z = seed;
#pragma unroll
		for (int j = 0; j < 16; j++)
		{
			asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(z));
		}
		if (z == 1234.5)
			out[37] = z;
		const long long t6 = tick();
		if (threadIdx.x == 0)
		{
			ticks[0] = t1 - t0;
			ticks[1] = t2 - t1;
			ticks[2] = t3 - t2;
			ticks[3] = t4 - t3;
			ticks[4] = t5 - t4;
			ticks[5] = t6 - t5;
		}
	}
	out[threadIdx.x] = Ld[threadIdx.x] + rsd[threadIdx.x & 7];
}
int main()
{
	long long* ticks;
	double* out;
	cudaMalloc(&ticks, 64);
	cudaMalloc(&out, 64 * 8);
	bench<<<1, 32>>>(ticks, out, 0.75);
	long long h[6];
	cudaMemcpy(h, ticks, sizeof(h), cudaMemcpyDeviceToHost);
	std::printf("36 loads + sum %lld | factor chain + sum %lld | 8 divergent row stores %lld | 8 x (rsqrt + 3 dependent ops) %lld | 64 dependent DMUL/DFMA %lld (%.1f each) | 16 dependent rsqrt.approx.f64 %lld (%.1f each)\n", h[0], h[1], h[2], h[3], h[4], h[4] / 64.0, h[5], h[5] / 16.0);
	return cudaGetLastError() != cudaSuccess;
}
