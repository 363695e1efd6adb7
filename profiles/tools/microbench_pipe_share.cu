// Do DMMA.8x8x4 and DFMA share one execution pipe on B200?  (profiles/r02_kstar_fusion_decision.md)
// Three runs of the same grid (148 x 4 CTAs of 512 threads, 16 warps per CTA): every warp DMMA; every warp DFMA; even warps
// DMMA + odd warps DFMA.  If the two instruction kinds ran on separate pipes, the mixed run would take about
// max(T_dmma, T_dfma) / 2; if they share the pipe it takes about (T_dmma + T_dfma) / 2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_pipe_share microbench_pipe_share.cu && ./microbench_pipe_share
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], const double a, const double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// mode 0: all warps DMMA, 1: all warps DFMA, 2: even warps DMMA / odd warps DFMA
__global__ void __launch_bounds__(512) mix(double* out, const int mode, const int iters, const double seed)
{
	const int warp = threadIdx.x >> 5;
	const bool do_mma = mode == 0 || (mode == 2 && (warp & 1) == 0);
	double acc[8][2];
#pragma unroll
	for (int i = 0; i < 8; i++)
	{
		acc[i][0] = seed + i;
		acc[i][1] = seed - i;
	}
	const double a = seed * 0.5, b = seed * 0.25;
	if (do_mma)
	{
		for (int it = 0; it < iters; it++)
		{
#pragma unroll
			for (int i = 0; i < 8; i++)
			{
				dmma884(acc[i], a, b); // 8 independent accumulator chains: 512 flops each
			}
		}
	}
	else
	{
		for (int it = 0; it < iters; it++)
		{
#pragma unroll
			for (int r = 0; r < 4; r++) // 64 DFMA per thread and iteration = 8 DMMA worth of flops per warp (8 * 512 = 32 * 64 * 2)
			{
#pragma unroll
				for (int i = 0; i < 8; i++)
				{
					acc[i][0] = fma(acc[i][0], a, b);
					acc[i][1] = fma(acc[i][1], a, b);
				}
			}
		}
	}
	double s = 0.0;
#pragma unroll
	for (int i = 0; i < 8; i++)
	{
		s += acc[i][0] + acc[i][1];
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
	int sms = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	const int grid = sms * 4, iters = 20000;
	double* out;
	cudaMalloc(&out, size_t(grid) * 512 * sizeof(double));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float ms[3];
	for (int mode = 0; mode < 3; mode++)
	{
		mix<<<grid, 512>>>(out, mode, 200, 0.5);
		cudaEventRecord(e0);
		mix<<<grid, 512>>>(out, mode, iters, 0.5);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms[mode], e0, e1);
	}
	const double flops_all = double(grid) * 16 * iters * 8 * 512; // every warp, either kind: 8 * 512 flops per iteration
	std::printf("SMs %d, grid %d x 512 threads, %d iterations\n", sms, grid, iters);
	std::printf("all warps DMMA.8x8x4          %8.3f ms  %6.2f TFLOP/s\n", ms[0], flops_all / ms[0] / 1e9);
	std::printf("all warps DFMA                %8.3f ms  %6.2f TFLOP/s\n", ms[1], flops_all / ms[1] / 1e9);
	std::printf("even warps DMMA, odd DFMA     %8.3f ms  %6.2f TFLOP/s (both kinds together)\n", ms[2], flops_all / ms[2] / 1e9);
	std::printf("shared pipe predicts %.3f ms, separate pipes predict %.3f ms\n", 0.5 * (ms[0] + ms[1]), 0.5 * (ms[0] > ms[1] ? ms[0] : ms[1]));
	return cudaGetLastError() != cudaSuccess;
}
