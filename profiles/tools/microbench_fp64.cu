// Latency micro-benchmarks behind the factorisation leaf design (profiles/r02_leaf_latency.md): dependent-issue latency of
// DFMA, the rsqrt seed chain, DMMA.8x8x4, shared-memory loads, __syncthreads with 256 threads, on one SM of a B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu && ./microbench_fp64
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#include "../../gaussian_process_liouville_equation_b200/csrc/chol.cu" // the leaf kernel in its TIMED instantiation

__device__ __forceinline__ void dmma884(double (&c)[2], const double a, const double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__global__ void bench(double* out, long long* ticks, double seed)
{
	__shared__ double sm[1024];
	const int tid = threadIdx.x;
	sm[tid] = seed + tid;
	sm[tid + 256] = seed;
	sm[tid + 512] = seed;
	sm[tid + 768] = seed;
	__syncthreads();
	constexpr int R = 512;
	double x = seed, y = seed * 0.5;
	long long t0, t1;
	// 1. dependent DFMA chain
	t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < R; i++)
	{
		x = fma(x, y, 1e-9);
	}
	t1 = clock64();
	if (tid == 0)
	{
		ticks[0] = t1 - t0;
	}
	// 2. rsqrt seed + 2 Newton steps, dependent
	t0 = clock64();
#pragma unroll 8
	for (int i = 0; i < R / 8; i++)
	{
		double d = x + 1.5;
		double r = double(rsqrtf(float(d)));
		const double h = 0.5 * d;
		r = r * fma(-h * r, r, 1.5);
		r = r * fma(-h * r, r, 1.5);
		x = r;
	}
	t1 = clock64();
	if (tid == 0)
	{
		ticks[1] = t1 - t0;
	}
	// 3. dependent DMMA chain (accumulator dependency)
	double c[2] = {x, y};
	t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < R; i++)
	{
		dmma884(c, y, y);
	}
	t1 = clock64();
	if (tid == 0)
	{
		ticks[2] = t1 - t0;
	}
	// 4. dependent shared-memory loads (pointer chase)
	int idx = tid;
	t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < R; i++)
	{
		idx = (int(sm[idx] * 0.0) + idx + 32) & 1023;
	}
	t1 = clock64();
	if (tid == 0)
	{
		ticks[3] = t1 - t0;
	}
	// 5. __syncthreads
	t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < R; i++)
	{
		__syncthreads();
	}
	t1 = clock64();
	if (tid == 0)
	{
		ticks[4] = t1 - t0;
	}
	// 6. double division, dependent
	t0 = clock64();
#pragma unroll 8
	for (int i = 0; i < R / 8; i++)
	{
		x = 1.0 / (x + 1.25);
	}
	t1 = clock64();
	if (tid == 0)
	{
		ticks[5] = t1 - t0;
	}
	// 7. DMMA with operands from a dependent shared-memory load: load -> dmma -> store -> (next reads it)
	t0 = clock64();
#pragma unroll 8
	for (int i = 0; i < R / 8; i++)
	{
		double cc[2] = {sm[tid], sm[tid + 256]};
		dmma884(cc, sm[tid + 512], y);
		sm[tid] = cc[0];
		sm[tid + 256] = cc[1];
		__syncwarp();
	}
	t1 = clock64();
	if (tid == 0)
	{
		ticks[6] = t1 - t0;
	}
	// 8. warp shuffle of a double, dependent
	t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < R; i++)
	{
		x = __shfl_sync(0xffffffffu, x, (tid + 1) & 31);
	}
	t1 = clock64();
	if (tid == 0)
	{
		ticks[7] = t1 - t0;
	}
	out[tid] = x + c[0] + c[1] + idx;
}

int main()
{
	double* out;
	long long* ticks;
	cudaMalloc(&out, 256 * sizeof(double));
	cudaMalloc(&ticks, 16 * sizeof(long long));
	for (int rep = 0; rep < 2; rep++)
	{
		bench<<<1, 256>>>(out, ticks, 0.75);
	}
	long long h[16];
	cudaMemcpy(h, ticks, sizeof(h), cudaMemcpyDeviceToHost);
	const char* names[8] = {"DFMA dependent", "rsqrtf seed + 2 Newton (per rsqrt)", "DMMA.8x8x4 dependent accumulator", "shared load pointer chase", "__syncthreads (256 threads)", "double division", "LDS -> DMMA -> STS round trip", "shfl.sync of a double"};
	const int per[8] = {512, 64, 512, 512, 512, 64, 64, 512};
	for (int i = 0; i < 8; i++)
	{
		std::printf("%-40s %8.1f cycles\n", names[i], double(h[i]) / per[i]);
	}
	// the factorisation leaf, phase by phase (clock64 of thread 0)
	{
		using namespace gple;
		const int n = 128;
		std::vector<double> hA(n * n);
		for (int i = 0; i < n; i++)
		{
			for (int j = 0; j < n; j++)
			{
				hA[i * n + j] = std::exp(-0.5 * (i - j) * (i - j) / 400.0) + (i == j ? 1e-2 : 0.0);
			}
		}
		double *dA, *dinv;
		int* info;
		cudaMalloc(&dA, n * n * sizeof(double));
		cudaMalloc(&dinv, n * n * sizeof(double));
		cudaMalloc(&info, sizeof(int));
		cudaMemset(info, 0, sizeof(int));
		cudaFuncSetAttribute(potrf_leaf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LEAF_SMEM));
		for (int rep = 0; rep < 3; rep++)
		{
			cudaMemcpy(dA, hA.data(), n * n * sizeof(double), cudaMemcpyHostToDevice);
			potrf_leaf_kernel<true><<<1, LEAF_THREADS, LEAF_SMEM>>>(dA, size_t(n), dinv, info, 0, ticks);
		}
		cudaMemcpy(h, ticks, sizeof(h), cudaMemcpyDeviceToHost);
		int hinfo = -1;
		cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
		const char* ph[10] = {"load", "panel loop (factor + hidden inverse)", "last inverse row block + copy-back", "store", "total", "  loop: own phase-1 work (thread 0)", "  loop: wait for the inverse warps", "  loop: rank-8 update in the super-panel", "  loop: rank-32 update", "  loop: closing barrier"};
		std::printf("\npotrf_leaf_kernel phases (info = %d):\n", hinfo);
		for (int i = 0; i < 10; i++)
		{
			std::printf("%-42s %8lld cycles\n", ph[i], h[i]);
		}
	}
	return cudaGetLastError() != cudaSuccess;
}
