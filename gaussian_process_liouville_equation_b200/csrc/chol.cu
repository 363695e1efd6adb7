// Blocked FP64 Cholesky (LL^T) and triangular inverse on DMMA for sm_100a.
//
// Replaces `Eigen::LDLT<MatrixXd>` + `solve(Identity)` of the reference (gple/kernel.cpp:281-283,
// gple/kernel.h:253).  The reference factorises with Eigen's unblocked, pivoted LDL^T and then forms the
// explicit inverse with 2 N^3 flops of triangular solves; K is symmetric positive definite by
// construction (sigma_n^2 > 0 on the diagonal), so here it is a recursive blocked Cholesky whose
// trailing updates, panel solves and the triangular inverse W = L^-1 are all 128-tiled DMMA GEMMs:
//   potrf(A):  potrf(A11); A21 <- A21 L11^-T (recursive, leaf = multiply by the inverted 128-block);
//              A22 <- A22 - A21 A21^T (lower tiles only); potrf(A22)
//   trtri(L):  W11 = trtri(L11); W22 = trtri(L22); W21 = -W22 (L21 W11)
// The only non-GEMM work is the 128 x 128 leaf (factor + inverse of the diagonal block in one CTA).
// Row-major, lower triangle, leading dimension ld; sizes are multiples of 128.
#include "chol.cuh"

namespace gple
{
namespace
{
constexpr int LEAF = 128;
constexpr int LP = 132;			 // pitch (doubles): 4 (mod 16) -> conflict-free for the (row, q) thread layout
constexpr int LEAF_THREADS = 512; // 4 threads per row / column
constexpr size_t LEAF_SMEM = (size_t(LEAF) * LP + LEAF) * sizeof(double);

/// Factor one 128 x 128 diagonal block in shared memory (left-looking Cholesky-Crout, 4 lanes per row)
/// and invert the resulting triangle (4 lanes per column).  L overwrites the block (upper part zeroed),
/// its inverse goes to `dinv` (row-major 128 x 128, upper part zero).
__global__ void __launch_bounds__(LEAF_THREADS, 1) potrf_leaf_kernel(double* __restrict__ A, const size_t ld, double* __restrict__ dinv, int* __restrict__ info, const int global_row0)
{
	extern __shared__ __align__(16) double sm[];
	double* S = sm;				  // [128][LP]: strictly-lower part = L; XT(c, i) = S[c][i + 1] (i >= c) = inverse, transposed
	double* dg = sm + LEAF * LP; // diagonal of L
	const int tid = threadIdx.x;
	for (int e = tid; e < LEAF * LEAF; e += LEAF_THREADS)
	{
		const int r = e >> 7, c = e & 127;
		if (c <= r)
		{
			S[r * LP + c] = A[size_t(r) * ld + c];
		}
	}
	__syncthreads();
	const int i = tid >> 2, q = tid & 3;
	for (int j = 0; j < LEAF; j++)
	{
		double s = 0.0;
		if (i >= j)
		{
			// 4 independent partial sums so that the shared-memory loads pipeline (the loop is latency-bound otherwise)
			double s1 = 0.0, s2 = 0.0, s3 = 0.0;
			const double* ri = S + i * LP;
			const double* rj = S + j * LP;
			int k = q;
			for (; k + 12 < j; k += 16)
			{
				s = fma(ri[k], rj[k], s);
				s1 = fma(ri[k + 4], rj[k + 4], s1);
				s2 = fma(ri[k + 8], rj[k + 8], s2);
				s3 = fma(ri[k + 12], rj[k + 12], s3);
			}
			for (; k < j; k += 4)
			{
				s = fma(ri[k], rj[k], s);
			}
			s = (s + s1) + (s2 + s3);
		}
		s += __shfl_xor_sync(0xffffffffu, s, 1);
		s += __shfl_xor_sync(0xffffffffu, s, 2);
		if (i >= j && q == 0)
		{
			S[i * LP + j] -= s; // column j is only read as S[.][k], k < j, by the dot products: no hazard
		}
		__syncthreads();
		double d = S[j * LP + j];
		if (!(d > 0.0))
		{
			if (tid == 0)
			{
				atomicCAS(info, 0, global_row0 + j + 1);
			}
			d = 1.0;
		}
		const double sd = sqrt(d);
		if (q == 0)
		{
			if (i > j)
			{
				S[i * LP + j] = S[i * LP + j] / sd;
			}
			else if (i == j)
			{
				dg[j] = sd;
			}
		}
		__syncthreads();
	}
	// inverse: column c of X = L^-1 by forward substitution, 4 lanes share one column
	{
		const int c = tid >> 2;
		const int c0 = (tid >> 5) << 3; // first column of this warp
		if (q == 0)
		{
			S[c * LP + c + 1] = 1.0 / dg[c];
		}
		__syncwarp();
		for (int r = c0 + 1; r < LEAF; r++)
		{
			double s = 0.0;
			if (r > c)
			{
				double s1 = 0.0, s2 = 0.0, s3 = 0.0;
				const double* lr = S + r * LP;
				const double* xc = S + c * LP + 1;
				int k = c + q;
				for (; k + 12 < r; k += 16)
				{
					s = fma(lr[k], xc[k], s);
					s1 = fma(lr[k + 4], xc[k + 4], s1);
					s2 = fma(lr[k + 8], xc[k + 8], s2);
					s3 = fma(lr[k + 12], xc[k + 12], s3);
				}
				for (; k < r; k += 4)
				{
					s = fma(lr[k], xc[k], s);
				}
				s = (s + s1) + (s2 + s3);
			}
			s += __shfl_xor_sync(0xffffffffu, s, 1);
			s += __shfl_xor_sync(0xffffffffu, s, 2);
			if (r > c && q == 0)
			{
				S[c * LP + r + 1] = -s / dg[r];
			}
			__syncwarp();
		}
	}
	__syncthreads();
	for (int e = tid; e < LEAF * LEAF; e += LEAF_THREADS)
	{
		const int r = e >> 7, c = e & 127;
		double l = 0.0, x = 0.0;
		if (c < r)
		{
			l = S[r * LP + c];
			x = S[c * LP + r + 1];
		}
		else if (c == r)
		{
			l = dg[r];
			x = S[c * LP + r + 1];
		}
		A[size_t(r) * ld + c] = l;
		dinv[e] = x;
	}
}

/// Zero the strictly-upper 128-blocks of a row-major matrix (the diagonal blocks are cleaned by the leaf).
__global__ void zero_upper_blocks_kernel(double* __restrict__ A, const size_t ld, const int n)
{
	const int bj = blockIdx.x, bi = blockIdx.y;
	if (bj <= bi)
	{
		return;
	}
	for (int e = threadIdx.x; e < LEAF * LEAF / 2; e += blockDim.x)
	{
		const int r = e >> 6, c2 = (e & 63) * 2;
		*reinterpret_cast<double2*>(A + size_t(bi * LEAF + r) * ld + bj * LEAF + c2) = make_double2(0.0, 0.0);
	}
}

/// Copy the inverted diagonal blocks into the block diagonal of W.
__global__ void copy_dinv_kernel(const double* __restrict__ dinv, double* __restrict__ W, const size_t ld)
{
	const int b = blockIdx.x;
	for (int e = threadIdx.x; e < LEAF * LEAF / 2; e += blockDim.x)
	{
		const int r = e >> 6, c2 = (e & 63) * 2;
		*reinterpret_cast<double2*>(W + size_t(b * LEAF + r) * ld + b * LEAF + c2) = *reinterpret_cast<const double2*>(dinv + size_t(b) * LEAF * LEAF + r * LEAF + c2);
	}
}

void run_gemm(gple_ctx* ctx, bool b_nn, const gemm::GemmArgs& a)
{
	if (a.M <= 0 || a.N <= 0)
	{
		return;
	}
	// 128 x 128 tiles once they fill at least about half of the SMs, 64 x 64 tiles below that (latency-bound regime)
	const long long tiles128 = (long long)(a.N / 128) * (a.M / 128) / (a.lower_only ? 2 : 1);
	if (tiles128 * 2 >= ctx->num_sms || a.in_place)
	{
		using C = gemm::DefaultConfig;
		const dim3 grid(a.N / C::BN, a.M / C::BM);
		if (b_nn)
		{
			GPLE_LAUNCH(ctx, (gemm::gemm_kernel<C, true>), grid, C::THREADS, C::SMEM_BYTES, a);
		}
		else
		{
			GPLE_LAUNCH(ctx, (gemm::gemm_kernel<C, false>), grid, C::THREADS, C::SMEM_BYTES, a);
		}
	}
	else
	{
		using C = gemm::SmallConfig;
		const dim3 grid(a.N / C::BN, a.M / C::BM);
		if (b_nn)
		{
			GPLE_LAUNCH(ctx, (gemm::gemm_kernel<C, true>), grid, C::THREADS, C::SMEM_BYTES, a);
		}
		else
		{
			GPLE_LAUNCH(ctx, (gemm::gemm_kernel<C, false>), grid, C::THREADS, C::SMEM_BYTES, a);
		}
	}
}

int split(const int n)
{
	return ((n / LEAF) / 2) * LEAF;
}

struct Chol
{
	gple_ctx* ctx;
	double* A; // whole matrix (row-major, lower)
	size_t ld;
	double* dinv; // [n / 128][128 * 128]
	int* info;

	double* at(const int r, const int c) const { return A + size_t(r) * ld + c; }

	/// B (m rows from row r0, columns c0 .. c0 + n) <- B L^-T with L the diagonal block at (c0, c0) of size n
	void trsm(const int r0, const int m, const int c0, const int n) const
	{
		if (n == LEAF)
		{
			gemm::GemmArgs g{};
			g.A = at(r0, c0);
			g.B = dinv + size_t(c0 / LEAF) * LEAF * LEAF;
			g.C = at(r0, c0);
			g.lda = ld;
			g.ldb = LEAF;
			g.ldc = ld;
			g.M = m;
			g.N = LEAF;
			g.K = LEAF;
			g.alpha = 1.0;
			g.beta = 0.0;
			g.tri = gemm::B_LOWER_NT;
			g.in_place = 1;
			run_gemm(ctx, false, g); // in place: one n-tile per row block, all reads precede the epilogue
			return;
		}
		const int n1 = split(n), n2 = n - n1;
		trsm(r0, m, c0, n1);
		gemm::GemmArgs g{};
		g.A = at(r0, c0);
		g.B = at(c0 + n1, c0); // L21 (n2 x n1)
		g.C = at(r0, c0 + n1);
		g.lda = g.ldb = g.ldc = ld;
		g.M = m;
		g.N = n2;
		g.K = n1;
		g.alpha = -1.0;
		g.beta = 1.0;
		run_gemm(ctx, false, g);
		trsm(r0, m, c0 + n1, n2);
	}

	void potrf(const int o, const int n) const
	{
		if (n == LEAF)
		{
			GPLE_LAUNCH(ctx, potrf_leaf_kernel, 1, LEAF_THREADS, LEAF_SMEM, at(o, o), ld, dinv + size_t(o / LEAF) * LEAF * LEAF, info, o);
			return;
		}
		const int n1 = split(n), n2 = n - n1;
		potrf(o, n1);
		trsm(o + n1, n2, o, n1);
		gemm::GemmArgs g{};
		g.A = at(o + n1, o);
		g.B = at(o + n1, o);
		g.C = at(o + n1, o + n1);
		g.lda = g.ldb = g.ldc = ld;
		g.M = n2;
		g.N = n2;
		g.K = n1;
		g.alpha = -1.0;
		g.beta = 1.0;
		g.lower_only = 1;
		run_gemm(ctx, false, g);
		potrf(o + n1, n2);
	}

	/// W (same layout, zero-initialised with the inverted diagonal blocks in place) <- L^-1
	void trtri(double* W, double* T, const int o, const int n) const
	{
		if (n == LEAF)
		{
			return;
		}
		const int n1 = split(n), n2 = n - n1;
		trtri(W, T, o, n1);
		trtri(W, T, o + n1, n2);
		// T (n2 x n1) = L21 * W11   (NN; W11[k][j] == 0 for j > k)
		gemm::GemmArgs g{};
		g.A = at(o + n1, o);
		g.B = W + size_t(o) * ld + o;
		g.C = T;
		g.lda = ld;
		g.ldb = ld;
		g.ldc = size_t(n1);
		g.M = n2;
		g.N = n1;
		g.K = n1;
		g.alpha = 1.0;
		g.beta = 0.0;
		g.tri = gemm::B_LOWER_NN;
		run_gemm(ctx, true, g);
		// W21 = -W22 * T            (NN; W22[i][k] == 0 for k > i)
		gemm::GemmArgs h{};
		h.A = W + size_t(o + n1) * ld + o + n1;
		h.B = T;
		h.C = W + size_t(o + n1) * ld + o;
		h.lda = ld;
		h.ldb = size_t(n1);
		h.ldc = ld;
		h.M = n2;
		h.N = n1;
		h.K = n2;
		h.alpha = -1.0;
		h.beta = 0.0;
		h.tri = gemm::A_LOWER;
		run_gemm(ctx, true, h);
	}
};
} // namespace

void chol_setup_attributes()
{
	static bool done = false;
	if (done)
	{
		return;
	}
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::DefaultConfig, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::DefaultConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::DefaultConfig, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::DefaultConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::SmallConfig, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::SmallConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::SmallConfig, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::SmallConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LEAF_SMEM)));
	done = true;
}

void gemm_nt(gple_ctx* ctx, const gemm::GemmArgs& a)
{
	chol_setup_attributes();
	run_gemm(ctx, false, a);
}
void gemm_nn(gple_ctx* ctx, const gemm::GemmArgs& a)
{
	chol_setup_attributes();
	run_gemm(ctx, true, a);
}

void potrf_trtri(gple_ctx* ctx, double* A, double* W, const int n, int* d_info)
{
	chol_setup_attributes();
	const size_t ld = size_t(n);
	double* dinv = ctx->ws.get<double>("chol.dinv", size_t(n / LEAF) * LEAF * LEAF);
	double* T = ctx->ws.get<double>("chol.T", size_t(n / 2 + LEAF) * size_t(n / 2 + LEAF));
	GPLE_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), ctx->stream));
	const Chol c{ctx, A, ld, dinv, d_info};
	c.potrf(0, n);
	if (n > LEAF)
	{
		GPLE_LAUNCH(ctx, zero_upper_blocks_kernel, dim3(n / LEAF, n / LEAF), 256, 0, A, ld, n);
	}
	if (W != nullptr)
	{
		GPLE_CUDA(cudaMemsetAsync(W, 0, size_t(n) * n * sizeof(double), ctx->stream));
		GPLE_LAUNCH(ctx, copy_dinv_kernel, n / LEAF, 256, 0, dinv, W, ld);
		c.trtri(W, T, 0, n);
	}
}

} // namespace gple
