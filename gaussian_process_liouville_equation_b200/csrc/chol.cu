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
constexpr int LP = 132;			 // pitch (doubles): 4 (mod 16) -> conflict-free DMMA fragment loads
constexpr int LEAF_THREADS = 256; // 8 warps
constexpr int PW = 16;			 // panel width / inverse block size
constexpr size_t LEAF_SMEM = (size_t(LEAF) * LP + 8 * 16 * 20) * sizeof(double);

/// 1 / sqrt(d) for the pivots: single-precision MUFU seed + two Newton steps in FP64 (error ~1 ulp), a much shorter
/// dependent chain than the library rsqrt(double); outside the float range it falls back to the library.
__device__ __forceinline__ double fast_rsqrt(const double d)
{
	if (!(d > 1e-30 && d < 1e30))
	{
		return rsqrt(d);
	}
	double y = double(rsqrtf(float(d)));
	const double h = 0.5 * d;
	y = y * fma(-h * y, y, 1.5);
	y = y * fma(-h * y, y, 1.5);
	return y;
}

/// Factor one 128 x 128 diagonal block and invert the resulting triangle, entirely in shared memory.
///   factor : right-looking with 8-wide panels -- the 8 x 8 diagonal block factorised redundantly in the registers of
///            every thread that owns a panel row (no shuffles, no barrier between block and panel), one row-wise
///            triangular solve per thread, and the rank-8 trailing update of the lower triangle by all warps on DMMA;
///   inverse: X = L^-1 by 16 x 16 blocks -- diagonal blocks by forward substitution (16 lanes each), then block
///            sub-diagonal after sub-diagonal  X_ij = -X_ii sum_{k=j}^{i-1} L_ik X_kj  on DMMA, one warp per block.
/// L lives in the lower triangle of S (row-major, pitch LP); X is kept TRANSPOSED in the strictly-upper part,
/// shifted by one column: X[a][b] (a >= b) = S[b][a + 1].
/// L overwrites the block (upper part zeroed); its inverse goes to `dinv` (row-major 128 x 128, upper part zero).
__global__ void __launch_bounds__(LEAF_THREADS, 1) potrf_leaf_kernel(double* __restrict__ A, const size_t ld, double* __restrict__ dinv, int* __restrict__ info, const int global_row0)
{
	extern __shared__ __align__(16) double sm[];
	double* S = sm;					 // [128][LP]
	double* scratch = sm + LEAF * LP; // [8 warps][16][20]
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
	for (int e = tid; e < LEAF * LEAF / 2; e += LEAF_THREADS)
	{
		const int r = e >> 6, c2 = (e & 63) * 2;
		if (c2 <= r)
		{
			*reinterpret_cast<double2*>(S + r * LP + c2) = *reinterpret_cast<const double2*>(A + size_t(r) * ld + c2);
		}
	}
	__syncthreads();

	// ---------------------------------------------------------------- factor
	// Right-looking with 8-wide panels.  The 8 x 8 diagonal block is factorised REDUNDANTLY by every thread that needs it,
	// entirely in registers (36 doubles, no shuffles, no barrier): the serial chain per column is rsqrt -> scale -> fma.
	// The same thread then solves its own panel row against its register copy of L_d (8-step substitution, the pivots'
	// reciprocals come for free from the factorisation), and all warps apply the rank-8 trailing update on DMMA.
	// The factorised diagonal blocks are parked in `scratch` (nobody reads them during the sweep) and copied back at the end,
	// so that no thread can overwrite a diagonal block another thread is still loading.
	constexpr int FW = 8;
	double* Ld = scratch; // [16 panels][8][8]
	for (int p = 0; p < LEAF / FW; p++)
	{
		const int c0 = p * FW;
		const int rows_below = LEAF - c0 - FW;
		if (tid < rows_below || tid == LEAF_THREADS - 1)
		{
			double l[FW][FW], rs[FW];
#pragma unroll
			for (int i = 0; i < FW; i++)
			{
#pragma unroll
				for (int k = 0; k <= i; k++)
				{
					l[i][k] = S[(c0 + i) * LP + c0 + k];
				}
			}
			bool bad = false;
#pragma unroll
			for (int j = 0; j < FW; j++)
			{
				double d = l[j][j];
				if (!(d > 0.0))
				{
					if (!bad && tid == LEAF_THREADS - 1)
					{
						atomicCAS(info, 0, global_row0 + c0 + j + 1);
					}
					bad = true;
					d = 1.0;
				}
				rs[j] = fast_rsqrt(d);
				l[j][j] = d * rs[j];
#pragma unroll
				for (int i = j + 1; i < FW; i++)
				{
					l[i][j] *= rs[j];
				}
#pragma unroll
				for (int k = j + 1; k < FW; k++)
				{
#pragma unroll
					for (int i = k; i < FW; i++)
					{
						l[i][k] = fma(-l[i][j], l[k][j], l[i][k]);
					}
				}
			}
			if (tid == LEAF_THREADS - 1)
			{
#pragma unroll
				for (int i = 0; i < FW; i++)
				{
#pragma unroll
					for (int k = 0; k <= i; k++)
					{
						Ld[p * FW * FW + i * FW + k] = l[i][k];
					}
				}
			}
			else
			{
				// panel row r: x L_d^T = a
				const int r = c0 + FW + tid;
				double x[FW];
#pragma unroll
				for (int k = 0; k < FW; k += 2)
				{
					const double2 v = *reinterpret_cast<const double2*>(S + r * LP + c0 + k);
					x[k] = v.x;
					x[k + 1] = v.y;
				}
#pragma unroll
				for (int j = 0; j < FW; j++)
				{
					double v = x[j];
#pragma unroll
					for (int k = 0; k < j; k++)
					{
						v = fma(-x[k], l[j][k], v);
					}
					x[j] = v * rs[j];
				}
#pragma unroll
				for (int k = 0; k < FW; k += 2)
				{
					*reinterpret_cast<double2*>(S + r * LP + c0 + k) = make_double2(x[k], x[k + 1]);
				}
			}
		}
		__syncthreads();
		// trailing update of the lower triangle: A22 -= P P^T (k = 8), 8 x 8 tiles dealt round-robin to the warps,
		// two tiles in flight per warp (the load -> DMMA -> store chain of one tile is pure latency)
		{
			const int t0 = c0 + FW;
			const int m = (LEAF - t0) / 8;
			const int ntiles = m * (m + 1) / 2;
			constexpr int NW = LEAF_THREADS / 32;
			// tile idx -> (ti, tj), tj <= ti, advanced incrementally
			int ti = 0, tj = warp;
			auto normalise = [&]()
			{
				while (tj > ti)
				{
					tj -= ti + 1;
					ti++;
				}
			};
			normalise();
			for (int idx = warp; idx < ntiles; idx += 2 * NW)
			{
				const int r0 = t0 + ti * 8, q0 = t0 + tj * 8;
				tj += NW;
				normalise();
				const bool second = idx + NW < ntiles;
				const int r1 = second ? t0 + ti * 8 : r0, q1 = second ? t0 + tj * 8 : q0;
				tj += NW;
				normalise();
				double2* cp0 = reinterpret_cast<double2*>(S + (r0 + g) * LP + q0 + 2 * t);
				double2* cp1 = reinterpret_cast<double2*>(S + (r1 + g) * LP + q1 + 2 * t);
				const double2 cv0 = *cp0, cv1 = *cp1;
				double c0v[2] = {-cv0.x, -cv0.y}, c1v[2] = {-cv1.x, -cv1.y}; // accumulate -(A22) + P P^T, negate back on store
#pragma unroll
				for (int kk = 0; kk < FW / 4; kk++)
				{
					const double a0 = S[(r0 + g) * LP + c0 + kk * 4 + t], b0 = S[(q0 + g) * LP + c0 + kk * 4 + t];
					const double a1 = S[(r1 + g) * LP + c0 + kk * 4 + t], b1 = S[(q1 + g) * LP + c0 + kk * 4 + t];
					gemm::dmma884(c0v, a0, b0);
					gemm::dmma884(c1v, a1, b1);
				}
				*cp0 = make_double2(-c0v[0], -c0v[1]);
				if (second)
				{
					*cp1 = make_double2(-c1v[0], -c1v[1]);
				}
			}
		}
		__syncthreads();
	}
	// factorised diagonal blocks back into place
	for (int e = tid; e < (LEAF / FW) * FW * FW; e += LEAF_THREADS)
	{
		const int p = e / (FW * FW), i = (e / FW) % FW, k = e % FW;
		if (k <= i)
		{
			S[(p * FW + i) * LP + p * FW + k] = Ld[e];
		}
	}
	__syncthreads();

	// ---------------------------------------------------------------- inverse
	// diagonal 16 x 16 blocks: block i by lanes 0..15 of warp i; lane b owns column b of X_ii
	if (lane < PW)
	{
		const int i0 = warp * PW, bcol = lane;
		double x[PW];
#pragma unroll
		for (int a = 0; a < PW; a++)
		{
			double v = (a == bcol) ? 1.0 : 0.0;
#pragma unroll
			for (int k = 0; k < a; k++)
			{
				v = fma(-S[(i0 + a) * LP + i0 + k], (k >= bcol) ? x[k] : 0.0, v);
			}
			x[a] = (a >= bcol) ? v / S[(i0 + a) * LP + i0 + a] : 0.0;
		}
#pragma unroll
		for (int a = 0; a < PW; a++)
		{
			if (a >= bcol)
			{
				S[(i0 + bcol) * LP + i0 + a + 1] = x[a]; // X[i0 + a][i0 + bcol], transposed + shifted
			}
		}
	}
	__syncthreads();
	// block sub-diagonals d = 1 .. 7: block (i, j) = (j + d, j) by warp j
	double* ws = scratch + warp * 16 * 20;
	for (int d = 1; d < LEAF / PW; d++)
	{
		const int j = warp, i = j + d;
		if (i < LEAF / PW)
		{
			// T = sum_{k = j}^{i - 1} L_ik X_kj   (16 x 16, as 2 x 2 DMMA tiles)
			double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
			for (int k = j; k < i; k++)
			{
#pragma unroll
				for (int kk = 0; kk < 4; kk++)
				{
					double av[2], bv[2];
#pragma unroll
					for (int mi = 0; mi < 2; mi++)
					{
						av[mi] = S[(i * PW + mi * 8 + g) * LP + k * PW + kk * 4 + t]; // L_ik[mi*8+g][kk*4+t]
					}
#pragma unroll
					for (int nj = 0; nj < 2; nj++)
					{
						// X_kj[a][b], a = kk*4+t, b = nj*8+g, stored at S[j*16 + b][k*16 + a + 1]; X_jj is lower triangular
						const int a = kk * 4 + t, b = nj * 8 + g;
						const double v = S[(j * PW + b) * LP + k * PW + a + 1];
						bv[nj] = (k > j || a >= b) ? v : 0.0;
					}
#pragma unroll
					for (int mi = 0; mi < 2; mi++)
					{
#pragma unroll
						for (int nj = 0; nj < 2; nj++)
						{
							gemm::dmma884(acc[mi][nj], av[mi], bv[nj]);
						}
					}
				}
			}
			// T -> scratch (row-major 16 x 20) so that it can be re-read as a B operand
#pragma unroll
			for (int mi = 0; mi < 2; mi++)
			{
#pragma unroll
				for (int nj = 0; nj < 2; nj++)
				{
					*reinterpret_cast<double2*>(ws + (mi * 8 + g) * 20 + nj * 8 + 2 * t) = make_double2(acc[mi][nj][0], acc[mi][nj][1]);
				}
			}
			__syncwarp();
			// R = X_ii T ; X_ij = -R
			double r[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
			for (int kk = 0; kk < 4; kk++)
			{
				double av[2], bv[2];
#pragma unroll
				for (int mi = 0; mi < 2; mi++)
				{
					// X_ii[a][b], a = mi*8+g, b = kk*4+t, stored at S[i*16 + b][i*16 + a + 1]; lower triangular
					const int a = mi * 8 + g, b = kk * 4 + t;
					const double v = S[(i * PW + b) * LP + i * PW + a + 1];
					av[mi] = (a >= b) ? v : 0.0;
				}
#pragma unroll
				for (int nj = 0; nj < 2; nj++)
				{
					bv[nj] = ws[(kk * 4 + t) * 20 + nj * 8 + g]; // T[kk*4+t][nj*8+g]
				}
#pragma unroll
				for (int mi = 0; mi < 2; mi++)
				{
#pragma unroll
					for (int nj = 0; nj < 2; nj++)
					{
						gemm::dmma884(r[mi][nj], av[mi], bv[nj]);
					}
				}
			}
			// X_ij[a][b] -> S[j*16 + b][i*16 + a + 1]
#pragma unroll
			for (int mi = 0; mi < 2; mi++)
			{
#pragma unroll
				for (int nj = 0; nj < 2; nj++)
				{
#pragma unroll
					for (int u = 0; u < 2; u++)
					{
						S[(j * PW + nj * 8 + 2 * t + u) * LP + i * PW + mi * 8 + g + 1] = -r[mi][nj][u];
					}
				}
			}
		}
		__syncthreads();
	}
	for (int e = tid; e < LEAF * LEAF; e += LEAF_THREADS)
	{
		const int r = e >> 7, c = e & 127;
		double l = 0.0, x = 0.0;
		if (c <= r)
		{
			l = S[r * LP + c];
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
	const unsigned nz = unsigned(a.batch > 1 ? a.batch : 1);
	if (a.in_place && !b_nn && a.N == 128 && a.M / 128 * 2 < ctx->num_sms)
	{
		// panel solve with few row blocks: 32-row strips quarter the per-CTA latency (a 128 x 128 x 128 tile is 17 us of DMMA)
		using C = gemm::StripConfig;
		GPLE_LAUNCH(ctx, (gemm::gemm_kernel<C, false>), dim3(1, a.M / C::BM, nz), C::THREADS, C::SMEM_BYTES, a);
		return;
	}
	const long long tiles128 = (long long)(a.N / 128) * (a.M / 128) / (a.lower_only ? 2 : 1) * nz;
	if (tiles128 * 2 >= ctx->num_sms || a.in_place)
	{
		using C = gemm::DefaultConfig;
		const dim3 grid(a.N / C::BN, a.M / C::BM, nz);
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
		const dim3 grid(a.N / C::BN, a.M / C::BM, nz);
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

/// potrf switches from the recursive form to the right-looking sweep at or below this block size (gple_set_potrf_flat)
int g_potrf_flat = 4096; // profiles/r01_tune_potrf.md

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

	void leaf(const int o) const
	{
		GPLE_LAUNCH(ctx, potrf_leaf_kernel, 1, LEAF_THREADS, LEAF_SMEM, at(o, o), ld, dinv + size_t(o / LEAF) * LEAF * LEAF, info, o);
	}

	/// rank-128 update C -= P P^T of the lower 128-tiles of a trailing block (C: m x nc at (r, c); P: the panel rows r.. and c..)
	void syrk(const int r, const int c, const int m, const int nc, const int pcol) const
	{
		gemm::GemmArgs g{};
		g.A = at(r, pcol);
		g.B = at(c, pcol);
		g.C = at(r, c);
		g.lda = g.ldb = g.ldc = ld;
		g.M = m;
		g.N = nc;
		g.K = LEAF;
		g.alpha = -1.0;
		g.beta = 1.0;
		g.lower_only = (r == c && m == nc) ? 1 : 0;
		run_gemm(ctx, false, g);
	}

	/// Right-looking sweep over 128-blocks (every GEMM has K = 128): the latency-optimal order for blocks of up to a few
	/// thousand rows, where the recursive form leaves most SMs idle behind a few long-K tiles.  With LOOK-AHEAD: after the
	/// panel solve of step k only block column k + 1 is updated on the main stream (that is all leaf k + 1 and its panel
	/// solve need); the bulk of the trailing update (columns k + 2 ...) runs on the context's auxiliary stream underneath
	/// the next leaf, which is a single-CTA kernel.  Ordering: bulk(k) waits for the panel of step k; the column update of
	/// step k + 1 waits for bulk(k) (both write column k + 2).
	void potrf_flat(const int o, const int n) const
	{
		const bool lookahead = n >= 4 * LEAF && ctx->aux_stream != nullptr;
		cudaStream_t main = ctx->stream;
		bool bulk_pending = false;
		for (int k = 0; k < n; k += LEAF)
		{
			leaf(o + k);
			const int rest = n - k - LEAF;
			if (rest <= 0)
			{
				break;
			}
			trsm(o + k + LEAF, rest, o + k, LEAF);
			if (!lookahead || rest == LEAF)
			{
				if (bulk_pending)
				{
					GPLE_CUDA(cudaStreamWaitEvent(main, ctx->ev_bulk, 0));
					bulk_pending = false;
				}
				syrk(o + k + LEAF, o + k + LEAF, rest, rest, o + k);
				continue;
			}
			GPLE_CUDA(cudaEventRecord(ctx->ev_panel, main));
			if (bulk_pending)
			{
				GPLE_CUDA(cudaStreamWaitEvent(main, ctx->ev_bulk, 0)); // bulk(k - 1) also wrote block column k + 1
			}
			syrk(o + k + LEAF, o + k + LEAF, rest, LEAF, o + k); // block column k + 1: all the next leaf and panel solve read
			ctx->stream = ctx->aux_stream;
			GPLE_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_panel, 0));
			syrk(o + k + 2 * LEAF, o + k + 2 * LEAF, rest - LEAF, rest - LEAF, o + k);
			GPLE_CUDA(cudaEventRecord(ctx->ev_bulk, ctx->aux_stream));
			ctx->stream = main;
			bulk_pending = true;
		}
		if (bulk_pending)
		{
			GPLE_CUDA(cudaStreamWaitEvent(main, ctx->ev_bulk, 0));
		}
	}

	void potrf(const int o, const int n) const
	{
		if (n == LEAF)
		{
			leaf(o);
			return;
		}
		if (n <= g_potrf_flat)
		{
			potrf_flat(o, n);
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

	/// W (same layout, zero-initialised with the inverted diagonal blocks in place) <- L^-1, bottom-up: at block size
	/// b = 128, 256, ... every pair of adjacent b-blocks [o, o + b), [o + b, o + b + n2) is combined,
	///     T = L21 W11 ;  W21 = -W22 T ,
	/// all pairs of a level in ONE batched launch per product (they are independent), so the whole inverse takes
	/// 2 log2(n / 128) launches that fill the chip instead of 2 (n / 128 - 1) small dependent ones.  A ragged last pair
	/// (n2 < b) gets its own launch.
	void trtri(double* W, double* T) const
	{
		const int n = int(ld);
		for (int b = LEAF; b < n; b *= 2)
		{
			const int full = n / (2 * b);						// pairs with two complete b-blocks
			const int tail = n - full * 2 * b;					// rows left over after them
			const int ragged_n2 = tail > b ? tail - b : 0;		// a last pair (b, ragged_n2) exists iff tail > b
			auto combine = [&](const int o, const int n2, const int batch)
			{
				// T (n2 x b) = L21 * W11   (NN; W11[k][j] == 0 for j > k)
				gemm::GemmArgs g{};
				g.A = at(o + b, o);
				g.B = W + size_t(o) * ld + o;
				g.C = T;
				g.lda = ld;
				g.ldb = ld;
				g.ldc = size_t(b);
				g.M = n2;
				g.N = b;
				g.K = b;
				g.alpha = 1.0;
				g.beta = 0.0;
				g.tri = gemm::B_LOWER_NN;
				g.batch = batch;
				g.strideA = g.strideB = size_t(2 * b) * ld + size_t(2 * b);
				g.strideC = size_t(b) * b;
				run_gemm(ctx, true, g);
				// W21 = -W22 * T            (NN; W22[i][k] == 0 for k > i)
				gemm::GemmArgs h{};
				h.A = W + size_t(o + b) * ld + o + b;
				h.B = T;
				h.C = W + size_t(o + b) * ld + o;
				h.lda = ld;
				h.ldb = size_t(b);
				h.ldc = ld;
				h.M = n2;
				h.N = b;
				h.K = n2;
				h.alpha = -1.0;
				h.beta = 0.0;
				h.tri = gemm::A_LOWER;
				h.batch = batch;
				h.strideA = h.strideC = size_t(2 * b) * ld + size_t(2 * b);
				h.strideB = size_t(b) * b;
				run_gemm(ctx, true, h);
			};
			if (full > 0)
			{
				combine(0, b, full);
			}
			if (ragged_n2 > 0)
			{
				combine(full * 2 * b, ragged_n2, 1);
			}
		}
	}
};
} // namespace

/// Opt-in shared-memory sizes are a per-DEVICE function attribute: gple_ctx_create calls this once for every context,
/// after cudaSetDevice, so that one process may hold contexts on several GPUs.
void chol_setup_attributes()
{
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::DefaultConfig, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::DefaultConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::DefaultConfig, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::DefaultConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::SmallConfig, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::SmallConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::SmallConfig, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::SmallConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(gemm::gemm_kernel<gemm::StripConfig, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm::StripConfig::SMEM_BYTES)));
	GPLE_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LEAF_SMEM)));
}

void gemm_nt(gple_ctx* ctx, const gemm::GemmArgs& a)
{
	run_gemm(ctx, false, a);
}
void gemm_nn(gple_ctx* ctx, const gemm::GemmArgs& a)
{
	run_gemm(ctx, true, a);
}

int set_potrf_flat(const int n)
{
	const int old = g_potrf_flat;
	g_potrf_flat = n;
	return old;
}

void potrf_trtri(gple_ctx* ctx, double* A, double* W, const int n, int* d_info)
{
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
		c.trtri(W, T);
	}
}

} // namespace gple
