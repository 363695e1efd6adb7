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
constexpr int LEAF_THREADS = 256; // 8 warps: 0-3 factorise, 4-7 build the inverse underneath
constexpr int FW = 8;			 // panel width = DMMA tile = granularity of the inverse
constexpr int SPW = 32;			 // super-panel: trailing matrix updated once per 32 columns (rank-32 DMMA update)
constexpr int TWS = 8 * 10;		 // per-warp 8 x 8 transposition scratch, pitch 10
// S[128][LP] | Ld[16][8][8] factorised diagonal blocks | rsd[128] reciprocal pivots | tws[8 warps][4 tiles][TWS]
constexpr size_t LEAF_SMEM = (size_t(LEAF) * LP + (LEAF / FW) * FW * FW + LEAF + 8 * 4 * TWS) * sizeof(double);

/// 1 / sqrt(d) for the pivots: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~22 bits) + two Newton steps in FP64 (error ~1 ulp):
/// 90 cycles of dependent latency against ~125 for the float seed with its two conversions and far more for the library
/// rsqrt(double) (profiles/r02_leaf_latency.md); outside the normal range it falls back to the library.
__device__ __forceinline__ double fast_rsqrt(const double d)
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

__device__ __forceinline__ void named_barrier(const int id, const int count)
{
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

/// X_ii = L_ii^-1 of one 8 x 8 diagonal block by forward substitution, lane b (< 8) owning column b, column-oriented (once
/// x[a] is final every later row gets its fma: 8 x (mul, independent fmas)); the reciprocal pivots come from the factorisation.
/// Stored with the rest of X: X[a][b] (a >= b) = S[b][a + 1].  The stores are UNCONDITIONAL: a lane-dependent store is a
/// divergent branch (1475 cycles for eight of them, profiles/r02_leaf_latency.md); the zeros of a < b land in the lower
/// triangle of the diagonal tile of S, which nobody reads during the sweep (the factor is parked in Ld and copied back at the end).
__device__ __forceinline__ void inverse_diagonal_block(double* __restrict__ S, const double* __restrict__ Lb, const double* __restrict__ rs, const int i, const int b)
{
	double x[FW];
#pragma unroll
	for (int a = 0; a < FW; a++)
	{
		x[a] = a == b ? 1.0 : 0.0; // right-hand side e_b
	}
#pragma unroll
	for (int a = 0; a < FW; a++)
	{
		x[a] *= rs[a];
#pragma unroll
		for (int k = a + 1; k < FW; k++)
		{
			x[k] = fma(-Lb[k * FW + a], x[a], x[k]);
		}
	}
#pragma unroll
	for (int a = 0; a < FW; a++)
	{
		S[(i * FW + b) * LP + i * FW + a + 1] = x[a];
	}
}

/// Rows 8 i .. 8 i + 7 of X = L^-1 below the diagonal block (which inverse_diagonal_block has stored already):
///     X_ij = -X_ii sum_{k = j}^{i - 1} L_ik X_kj   (j < i)
/// on DMMA.  Warp w of nw takes the column tiles j = w, w + nw, ... (at most four), all in flight together: the L_ik
/// fragment of a k-step is loaded once and feeds every tile that has reached k; the two k-halves of a step accumulate
/// separately, so a warp runs up to eight independent DMMA chains and the tensor pipe, not the chain latency, is the limit.
/// L lives in the lower triangle of S; X is kept TRANSPOSED in the strictly-upper part, shifted by one column.
__device__ __forceinline__ void inverse_row_block(double* __restrict__ S, double* __restrict__ ws, const int i, const int w, const int nw, const int lane)
{
	const int g = lane >> 2, t = lane & 3;
	constexpr int MAXT = 4;
	if (w >= i)
	{
		return;
	}
	double acc[MAXT][2][2];
#pragma unroll
	for (int u = 0; u < MAXT; u++)
	{
		acc[u][0][0] = acc[u][0][1] = acc[u][1][0] = acc[u][1][1] = 0.0;
	}
	const double* __restrict__ Lrow = S + (i * FW + g) * LP + t; // L_ik[g][kk * 4 + t] at Lrow[k * 8 + kk * 4]
	for (int k = w; k < i; k++)
	{
		const double a0 = Lrow[k * FW], a1 = Lrow[k * FW + 4];
#pragma unroll
		for (int u = 0; u < MAXT; u++)
		{
			const int j = w + u * nw;
			if (j <= k) // warp-uniform; j < i follows from k < i
			{
				const double* __restrict__ Xc = S + (j * FW + g) * LP + k * FW + t + 1; // X_kj[kk * 4 + t][g]
				double b0 = Xc[0], b1 = Xc[4];
				if (k == j)
				{
					// X_jj is lower triangular (the slots above hold entries of L)
					b0 = t < g ? 0.0 : b0;
					b1 = t + 4 < g ? 0.0 : b1;
				}
				gemm::dmma884(acc[u][0], a0, b0);
				gemm::dmma884(acc[u][1], a1, b1);
			}
		}
	}
	// T_j (accumulator layout C[g][2t .. 2t + 1]) -> scratch, re-read as the B operand T[tp][g] of  R = X_ii T_j
	__syncwarp();
#pragma unroll
	for (int u = 0; u < MAXT; u++)
	{
		if (w + u * nw < i)
		{
			ws[u * TWS + g * 10 + 2 * t] = acc[u][0][0] + acc[u][1][0];
			ws[u * TWS + g * 10 + 2 * t + 1] = acc[u][0][1] + acc[u][1][1];
		}
	}
	__syncwarp();
	const double xa0 = g >= t ? S[(i * FW + t) * LP + i * FW + g + 1] : 0.0;		 // X_ii[g][t]
	const double xa1 = g >= t + 4 ? S[(i * FW + t + 4) * LP + i * FW + g + 1] : 0.0; // X_ii[g][t + 4]
#pragma unroll
	for (int u = 0; u < MAXT; u++)
	{
		const int j = w + u * nw;
		if (j < i)
		{
			double r[2] = {0.0, 0.0};
			gemm::dmma884(r, xa0, ws[u * TWS + t * 10 + g]);
			gemm::dmma884(r, xa1, ws[u * TWS + (t + 4) * 10 + g]);
			// X_ij[g][2t + u'] -> S[8 j + 2t + u'][8 i + g + 1]
			S[(j * FW + 2 * t) * LP + i * FW + g + 1] = -r[0];
			S[(j * FW + 2 * t + 1) * LP + i * FW + g + 1] = -r[1];
		}
	}
}

/// Factor one 128 x 128 diagonal block and invert the resulting triangle, entirely in shared memory, in one CTA.
///   factor : right-looking over 8-wide panels inside 32-wide super-panels.  Per panel, phase 1: the 8 x 8 diagonal block is
///            factorised REDUNDANTLY in the registers of every thread that owns a row below it (no shuffles, no barrier between
///            block and row solve: the serial chain per column is rsqrt -> scale -> fma), then the thread solves its own row;
///            phase 2: all warps apply the rank-8 update on DMMA, but only to the rest of the 32-wide super-panel; the trailing
///            matrix beyond it gets ONE rank-32 update per super-panel (16 x 16 warp tiles, 8 k-steps: tensor-pipe bound
///            instead of the latency-bound rank-8 sweeps over the whole trailing triangle).
///   inverse: hidden underneath.  Phase 1 keeps at most 120 threads busy, so warps 4-7 build rows 8 (p - 1) .. of X = L^-1
///            (inverse_row_block) while warps 0-3 factorise panel p; only the last row block is left for after the loop.
/// The factorised diagonal blocks are parked in Ld (nobody reads them from S during the sweep) and copied back at the end,
/// so that no thread can overwrite a diagonal block another thread is still loading.
/// L overwrites the block (upper part zeroed); its inverse goes to `dinv` (row-major 128 x 128, upper part zero).
template <bool TIMED>
__global__ void __launch_bounds__(LEAF_THREADS, 1) potrf_leaf_kernel(double* __restrict__ A, const size_t ld, double* __restrict__ dinv, int* __restrict__ info, const int global_row0, long long* __restrict__ ticks)
{
	extern __shared__ __align__(16) double sm[];
	double* S = sm;						   // [128][LP]
	double* Ld = sm + LEAF * LP;		   // [16][8][8]
	double* rsd = Ld + (LEAF / FW) * FW * FW; // [128]
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
	double* ws = rsd + LEAF + warp * 4 * TWS;
	// TIMED (profiles/tools/microbench_fp64.cu only): clock64 at the phase boundaries, as seen by thread 0
	long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	auto stamp = [&](const int i)
	{
		if (TIMED)
		{
			tk[i] = clock64();
		}
	};
	stamp(0);
	// lower triangle -> shared memory: every 16-byte chunk in flight at once (LDGSTS)
	for (int e = tid; e < LEAF * LEAF / 2; e += LEAF_THREADS)
	{
		const int r = e >> 6, c2 = (e & 63) * 2;
		if (c2 <= r)
		{
			gemm::cp_async16(S + r * LP + c2, A + size_t(r) * ld + c2);
		}
	}
	gemm::cp_async_commit();
	gemm::cp_async_wait<0>();
	__syncthreads();
	stamp(1);

	long long part[5] = {0, 0, 0, 0, 0}, c_start = TIMED ? clock64() : 0;
	for (int p = 0; p < LEAF / FW; p++)
	{
		const int c0 = p * FW;
		const int rows_below = LEAF - c0 - FW; // <= 120: thread 127 never owns a row and keeps the block for everybody
		// ------------------------------------------------------------ phase 1
		if (warp < 4)
		{
			// threads 120 .. 127 never own a row (rows_below <= 120): they factorise the block as well and each of them
			// inverts one column of it (X_pp), which the inverse warps need one panel later; 127 also parks the block in Ld
			if (tid < rows_below || tid >= 120)
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
						if (!bad && tid == 127)
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
				if (tid >= 120)
				{
					if (tid == 127)
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
					inverse_diagonal_block(S, &l[0][0], rs, p, tid - 120);
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
		}
		else if (p > 0)
		{
			inverse_row_block(S, ws, p - 1, warp - 4, 4, lane);
		}
		long long c1 = 0, c2 = 0, c3 = 0, c4 = 0;
		if (TIMED)
		{
			c1 = clock64();
			part[0] += c1 - c_start; // thread 0's own phase-1 work
			if (__syncthreads_count(1) < 0) // a barrier whose result is consumed: the clock below is read after it completed
			{
				c1 = 0;
			}
		}
		else
		{
			__syncthreads();
		}
		if (TIMED)
		{
			c2 = clock64();
			part[1] += c2 - c1; // waiting for the slowest thread of phase 1 (the inverse group)
		}
		// ------------------------------------------------------------ phase 2
		const int q = p & 3, s0 = c0 - q * FW, t0 = c0 + FW;
		{
			// rank-8 update of the rest of the super-panel: rows >= t0, columns [t0, s0 + 32), lower tiles only.  A warp owns the
			// tile rows `warp` and `warp + 8` (there are at most 15); the up to three 8 x 8 tiles of a row share the A fragment,
			// and all (up to six) load -> 2 DMMA -> store chains of the warp are in flight together
			const int m = (LEAF - t0) / FW, nc = min(3 - q, m);
			if (nc > 0 && warp < m)
			{
				int r0[2], cnt[2];
				double a[2][2];
				double2 cv[2][3];
#pragma unroll
				for (int h = 0; h < 2; h++)
				{
					const int ti = warp + 8 * h;
					cnt[h] = ti < m ? min(ti + 1, nc) : 0;
					r0[h] = t0 + (ti < m ? ti : warp) * FW;
#pragma unroll
					for (int kk = 0; kk < 2; kk++)
					{
						a[h][kk] = -S[(r0[h] + g) * LP + c0 + kk * 4 + t];
					}
#pragma unroll
					for (int tj = 0; tj < 3; tj++)
					{
						cv[h][tj] = tj < cnt[h] ? *reinterpret_cast<const double2*>(S + (r0[h] + g) * LP + t0 + tj * FW + 2 * t) : make_double2(0.0, 0.0);
					}
				}
				double b[3][2];
#pragma unroll
				for (int tj = 0; tj < 3; tj++)
				{
#pragma unroll
					for (int kk = 0; kk < 2; kk++)
					{
						b[tj][kk] = tj < nc ? S[(t0 + tj * FW + g) * LP + c0 + kk * 4 + t] : 0.0;
					}
				}
#pragma unroll
				for (int kk = 0; kk < 2; kk++)
				{
#pragma unroll
					for (int h = 0; h < 2; h++)
					{
#pragma unroll
						for (int tj = 0; tj < 3; tj++)
						{
							if (tj < cnt[h])
							{
								double c[2] = {cv[h][tj].x, cv[h][tj].y};
								gemm::dmma884(c, a[h][kk], b[tj][kk]);
								cv[h][tj] = make_double2(c[0], c[1]);
							}
						}
					}
				}
#pragma unroll
				for (int h = 0; h < 2; h++)
				{
#pragma unroll
					for (int tj = 0; tj < 3; tj++)
					{
						if (tj < cnt[h])
						{
							*reinterpret_cast<double2*>(S + (r0[h] + g) * LP + t0 + tj * FW + 2 * t) = cv[h][tj];
						}
					}
				}
			}
		}
		if (TIMED)
		{
			c3 = clock64();
			part[2] += c3 - c2; // rank-8 update inside the super-panel
		}
		if (q == 3 && s0 + SPW < LEAF)
		{
			// the super-panel is complete: rank-32 update of everything beyond it, 16 x 16 warp tiles of the lower triangle
			__syncthreads();
			const int u0 = s0 + SPW, m2 = (LEAF - u0) / 16, ntiles = m2 * (m2 + 1) / 2;
			for (int idx = warp; idx < ntiles; idx += 8)
			{
				int ti = 0, tj = idx;
				while (tj > ti)
				{
					tj -= ti + 1;
					ti++;
				}
				const int r0 = u0 + 16 * ti, q0 = u0 + 16 * tj;
				const bool diag = ti == tj;
				double acc[2][2][2];
#pragma unroll
				for (int mi = 0; mi < 2; mi++)
				{
#pragma unroll
					for (int nj = 0; nj < 2; nj++)
					{
						const double2 v = *reinterpret_cast<const double2*>(S + (r0 + 8 * mi + g) * LP + q0 + 8 * nj + 2 * t);
						acc[mi][nj][0] = v.x;
						acc[mi][nj][1] = v.y;
					}
				}
#pragma unroll
				for (int kk = 0; kk < SPW / 4; kk++)
				{
					double a[2], b[2];
#pragma unroll
					for (int mi = 0; mi < 2; mi++)
					{
						a[mi] = -S[(r0 + 8 * mi + g) * LP + s0 + kk * 4 + t];
						b[mi] = S[(q0 + 8 * mi + g) * LP + s0 + kk * 4 + t];
					}
					gemm::dmma884(acc[0][0], a[0], b[0]);
					gemm::dmma884(acc[1][0], a[1], b[0]);
					gemm::dmma884(acc[1][1], a[1], b[1]);
					if (!diag)
					{
						gemm::dmma884(acc[0][1], a[0], b[1]);
					}
				}
#pragma unroll
				for (int mi = 0; mi < 2; mi++)
				{
#pragma unroll
					for (int nj = 0; nj < 2; nj++)
					{
						if (!(diag && mi == 0 && nj == 1))
						{
							*reinterpret_cast<double2*>(S + (r0 + 8 * mi + g) * LP + q0 + 8 * nj + 2 * t) = make_double2(acc[mi][nj][0], acc[mi][nj][1]);
						}
					}
				}
			}
		}
		if (TIMED)
		{
			c4 = clock64();
			part[3] += c4 - c3; // rank-32 update
		}
		if (TIMED)
		{
			if (__syncthreads_count(1) < 0)
			{
				c4 = 0;
			}
			c_start = clock64();
			part[4] += c_start - c4; // barrier at the end of the panel
		}
		else
		{
			__syncthreads();
		}
	}
	stamp(2);
	// last row block of the inverse, all warps
	inverse_row_block(S, ws, LEAF / FW - 1, warp, 8, lane);
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
	stamp(3);
	stamp(4);
	// L (upper part zero) and X = L^-1 (transposed + shifted in S) out, 16 bytes per store
	for (int e = tid; e < LEAF * LEAF / 2; e += LEAF_THREADS)
	{
		const int r = e >> 6, c = (e & 63) * 2;
		double2 l = make_double2(0.0, 0.0), x = make_double2(0.0, 0.0);
		if (c <= r)
		{
			l.x = S[r * LP + c];
			x.x = S[c * LP + r + 1];
		}
		if (c + 1 <= r)
		{
			l.y = S[r * LP + c + 1];
			x.y = S[(c + 1) * LP + r + 1];
		}
		*reinterpret_cast<double2*>(A + size_t(r) * ld + c) = l;
		*reinterpret_cast<double2*>(dinv + r * LEAF + c) = x;
	}
	if (TIMED)
	{
		stamp(5);
		if (tid == 0)
		{
			// load, panel loop, last inverse row block + copy-back, store, total; then the loop split into its five parts
			ticks[0] = tk[1] - tk[0];
			ticks[1] = tk[2] - tk[1];
			ticks[2] = tk[3] - tk[2];
			ticks[3] = tk[5] - tk[4];
			ticks[4] = tk[5] - tk[0];
			for (int i = 0; i < 5; i++)
			{
				ticks[5 + i] = part[i];
			}
		}
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
		GPLE_LAUNCH(ctx, potrf_leaf_kernel<false>, 1, LEAF_THREADS, LEAF_SMEM, at(o, o), ld, dinv + size_t(o / LEAF) * LEAF * LEAF, info, o, static_cast<long long*>(nullptr));
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
	GPLE_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LEAF_SMEM)));
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
	auto enqueue = [&]()
	{
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
	};
	// CUDA graph: at the sizes where the factorisation is a chain of short launches (a 128-leaf is 39 us, the GEMMs between two
	// leaves 10-30 us, the look-ahead crosses streams through events) the whole schedule is captured once per (size, buffers) --
	// the workspace and the model pool hand out the same buffers evaluation after evaluation -- and replayed with one launch.
	if (!ctx->factorise_graphs || n > 8192 || n <= LEAF || ctx->own_stream == nullptr)
	{
		enqueue();
		return;
	}
	gple_ctx::FactoriseGraph* hit = nullptr;
	for (auto& g : ctx->factorise_graph_cache)
	{
		if (g.n == n && g.A == A && g.W == W && g.info == d_info && g.dinv == dinv && g.T == T && g.potrf_flat == g_potrf_flat)
		{
			hit = &g;
			break;
		}
	}
	cudaStream_t user = ctx->stream, own = ctx->own_stream;
	if (hit == nullptr)
	{
		// capture on the context's own stream (the caller's may be the legacy default stream, which cannot capture); relaxed
		// mode: other host threads of this process keep calling the runtime for their own contexts
		// One capture at a time per process: it happens once per (size, buffers), and tools that interpose the runtime (ncu) do not
		// survive simultaneous captures from several host threads.
		static std::mutex capture_mutex;
		const std::lock_guard<std::mutex> capture_lock(capture_mutex);
		const unsigned long long before = ctx->launches;
		cudaGraph_t graph = nullptr;
		if (cudaStreamBeginCapture(own, cudaStreamCaptureModeRelaxed) != cudaSuccess)
		{
			cudaGetLastError();
			enqueue();
			return;
		}
		ctx->stream = own;
		try
		{
			enqueue();
		}
		catch (...)
		{
			ctx->stream = user;
			cudaStreamEndCapture(own, &graph);
			if (graph != nullptr)
			{
				cudaGraphDestroy(graph);
			}
			throw;
		}
		ctx->stream = user;
		GPLE_CUDA(cudaStreamEndCapture(own, &graph));
		gple_ctx::FactoriseGraph g;
		g.n = n;
		g.A = A;
		g.W = W;
		g.info = d_info;
		g.dinv = dinv;
		g.T = T;
		g.potrf_flat = g_potrf_flat;
		g.launches = ctx->launches - before;
		const cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
		cudaGraphDestroy(graph);
		if (e != cudaSuccess)
		{
			cudaGetLastError();
			ctx->launches = before;
			enqueue();
			return;
		}
		ctx->launches = before;
		if (ctx->factorise_graph_cache.size() >= 12) // least used out
		{
			size_t worst = 0;
			for (size_t i = 1; i < ctx->factorise_graph_cache.size(); i++)
			{
				if (ctx->factorise_graph_cache[i].uses < ctx->factorise_graph_cache[worst].uses)
				{
					worst = i;
				}
			}
			cudaGraphExecDestroy(ctx->factorise_graph_cache[worst].exec);
			ctx->factorise_graph_cache.erase(ctx->factorise_graph_cache.begin() + long(worst));
		}
		ctx->factorise_graph_cache.push_back(g);
		hit = &ctx->factorise_graph_cache.back();
		ctx->graph_captures++;
	}
	if (user != own)
	{
		GPLE_CUDA(cudaEventRecord(ctx->ev_graph, user));
		GPLE_CUDA(cudaStreamWaitEvent(own, ctx->ev_graph, 0));
	}
	GPLE_CUDA(cudaGraphLaunch(hit->exec, own));
	if (user != own)
	{
		GPLE_CUDA(cudaEventRecord(ctx->ev_graph, own));
		GPLE_CUDA(cudaStreamWaitEvent(user, ctx->ev_graph, 0));
	}
	hit->uses++;
	ctx->launches += hit->launches;
	ctx->graph_replays++;
}

} // namespace gple
