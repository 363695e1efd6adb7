// Training and batched prediction of one density-matrix element on B200 (sm_100a).
//
// Reference path replaced (gple/ = gaussian_process_liouville_equation/ of the reference):
//   TrainingKernel          gple/kernel.cpp:244-479        PredictiveKernel         gple/kernel.cpp:481-544
//   TrainingComplexKernel   gple/complex_kernel.cpp:221-592 PredictiveComplexKernel gple/complex_kernel.cpp:594-670
//
// Design (not a port):
//  * K = L L^T by blocked DMMA Cholesky, W = L^-1 by blocked DMMA triangular inverse (chol.cu); K^-1 = W^T W
//    is never formed on the prediction path.  diag(K^-1) (needed by the LOOCV error, kernel.cpp:285) is the
//    column sum of squares of W; v = K^-1 y' = W^T (W y').
//  * The per-query variance k** - k K^-1 k^T (kernel.cpp:496-518: one GEMV over the N x N inverse PER
//    QUERY in the reference) becomes  k** - || W k^T ||^2 : one triangular DMMA GEMM  Z = K* W^T  over a
//    chunk of queries with the row sum of squares fused into the epilogue -- half the flops of K* K^-1 and
//    no cancellation inside the quadratic form.
//  * The widely-linear complex GPR (K, K-tilde, augmented inverse blocks P, Q) is evaluated as the equivalent
//    real process over [Re f; Im f] of order 2N, so it reuses the same Cholesky / inverse / variance kernels;
//    P_ii, Q_ii and v of complex_kernel.cpp:264-286 are recovered from W by O(N^2) column passes.
#include "chol.cuh"
#include "gpr.cuh"
#include "gpr_kernels.cuh"
#include "mc.cuh"

namespace gple
{
namespace
{
// ---------------------------------------------------------------------------------------------------
// covariance construction
// ---------------------------------------------------------------------------------------------------

// exp(-s / 2), s >= 0, for the test-vs-training covariance rows (kmean_kernel / kstar_kernel): those two kernels evaluate
// 8 Q N exponentials per element and step and are bound by the FP64 pipe, which DFMA shares with DMMA on B200
// (profiles/r02_kstar_fusion_decision.md), so what counts is the number of FP64 instructions per kernel value -- and, right
// behind it, the number of instructions of any kind (one issue slot per cycle and scheduler against two cycles per FP64
// instruction).  The library exp() spends about 17 FP64 instructions (degree-11 polynomial); this one 12:
//   k = rint(-s * 16 / ln 2),  exp(-s / 2) = 2^(k >> 5) * T[k & 31] * exp(r),  T[j] = 2^(j / 32) from shared memory (32 x 8 bytes:
//   at most a two-way bank conflict),  |r| <= ln 2 / 64,  degree-6 Taylor polynomial (remainder 3.5e-18) in Estrin form;
// the power of two goes into the exponent field of the prefactor sigma_f^2 on the integer pipe, so that scaling and prefactor
// are ONE multiplication.  Error <= 2 ulp.  The hot loop carries no special cases at all: it only tracks, per row, whether
// some s was exactly zero (a query that coincides with a training point: the delta term of kernel.cpp:8-31) or outside
// [2^-1022, 1270) (underflow of the result, Inf, NaN); such a row -- rare -- is recomputed as a whole with the library exp and the
// exact coincidence predicate (gauss_value), in the same summation order.  Whether a row is recomputed depends on its data
// only, never on which kernel evaluates it, so predictions stay bit-identical across kmean_kernel<*> and kstar_kernel.
__constant__ double c_exp2_tab[32] = {
	0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
	0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, 0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
	0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0, 0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
	0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

/// every thread of the CTA calls this before the first fast_kernel_value; includes the barrier
__device__ __forceinline__ void load_exp_table(double* __restrict__ tab)
{
	if (threadIdx.x < 32)
	{
		tab[threadIdx.x] = c_exp2_tab[threadIdx.x];
	}
	__syncthreads();
}

/// `track` trips (>= FAST_TRIP) when the high word of some s was 0 (s == 0 or denormal) or >= that of 1270.0 (large, Inf, NaN,
/// negative): unsigned(hi) - 1 maps 0 to 0xFFFFFFFF
constexpr unsigned FAST_TRIP = 0x4093D800u - 1u;

/// a block's prefactor must leave room in the exponent field for 2^(k >> 5) >= 2^-917 and for the table value: 2^-100 .. 2^100
__device__ __forceinline__ bool fast_prefactor_ok(const GaussBlock& g)
{
	const unsigned e = (unsigned(__double2hiint(g.mag2)) >> 20) & 0x7FFu; // sign bit dropped: mag2 may be negative (cross block)
	return e >= 1023u - 100u && e <= 1023u + 100u;
}

/// sigma_f^2 exp(-r^2 / 2) without the coincidence term, valid when `track` does not trip for the row (else garbage)
__device__ __forceinline__ double fast_kernel_value(const GaussBlock& g, const double2 a, const double2 c, const double* __restrict__ tab, unsigned& track)
{
	const double dx = (a.x - c.x) * g.inv_lx, dp = (a.y - c.y) * g.inv_lp;
	const double s = fma(dp, dp, dx * dx);
	track = max(track, unsigned(__double2hiint(s)) - 1u);
	const double t = fma(s, -0x1.71547652b82fep+4, 6755399441055744.0); // -16 / ln 2; 1.5 * 2^52: the low word of t is k
	const int k = __double2loint(t);
	const double kf = t - 6755399441055744.0;
	double r = fma(kf, 0x1.62e42fee00000p-5, s); // r = s + k * (2 ln 2 / 32) = -2 * (reduced argument); the product is exact
	r = fma(kf, 0x1.a39ef35793c76p-37, r);
	// sum_n (-1/2)^n r^(n-1) / n!, n = 1 .. 6, by Estrin's scheme: three dependent levels instead of Horner's five (the loop is
	// bound by the latency between dependent FP64 instructions, ncu: stall `wait`), one multiplication more
	const double r2 = r * r;
	const double pa = fma(0x1.0000000000000p-3, r, -0.5);
	const double pb = fma(0x1.5555555555555p-9, r, -0x1.5555555555555p-6);
	const double pc = fma(0x1.6c16c16c16c17p-16, r, -0x1.1111111111111p-12);
	const double p = fma(fma(pc, r2, pb), r2, pa);
	const double T = tab[k & 31];
	const double v = fma(T, p * r, T);
	// sigma_f^2 * 2^(k >> 5): (k >> 5) << 20 == (k << 15) & 0xFFF00000 added to the high word (k <= 0; no carry out of the field)
	const double m = __hiloint2double(__double2hiint(g.mag2) + ((k << 15) & 0xFFF00000), __double2loint(g.mag2));
	return v * m;
}

/// the exact evaluation (library exp, coincidence term of kernel.cpp:8-31) used for the rows the fast path flags
__device__ __forceinline__ double exact_kernel_value(const GaussBlock& g, const double2 a, const double2 c, bool& same)
{
	same = a.x == c.x && a.y == c.y;
	return gauss_value(g, a, c) + (same ? g.diag_add : 0.0);
}

/// Lower 128-tiles of the (composite) training covariance, row-major n x n with n = nb * Np.
/// Padding rows/columns (point index >= N) carry the identity so the factorisation stays decoupled.
__global__ void __launch_bounds__(256) build_cov_lower_kernel(const BlockSpec spec, const double2* __restrict__ X, const int N, const int Np, const int n, double* __restrict__ K)
{
	const int bj = blockIdx.x, bi = blockIdx.y;
	if (bj > bi)
	{
		return;
	}
	__shared__ double2 xr[128], xc[128];
	const int I0 = bi * 128, J0 = bj * 128;
	const int rb = I0 / Np, cb = J0 / Np; // a 128-tile never straddles blocks (Np is a multiple of 128)
	const int i0 = I0 - rb * Np, j0 = J0 - cb * Np;
	if (threadIdx.x < 128)
	{
		xr[threadIdx.x] = X[i0 + threadIdx.x];
	}
	else
	{
		xc[threadIdx.x - 128] = X[j0 + threadIdx.x - 128];
	}
	__syncthreads();
	const GaussBlock g = spec.b[rb][cb];
	const int c2 = (threadIdx.x & 63) * 2;
	for (int r = threadIdx.x >> 6; r < 128; r += 4)
	{
		const int i = i0 + r;
		double2 out;
		double* o = &out.x;
#pragma unroll
		for (int u = 0; u < 2; u++)
		{
			const int j = j0 + c2 + u;
			double val;
			if (i < N && j < N)
			{
				val = gauss_value(g, xr[r], xc[c2 + u]) + (i == j ? g.diag_add : 0.0);
			}
			else
			{
				val = (I0 + r == J0 + c2 + u) ? 1.0 : 0.0;
			}
			o[u] = val;
		}
		*reinterpret_cast<double2*>(K + size_t(I0 + r) * n + J0 + c2) = out;
	}
}

/// One row of K* by one warp: lane l takes the column pairs 2 l, 2 l + 64, ...; returns the lane's share of the fused mean.
/// FAST: fast_kernel_value, no coincidence term (valid unless `track` trips); else the exact evaluation.
template <bool FAST>
__device__ __forceinline__ double kstar_row(const BlockSpec& spec, const int rb, const double2 xq, const double2* __restrict__ Xt, const int N, const int Np, const int col_limit, const double* __restrict__ w, double* __restrict__ out, const int lane, const double* __restrict__ etab, unsigned& track)
{
	double acc = 0.0;
	for (int cb = 0; cb < spec.nb; cb++)
	{
		const GaussBlock g = spec.b[rb][cb];
		const double* __restrict__ wb = w + cb * Np;
		double* __restrict__ ob = out + cb * Np;
		const int jend = min(Np, col_limit - cb * Np); // col_limit is a multiple of 128
		const int jpair = min(jend, N & ~1);		   // column pairs that are training points both
		int j = lane * 2;
#pragma unroll 4
		for (; j < jpair; j += 64) // no bounds checks, no branches: ptxas interleaves the unrolled pairs
		{
			const double4 x = *reinterpret_cast<const double4*>(Xt + j);
			const double2 wv = *reinterpret_cast<const double2*>(wb + j);
			double2 val;
			if (FAST)
			{
				val.x = fast_kernel_value(g, xq, make_double2(x.x, x.y), etab, track);
				val.y = fast_kernel_value(g, xq, make_double2(x.z, x.w), etab, track);
			}
			else
			{
				bool same;
				val.x = exact_kernel_value(g, xq, make_double2(x.x, x.y), same);
				val.y = exact_kernel_value(g, xq, make_double2(x.z, x.w), same);
			}
			acc += val.x * wv.x;
			acc += val.y * wv.y;
			*reinterpret_cast<double2*>(ob + j) = val;
		}
		for (; j < jend; j += 64) // the odd last training point and the padding columns
		{
			double2 val = make_double2(0.0, 0.0);
			if (j < N)
			{
				bool same;
				val.x = FAST ? fast_kernel_value(g, xq, Xt[j], etab, track) : exact_kernel_value(g, xq, Xt[j], same);
				acc += val.x * wb[j];
			}
			*reinterpret_cast<double2*>(ob + j) = val;
		}
	}
	return acc;
}

/// Rows of the test-vs-training covariance for a chunk of `rows` composite rows starting at row0, written
/// K-contiguous (row-major rows x n) as the A operand of the variance GEMM, fused with the mean K* v
/// (kernel.cpp:495 / complex_kernel.cpp:608).  One warp per row; lanes stream 16-byte stores.
__global__ void __launch_bounds__(256) kstar_kernel(
	const BlockSpec spec,
	const double2* __restrict__ Xq,
	const long long Q,
	const long long row0,
	const int rows,
	const int* __restrict__ row_list, // nullptr: composite row = row0 + r; else row_list[row0 + r] (r < list_count)
	const long long list_count,
	const double2* __restrict__ Xt,
	const int N,
	const int Np,
	const int n,
	const int col_limit, // only composite columns [0, col_limit) are generated (the stage-A tile set reads no others)
	const double* __restrict__ w,
	double* __restrict__ A,
	double* __restrict__ pred
)
{
	__shared__ double etab[32];
	load_exp_table(etab);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int r = blockIdx.x * 8 + warp;
	if (r >= rows)
	{
		return;
	}
	long long R = row0 + r;
	bool valid = true;
	if (row_list != nullptr)
	{
		valid = R < list_count;
		R = valid ? row_list[R] : 0;
	}
	const long long m = R / spec.nb;
	const int rb = int(R - m * spec.nb);
	double* __restrict__ out = A + size_t(r) * n;
	if (m >= Q || !valid)
	{
		for (int J = lane * 2; J < min(n, col_limit); J += 64)
		{
			*reinterpret_cast<double2*>(out + J) = make_double2(0.0, 0.0);
		}
		return;
	}
	const double2 xq = Xq[m];
	unsigned track = fast_prefactor_ok(spec.b[rb][0]) && (spec.nb == 1 || fast_prefactor_ok(spec.b[rb][1])) ? 0u : 0xFFFFFFFFu;
	double acc = 0.0;
	if (track == 0u)
	{
		acc = kstar_row<true>(spec, rb, xq, Xt, N, Np, col_limit, w, out, lane, etab, track);
	}
	if (__any_sync(0xffffffffu, track >= FAST_TRIP)) // a coincidence or an out-of-range exponent somewhere in the row: redo it exactly
	{
		acc = kstar_row<false>(spec, rb, xq, Xt, N, Np, col_limit, w, out, lane, etab, track);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		acc += __shfl_xor_sync(0xffffffffu, acc, o);
	}
	if (lane == 0 && pred != nullptr)
	{
		pred[r] = acc;
	}
}

/// Mean only: pred[R] = sum_j k(x_q, x_j) v_j for composite rows R = 0 .. total_rows-1 (kernel.cpp:495 /
/// complex_kernel.cpp:608) without ever storing K*.  exp-bound: a tile of 256 training points (x, p, v) is staged in
/// shared memory.  WARP_PER_ROW = false (large batches): one row per thread with 32 independent partial sums in
/// registers, tile entries read as broadcasts, no shuffles; true (small batches): one warp per row.  Both walk the
/// training points in exactly the order of kstar_kernel's fused mean (slot s takes columns 2s, 2s+1, 2s+64, ...; the
/// slots are combined by the xor butterfly), so a prediction is bit-identical whichever kernel produced it.
/// `coincident[R]` = 1 when the query equals a training point (the delta term of kernel.cpp:8-31 fired).
template <bool WARP_PER_ROW>
__global__ void __launch_bounds__(256, WARP_PER_ROW ? 1 : 2) kmean_kernel(
	const BlockSpec spec,
	const double2* __restrict__ Xq,
	const long long total_rows,
	const double2* __restrict__ Xt,
	const int N,
	const int Np,
	const double* __restrict__ w,
	double* __restrict__ pred,
	unsigned char* __restrict__ coincident
)
{
	constexpr int TJ = 256, ROWS = WARP_PER_ROW ? 8 : 256, NACC = WARP_PER_ROW ? 1 : 32;
	__shared__ double4 tile[TJ];
	__shared__ double etab[32];
	load_exp_table(etab);
	const int lane = threadIdx.x & 31;
	const long long R = (long long)blockIdx.x * ROWS + (WARP_PER_ROW ? threadIdx.x >> 5 : threadIdx.x);
	const bool live = R < total_rows;
	const long long m = live ? R / spec.nb : 0;
	const int rb = int(live ? R - m * spec.nb : 0);
	const double2 xq = Xq[m];
	double acc[NACC];
#pragma unroll
	for (int i = 0; i < NACC; i++)
	{
		acc[i] = 0.0;
	}
	// fast pass (fast_kernel_value; see the comment at c_exp2_tab): no coincidence term, no special cases, only `track`
	unsigned track = fast_prefactor_ok(spec.b[rb][0]) && (spec.nb == 1 || fast_prefactor_ok(spec.b[rb][1])) ? 0u : 0xFFFFFFFFu;
	for (int cb = 0; cb < spec.nb; cb++)
	{
		const GaussBlock g = spec.b[rb][cb];
		for (int j0 = 0; j0 < Np; j0 += TJ)
		{
			__syncthreads();
			{
				// padding entries repeat the last training point with weight zero: they add exactly 0 to a sum and can only flag a
				// coincidence that the real entry flags as well, so the loops below need no per-pair bounds check
				const int j = j0 + threadIdx.x;
				const double2 x = Xt[min(j, N - 1)];
				tile[threadIdx.x] = make_double4(x.x, x.y, j < N ? w[cb * Np + j] : 0.0, 0.0);
			}
			__syncthreads();
			for (int k = 0; k < TJ / 64; k++)
			{
				if (j0 + 64 * k >= N)
				{
					break;
				}
				if (WARP_PER_ROW)
				{
#pragma unroll
					for (int u = 0; u < 2; u++)
					{
						const double4 t = tile[64 * k + 2 * lane + u];
						acc[0] += fast_kernel_value(g, xq, make_double2(t.x, t.y), etab, track) * t.z;
					}
				}
				else
				{
#pragma unroll
					for (int sl = 0; sl < 32; sl++)
					{
#pragma unroll
						for (int u = 0; u < 2; u++)
						{
							const double4 t = tile[64 * k + 2 * sl + u];
							acc[sl < NACC ? sl : 0] += fast_kernel_value(g, xq, make_double2(t.x, t.y), etab, track) * t.z;
						}
					}
				}
			}
		}
	}
	// exact pass for the flagged rows (a coincidence or an out-of-range exponent somewhere in the row), same summation order,
	// training points straight from global memory (no barrier may sit inside a divergent region)
	bool hit = false;
	if (WARP_PER_ROW ? __any_sync(0xffffffffu, track >= FAST_TRIP) : track >= FAST_TRIP)
	{
#pragma unroll
		for (int i = 0; i < NACC; i++)
		{
			acc[i] = 0.0;
		}
		for (int cb = 0; cb < spec.nb; cb++)
		{
			const GaussBlock g = spec.b[rb][cb];
			const double* __restrict__ wb = w + cb * Np;
			for (int jb = 0; jb < N; jb += 64)
			{
#pragma unroll
				for (int sl = 0; sl < NACC; sl++)
				{
#pragma unroll
					for (int u = 0; u < 2; u++)
					{
						const int j = jb + 2 * (WARP_PER_ROW ? lane : sl) + u;
						if (j < N)
						{
							bool same;
							acc[sl] += exact_kernel_value(g, xq, Xt[j], same) * wb[j];
							hit |= same;
						}
					}
				}
			}
		}
	}
	double total;
	if (WARP_PER_ROW)
	{
		total = acc[0];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			total += __shfl_xor_sync(0xffffffffu, total, o);
			hit |= __shfl_xor_sync(0xffffffffu, int(hit), o) != 0;
		}
	}
	else
	{
		// the xor butterfly as seen by lane 0: a[i] += a[i ^ o] for o = 16, 8, 4, 2, 1
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
#pragma unroll
			for (int i = 0; i < o; i++)
			{
				acc[i < NACC ? i : 0] += acc[i + o < NACC ? i + o : 0];
			}
		}
		total = acc[0];
	}
	if (live && (!WARP_PER_ROW || lane == 0))
	{
		pred[R] = total;
		if (coincident != nullptr)
		{
			coincident[R] = hit ? 1 : 0;
		}
	}
}

void launch_kmean(gple_ctx* ctx, const BlockSpec& spec, const double2* Xq, long long total_rows, const double2* Xt, int N, int Np, const double* w, double* pred, unsigned char* coincident)
{
	if (total_rows >= 148ll * 256)
	{
		GPLE_LAUNCH(ctx, kmean_kernel<false>, unsigned((total_rows + 255) / 256), 256, 0, spec, Xq, total_rows, Xt, N, Np, w, pred, coincident);
	}
	else
	{
		GPLE_LAUNCH(ctx, kmean_kernel<true>, unsigned((total_rows + 7) / 8), 256, 0, spec, Xq, total_rows, Xt, N, Np, w, pred, coincident);
	}
}

// ---------------------------------------------------------------------------------------------------
// variance GEMM:  q[r] = sum_n ( sum_{k <= n} A[r][k] W[n][k] )^2      (Z = A W^T, W lower triangular)
// ---------------------------------------------------------------------------------------------------

/// One CTA owns 128 query rows and walks all n-tiles; the cp.async pipeline runs across n-tile boundaries
/// (flattened (n-tile, k-step) sequence) so the tensor pipe never drains.  The squared accumulators are
/// folded into per-thread row sums; no Z is ever written.
/// The n-tiles a launch walks: two runs of consecutive 128-wide tiles, [b0, b0 + c0) then [b1, b1 + c1).  Any subset of the
/// columns of Z gives a LOWER bound of sum Z^2, i.e. an upper bound of the variance -- the staged gate uses that.
struct TileSet
{
	int b0, c0, b1, c1;
	__host__ __device__ int count() const { return c0 + c1; }
	__host__ __device__ int tile(const int v) const { return v < c0 ? b0 + v : b1 + (v - c0); }
	/// sum over the set of (tile + 1): the k-length of the triangular products in units of 128
	double weight() const
	{
		double w = 0.0;
		for (int v = 0; v < count(); v++)
		{
			w += double(tile(v) + 1);
		}
		return w;
	}
};

template <typename C>
__global__ void __launch_bounds__(C::THREADS, 1) var_gemm_kernel(const double* __restrict__ A, const int lda, const double* __restrict__ W, const int n, double* __restrict__ part, const int rows, const TileSet ts)
{
	using namespace gemm;
	extern __shared__ __align__(16) double smem[];
	double* As = smem;
	double* Bs = smem + C::STAGES * C::A_STAGE;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N, g = lane >> 2, t = lane & 3;
	constexpr int BM = 128, BN = 128;
	static_assert(C::BM == 128 && C::BN == 128, "the variance GEMM walks 128-wide n-tiles");
	const int m0 = blockIdx.x * BM;
	const int V = ts.count();
	const double* Ag = A + size_t(m0) * lda;
	// When a chunk has fewer than ~148 row blocks the n-tiles are dealt round-robin to gridDim.y CTAs per row
	// block (interleaving balances the triangular work).  Whatever the deal, every (n-tile, column-warp) pair writes its
	// own partial row sums to part[(v * WARPS_N + wn)][rows]; var_reduce_kernel adds them in that fixed order, so a row's
	// result does not depend on the size or composition of the batch it was computed in (nor on the number of GPUs).
	const int nt0 = blockIdx.y, nts = gridDim.y;

	int l_v = nt0, l_nt = ts.tile(nt0), l_kt = 0, l_slot = 0;
	auto issue = [&]()
	{
		if (l_v < V)
		{
			load_tile_kmajor<C, 128>(As + l_slot * C::A_STAGE, Ag + size_t(l_kt) * C::BK, size_t(lda), tid);
			load_tile_kmajor<C, 128>(Bs + l_slot * C::B_STAGE, W + size_t(l_nt) * BN * n + size_t(l_kt) * C::BK, size_t(n), tid);
			l_slot = (l_slot + 1 == C::STAGES) ? 0 : l_slot + 1;
			if (++l_kt == (l_nt + 1) * (BN / C::BK))
			{
				l_kt = 0;
				l_v += nts;
				l_nt = ts.tile(l_v);
			}
		}
		cp_async_commit();
	};
#pragma unroll
	for (int s = 0; s < C::STAGES - 1; s++)
	{
		issue();
	}
	double acc[C::MI][C::NJ][2];
	int c_slot = 0;
	for (int v = nt0; v < V; v += nts)
	{
		zero_acc<C>(acc);
		const int steps = (ts.tile(v) + 1) * (BN / C::BK);
		for (int kt = 0; kt < steps; kt++)
		{
			cp_async_wait<C::STAGES - 2>();
			__syncthreads();
			issue();
			compute_stage<C, false>(acc, As + c_slot * C::A_STAGE, Bs + c_slot * C::B_STAGE, wm, wn, g, t);
			c_slot = (c_slot + 1 == C::STAGES) ? 0 : c_slot + 1;
		}
		// this warp's share of the tile: sum of squares over its C::NJ * 8 columns, per row (quad reduction over the columns)
		double* dst = part + (size_t(v) * C::WARPS_N + wn) * rows + m0 + wm * C::WTM + g;
#pragma unroll
		for (int i = 0; i < C::MI; i++)
		{
			double s = 0.0;
#pragma unroll
			for (int j = 0; j < C::NJ; j++)
			{
				s = fma(acc[i][j][0], acc[i][j][0], s);
				s = fma(acc[i][j][1], acc[i][j][1], s);
			}
			s += __shfl_xor_sync(0xffffffffu, s, 1);
			s += __shfl_xor_sync(0xffffffffu, s, 2);
			if (t == 0)
			{
				dst[i * 8] = s;
			}
		}
	}
	cp_async_wait<0>();
}

/// q[r] = sum of the `count` partial row sums of a launch (count = n-tiles of the set x column-warps), in storage order
__global__ void var_reduce_kernel(const double* __restrict__ part, double* __restrict__ q, const int rows, const int count)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= rows)
	{
		return;
	}
	double s = 0.0;
	for (int k = 0; k < count; k++)
	{
		s += part[size_t(k) * rows + r];
	}
	q[r] = s;
}

/// Tile configurations of the variance GEMM selectable at run time (tuned on the B200, see profiles/).
using VarCfg0 = gemm::Config<16, 4, 2, 4>;
using VarCfg1 = gemm::Config<32, 3, 2, 4>;
using VarCfg2 = gemm::Config<16, 4, 4, 4>;
using VarCfg3 = gemm::Config<32, 3, 4, 4>;
using VarCfg4 = gemm::Config<16, 5, 2, 4>;
using VarCfg5 = gemm::Config<16, 4, 2, 8>;
using VarCfg6 = gemm::Config<16, 4, 4, 2>;
constexpr int NUM_VAR_VARIANTS = 7;
constexpr int MAX_VAR_SPLITS = 16;
int g_var_variant = 1; // BK = 32, 3 stages, 2 x 4 warps: 88.8 % of the DMMA peak on B200 (profiles/r01_tune_var_gemm.txt)

template <typename C>
void launch_var_gemm_cfg(gple_ctx* ctx, const double* A, int lda, const double* W, int n, int rows, double* q, const TileSet& ts)
{
	// choose the number of n-splits S that minimises the makespan ceil(blocks * S / SMs) / S
	const int mb = rows / 128, T = ts.count();
	int best = 1;
	double best_cost = double((mb + ctx->num_sms - 1) / ctx->num_sms);
	for (int S = 2; S <= std::min(T, MAX_VAR_SPLITS); S++)
	{
		const double cost = double((mb * S + ctx->num_sms - 1) / ctx->num_sms) / S * (1.0 + 0.01 * S); // small penalty per split
		if (cost < best_cost * 0.97)
		{
			best_cost = cost;
			best = S;
		}
	}
	double* part = ctx->ws.get<double>("pred.qpart", size_t(rows) * size_t(T) * C::WARPS_N);
	GPLE_LAUNCH(ctx, var_gemm_kernel<C>, dim3(mb, best), C::THREADS, C::SMEM_BYTES, A, lda, W, n, part, rows, ts);
	GPLE_LAUNCH(ctx, var_reduce_kernel, (rows + 255) / 256, 256, 0, part, q, rows, T * C::WARPS_N);
}

/// flops executed by a launch over `rows` rows: 2 * 128^2 * sum over the tile set of (tile + 1) per row
double var_gemm_flops(const TileSet& ts, int rows)
{
	return 2.0 * 128.0 * 128.0 * ts.weight() * double(rows);
}

/// A: rows x (at least the columns the tile set reads), row pitch lda; W: the n x n lower-triangular inverse factor
void launch_var_gemm(gple_ctx* ctx, int variant, const double* A, int lda, const double* W, int n, int rows, double* q, const TileSet& ts)
{
	switch (variant)
	{
	case 1:
		return launch_var_gemm_cfg<VarCfg1>(ctx, A, lda, W, n, rows, q, ts);
	case 2:
		return launch_var_gemm_cfg<VarCfg2>(ctx, A, lda, W, n, rows, q, ts);
	case 3:
		return launch_var_gemm_cfg<VarCfg3>(ctx, A, lda, W, n, rows, q, ts);
	case 4:
		return launch_var_gemm_cfg<VarCfg4>(ctx, A, lda, W, n, rows, q, ts);
	case 5:
		return launch_var_gemm_cfg<VarCfg5>(ctx, A, lda, W, n, rows, q, ts);
	case 6:
		return launch_var_gemm_cfg<VarCfg6>(ctx, A, lda, W, n, rows, q, ts);
	default:
		return launch_var_gemm_cfg<VarCfg0>(ctx, A, lda, W, n, rows, q, ts);
	}
}

/// Bound-gated variance.  The cutoff gate of kernel.h:301-332 is decided without the variance for two sets of queries:
///  * |f|^2 >= 4 k**: the computed variance k** - sum Z^2 can never exceed the prior k**, so the gate is exactly 1;
///  * |f|^2 <= noise/2 where noise = sigma_f^2 sigma_n^2 (complex: sigma^2 sigma_n^2) and the query coincides with no
///    training point: k K^-1 k^T <= the latent prior variance (Schur complement of the noise-free covariance, for the
///    complex element of the composite [Re; Im] process), so var >= noise > |f|^2 and the gate is exactly 0.  Used only
///    when noise >= 1e-9 k**, far above the rounding error of the computed variance.
/// Collects the rows of the remaining queries (both rows of a complex query) into `idx` and records each row's slot
/// (-1: gate 1, -2: gate 0).  counter[0] = rows collected, counter[1] = rows decided 0.
__global__ void __launch_bounds__(256) classify_kernel(const double* __restrict__ pred, const unsigned char* __restrict__ coincident, const int points, const int nb, const double four_prior, const double half_noise, int* __restrict__ idx, int* __restrict__ slot, int* __restrict__ counter)
{
	const int p = blockIdx.x * 256 + threadIdx.x;
	bool need = false, zero = false;
	if (p < points)
	{
		double f2 = 0.0;
		bool hit = false;
		for (int k = 0; k < nb; k++)
		{
			const double f = pred[p * nb + k];
			f2 = fma(f, f, f2);
			hit |= coincident[p * nb + k] != 0;
		}
		zero = !hit && f2 <= half_noise;	   // half_noise < 0 disables the lower bound
		need = !(f2 >= four_prior) && !zero; // NaN keeps the full path
	}
	const unsigned ballot = __ballot_sync(0xffffffffu, need), zballot = __ballot_sync(0xffffffffu, zero);
	const int lane = threadIdx.x & 31;
	int base = 0;
	if (lane == 0)
	{
		if (ballot != 0u)
		{
			base = atomicAdd(counter, __popc(ballot) * nb);
		}
		if (zballot != 0u)
		{
			atomicAdd(counter + 1, __popc(zballot) * nb);
		}
	}
	base = __shfl_sync(0xffffffffu, base, 0);
	if (p < points)
	{
		const int mine = base + __popc(ballot & ((1u << lane) - 1u)) * nb;
		for (int k = 0; k < nb; k++)
		{
			slot[p * nb + k] = need ? mine + k : (zero ? -2 : -1);
			if (need)
			{
				idx[mine + k] = p * nb + k;
			}
		}
	}
}

/// A stage of the bound-gated variance: `qs` = sum Z^2 of the listed rows over this stage's tile set is added to the running
/// per-row sum `qsum` (dense, indexed by composite row; assigned when `first`), and var <= k** - qsum decides gate == 1 exactly
/// where |f|^2 >= 4 (k** - qsum).  One thread per listed query (nb consecutive list entries).  Decided rows get slot = -1; the
/// rows of the others are compacted into `next` (composite rows); *counter = rows kept.
__global__ void __launch_bounds__(256) classify_stage_kernel(const double* __restrict__ pred, const int* __restrict__ list, const double* __restrict__ qs, double* __restrict__ qsum, const int first, const int queries, const int nb, const double prior, int* __restrict__ next, int* __restrict__ slot, int* __restrict__ counter)
{
	const int e = blockIdx.x * 256 + threadIdx.x;
	bool need = false;
	if (e < queries)
	{
		double f2 = 0.0, total = 0.0;
		for (int k = 0; k < nb; k++)
		{
			const int r = list[e * nb + k];
			const double f = pred[r];
			f2 = fma(f, f, f2);
			const double t = (first ? 0.0 : qsum[r]) + qs[e * nb + k];
			qsum[r] = t;
			total += t;
		}
		need = !(f2 >= 4.0 * (prior - total)); // NaN keeps the full path
	}
	const unsigned ballot = __ballot_sync(0xffffffffu, need);
	const int lane = threadIdx.x & 31;
	int base = 0;
	if (lane == 0 && ballot != 0u)
	{
		base = atomicAdd(counter, __popc(ballot) * nb);
	}
	base = __shfl_sync(0xffffffffu, base, 0);
	if (e < queries)
	{
		const int mine = base + __popc(ballot & ((1u << lane) - 1u)) * nb;
		for (int k = 0; k < nb; k++)
		{
			const int r = list[e * nb + k];
			if (need)
			{
				next[mine + k] = r;
			}
			else
			{
				slot[r] = -1;
			}
		}
	}
}

/// Last stage: the listed rows have seen every tile; their sums are completed (the gate is then evaluated from the variance).
__global__ void __launch_bounds__(256) accumulate_stage_kernel(const int* __restrict__ list, const double* __restrict__ qs, double* __restrict__ qsum, const int first, const int count)
{
	const int e = blockIdx.x * 256 + threadIdx.x;
	if (e < count)
	{
		const int r = list[e];
		qsum[r] = (first ? 0.0 : qsum[r]) + qs[e];
	}
}

__device__ __forceinline__ double gate_factor(const double pred_sq, const double abs_pred, const double var)
{
	if (pred_sq >= 4.0 * var) // ConnectingPoint^2 (also taken for var < 0: quirk q5)
	{
		return 1.0;
	}
	if (pred_sq <= var)
	{
		return 0.0;
	}
	const double a = abs_pred / sqrt(var);
	return (5.0 - 2.0 * a) * (a - 1.0) * (a - 1.0);
}

/// kernel.cpp:496-519 + kernel.h:301-332: variance, cubic gate, cutoff prediction (real element)
__global__ void finalize_real_kernel(const double* __restrict__ pred, const double* __restrict__ q, const int* __restrict__ slot, const int rows, const long long row0, const long long Q, const double prior, const double rescale, double* __restrict__ pred_out, double* __restrict__ var_out, double* __restrict__ cut_out)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= rows || row0 + r >= Q)
	{
		return;
	}
	const double f = pred[r];
	// slot (bound-gated schedule only): >= 0 the row went through every tile and q holds its sum Z^2; -1 / -2: the gate was decided
	// 1 / 0 by a bound (classify_kernel, classify_stage_kernel)
	const int sl = slot != nullptr ? slot[r] : 0;
	const double var = sl >= 0 ? prior - q[r] : prior;
	const double gate = sl >= 0 ? gate_factor(f * f, fabs(f), var) : (sl == -1 ? 1.0 : 0.0);
	if (pred_out != nullptr)
	{
		pred_out[row0 + r] = f;
	}
	if (var_out != nullptr)
	{
		var_out[row0 + r] = var;
	}
	if (cut_out != nullptr)
	{
		cut_out[row0 + r] = f * gate / rescale;
	}
}

/// complex_kernel.cpp:608-643 in composite form: rows (2m, 2m+1) = (Re, Im) parts of query m
__global__ void finalize_complex_kernel(const double* __restrict__ pred, const double* __restrict__ q, const int* __restrict__ slot, const int rows, const long long row0, const long long Q, const double prior, const double rescale, double2* __restrict__ pred_out, double* __restrict__ var_out, double2* __restrict__ cut_out)
{
	const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
	const long long m = row0 / 2 + pidx;
	if (2 * pidx + 1 >= rows || m >= Q)
	{
		return;
	}
	const double fr = pred[2 * pidx], fi = pred[2 * pidx + 1];
	const int sl = slot != nullptr ? slot[2 * pidx] : 0;
	const double var = sl >= 0 ? prior - q[2 * pidx] - q[2 * pidx + 1] : prior;
	const double ps = fr * fr + fi * fi;
	const double gate = sl >= 0 ? gate_factor(ps, hypot(fr, fi), var) : (sl == -1 ? 1.0 : 0.0);
	if (pred_out != nullptr)
	{
		pred_out[m] = make_double2(fr, fi);
	}
	if (var_out != nullptr)
	{
		var_out[m] = var;
	}
	if (cut_out != nullptr)
	{
		cut_out[m] = make_double2(fr * gate / rescale, fi * gate / rescale);
	}
}

/// partial sums of (pred - rescale * y)^2 over a chunk, accumulated into acc[0] by a single block
/// (deterministic order: chunks are processed sequentially on the stream).
__global__ void __launch_bounds__(1024) sqerr_kernel(const double* __restrict__ pred, const int rows, const long long row0, const long long total_rows, const double* __restrict__ yq, const int y_stride, const int nb, const double rescale, double* __restrict__ acc)
{
	__shared__ double scratch[32];
	double s[1] = {0.0};
	for (int r = threadIdx.x; r < rows && row0 + r < total_rows; r += 1024)
	{
		const long long R = row0 + r;
		const long long m = R / nb;
		const int part = int(R - m * nb);
		const double d = pred[r] - rescale * yq[m * y_stride + part];
		s[0] = fma(d, d, s[0]);
	}
	block_reduce<1, 1024>(s, scratch);
	if (threadIdx.x == 0)
	{
		acc[0] += s[0];
	}
}

// ---------------------------------------------------------------------------------------------------
// training: labels, v = W^T W y', diag(K^-1), scalar reductions
// ---------------------------------------------------------------------------------------------------

/// kernel.cpp:279-280 / complex_kernel.cpp:262-263.  out[0] = rescale.  label = [s Re y ; s Im y (complex only)].
__global__ void __launch_bounds__(1024) label_kernel(const double2* __restrict__ y, const int N, const int Np, const int is_complex, double* __restrict__ label, double* __restrict__ scal)
{
	__shared__ double red[32];
	__shared__ double s_rescale;
	double mx = 0.0;
	for (int i = threadIdx.x; i < N; i += 1024)
	{
		const double2 v = y[i];
		mx = fmax(mx, is_complex ? hypot(v.x, v.y) : fabs(v.x));
	}
	for (int o = 16; o > 0; o >>= 1)
	{
		mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	}
	if ((threadIdx.x & 31) == 0)
	{
		red[threadIdx.x >> 5] = mx;
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		double m = 0.0;
		for (int w = 0; w < 32; w++)
		{
			m = fmax(m, red[w]);
		}
		s_rescale = 10.0 / m; // RescaleMaximum, kernel.h:37
		scal[0] = s_rescale;
	}
	__syncthreads();
	const double s = s_rescale;
	for (int i = threadIdx.x; i < Np; i += 1024)
	{
		const double2 v = i < N ? y[i] : make_double2(0.0, 0.0);
		label[i] = v.x * s;
		if (is_complex)
		{
			label[Np + i] = v.y * s;
		}
	}
}

/// z = W y (W lower triangular, row-major): one warp per row.
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* __restrict__ W, const int n, const double* __restrict__ y, double* __restrict__ z)
{
	const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (row >= n)
	{
		return;
	}
	const double* __restrict__ wr = W + size_t(row) * n;
	double s = 0.0;
	for (int k = lane * 2; k <= row; k += 64)
	{
		const double2 a = *reinterpret_cast<const double2*>(wr + k), b = *reinterpret_cast<const double2*>(y + k);
		s = fma(a.x, b.x, s);
		s = fma(a.y, b.y, s); // W[row][row + 1] == 0
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		s += __shfl_xor_sync(0xffffffffu, s, o);
	}
	if (lane == 0)
	{
		z[row] = s;
	}
}

/// Column pass over W: v_k = sum_i W[i][k] z_i, ss_k = sum_i W[i][k]^2 (= [K^-1]_kk), and for the composite
/// model cross_k = sum_i W[i][k] W[i][Np + k], k < Np (= [M_ri]_kk).  Rows are split over blockIdx.y and
/// the per-slab partials are reduced by `col_reduce_kernel` (deterministic, no atomics).
__global__ void __launch_bounds__(128) col_pass_kernel(const double* __restrict__ W, const int n, const int Np, const int want_cross, const double* __restrict__ z, const int slab, double* __restrict__ part /* [gridDim.y][3][n] */)
{
	const int k = blockIdx.x * 128 + threadIdx.x;
	const int i_begin = max(blockIdx.y * slab, (k / 128) * 128), i_end = min(n, (blockIdx.y + 1) * slab);
	double sv = 0.0, ss = 0.0, sc = 0.0;
	const bool cross = want_cross && k < Np;
#pragma unroll 8
	for (int i = i_begin; i < i_end; i++)
	{
		const double a = W[size_t(i) * n + k];
		sv = fma(a, z[i], sv);
		ss = fma(a, a, ss);
		if (cross)
		{
			sc = fma(a, W[size_t(i) * n + Np + k], sc);
		}
	}
	double* p = part + size_t(blockIdx.y) * 3 * n;
	p[k] = sv;
	p[n + k] = ss;
	p[2 * n + k] = sc;
}

/// Residual of the solve, r = y' - C v, with the covariance regenerated from the points (the factorisation overwrote it):
/// one warp per composite row, entries evaluated exactly as build_cov_lower_kernel evaluates them.  One step of iterative
/// refinement v <- v + W^T W r brings K^-1 y' from the accuracy of the explicit triangular inverse (a few 1e-8 relative at
/// cond(K) ~ 1e7, measured against a double-double evaluation, tests/golden/arbiter_dd.cpp) to that of a backward-stable solve
/// (1e-10), which is what the reference's LDLT + solve delivers (kernel.cpp:281-284).
__global__ void __launch_bounds__(256) residual_kernel(const BlockSpec spec, const double2* __restrict__ X, const int N, const int Np, const int n, const double* __restrict__ label, const double* __restrict__ v, double* __restrict__ r)
{
	const int I = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (I >= n)
	{
		return;
	}
	const int rb = I / Np, i = I - rb * Np;
	double acc = 0.0;
	if (i < N)
	{
		const double2 xi = X[i];
		for (int cb = 0; cb < spec.nb; cb++)
		{
			const GaussBlock g = spec.b[rb][cb];
			const double* __restrict__ vb = v + cb * Np;
			for (int j = lane; j < N; j += 32)
			{
				const double val = gauss_value(g, xi, X[j]) + (i == j ? g.diag_add : 0.0);
				acc = fma(-val, vb[j], acc);
			}
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		acc += __shfl_xor_sync(0xffffffffu, acc, o);
	}
	if (lane == 0)
	{
		r[I] = i < N ? label[I] + acc : 0.0;
	}
}
/// v += sum over the slabs of the solution partials of a column pass
__global__ void refine_add_kernel(const double* __restrict__ part, const int n, const int slabs, double* __restrict__ v)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n)
	{
		return;
	}
	double a = 0.0;
	for (int s = 0; s < slabs; s++)
	{
		a += part[size_t(s) * 3 * n + k];
	}
	v[k] += a;
}
__global__ void col_reduce_kernel(const double* __restrict__ part, const int n, const int slabs, double* __restrict__ v, double* __restrict__ ss, double* __restrict__ cross)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n)
	{
		return;
	}
	double a = 0.0, b = 0.0, c = 0.0;
	for (int s = 0; s < slabs; s++)
	{
		const double* p = part + size_t(s) * 3 * n;
		a += p[k];
		b += p[n + k];
		c += p[2 * n + k];
	}
	v[k] = a;
	ss[k] = b;
	if (cross != nullptr)
	{
		cross[k] = c;
	}
}

/// Real element reductions.  scal[1..6] = LOOCV error, sum v, sum x v, sum p v, y'^T v, unused
__global__ void __launch_bounds__(1024) real_scalars_kernel(const double2* __restrict__ X, const double* __restrict__ label, const double* __restrict__ v, const double* __restrict__ dinv, const int N, double* __restrict__ scal)
{
	__shared__ double scratch[5 * 32];
	double s[5] = {0, 0, 0, 0, 0};
	for (int i = threadIdx.x; i < N; i += 1024)
	{
		const double vi = v[i], r = vi / dinv[i];
		const double2 x = X[i];
		s[0] = fma(r, r, s[0]);
		s[1] += vi;
		s[2] = fma(x.x, vi, s[2]);
		s[3] = fma(x.y, vi, s[3]);
		s[4] = fma(label[i], vi, s[4]);
	}
	block_reduce<5, 1024>(s, scratch);
	if (threadIdx.x == 0)
	{
		for (int i = 0; i < 5; i++)
		{
			scal[1 + i] = s[i];
		}
	}
}

/// Complex element reductions from the composite solution (complex_kernel.cpp:270-286, complex_kernel.h:192-204):
/// P_ii = (Mrr + Mii) / 4, Q_ii = (Mrr - Mii) / 4 - i Mri / 2, v_i = (wr + i wi) / 2.
/// scal[1] = LOOCV error, scal[2] = Re(y'^H v)
__global__ void __launch_bounds__(1024) complex_scalars_kernel(const double* __restrict__ label, const double* __restrict__ w, const double* __restrict__ ss, const double* __restrict__ cross, const int N, const int Np, double* __restrict__ scal)
{
	__shared__ double scratch[2 * 32];
	double s[2] = {0, 0};
	for (int i = threadIdx.x; i < N; i += 1024)
	{
		const double mrr = ss[i], mii = ss[Np + i], mri = cross[i];
		const double p = 0.25 * (mrr + mii);
		const double qr = 0.25 * (mrr - mii), qi = -0.5 * mri;
		const double vr = 0.5 * w[i], vi = 0.5 * w[Np + i];
		// numerator P v - conj(Q v)
		const double qvr = qr * vr - qi * vi, qvi = qr * vi + qi * vr;
		const double nr = p * vr - qvr, ni = p * vi + qvi;
		const double den = p * p - (qr * qr + qi * qi);
		const double dr = nr / den, di = ni / den;
		s[0] += dr * dr + di * di;
		s[1] += label[i] * vr + label[Np + i] * vi;
	}
	block_reduce<2, 1024>(s, scratch);
	if (threadIdx.x == 0)
	{
		scal[1] = s[0];
		scal[2] = s[1];
	}
}

/// Quadratic form  sum_{I,J} w_I Kaux(I,J) w_J  with Kaux generated on the fly (never stored): the purity
/// integrals of kernel.cpp:313-335 and complex_kernel.cpp:357-377.  Each block of the covariance may be the
/// sum of two Gaussians.  One CTA per 128 x 128 tile; per-tile partials are summed by quad_reduce_kernel.
struct QuadSpec
{
	int nb;
	GaussBlock g[2][2][2];
};
__global__ void __launch_bounds__(256) quadform_kernel(const QuadSpec spec, const double2* __restrict__ X, const int N, const int Np, const double* __restrict__ w, double* __restrict__ part)
{
	__shared__ double2 xr[128], xc[128];
	__shared__ double wr[128], wc[128];
	__shared__ double scratch[8];
	const int tiles = Np / 128;
	const int bi = blockIdx.y, bj = blockIdx.x;
	const int rb = bi / tiles, cb = bj / tiles;
	const int i0 = (bi - rb * tiles) * 128, j0 = (bj - cb * tiles) * 128;
	if (threadIdx.x < 128)
	{
		xr[threadIdx.x] = X[i0 + threadIdx.x];
		wr[threadIdx.x] = (i0 + threadIdx.x < N) ? w[rb * Np + i0 + threadIdx.x] : 0.0;
	}
	else
	{
		const int c = threadIdx.x - 128;
		xc[c] = X[j0 + c];
		wc[c] = (j0 + c < N) ? w[cb * Np + j0 + c] : 0.0;
	}
	__syncthreads();
	const GaussBlock ga = spec.g[rb][cb][0], gb = spec.g[rb][cb][1];
	double s[1] = {0.0};
	const int c = threadIdx.x & 127;
	for (int r = threadIdx.x >> 7; r < 128; r += 2)
	{
		double k = gauss_value(ga, xr[r], xc[c]);
		if (gb.mag2 != 0.0)
		{
			k += gauss_value(gb, xr[r], xc[c]);
		}
		s[0] = fma(wr[r] * k, wc[c], s[0]);
	}
	block_reduce<1, 256>(s, scratch);
	if (threadIdx.x == 0)
	{
		part[blockIdx.y * gridDim.x + blockIdx.x] = s[0];
	}
}
__global__ void __launch_bounds__(1024) sum_kernel(const double* __restrict__ part, const int count, double* __restrict__ out)
{
	__shared__ double scratch[32];
	double s[1] = {0.0};
	for (int i = threadIdx.x; i < count; i += 1024)
	{
		s[0] += part[i];
	}
	block_reduce<1, 1024>(s, scratch);
	if (threadIdx.x == 0)
	{
		out[0] = s[0];
	}
}

/// out (n x n, row-major) = transpose of in
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, const int n)
{
	__shared__ double tile[32][33];
	const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
	for (int r = threadIdx.y; r < 32; r += 8)
	{
		tile[r][threadIdx.x] = in[size_t(by + r) * n + bx + threadIdx.x];
	}
	__syncthreads();
	for (int r = threadIdx.y; r < 32; r += 8)
	{
		out[size_t(bx + r) * n + by + threadIdx.x] = tile[threadIdx.x][r];
	}
}

// ---------------------------------------------------------------------------------------------------
// general kernel-matrix builder for the C-ABI (column-major output like Eigen)
// ---------------------------------------------------------------------------------------------------

/// KernelBase (kernel.cpp:217-242) incl. delta_kernel (kernel.cpp:8-31) and calculate_derivative
/// (kernel.cpp:168-215): K and the four dK/dtheta, column-major nL x nR.
__global__ void __launch_bounds__(256) kernel_real_kernel(const double2* __restrict__ XL, const int nL, const double2* __restrict__ XR, const int nR, const double mag, const double lx, const double lp, const double noise, const int same, double* __restrict__ K, double* __restrict__ dK)
{
	const int r = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
	if (r >= nL)
	{
		return;
	}
	const double2 a = XL[r], b = XR[c];
	const double dx = (a.x - b.x) / lx, dp = (a.y - b.y) / lp;
	const double gsn = exp(-(dx * dx + dp * dp) / 2.0);
	const double delta = same ? (r == c ? 1.0 : 0.0) : ((a.x == b.x && a.y == b.y) ? 1.0 : 0.0);
	const double k = mag * mag * (gsn + noise * noise * delta);
	const size_t o = size_t(c) * nL + r, sz = size_t(nL) * nR;
	K[o] = k;
	if (dK != nullptr)
	{
		dK[o] = k * (2.0 / mag);
		const bool diag = same && r == c;
		const double gp = same ? k - (diag ? (mag * noise) * (mag * noise) : 0.0) : k;
		dK[sz + o] = diag ? 0.0 : gp * (dx * dx / lx);
		dK[2 * sz + o] = diag ? 0.0 : gp * (dp * dp / lp);
		dK[3 * sz + o] = diag ? 2.0 * mag * mag * noise : 0.0;
	}
}

/// ComplexKernelBase (complex_kernel.cpp:134-200): K (real) and the pseudo-covariance Kt (complex), column-major.
__global__ void __launch_bounds__(256) kernel_complex_kernel(const double2* __restrict__ XL, const int nL, const double2* __restrict__ XR, const int nR, const GaussBlock gr, const GaussBlock gi, const GaussBlock gc, const double mag2, const double noise2, const int same, double* __restrict__ K, double2* __restrict__ Kt)
{
	const int r = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
	if (r >= nL)
	{
		return;
	}
	const double2 a = XL[r], b = XR[c];
	const double kr = gauss_value(gr, a, b), ki = gauss_value(gi, a, b), kc = gauss_value(gc, a, b);
	const double delta = same ? (r == c ? 1.0 : 0.0) : ((a.x == b.x && a.y == b.y) ? 1.0 : 0.0);
	const size_t o = size_t(c) * nL + r;
	K[o] = mag2 * (kr + ki + noise2 * delta);
	Kt[o] = make_double2(mag2 * (kr - ki), mag2 * (2.0 * kc));
}

/// calculate_derivative / calculate_pseudo_derivative of the complex kernel (complex_kernel.cpp:20-59, 74-132), materialised:
/// dK[p] (real) and dKt[p] (complex), p = sigma, sigma_R, l_Rx, l_Rp, sigma_I, l_Ix, l_Ip, sigma_n, column-major nL x nR each.
/// As in the reference the sub-kernel blocks carry NO sigma^2 and the noise entry is 2 sigma_n on the diagonal of a same-set
/// matrix (quirk q2); the correlation kernel's parameters are functions of the R / I ones (chain rule, :96-126).
struct ComplexDerivSpec
{
	double sigma, noise, sr, si, lr[2], li[2], lc[2];
};
__global__ void __launch_bounds__(256) kernel_complex_deriv_kernel(const double2* __restrict__ XL, const int nL, const double2* __restrict__ XR, const int nR, const GaussBlock gr, const GaussBlock gi, const GaussBlock gc, const ComplexDerivSpec sp, const int same, double* __restrict__ dK, double2* __restrict__ dKt)
{
	const int r = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
	if (r >= nL)
	{
		return;
	}
	const double2 a = XL[r], b = XR[c];
	const double kr = gauss_value(gr, a, b), ki = gauss_value(gi, a, b), kc = gauss_value(gc, a, b);
	const double delta = same ? (r == c ? 1.0 : 0.0) : ((a.x == b.x && a.y == b.y) ? 1.0 : 0.0);
	const double d2[2] = {(a.x - b.x) * (a.x - b.x), (a.y - b.y) * (a.y - b.y)};
	const size_t plane = size_t(nL) * nR, o = size_t(c) * nL + r;
	const double m2 = sp.sigma * sp.sigma;
	double D[8];
	double2 Dt[8];
	D[0] = 2.0 / sp.sigma * (m2 * (kr + ki + sp.noise * sp.noise * delta));
	Dt[0] = make_double2(2.0 / sp.sigma * (m2 * (kr - ki)), 2.0 / sp.sigma * (m2 * (2.0 * kc)));
	D[1] = 2.0 / sp.sr * kr;
	D[4] = 2.0 / sp.si * ki;
	Dt[1] = make_double2(D[1], 2.0 / sp.sr * kc);
	Dt[4] = make_double2(-D[4], 2.0 / sp.si * kc);
#pragma unroll
	for (int d = 0; d < 2; d++)
	{
		const double dc = kc * d2[d] / (sp.lc[d] * sp.lc[d] * sp.lc[d]); // derivative of the correlation kernel over its own length
		D[2 + d] = kr * d2[d] / (sp.lr[d] * sp.lr[d] * sp.lr[d]);
		D[5 + d] = ki * d2[d] / (sp.li[d] * sp.li[d] * sp.li[d]);
		Dt[2 + d] = make_double2(D[2 + d], 2.0 * (1.0 / sp.lr[d] - sp.lr[d] / (sp.lc[d] * sp.lc[d])) * kc + sp.lr[d] / sp.lc[d] * dc);
		Dt[5 + d] = make_double2(-D[5 + d], 2.0 * (1.0 / sp.li[d] - sp.li[d] / (sp.lc[d] * sp.lc[d])) * kc + sp.li[d] / sp.lc[d] * dc);
	}
	D[7] = (same && r == c) ? 2.0 * sp.noise : 0.0;
	Dt[7] = make_double2(0.0, 0.0);
#pragma unroll
	for (int p = 0; p < 8; p++)
	{
		if (dK != nullptr)
		{
			dK[p * plane + o] = D[p];
		}
		if (dKt != nullptr)
		{
			dKt[p * plane + o] = Dt[p];
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// host-side orchestration
// ---------------------------------------------------------------------------------------------------

BlockSpec real_spec(const double* th)
{
	BlockSpec s{};
	s.nb = 1;
	s.b[0][0] = GaussBlock{th[0] * th[0], 1.0 / th[1], 1.0 / th[2], th[0] * th[0] * th[3] * th[3]};
	return s;
}

struct ComplexSub
{
	double sr, si, sc;
	double lr[2], li[2], lc[2];
};
/// complex_kernel.cpp:142-157
ComplexSub complex_sub(const double* th)
{
	ComplexSub c{};
	c.sr = th[1];
	c.lr[0] = th[2];
	c.lr[1] = th[3];
	c.si = th[4];
	c.li[0] = th[5];
	c.li[1] = th[6];
	double prod = 1.0;
	for (int d = 0; d < 2; d++)
	{
		const double ss = c.lr[d] * c.lr[d] + c.li[d] * c.li[d];
		prod *= 2.0 * c.lr[d] * c.li[d] / ss;
		c.lc[d] = std::sqrt(ss / 2.0);
	}
	c.sc = std::sqrt(c.sr * c.si * prod);
	return c;
}
BlockSpec complex_spec(const double* th)
{
	const ComplexSub c = complex_sub(th);
	const double s2 = th[0] * th[0], hn = 0.5 * s2 * th[7] * th[7];
	BlockSpec s{};
	s.nb = 2;
	s.b[0][0] = GaussBlock{s2 * c.sr * c.sr, 1.0 / c.lr[0], 1.0 / c.lr[1], hn};
	s.b[1][1] = GaussBlock{s2 * c.si * c.si, 1.0 / c.li[0], 1.0 / c.li[1], hn};
	s.b[0][1] = s.b[1][0] = GaussBlock{s2 * c.sc * c.sc, 1.0 / c.lc[0], 1.0 / c.lc[1], 0.0};
	return s;
}

/// Gaussian block of a "purity auxiliary" kernel (kernel.h:285-294): magnitude mag^2 sqrt(lx lp), lengths sqrt(2) l
GaussBlock aux_block(const double mag, const double lx, const double lp, const double scale)
{
	const double m = mag * mag * std::sqrt(lx * lp);
	return GaussBlock{scale * m * m, 1.0 / (std::sqrt(2.0) * lx), 1.0 / (std::sqrt(2.0) * lp), 0.0};
}
/// mixed auxiliary kernel (complex_kernel.cpp:206-219)
GaussBlock mixed_block(const double ma, const double* la, const double mb, const double* lb, const double scale)
{
	double prod = 1.0, l[2];
	for (int d = 0; d < 2; d++)
	{
		prod *= 0.5 * (1.0 / (la[d] * la[d]) + 1.0 / (lb[d] * lb[d]));
		l[d] = std::sqrt(la[d] * la[d] + lb[d] * lb[d]);
	}
	const double m = ma * mb / std::sqrt(std::sqrt(prod));
	return GaussBlock{scale * m * m, 1.0 / l[0], 1.0 / l[1], 0.0};
}

double read_scalars(gple_ctx* ctx, const double* d_scal, const int count, double* h)
{
	GPLE_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_scal, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	GPLE_CUDA(cudaStreamSynchronize(ctx->stream));
	std::memcpy(h, ctx->h_pinned, count * sizeof(double));
	return h[0];
}

constexpr int SCAL_COUNT = 16;

/// Steps shared by both element kinds: build covariance, factorise, invert, solve, column pass.
/// Returns GPLE_OK or GPLE_ERR_NOT_SPD.
int train_common(gple_ctx* ctx, gple_model* m, const BlockSpec& spec, const DeviceArray<double>& X, const DeviceArray<double>& y, double* d_scal)
{
	const int N = int(m->N), Np = m->Np, n = m->n;
	m->X = static_cast<double*>(ctx->pool.alloc(size_t(2) * Np * sizeof(double)));
	GPLE_CUDA(cudaMemsetAsync(m->X, 0, size_t(2) * Np * sizeof(double), ctx->stream));
	GPLE_CUDA(cudaMemcpyAsync(m->X, X.dev, size_t(2) * N * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
	m->W = static_cast<double*>(ctx->pool.alloc(size_t(n) * n * sizeof(double)));
	m->v = static_cast<double*>(ctx->pool.alloc(size_t(n) * sizeof(double)));
	m->label = static_cast<double*>(ctx->pool.alloc(size_t(n) * sizeof(double)));
	m->kinv_diag = static_cast<double*>(ctx->pool.alloc(size_t(3) * n * sizeof(double)));
	double* K = ctx->ws.get<double>("train.K", size_t(n) * n);
	int* d_info = ctx->ws.get<int>("train.info", 4);
	const int tiles = n / 128;
	GPLE_LAUNCH(ctx, build_cov_lower_kernel, dim3(tiles, tiles), 256, 0, spec, reinterpret_cast<const double2*>(m->X), N, Np, n, K);
	{
		const unsigned long long before = ctx->launches;
		ProfScope prof(ctx, GPLE_PROF_FACTORISE, 2.0 * double(n) * n * n / 3.0, 0);
		potrf_trtri(ctx, K, m->W, n, d_info);
		if (ctx->prof_on)
		{
			ctx->prof[GPLE_PROF_FACTORISE].launches += ctx->launches - before;
		}
	}
	GPLE_LAUNCH(ctx, label_kernel, 1, 1024, 0, reinterpret_cast<const double2*>(y.dev), N, Np, m->is_complex, m->label, d_scal);
	double* z = ctx->ws.get<double>("train.z", size_t(n));
	GPLE_LAUNCH(ctx, trmv_lower_kernel, (n + 7) / 8, 256, 0, m->W, n, m->label, z);
	// row slabs of 128 (capped at 64 slabs): enough CTAs in flight to cover the load latency of this O(n^2) pass
	const int slabs = std::max(1, std::min(64, n / 128));
	const int slab = int(round_up(size_t((n + slabs - 1) / slabs), 128));
	const int nslab = (n + slab - 1) / slab;
	double* part = ctx->ws.get<double>("train.colpart", size_t(nslab) * 3 * n);
	GPLE_LAUNCH(ctx, col_pass_kernel, dim3(n / 128, nslab), 128, 0, m->W, n, Np, m->is_complex, z, slab, part);
	GPLE_LAUNCH(ctx, col_reduce_kernel, (n + 255) / 256, 256, 0, part, n, nslab, m->v, m->kinv_diag, m->is_complex ? m->kinv_diag + n : nullptr);
	if (ctx->refine_solution)
	{
		// one step of iterative refinement of v = K^-1 y' (see residual_kernel)
		double* r = ctx->ws.get<double>("train.r", size_t(n));
		GPLE_LAUNCH(ctx, residual_kernel, (n + 7) / 8, 256, 0, spec, reinterpret_cast<const double2*>(m->X), N, Np, n, m->label, m->v, r);
		GPLE_LAUNCH(ctx, trmv_lower_kernel, (n + 7) / 8, 256, 0, m->W, n, r, z);
		GPLE_LAUNCH(ctx, col_pass_kernel, dim3(n / 128, nslab), 128, 0, m->W, n, Np, 0, z, slab, part);
		GPLE_LAUNCH(ctx, refine_add_kernel, (n + 255) / 256, 256, 0, part, n, nslab, m->v);
	}
	return GPLE_OK; // nothing has been waited for: the pivot status comes back with the scalars (finish_train)
}

/// The ONE host synchronisation of a training call: the first `count` scalars and the factorisation's pivot status.
/// Returns GPLE_OK or GPLE_ERR_NOT_SPD.
int finish_train(gple_ctx* ctx, const double* d_scal, const int count, double* h)
{
	const int* d_info = ctx->ws.get<int>("train.info", 4);
	GPLE_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_scal, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	GPLE_CUDA(cudaMemcpyAsync(ctx->h_pinned + 64, d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	GPLE_CUDA(cudaStreamSynchronize(ctx->stream));
	std::memcpy(h, ctx->h_pinned, count * sizeof(double));
	int info = 0;
	std::memcpy(&info, ctx->h_pinned + 64, sizeof(int));
	return info == 0 ? GPLE_OK : GPLE_ERR_NOT_SPD;
}

/// w^T K1 w into *d_out, enqueued only
void quadform_enqueue(gple_ctx* ctx, const gple_model* m, const QuadSpec& qs, const double* w, double* d_out)
{
	const int tiles = m->n / 128;
	double* part = ctx->ws.get<double>("train.quadpart", size_t(tiles) * tiles);
	GPLE_LAUNCH(ctx, quadform_kernel, dim3(tiles, tiles), 256, 0, qs, reinterpret_cast<const double2*>(m->X), int(m->N), m->Np, w, part);
	GPLE_LAUNCH(ctx, sum_kernel, 1, 1024, 0, part, tiles * tiles, d_out);
}
double quadform(gple_ctx* ctx, const gple_model* m, const QuadSpec& qs, const double* w, double* d_out)
{
	quadform_enqueue(ctx, m, qs, w, d_out);
	double h = 0.0;
	read_scalars(ctx, d_out, 1, &h);
	return h;
}

constexpr int CHUNK_ROWS = 148 * 128;

void require_rows(long long total_rows)
{
	if (total_rows >= (1ll << 31) - 256)
	{
		throw ArgError{"gated prediction: more than 2^31 composite rows in one call"};
	}
}

} // namespace

namespace
{
template <typename C>
void var_gemm_attribute()
{
	GPLE_CUDA(cudaFuncSetAttribute(var_gemm_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(C::SMEM_BYTES)));
}
} // namespace

/// Per-device function attributes of every kernel with opt-in shared memory (called by gple_ctx_create on the context's device).
void gpr_setup_attributes()
{
	chol_setup_attributes();
	var_gemm_attribute<VarCfg0>();
	var_gemm_attribute<VarCfg1>();
	var_gemm_attribute<VarCfg2>();
	var_gemm_attribute<VarCfg3>();
	var_gemm_attribute<VarCfg4>();
	var_gemm_attribute<VarCfg5>();
	var_gemm_attribute<VarCfg6>();
	mc_setup_attributes();
}

void free_model(gple_ctx* ctx, gple_model* m)
{
	if (m == nullptr)
	{
		return;
	}
	gple_ctx* owner = m->owner != nullptr ? m->owner : ctx;
	for (double* p : {m->X, m->W, m->v, m->label, m->kinv_diag, m->Kinv, m->dv})
	{
		owner->pool.free(p); // reuse is ordered on the owner's stream; callers sync before handing a model to another context
	}
	delete m;
}

void ensure_full_inverse(gple_ctx* ctx, gple_model* m)
{
	if (m->Kinv != nullptr)
	{
		return;
	}
	const int n = m->n;
	m->Kinv = static_cast<double*>((m->owner != nullptr ? m->owner : ctx)->pool.alloc(size_t(n) * n * sizeof(double))); // freed through the owner's pool
	double* U = ctx->ws.get<double>("inv.U", size_t(n) * n);
	GPLE_LAUNCH(ctx, transpose_kernel, dim3(n / 32, n / 32), dim3(32, 8), 0, m->W, U, n);
	gemm::GemmArgs g{};
	g.A = U;
	g.B = U;
	g.C = m->Kinv;
	g.lda = g.ldb = g.ldc = size_t(n);
	g.M = g.N = g.K = n;
	g.alpha = 1.0;
	g.beta = 0.0;
	g.tri = gemm::A_UPPER | gemm::B_UPPER_NT;
	gemm_nt(ctx, g);
}

int train_real(gple_ctx* ctx, const double* X_, const double* y_, size_t N, const double* theta, unsigned flags, gple_model** out_model, gple_real_scalars* out)
{
	DeviceArray<double> X(ctx, X_, 2 * N, false), y(ctx, y_, 2 * N, false);
	gple_model* m = new gple_model();
	m->owner = ctx;
	m->is_complex = 0;
	m->N = N;
	m->Np = int(round_up(N, 128));
	m->n = m->Np;
	std::memcpy(m->theta, theta, 4 * sizeof(double));
	m->flags = flags;
	m->prior = theta[0] * theta[0] * (1.0 + theta[3] * theta[3]); // kernel.cpp:512 (quirk q3)
	double* d_scal = ctx->ws.get<double>("train.scal", SCAL_COUNT);
	int status = GPLE_OK;
	try
	{
		train_common(ctx, m, real_spec(theta), X, y, d_scal);
		GPLE_LAUNCH(ctx, real_scalars_kernel, 1, 1024, 0, reinterpret_cast<const double2*>(m->X), m->label, m->v, m->kinv_diag, int(N), d_scal);
		if (flags & GPLE_CALC_AVERAGE)
		{
			QuadSpec qs{};
			qs.nb = 1;
			qs.g[0][0][0] = aux_block(theta[0], theta[1], theta[2], 1.0);
			quadform_enqueue(ctx, m, qs, m->v, d_scal + 8);
		}
		double h[SCAL_COUNT];
		status = finish_train(ctx, d_scal, 9, h); // the one synchronisation of the call
		m->rescale = h[0];
		const double nan = std::nan("");
		gple_real_scalars r{};
		r.rescale = h[0];
		r.error = (flags & GPLE_CALC_ERROR) ? h[1] : nan;
		const double w = h[5] / double(N); // kernel.h:167-179
		r.magnitude = std::sqrt(std::fabs(w));
		r.population = r.first_order[0] = r.first_order[1] = r.purity = nan;
		for (int p = 0; p < 4; p++)
		{
			r.d_error[p] = r.d_population[p] = r.d_purity[p] = nan;
		}
		if (flags & GPLE_CALC_AVERAGE)
		{
			const double f = 2.0 * M_PI * theta[0] * theta[0] * theta[1] * theta[2];
			r.population = f * h[2] / r.rescale;	   // kernel.cpp:286-297
			r.first_order[0] = f * h[3] / r.rescale; // kernel.cpp:298-312
			r.first_order[1] = f * h[4] / r.rescale;
			const double quad = h[8];
			r.purity = (2.0 * M_PI) * M_PI * quad / (r.rescale * r.rescale); // kernel.cpp:325-335
		}
		if (flags & GPLE_CALC_DERIVATIVE)
		{
			real_derivatives(ctx, m, flags, h, &r);
		}
		if (status == GPLE_ERR_NOT_SPD)
		{
			r.error = r.population = r.purity = nan;
		}
		if (out != nullptr)
		{
			*out = r;
		}
	}
	catch (...)
	{
		free_model(ctx, m);
		throw;
	}
	*out_model = m;
	return status;
}

int train_complex(gple_ctx* ctx, const double* X_, const double* y_, size_t N, const double* theta, unsigned flags, gple_model** out_model, gple_complex_scalars* out)
{
	DeviceArray<double> X(ctx, X_, 2 * N, false), y(ctx, y_, 2 * N, false);
	gple_model* m = new gple_model();
	m->owner = ctx;
	m->is_complex = 1;
	m->N = N;
	m->Np = int(round_up(N, 128));
	m->n = 2 * m->Np;
	std::memcpy(m->theta, theta, 8 * sizeof(double));
	m->flags = flags;
	m->prior = theta[0] * theta[0] * (theta[1] * theta[1] + theta[4] * theta[4] + theta[7] * theta[7]); // complex_kernel.cpp:632
	double* d_scal = ctx->ws.get<double>("train.scal", SCAL_COUNT);
	int status = GPLE_OK;
	try
	{
		train_common(ctx, m, complex_spec(theta), X, y, d_scal);
		GPLE_LAUNCH(ctx, complex_scalars_kernel, 1, 1024, 0, m->label, m->v, m->kinv_diag, m->kinv_diag + m->n, int(N), m->Np, d_scal);
		if (flags & GPLE_CALC_AVERAGE)
		{
			// complex_kernel.cpp:357-377 with v = (wr + i wi) / 2:
			//   v^H K1 v + Re v^T K2 v = 1/2 [ wr^T (KR' + KC') wr + wi^T (KI' + KC') wi + 2 wr^T (KRC + KIC) wi ]
			const ComplexSub c = complex_sub(theta);
			QuadSpec qs{};
			qs.nb = 2;
			qs.g[0][0][0] = aux_block(c.sr, c.lr[0], c.lr[1], 1.0);
			qs.g[0][0][1] = aux_block(c.sc, c.lc[0], c.lc[1], 1.0);
			qs.g[1][1][0] = aux_block(c.si, c.li[0], c.li[1], 1.0);
			qs.g[1][1][1] = aux_block(c.sc, c.lc[0], c.lc[1], 1.0);
			qs.g[0][1][0] = qs.g[1][0][0] = mixed_block(c.sr, c.lr, c.sc, c.lc, 1.0);
			qs.g[0][1][1] = qs.g[1][0][1] = mixed_block(c.si, c.li, c.sc, c.lc, 1.0);
			quadform_enqueue(ctx, m, qs, m->v, d_scal + 8);
		}
		double h[SCAL_COUNT];
		status = finish_train(ctx, d_scal, 9, h); // the one synchronisation of the call
		m->rescale = h[0];
		const double nan = std::nan("");
		gple_complex_scalars r{};
		r.rescale = h[0];
		r.error = (flags & GPLE_CALC_ERROR) ? h[1] : nan;
		r.magnitude = std::sqrt(std::fabs(h[2] / double(N))); // complex_kernel.h:192-204
		r.purity = nan;
		for (int p = 0; p < 8; p++)
		{
			r.d_error[p] = r.d_purity[p] = nan;
		}
		if (flags & GPLE_CALC_AVERAGE)
		{
			const double quad = 0.5 * h[8];
			const double s4 = theta[0] * theta[0] * theta[0] * theta[0];
			r.purity = (2.0 * M_PI) * 2.0 * M_PI * s4 * quad / (r.rescale * r.rescale);
		}
		if (flags & GPLE_CALC_DERIVATIVE)
		{
			complex_derivatives(ctx, m, flags, h, &r);
		}
		if (status == GPLE_ERR_NOT_SPD)
		{
			r.error = r.purity = nan;
		}
		if (out != nullptr)
		{
			*out = r;
		}
	}
	catch (...)
	{
		free_model(ctx, m);
		throw;
	}
	*out_model = m;
	return status;
}

/// One stage of the bound-gated variance: the tiles up to block `re_end` of the (Re) rows and up to block `im_end` of the Im
/// rows (complex element; cumulative, in 128-blocks of training points).
struct GateStage
{
	int re_end, im_end;
};
constexpr int MAX_GATE_STAGES = 16;

/// The stages before the last one (which always completes the tile set).
///  * gple_ctx_set_gate_schedule: explicit;
///  * GPLE_OPT_GATE_STAGE_TILES >= 0: the round-1 / early round-2 schedule (0: single stage; t: [t (+ t/4 Im blocks)], then
///    GPLE_OPT_GATE_STAGE2_TILES);
///  * automatic: Re boundaries 1, then a geometric progression (ratio about 2.4) up to 21/32 of the blocks -- 1, 2, 5, 10 of 16;
///    1, 3, 8, 21 of 32; 1, 2, 6, 14, 35, 84 of 128: the first block alone decides more than half of the open queries for 1/136
///    of the full product, and each later boundary sits where the rows it removes pay for the extra pass (dynamic programme
///    over the survival curves of the bench workload, profiles/gate_schedule_sim.py, profiles/r02_gate_schedule.md).
///    Complex element: no Im block in the first stage (an Im tile costs a product over ALL Re columns), then one Im block per
///    five Re blocks, and a stage with all Re blocks and 3/8 of the Im blocks before the last one.
/// cumulative, strictly growing, and something left for the last stage
std::vector<GateStage> normalise_schedule(const std::vector<GateStage>& raw, const bool is_complex, const int Th)
{
	std::vector<GateStage> out;
	const int Ti = is_complex ? Th : 0;
	int re = 0, im = 0;
	for (const GateStage& g : raw)
	{
		const int r1 = std::max(re, std::min(g.re_end, Th)), i1 = std::max(im, std::min(g.im_end, Ti));
		if ((r1 > re || i1 > im) && (r1 < Th || i1 < Ti) && int(out.size()) < MAX_GATE_STAGES - 1)
		{
			out.push_back(GateStage{r1, i1});
			re = r1;
			im = i1;
		}
	}
	return out;
}

std::vector<GateStage> automatic_schedule(const bool is_complex, const int Th)
{
	std::vector<GateStage> raw;
	// the first block alone, then boundaries in geometric progression (ratio ~2.4) up to `late` = 21/32 of the blocks
	const int late = 21 * Th / 32;
	if (late >= 1)
	{
		raw.push_back(GateStage{1, 0});
		const int k = std::max(1, int(std::lround(std::log(double(late)) / std::log(2.4))));
		for (int i = 1; i <= k; i++)
		{
			const int b = std::max(1, int(std::lround(std::pow(double(late), double(i) / k))));
			raw.push_back(GateStage{b, std::max(1, b / 5)});
		}
	}
	if (is_complex)
	{
		raw.push_back(GateStage{Th, std::max(1, 3 * Th / 8)});
	}
	return normalise_schedule(raw, is_complex, Th);
}

std::vector<GateStage> gate_schedule(const gple_ctx* ctx, const bool is_complex, const int Th)
{
	std::vector<GateStage> raw;
	const std::vector<int>& ex = ctx->gate_schedule[is_complex ? 1 : 0];
	if (!ex.empty())
	{
		for (size_t k = 0; k + 1 < ex.size(); k += 2)
		{
			raw.push_back(GateStage{ex[k], ex[k + 1]});
		}
	}
	else if (ctx->gate_stage_tiles >= 0)
	{
		const int stage = std::min(ctx->gate_stage_tiles, Th);
		if (stage > 0 && stage < Th)
		{
			const int stage_im = std::min(ctx->gate_stage_tiles_im >= 0 ? ctx->gate_stage_tiles_im : std::max(1, stage / 4), Th);
			raw.push_back(GateStage{stage, stage_im});
			const int stage2 = ctx->gate_stage2_tiles >= 0 ? ctx->gate_stage2_tiles : std::max(stage + 1, 5 * Th / 8);
			raw.push_back(GateStage{stage2, stage_im});
		}
	}
	else
	{
		return automatic_schedule(is_complex, Th);
	}
	return normalise_schedule(raw, is_complex, Th);
}

/// gple_gate_schedule_automatic of the C-ABI (host only)
int gate_schedule_automatic_host(const int complex_element, const int blocks, int* re_end, int* im_end, const int capacity)
{
	const std::vector<GateStage> s = automatic_schedule(complex_element != 0, blocks);
	for (size_t k = 0; k < s.size() && int(k) < capacity; k++)
	{
		re_end[k] = s[k].re_end;
		im_end[k] = s[k].im_end;
	}
	return int(s.size());
}

/// Batched prediction of `Q` points on the device.  d_pred / d_cut hold nb doubles per point.
///
/// Two schedules:
///  * every variance (var_out requested, or GPLE_OPT_GATED_VARIANCE off): chunks of 18944 composite rows, per chunk
///    K* rows + fused mean -> triangular variance GEMM -> gate;
///  * bound-gated (only the cutoff prediction is wanted): one mean-only pass over all queries (K* never stored),
///    classification by |f|^2 >= 4 k**, then K* rows + variance GEMM in FULL chunks over the compacted list of the
///    remaining queries, so the GEMM always runs at full occupancy.
void predict_device(gple_ctx* ctx, const gple_model* m, const double* d_Xq, size_t Q, const double* d_yq, double* d_pred, double* d_var, double* d_cut, double* d_err)
{
	const int nb = m->is_complex ? 2 : 1;
	const BlockSpec spec = m->is_complex ? complex_spec(m->theta) : real_spec(m->theta);
	const long long total_rows = (long long)(Q)*nb;
	const int n = m->n;
	const int max_rows = int(std::min<long long>(CHUNK_ROWS, (long long)round_up(size_t(total_rows), 128)));
	double* A = ctx->ws.get<double>("pred.A", size_t(max_rows) * n);
	const double2* Xq2 = reinterpret_cast<const double2*>(d_Xq);
	const double2* Xt2 = reinterpret_cast<const double2*>(m->X);
	if (d_err != nullptr)
	{
		GPLE_CUDA(cudaMemsetAsync(d_err, 0, sizeof(double), ctx->stream));
	}
	const bool want_gate = d_var != nullptr || d_cut != nullptr;
	if (d_var == nullptr && d_cut != nullptr && ctx->gated_variance)
	{
		const size_t rows_pad = round_up(size_t(total_rows), 128);
		double* pred = ctx->ws.get<double>("pred.f_all", rows_pad);
		unsigned char* coincident = ctx->ws.get<unsigned char>("pred.coincident", rows_pad);
		int* gate_idx = ctx->ws.get<int>("pred.gate_idx", rows_pad);
		int* gate_slot = ctx->ws.get<int>("pred.gate_slot", rows_pad);
		int* gate_cnt = ctx->ws.get<int>("pred.gate_cnt", 4);
		require_rows(total_rows);
		{
			ProfScope prof(ctx, GPLE_PROF_MEAN, double(total_rows) * double(m->N) * nb, 1);
			launch_kmean(ctx, spec, Xq2, total_rows, Xt2, int(m->N), m->Np, m->v, pred, coincident);
		}
		const double noise = m->theta[0] * m->theta[0] * m->theta[m->is_complex ? 7 : 3] * m->theta[m->is_complex ? 7 : 3];
		const double half_noise = noise >= 1e-9 * m->prior ? 0.5 * noise : -1.0;
		GPLE_CUDA(cudaMemsetAsync(gate_cnt, 0, 4 * sizeof(int), ctx->stream));
		GPLE_LAUNCH(ctx, classify_kernel, unsigned((Q + 255) / 256), 256, 0, pred, coincident, int(Q), nb, 4.0 * m->prior, half_noise, gate_idx, gate_slot, gate_cnt);
		GPLE_CUDA(cudaMemcpyAsync(ctx->h_pinned + 200, gate_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		GPLE_CUDA(cudaStreamSynchronize(ctx->stream));
		int counts[2] = {0, 0};
		std::memcpy(counts, ctx->h_pinned + 200, sizeof(counts));
		const int count = counts[0];
		ctx->gate_rows_total += (unsigned long long)total_rows;
		ctx->gate_rows_variance += (unsigned long long)count;
		ctx->gate_rows_zero += (unsigned long long)counts[1];
		// ---- the staged bound: k** - sum Z^2 over a subset of Z's columns is the posterior variance given only those observations,
		// an upper bound of the variance that is already close to the noise floor inside the point cloud.  The listed rows go
		// through the tile sets of gate_schedule() one after the other; after each but the last, the queries whose gate the bound
		// decides (== 1) leave the list.  W is lower triangular, so the columns of the first training blocks are the cheap ones.
		const int T = n / 128, Th = m->is_complex ? T / 2 : T, Ti = m->is_complex ? Th : 0;
		const std::vector<GateStage> sched = gate_schedule(ctx, m->is_complex, Th);
		double* qs = ctx->ws.get<double>("pred.q_all", round_up(size_t(count), 128) + 128);
		double* qsum = ctx->ws.get<double>("pred.qsum", rows_pad);
		int* lists[2] = {ctx->ws.get<int>("pred.gate_idx2", round_up(size_t(count), 128)), ctx->ws.get<int>("pred.gate_idx3", round_up(size_t(count), 128))};
		int* stage_cnt = ctx->ws.get<int>("pred.gate_stage_cnt", MAX_GATE_STAGES);
		GPLE_CUDA(cudaMemsetAsync(stage_cnt, 0, MAX_GATE_STAGES * sizeof(int), ctx->stream));
		auto sweep = [&](const int* list, const int list_count, const TileSet& ts, double* qout)
		{
			// the triangular products of tile t read K* columns [0, (t + 1) * 128): generate no more than the set needs, packed at
			// that pitch, so that a launch of an early stage (few columns) takes up to eight waves of row blocks
			const int cols = 128 * (1 + std::max(ts.c0 > 0 ? ts.b0 + ts.c0 - 1 : 0, ts.c1 > 0 ? ts.b1 + ts.c1 - 1 : 0));
			const long long chunk = (long long)max_rows * std::max(1, std::min(8, n / cols));
			for (long long c0 = 0; c0 < list_count; c0 += chunk)
			{
				const int rows_real = int(std::min<long long>(chunk, list_count - c0));
				const int rows = int(round_up(size_t(rows_real), 128));
				{
					ProfScope prof(ctx, GPLE_PROF_KERNEL_BUILD, 8.0 * double(rows) * cols, 1);
					GPLE_LAUNCH(ctx, kstar_kernel, (rows + 7) / 8, 256, 0, spec, Xq2, (long long)Q, c0, rows, list, (long long)list_count, Xt2, int(m->N), m->Np, cols, cols, m->v, A, nullptr);
				}
				ProfScope prof(ctx, GPLE_PROF_VARIANCE_GEMM, var_gemm_flops(ts, rows), 1);
				launch_var_gemm(ctx, g_var_variant, A, cols, m->W, n, rows, qout + c0, ts);
			}
		};
		const int* list = gate_idx;
		int list_count = count, re0 = 0, im0 = 0;
		bool first = true;
		for (size_t s = 0; s <= sched.size() && list_count > 0; s++)
		{
			const bool last = s == sched.size();
			const int re1 = last ? Th : sched[s].re_end, im1 = last ? Ti : sched[s].im_end;
			// the tiles [re0, re1) of the (Re) rows and [Th + im0, Th + im1) of the Im rows, as (up to) two runs
			const TileSet ts = re1 > re0 ? TileSet{re0, re1 - re0, Th + im0, im1 - im0} : TileSet{Th + im0, im1 - im0, 0, 0};
			if (last)
			{
				ctx->gate_rows_stage_b += (unsigned long long)list_count; // rows that needed the full variance
			}
			if (ts.count() > 0)
			{
				sweep(list, list_count, ts, qs);
				if (last)
				{
					GPLE_LAUNCH(ctx, accumulate_stage_kernel, unsigned((list_count + 255) / 256), 256, 0, list, qs, qsum, int(first), list_count);
				}
				else
				{
					int* next = lists[s & 1];
					GPLE_LAUNCH(ctx, classify_stage_kernel, unsigned((list_count / nb + 255) / 256), 256, 0, pred, list, qs, qsum, int(first), list_count / nb, nb, m->prior, next, gate_slot, stage_cnt + s);
					GPLE_CUDA(cudaMemcpyAsync(ctx->h_pinned + 204, stage_cnt + s, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
					GPLE_CUDA(cudaStreamSynchronize(ctx->stream));
					std::memcpy(&list_count, ctx->h_pinned + 204, sizeof(int));
					list = next;
				}
				first = false;
			}
			re0 = re1;
			im0 = im1;
		}
		if (m->is_complex)
		{
			GPLE_LAUNCH(ctx, finalize_complex_kernel, unsigned((total_rows / 2 + 255) / 256), 256, 0, pred, qsum, gate_slot, int(total_rows), 0ll, (long long)Q, m->prior, m->rescale, reinterpret_cast<double2*>(d_pred), d_var, reinterpret_cast<double2*>(d_cut));
		}
		else
		{
			GPLE_LAUNCH(ctx, finalize_real_kernel, unsigned((total_rows + 255) / 256), 256, 0, pred, qsum, gate_slot, int(total_rows), 0ll, (long long)Q, m->prior, m->rescale, d_pred, d_var, d_cut);
		}
		if (d_err != nullptr && d_yq != nullptr)
		{
			for (long long row0 = 0; row0 < total_rows; row0 += CHUNK_ROWS)
			{
				const int rows = int(std::min<long long>(CHUNK_ROWS, total_rows - row0));
				GPLE_LAUNCH(ctx, sqerr_kernel, 1, 1024, 0, pred + row0, rows, row0, total_rows, d_yq, nb, nb, m->rescale, d_err);
			}
		}
		return;
	}
	double* pred = ctx->ws.get<double>("pred.f", size_t(max_rows));
	double* q = ctx->ws.get<double>("pred.q", size_t(max_rows) * MAX_VAR_SPLITS);
	for (long long row0 = 0; row0 < total_rows; row0 += CHUNK_ROWS)
	{
		const int rows_real = int(std::min<long long>(CHUNK_ROWS, total_rows - row0));
		const int rows = int(round_up(size_t(rows_real), 128));
		if (want_gate)
		{
			{
				ProfScope prof(ctx, GPLE_PROF_KERNEL_BUILD, 8.0 * double(rows) * n, 1);
				GPLE_LAUNCH(ctx, kstar_kernel, (rows + 7) / 8, 256, 0, spec, Xq2, (long long)Q, row0, rows, nullptr, 0ll, Xt2, int(m->N), m->Np, n, n, m->v, A, pred);
			}
			ProfScope prof(ctx, GPLE_PROF_VARIANCE_GEMM, double(rows) * n * (double(n) + 128.0), 1);
			launch_var_gemm(ctx, g_var_variant, A, n, m->W, n, rows, q, TileSet{0, n / 128, 0, 0});
		}
		else
		{
			// mean only (e.g. the loss evaluation of opt.cpp:441-482 without gradient): K* is never stored
			ProfScope prof(ctx, GPLE_PROF_MEAN, double(rows_real) * double(m->N) * nb, 1);
			launch_kmean(ctx, spec, Xq2 + row0 / nb, rows_real, Xt2, int(m->N), m->Np, m->v, pred, nullptr);
			GPLE_CUDA(cudaMemsetAsync(q, 0, size_t(rows) * sizeof(double), ctx->stream));
		}
		if (m->is_complex)
		{
			GPLE_LAUNCH(ctx, finalize_complex_kernel, (rows / 2 + 255) / 256, 256, 0, pred, q, nullptr, rows, row0, (long long)Q, m->prior, m->rescale, reinterpret_cast<double2*>(d_pred), d_var, reinterpret_cast<double2*>(d_cut));
		}
		else
		{
			GPLE_LAUNCH(ctx, finalize_real_kernel, (rows + 255) / 256, 256, 0, pred, q, nullptr, rows, row0, (long long)Q, m->prior, m->rescale, d_pred, d_var, d_cut);
		}
		if (d_err != nullptr && d_yq != nullptr)
		{
			GPLE_LAUNCH(ctx, sqerr_kernel, 1, 1024, 0, pred, rows, row0, total_rows, d_yq, nb, nb, m->rescale, d_err);
		}
	}
}

int set_variance_gemm_variant(int variant)
{
	if (variant < 0 || variant >= NUM_VAR_VARIANTS)
	{
		return -1;
	}
	g_var_variant = variant;
	return 0;
}

/// Times `iters` launches of one variant on synthetic operands (tuning / roofline helper)
double bench_variance_gemm(gple_ctx* ctx, int variant, int rows, int n, int iters)
{
	double* A = ctx->ws.get<double>("pred.A", size_t(rows) * n);
	double* W = ctx->ws.get<double>("bench.W", size_t(n) * n);
	double* q = ctx->ws.get<double>("pred.q", size_t(rows) * MAX_VAR_SPLITS);
	GPLE_CUDA(cudaMemsetAsync(A, 0, size_t(rows) * n * sizeof(double), ctx->stream));
	GPLE_CUDA(cudaMemsetAsync(W, 0, size_t(n) * n * sizeof(double), ctx->stream));
	launch_var_gemm(ctx, variant, A, n, W, n, rows, q, TileSet{0, n / 128, 0, 0});
	cudaEvent_t e0, e1;
	GPLE_CUDA(cudaEventCreate(&e0));
	GPLE_CUDA(cudaEventCreate(&e1));
	GPLE_CUDA(cudaEventRecord(e0, ctx->stream));
	for (int i = 0; i < iters; i++)
	{
		launch_var_gemm(ctx, variant, A, n, W, n, rows, q, TileSet{0, n / 128, 0, 0});
	}
	GPLE_CUDA(cudaEventRecord(e1, ctx->stream));
	GPLE_CUDA(cudaEventSynchronize(e1));
	float ms = 0.f;
	GPLE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return double(ms) / iters;
}

void kernel_real_device(gple_ctx* ctx, const double* XL, int nL, const double* XR, int nR, const double* th, int same, double* K, double* dK)
{
	GPLE_LAUNCH(ctx, kernel_real_kernel, dim3((nL + 255) / 256, nR), 256, 0, reinterpret_cast<const double2*>(XL), nL, reinterpret_cast<const double2*>(XR), nR, th[0], th[1], th[2], th[3], same, K, dK);
}

void kernel_complex_device(gple_ctx* ctx, const double* XL, int nL, const double* XR, int nR, const double* th, int same, double* K, double* Kt)
{
	const ComplexSub c = complex_sub(th);
	const GaussBlock gr{c.sr * c.sr, 1.0 / c.lr[0], 1.0 / c.lr[1], 0.0}, gi{c.si * c.si, 1.0 / c.li[0], 1.0 / c.li[1], 0.0}, gc{c.sc * c.sc, 1.0 / c.lc[0], 1.0 / c.lc[1], 0.0};
	GPLE_LAUNCH(ctx, kernel_complex_kernel, dim3((nL + 255) / 256, nR), 256, 0, reinterpret_cast<const double2*>(XL), nL, reinterpret_cast<const double2*>(XR), nR, gr, gi, gc, th[0] * th[0], th[7] * th[7], same, K, reinterpret_cast<double2*>(Kt));
}

void kernel_complex_derivatives_device(gple_ctx* ctx, const double* XL, int nL, const double* XR, int nR, const double* th, int same, double* dK, double* dKt)
{
	const ComplexSub c = complex_sub(th);
	const GaussBlock gr{c.sr * c.sr, 1.0 / c.lr[0], 1.0 / c.lr[1], 0.0}, gi{c.si * c.si, 1.0 / c.li[0], 1.0 / c.li[1], 0.0}, gc{c.sc * c.sc, 1.0 / c.lc[0], 1.0 / c.lc[1], 0.0};
	const ComplexDerivSpec sp{th[0], th[7], c.sr, c.si, {c.lr[0], c.lr[1]}, {c.li[0], c.li[1]}, {c.lc[0], c.lc[1]}};
	GPLE_LAUNCH(ctx, kernel_complex_deriv_kernel, dim3((nL + 255) / 256, nR), 256, 0, reinterpret_cast<const double2*>(XL), nL, reinterpret_cast<const double2*>(XR), nR, gr, gi, gc, sp, same, dK, reinterpret_cast<double2*>(dKt));
}

} // namespace gple
