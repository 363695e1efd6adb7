// FP64 tensor-core (DMMA) GEMM building blocks for sm_100a.
//
// On B200 every f64 `mma.sync` shape lowers to the native DMMA.8x8x4 (checked with cuobjdump), and
// tcgen05 has no f64 kind, so the FP64 tensor path is warp-level m8n8k4 with register accumulators.
// CTA tile 128 x 128 x BK, WARPS_M x WARPS_N warps, operands staged in shared memory by a multi-stage
// cp.async (LDGSTS) pipeline.  Shared-memory pitches are chosen so that every fragment load (8 rows x 4 k,
// one double per lane) hits 16 distinct 8-byte banks per half-warp: pitch = 4 (mod 16) doubles.
//
// All matrices are ROW-MAJOR with leading dimension `ld` and padded to multiples of 128 (symmetric
// matrices make the reference's column-major layout identical).  Two operand forms:
//   NT: C[m][n] (+)= alpha * sum_k A[m][k] * B[n][k]      (both operands K-contiguous)
//   NN: C[m][n] (+)= alpha * sum_k A[m][k] * B[k][n]      (B is N-contiguous)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace gple
{
namespace gemm
{
/// Tile configuration.  CTA tile BM_ x BN_ x BK_ (BK_ in {16, 32}); warp tile = (BM / WARPS_M_) x (BN / WARPS_N_).
template <int BK_, int STAGES_, int WARPS_M_, int WARPS_N_, int BM_ = 128, int BN_ = 128>
struct Config
{
	static constexpr int BM = BM_, BN = BN_;
	static constexpr int BK = BK_, STAGES = STAGES_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
	static constexpr int THREADS = WARPS_M * WARPS_N * 32;
	static constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N; // warp tile
	static constexpr int MI = WTM / 8, NJ = WTN / 8;			 // DMMA tiles per warp
	static constexpr int KP = BK + 4;							 // pitch of a K-contiguous tile [rows][KP]
	static constexpr int NP = BN + 4;							 // pitch of an N-contiguous B tile [BK][NP]
	static constexpr int A_STAGE = BM * KP;						 // doubles
	static constexpr int B_STAGE_NT = BN * KP, B_STAGE_NN = BK * NP;
	static constexpr int B_STAGE = B_STAGE_NT > B_STAGE_NN ? B_STAGE_NT : B_STAGE_NN;
	static constexpr size_t SMEM_BYTES = size_t(STAGES) * (A_STAGE + B_STAGE) * sizeof(double);
	static_assert(KP % 16 == 4 && NP % 16 == 4, "pitch must be 4 mod 16 doubles for conflict-free fragment loads");
};
/// general GEMMs (Cholesky / inverse / gradients): 128 x 128 tiles when the grid fills the chip ...
using DefaultConfig = Config<16, 4, 2, 4>;
/// ... and 64 x 64 tiles (4 warps, 4x more CTAs, 4x shorter per-tile latency) for the small, latency-bound
/// GEMMs on the critical path of the recursive factorisation
using SmallConfig = Config<16, 4, 2, 2, 64, 64>;
/// 32-row strips of a 128-wide panel: the in-place panel solve A21 <- A21 L^-T (one CTA must own a whole row strip)
using StripConfig = Config<16, 4, 1, 4, 32, 128>;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
	const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit()
{
	asm volatile("cp.async.commit_group;\n" ::);
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

/// D(8x8) += A(8x4, row) * B(4x8, col).  Lane = 4 * g + t: a = A[g][t], b = B[t][g], c = C[g][2t .. 2t+1].
__device__ __forceinline__ void dmma884(double (&c)[2], const double a, const double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

/// Copy a K-contiguous ROWS x BK tile of a row-major matrix into [ROWS][KP].
template <typename C, int ROWS>
__device__ __forceinline__ void load_tile_kmajor(double* s, const double* __restrict__ g, const size_t ld, const int tid)
{
	constexpr int CH = C::BK / 2; // 16-byte chunks per row
#pragma unroll
	for (int i = 0; i < (ROWS * CH) / C::THREADS; i++)
	{
		const int c = tid + i * C::THREADS;
		const int row = c / CH, ch = c % CH;
		cp_async16(s + row * C::KP + ch * 2, g + size_t(row) * ld + ch * 2);
	}
}
/// Copy an N-contiguous BK x BN tile (k rows) into [BK][NP].
template <typename C>
__device__ __forceinline__ void load_tile_nmajor(double* s, const double* __restrict__ g, const size_t ld, const int tid)
{
	constexpr int CH = C::BN / 2;
#pragma unroll
	for (int i = 0; i < (C::BK * CH) / C::THREADS; i++)
	{
		const int c = tid + i * C::THREADS;
		const int row = c / CH, ch = c % CH;
		cp_async16(s + row * C::NP + ch * 2, g + size_t(row) * ld + ch * 2);
	}
}

/// One BK-deep stage of the warp tile: BK / 4 k-steps x (MI + NJ fragment loads, MI * NJ DMMA).
template <typename C, bool B_NN>
__device__ __forceinline__ void compute_stage(double (&acc)[C::MI][C::NJ][2], const double* __restrict__ As, const double* __restrict__ Bs, const int wm, const int wn, const int g, const int t)
{
#pragma unroll
	for (int kk = 0; kk < C::BK / 4; kk++)
	{
		double a[C::MI], b[C::NJ];
#pragma unroll
		for (int i = 0; i < C::MI; i++)
		{
			a[i] = As[(wm * C::WTM + i * 8 + g) * C::KP + kk * 4 + t];
		}
#pragma unroll
		for (int j = 0; j < C::NJ; j++)
		{
			if (B_NN)
			{
				b[j] = Bs[(kk * 4 + t) * C::NP + wn * C::WTN + j * 8 + g];
			}
			else
			{
				b[j] = Bs[(wn * C::WTN + j * 8 + g) * C::KP + kk * 4 + t];
			}
		}
#pragma unroll
		for (int i = 0; i < C::MI; i++)
		{
#pragma unroll
			for (int j = 0; j < C::NJ; j++)
			{
				dmma884(acc[i][j], a[i], b[j]);
			}
		}
	}
}

template <typename C>
__device__ __forceinline__ void zero_acc(double (&acc)[C::MI][C::NJ][2])
{
#pragma unroll
	for (int i = 0; i < C::MI; i++)
	{
#pragma unroll
		for (int j = 0; j < C::NJ; j++)
		{
			acc[i][j][0] = 0.0;
			acc[i][j][1] = 0.0;
		}
	}
}

/// Triangular structure of an operand, used to skip all-zero 128-blocks of the k range.
enum Tri : int
{
	FULL = 0,
	A_LOWER = 1,	// A[m][k] == 0 for k > m   (k_end = m0 + BM)
	B_LOWER_NT = 2, // NT: B[n][k] == 0 for k > n   (k_end = n0 + BN)
	B_LOWER_NN = 4, // NN: B[k][n] == 0 for n > k   (k_begin = n0)
	A_UPPER = 8,	// A[m][k] == 0 for k < m   (k_begin = m0)
	B_UPPER_NT = 16 // NT: B[n][k] == 0 for k < n  (k_begin = n0)
};

struct GemmArgs
{
	const double* A;
	const double* B;
	double* C;
	size_t lda, ldb, ldc;
	int M, N, K; // multiples of 128 (tiles of 128 or 64)
	double alpha, beta;
	int tri;		// bitmask of Tri
	int lower_only; // skip output tiles strictly above the diagonal (SYRK-style)
	int in_place;	// C aliases A with N == 128: one CTA must own a whole row strip (forces the 128 x 128 tiling)
	// batched form: blockIdx.z = item, operands advance by the strides (in doubles); batch <= 1: a single GEMM
	int batch;
	size_t strideA, strideB, strideC;
};

/// C = beta * C + alpha * A * op(B)
template <typename C, bool B_NN>
__global__ void __launch_bounds__(C::THREADS, 1) gemm_kernel(const GemmArgs p)
{
	extern __shared__ __align__(16) double smem[];
	double* As = smem;
	double* Bs = smem + C::STAGES * C::A_STAGE;
	constexpr int BM = C::BM, BN = C::BN;
	const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
	if (p.lower_only && n0 > m0)
	{
		return;
	}
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N, g = lane >> 2, t = lane & 3;
	int kb = 0, ke = p.K;
	if (p.tri & A_LOWER)
	{
		ke = min(ke, m0 + BM);
	}
	if (p.tri & B_LOWER_NT)
	{
		ke = min(ke, n0 + BN);
	}
	if (p.tri & B_LOWER_NN)
	{
		kb = max(kb, n0);
	}
	if (p.tri & A_UPPER)
	{
		kb = max(kb, m0);
	}
	if (p.tri & B_UPPER_NT)
	{
		kb = max(kb, n0);
	}
	const int nk = ke > kb ? (ke - kb) / C::BK : 0;
	const size_t z = blockIdx.z;
	const double* Ag = p.A + z * p.strideA + size_t(m0) * p.lda + kb;
	const double* Bz = p.B + z * p.strideB;
	const double* Bg = B_NN ? Bz + size_t(kb) * p.ldb + n0 : Bz + size_t(n0) * p.ldb + kb;
	double* Cz = p.C + z * p.strideC;

	double acc[C::MI][C::NJ][2];
	zero_acc<C>(acc);

	auto issue = [&](const int kt)
	{
		if (kt < nk)
		{
			const int slot = kt % C::STAGES;
			load_tile_kmajor<C, C::BM>(As + slot * C::A_STAGE, Ag + size_t(kt) * C::BK, p.lda, tid);
			if (B_NN)
			{
				load_tile_nmajor<C>(Bs + slot * C::B_STAGE, Bg + size_t(kt) * C::BK * p.ldb, p.ldb, tid);
			}
			else
			{
				load_tile_kmajor<C, C::BN>(Bs + slot * C::B_STAGE, Bg + size_t(kt) * C::BK, p.ldb, tid);
			}
		}
		cp_async_commit();
	};
#pragma unroll
	for (int s = 0; s < C::STAGES - 1; s++)
	{
		issue(s);
	}
	for (int kt = 0; kt < nk; kt++)
	{
		cp_async_wait<C::STAGES - 2>();
		__syncthreads();
		issue(kt + C::STAGES - 1);
		const int slot = kt % C::STAGES;
		compute_stage<C, B_NN>(acc, As + slot * C::A_STAGE, Bs + slot * C::B_STAGE, wm, wn, g, t);
	}
	cp_async_wait<0>();

	// epilogue: each lane owns (row, 2t..2t+1) of every 8x8 tile -> 16-byte row-major stores
#pragma unroll
	for (int i = 0; i < C::MI; i++)
	{
		const int row = m0 + wm * C::WTM + i * 8 + g;
#pragma unroll
		for (int j = 0; j < C::NJ; j++)
		{
			const int col = n0 + wn * C::WTN + j * 8 + 2 * t;
			double2* dst = reinterpret_cast<double2*>(Cz + size_t(row) * p.ldc + col);
			double2 v = make_double2(p.alpha * acc[i][j][0], p.alpha * acc[i][j][1]);
			if (p.beta != 0.0)
			{
				const double2 o = *dst;
				v.x += p.beta * o.x;
				v.y += p.beta * o.y;
			}
			*dst = v;
		}
	}
}

} // namespace gemm
} // namespace gple
