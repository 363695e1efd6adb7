// Hyper-parameter gradients of the training-set quantities and of the validation error (real element).
//
// Reference: gple/kernel.cpp:337-477 (InverseDerivatives ... PurityDerivatives) and :524-541.
// The reference materialises d K^-1 / d theta = -K^-1 (dK) K^-1 for every parameter (5 N x N GEMMs and
// 8 N x N temporaries), but only consumes its DIAGONAL (kernel.cpp:394) and its product with the label
// (kernel.cpp:373).  Here per length parameter there is one DMMA GEMM  G = K^-1 dK  followed by a row-dot
// with K^-1 for the diagonal, and  d v = -K^-1 (dK v)  by two matrix-vector products; the magnitude and
// noise parameters need no GEMM at all (dK is a multiple of K or of the identity).
#include "chol.cuh"
#include "gpr.cuh"
#include "gpr_kernels.cuh"

namespace gple
{
namespace
{
/// dK/dl_d (kernel.cpp:99-160 via :184-202): sigma_f^2 exp(-r^2/2) * ((x_i - x_j)_d / l_d)^2 / l_d, zero diagonal,
/// zero in the padding.  Full n x n, row-major.
__global__ void __launch_bounds__(256) build_dk_kernel(const GaussBlock g, const int d, const double2* __restrict__ X, const int N, const int n, double* __restrict__ dK)
{
	__shared__ double2 xr[128], xc[128];
	const int i0 = blockIdx.y * 128, j0 = blockIdx.x * 128;
	if (threadIdx.x < 128)
	{
		xr[threadIdx.x] = X[i0 + threadIdx.x];
	}
	else
	{
		xc[threadIdx.x - 128] = X[j0 + threadIdx.x - 128];
	}
	__syncthreads();
	const int c2 = (threadIdx.x & 63) * 2;
	for (int r = threadIdx.x >> 6; r < 128; r += 4)
	{
		double2 out;
		double* o = &out.x;
#pragma unroll
		for (int u = 0; u < 2; u++)
		{
			const int i = i0 + r, j = j0 + c2 + u;
			double val = 0.0;
			if (i < N && j < N && i != j)
			{
				const double2 a = xr[r], b = xc[c2 + u];
				const double dx = (a.x - b.x) * g.inv_lx, dp = (a.y - b.y) * g.inv_lp;
				const double e = g.mag2 * exp(-0.5 * (dx * dx + dp * dp));
				val = (d == 0) ? e * (dx * dx * g.inv_lx) : e * (dp * dp * g.inv_lp);
			}
			o[u] = val;
		}
		*reinterpret_cast<double2*>(dK + size_t(i0 + r) * n + j0 + c2) = out;
	}
}

/// out[i] = sum_b A[i][b] * B[i][b]   (one warp per row)
__global__ void __launch_bounds__(256) rowdot_kernel(const double* __restrict__ A, const double* __restrict__ B, const int n, double* __restrict__ out)
{
	const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (row >= n)
	{
		return;
	}
	const double2* a = reinterpret_cast<const double2*>(A + size_t(row) * n);
	const double2* b = reinterpret_cast<const double2*>(B + size_t(row) * n);
	double s = 0.0;
	for (int k = lane; k < n / 2; k += 32)
	{
		const double2 u = a[k], v = b[k];
		s = fma(u.x, v.x, s);
		s = fma(u.y, v.y, s);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		s += __shfl_xor_sync(0xffffffffu, s, o);
	}
	if (lane == 0)
	{
		out[row] = s;
	}
}

/// y = alpha * M x   (full n x n row-major, one warp per row)
__global__ void __launch_bounds__(256) matvec_kernel(const double* __restrict__ M, const int n, const double* __restrict__ x, const double alpha, double* __restrict__ y)
{
	const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (row >= n)
	{
		return;
	}
	const double2* a = reinterpret_cast<const double2*>(M + size_t(row) * n);
	const double2* b = reinterpret_cast<const double2*>(x);
	double s = 0.0;
	for (int k = lane; k < n / 2; k += 32)
	{
		const double2 u = a[k], v = b[k];
		s = fma(u.x, v.x, s);
		s = fma(u.y, v.y, s);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		s += __shfl_xor_sync(0xffffffffu, s, o);
	}
	if (lane == 0)
	{
		y[row] = alpha * s;
	}
}

/// On-the-fly products with a Gaussian kernel matrix that is never stored (one warp per row):
///   o0 = K u,  o1 = (K o dx^2 / l_x) u,  o2 = (K o dp^2 / l_p) u,   dx = (x_i - x_j) / l_x  (zero diagonal for o1, o2)
__global__ void __launch_bounds__(256) gauss_matvec_kernel(const GaussBlock g, const double2* __restrict__ X, const int N, const double* __restrict__ u, double* __restrict__ o0, double* __restrict__ o1, double* __restrict__ o2)
{
	const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (row >= N)
	{
		return;
	}
	const double2 a = X[row];
	double s0 = 0.0, s1 = 0.0, s2 = 0.0;
	for (int j = lane; j < N; j += 32)
	{
		const double2 b = X[j];
		const double dx = (a.x - b.x) * g.inv_lx, dp = (a.y - b.y) * g.inv_lp;
		const double k = g.mag2 * exp(-0.5 * (dx * dx + dp * dp)) * u[j];
		s0 += k;
		s1 = fma(k, dx * dx * g.inv_lx, s1);
		s2 = fma(k, dp * dp * g.inv_lp, s2);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		s0 += __shfl_xor_sync(0xffffffffu, s0, o);
		s1 += __shfl_xor_sync(0xffffffffu, s1, o);
		s2 += __shfl_xor_sync(0xffffffffu, s2, o);
	}
	if (lane == 0)
	{
		o0[row] = s0;
		o1[row] = s1;
		o2[row] = s2;
	}
}

/// Final scalar reductions of the real-element gradients.  Inputs per parameter p: dv[p] (n) and the
/// diagonal of dK^-1/dtheta_p, dd[p] (n).  out[0..3] = d error (kernel.cpp:394), out[4..7] = sum dv[p],
/// out[8..11] = dv[p]^T (K1 v), out[12] = v^T (K1 o Dx') v, out[13] = v^T (K1 o Dp') v
__global__ void __launch_bounds__(1024) real_grad_scalars_kernel(const double* __restrict__ v, const double* __restrict__ kd, const double* __restrict__ dv, const double* __restrict__ dd, const double* __restrict__ k1v, const double* __restrict__ k1xv, const double* __restrict__ k1pv, const int N, const int n, const int want_avg, double* __restrict__ out)
{
	__shared__ double scratch[14 * 32];
	double s[14];
#pragma unroll
	for (int i = 0; i < 14; i++)
	{
		s[i] = 0.0;
	}
	for (int i = threadIdx.x; i < N; i += 1024)
	{
		const double d = kd[i], vi = v[i], r = vi / d;
#pragma unroll
		for (int p = 0; p < 4; p++)
		{
			const double dvi = dv[size_t(p) * n + i];
			s[p] += r / d * (dvi - r * dd[size_t(p) * n + i]);
			s[4 + p] += dvi;
			if (want_avg)
			{
				s[8 + p] = fma(dvi, k1v[i], s[8 + p]);
			}
		}
		if (want_avg)
		{
			s[12] = fma(vi, k1xv[i], s[12]);
			s[13] = fma(vi, k1pv[i], s[13]);
		}
	}
	block_reduce<14, 1024>(s, scratch);
	if (threadIdx.x == 0)
	{
		for (int i = 0; i < 14; i++)
		{
			out[i] = s[i];
		}
	}
}

/// dv[0] = -2 v / sigma_f, dd[0] = -2 d / sigma_f (kernel.cpp:349); dv[3] = c * (K^-1 v), dd[3] = c * rownorm2 (kernel.cpp:358)
__global__ void simple_param_kernel(const double* __restrict__ v, const double* __restrict__ kd, const double* __restrict__ kinv_v, const double* __restrict__ rn2, const int n, const double mag, const double c_noise, double* __restrict__ dv, double* __restrict__ dd)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
	{
		return;
	}
	dv[i] = -2.0 / mag * v[i];
	dd[i] = -2.0 / mag * kd[i];
	dv[size_t(3) * n + i] = c_noise * kinv_v[i];
	dd[size_t(3) * n + i] = c_noise * rn2[i];
}

constexpr int VG_SUMS = 6;
/// kernel.cpp:524-541 for a chunk of queries: per query the six sums
///   (dK*/dl_x) v, (dK*/dl_p) v, K* dv[0..3]   -- K* is regenerated on the fly, never stored --
/// weighted by diff = rescale * (cutoff - y) and reduced per block into part[block][6 + 1] (last = diff * K* v).
__global__ void __launch_bounds__(256) valgrad_kernel(const GaussBlock g, const double2* __restrict__ Xq, const long long Q, const double* __restrict__ yq, const double* __restrict__ cut, const double rescale, const double2* __restrict__ Xt, const int N, const int n, const double* __restrict__ v, const double* __restrict__ dv, double* __restrict__ part)
{
	__shared__ double scratch[(VG_SUMS + 1) * 8];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	double tot[VG_SUMS + 1];
#pragma unroll
	for (int i = 0; i <= VG_SUMS; i++)
	{
		tot[i] = 0.0;
	}
	for (long long m = blockIdx.x * 8ll + warp; m < Q; m += 8ll * gridDim.x)
	{
		const double2 a = Xq[m];
		double s[VG_SUMS + 1];
#pragma unroll
		for (int i = 0; i <= VG_SUMS; i++)
		{
			s[i] = 0.0;
		}
		for (int j = lane; j < N; j += 32)
		{
			const double2 b = Xt[j];
			const double dx = (a.x - b.x) * g.inv_lx, dp = (a.y - b.y) * g.inv_lp;
			const double k = g.mag2 * exp(-0.5 * (dx * dx + dp * dp)) + ((a.x == b.x && a.y == b.y) ? g.diag_add : 0.0);
			const double kv = k * v[j];
			s[0] = fma(kv, dx * dx * g.inv_lx, s[0]);
			s[1] = fma(kv, dp * dp * g.inv_lp, s[1]);
#pragma unroll
			for (int p = 0; p < 4; p++)
			{
				s[2 + p] = fma(k, dv[size_t(p) * n + j], s[2 + p]);
			}
			s[6] += kv;
		}
		const double diff = rescale * (cut[m] - yq[m]); // CutoffPrediction * Rescale - Label (quirk q1)
#pragma unroll
		for (int i = 0; i <= VG_SUMS; i++)
		{
			tot[i] = fma(diff, s[i], tot[i]); // lanes hold partial sums; reduced below
		}
	}
	block_reduce<VG_SUMS + 1, 256>(tot, scratch);
	if (threadIdx.x == 0)
	{
		for (int i = 0; i <= VG_SUMS; i++)
		{
			part[blockIdx.x * (VG_SUMS + 1) + i] = tot[i];
		}
	}
}
__global__ void valgrad_final_kernel(const double* __restrict__ part, const int blocks, double* __restrict__ out)
{
	const int i = threadIdx.x;
	if (i <= VG_SUMS)
	{
		double s = 0.0;
		for (int b = 0; b < blocks; b++)
		{
			s += part[b * (VG_SUMS + 1) + i];
		}
		out[i] = s;
	}
}

void read_back(gple_ctx* ctx, const double* d, int count, double* h)
{
	GPLE_CUDA(cudaMemcpyAsync(ctx->h_pinned, d, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	GPLE_CUDA(cudaStreamSynchronize(ctx->stream));
	std::memcpy(h, ctx->h_pinned, count * sizeof(double));
}
} // namespace

void real_derivatives(gple_ctx* ctx, gple_model* m, unsigned flags, const double* h_scal, gple_real_scalars* r)
{
	const int N = int(m->N), n = m->n;
	const double mag = m->theta[0], lx = m->theta[1], lp = m->theta[2], noise = m->theta[3];
	const double2* X = reinterpret_cast<const double2*>(m->X);
	ensure_full_inverse(ctx, m);
	if (m->dv == nullptr)
	{
		m->dv = static_cast<double*>(ctx->pool.alloc(size_t(4) * n * sizeof(double)));
	}
	double* dd = ctx->ws.get<double>("deriv.dd", size_t(4) * n);
	double* tmp = ctx->ws.get<double>("deriv.tmp", size_t(6) * n);
	double* kinv_v = tmp;
	double* rn2 = tmp + n;
	double* t = tmp + 2 * n;
	double* k1v = tmp + 3 * n;
	double* k1xv = tmp + 4 * n;
	double* k1pv = tmp + 5 * n;
	GPLE_CUDA(cudaMemsetAsync(tmp, 0, size_t(6) * n * sizeof(double), ctx->stream));
	const int rb = (n + 7) / 8;
	GPLE_LAUNCH(ctx, matvec_kernel, rb, 256, 0, m->Kinv, n, m->v, 1.0, kinv_v);
	GPLE_LAUNCH(ctx, rowdot_kernel, rb, 256, 0, m->Kinv, m->Kinv, n, rn2);
	GPLE_LAUNCH(ctx, simple_param_kernel, (n + 255) / 256, 256, 0, m->v, m->kinv_diag, kinv_v, rn2, n, mag, -2.0 * mag * mag * noise, m->dv, dd);
	double* dK = ctx->ws.get<double>("deriv.dK", size_t(n) * n);
	double* G = ctx->ws.get<double>("deriv.G", size_t(n) * n);
	const GaussBlock g{mag * mag, 1.0 / lx, 1.0 / lp, 0.0};
	for (int d = 0; d < 2; d++)
	{
		GPLE_LAUNCH(ctx, build_dk_kernel, dim3(n / 128, n / 128), 256, 0, g, d, X, N, n, dK);
		gemm::GemmArgs a{};
		a.A = m->Kinv;
		a.B = dK; // symmetric: K^-1 dK = K^-1 dK^T
		a.C = G;
		a.lda = a.ldb = a.ldc = size_t(n);
		a.M = a.N = a.K = n;
		a.alpha = -1.0;
		a.beta = 0.0;
		gemm_nt(ctx, a);
		GPLE_LAUNCH(ctx, rowdot_kernel, rb, 256, 0, G, m->Kinv, n, dd + size_t(1 + d) * n); // diag(-K^-1 dK K^-1)
		GPLE_LAUNCH(ctx, matvec_kernel, rb, 256, 0, dK, n, m->v, 1.0, t);
		GPLE_LAUNCH(ctx, matvec_kernel, rb, 256, 0, m->Kinv, n, t, -1.0, m->dv + size_t(1 + d) * n);
	}
	const bool avg = (flags & GPLE_CALC_AVERAGE) != 0;
	const double lxa = std::sqrt(2.0) * lx, lpa = std::sqrt(2.0) * lp;
	if (avg)
	{
		const double ma = mag * mag * std::sqrt(lx * lp);
		const GaussBlock ga{ma * ma, 1.0 / lxa, 1.0 / lpa, 0.0};
		GPLE_LAUNCH(ctx, gauss_matvec_kernel, (N + 7) / 8, 256, 0, ga, X, N, m->v, k1v, k1xv, k1pv);
	}
	double* d_out = ctx->ws.get<double>("deriv.out", 16);
	GPLE_LAUNCH(ctx, real_grad_scalars_kernel, 1, 1024, 0, m->v, m->kinv_diag, m->dv, dd, k1v, k1xv, k1pv, N, n, int(avg), d_out);
	double h[16];
	read_back(ctx, d_out, 14, h);
	const double s = r->rescale;
	if (flags & GPLE_CALC_ERROR)
	{
		for (int p = 0; p < 4; p++)
		{
			r->d_error[p] = 2.0 * h[p];
		}
	}
	if (avg)
	{
		// kernel.cpp:401-435
		const double f = 2.0 * M_PI * mag * mag * lx * lp;
		const double sum_v = h_scal[2];
		r->d_population[0] = 0.0; // quirk q4
		r->d_population[1] = f * (sum_v / lx + h[4 + 1]) / s;
		r->d_population[2] = f * (sum_v / lp + h[4 + 2]) / s;
		r->d_population[3] = f * h[4 + 3] / s;
		// kernel.cpp:436-477: v^T (K1 / l_d + sqrt2 dK1/dl'_d) v + 2 dv_d^T K1 v
		const double gf = (2.0 * M_PI) * M_PI;
		const double quad = r->purity * s * s / gf; // v^T K1 v
		r->d_purity[0] = 0.0;
		r->d_purity[1] = gf * (quad / lx + std::sqrt(2.0) * h[12] + 2.0 * h[8 + 1]) / (s * s);
		r->d_purity[2] = gf * (quad / lp + std::sqrt(2.0) * h[13] + 2.0 * h[8 + 2]) / (s * s);
		r->d_purity[3] = 2.0 * gf * h[8 + 3] / (s * s);
	}
}

void complex_derivatives(gple_ctx*, gple_model*, unsigned, const double*, gple_complex_scalars*)
{
	// TODO(round 2): composite-form gradients of complex_kernel.cpp:379-590; the scalars stay NaN.
}

void validation_gradient(gple_ctx* ctx, const gple_model* m, const double* d_Xq, size_t Q, const double* d_yq, const double* d_cut, double* h_grad)
{
	if (m->is_complex || m->dv == nullptr)
	{
		for (int p = 0; p < m->nparam(); p++)
		{
			h_grad[p] = std::nan("");
		}
		return;
	}
	const double mag = m->theta[0];
	const GaussBlock g{mag * mag, 1.0 / m->theta[1], 1.0 / m->theta[2], mag * mag * m->theta[3] * m->theta[3]};
	const int blocks = int(std::min<size_t>(592, (Q + 7) / 8));
	double* part = ctx->ws.get<double>("valgrad.part", size_t(blocks) * (VG_SUMS + 1) + 8);
	GPLE_LAUNCH(ctx, valgrad_kernel, blocks, 256, 0, g, reinterpret_cast<const double2*>(d_Xq), (long long)Q, d_yq, d_cut, m->rescale, reinterpret_cast<const double2*>(m->X), int(m->N), m->n, m->v, m->dv, part);
	double* out = part + size_t(blocks) * (VG_SUMS + 1);
	GPLE_LAUNCH(ctx, valgrad_final_kernel, 1, 32, 0, part, blocks, out);
	double h[VG_SUMS + 1];
	read_back(ctx, out, VG_SUMS + 1, h);
	// dK*/d sigma_f = 2 K* / sigma_f ; dK*/d sigma_n = 0 for distinct buffers (kernel.cpp:181,211)
	h_grad[0] = 2.0 * (2.0 / mag * h[6] + h[2]);
	h_grad[1] = 2.0 * (h[0] + h[3]);
	h_grad[2] = 2.0 * (h[1] + h[4]);
	h_grad[3] = 2.0 * (0.0 + h[5]);
}

} // namespace gple
