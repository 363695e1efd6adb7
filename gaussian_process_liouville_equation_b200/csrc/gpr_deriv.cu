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


// ---------------------------------------------------------------------------------------------------
// complex element: gradients in the composite real form (see gpr.cu header and DESIGN.md)
// ---------------------------------------------------------------------------------------------------
// The reference's dP, dQ (complex_kernel.cpp:379-425) are the blocks of -Kaug^-1 Daug Kaug^-1 with
// Daug = [[D, Dt], [conj(Dt), D]] built from its (quirky, q2) "derivative" matrices D = Derivatives[p],
// Dt = PseudoDerivatives[p].  Through the same transform that maps C^-1 = M to (P, Q), they are the blocks of
// dM = -M Dc M with the real composite Dc = [[(D + Re Dt)/2, Im Dt / 2], [Im Dt / 2, (D - Re Dt)/2]], and
// dv = (dw_r + i dw_i) / 2 with dw = -M Dc w.  Only diagonals of dM's blocks and dw are consumed.

/// One term of a block: coef * mag2 * exp(-r^2/2) * weight, weight = 1 | dx^2/l_x | dp^2/l_p  (dx = (x_i - x_j)/l_x)
struct Term
{
	double mag2, inv_lx, inv_lp, coef;
	int mode;
};
struct BlockTerms
{
	int n;
	Term t[4];
};
/// blocks: 0 = (Re, Re), 1 = (Re, Im) = (Im, Re), 2 = (Im, Im)
struct CompTerms
{
	BlockTerms b[3];
};

__device__ __forceinline__ double eval_terms(const BlockTerms& bt, const double2 a, const double2 c)
{
	double s = 0.0;
	for (int k = 0; k < bt.n; k++)
	{
		const Term& t = bt.t[k];
		const double dx = (a.x - c.x) * t.inv_lx, dp = (a.y - c.y) * t.inv_lp;
		double val = t.coef * t.mag2 * exp(-0.5 * (dx * dx + dp * dp));
		if (t.mode == 1)
		{
			val *= dx * dx * t.inv_lx;
		}
		else if (t.mode == 2)
		{
			val *= dp * dp * t.inv_lp;
		}
		s += val;
	}
	return s;
}

/// Materialise a composite matrix (n = 2 Np, row-major, zero padding) for use as a GEMM operand.
__global__ void __launch_bounds__(256) build_comp_kernel(const CompTerms ct, const double2* __restrict__ X, const int N, const int Np, const int n, double* __restrict__ out)
{
	__shared__ double2 xr[128], xc[128];
	const int I0 = blockIdx.y * 128, J0 = blockIdx.x * 128;
	const int rb = I0 / Np, cb = J0 / Np;
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
	const BlockTerms& bt = ct.b[rb + cb];
	const int c2 = (threadIdx.x & 63) * 2;
	for (int r = threadIdx.x >> 6; r < 128; r += 4)
	{
		double2 o = make_double2(0.0, 0.0);
		if (i0 + r < N)
		{
			if (j0 + c2 < N)
			{
				o.x = eval_terms(bt, xr[r], xc[c2]);
			}
			if (j0 + c2 + 1 < N)
			{
				o.y = eval_terms(bt, xr[r], xc[c2 + 1]);
			}
		}
		*reinterpret_cast<double2*>(out + size_t(I0 + r) * n + J0 + c2) = o;
	}
}

/// out = Comp(ct) * u for a composite matrix generated on the fly (one warp per composite row)
__global__ void __launch_bounds__(256) comp_matvec_kernel(const CompTerms ct, const double2* __restrict__ X, const int N, const int Np, const double* __restrict__ u, double* __restrict__ out)
{
	const int R = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (R >= 2 * Np)
	{
		return;
	}
	const int rb = R / Np, i = R - rb * Np;
	double s = 0.0;
	if (i < N)
	{
		const double2 a = X[i];
		for (int cb = 0; cb < 2; cb++)
		{
			const BlockTerms& bt = ct.b[rb + cb];
			if (bt.n == 0)
			{
				continue;
			}
			for (int j = lane; j < N; j += 32)
			{
				s = fma(eval_terms(bt, a, X[j]), u[cb * Np + j], s);
			}
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		s += __shfl_xor_sync(0xffffffffu, s, o);
	}
	if (lane == 0)
	{
		out[R] = s;
	}
}

/// Diagonals of the (Re,Re), (Im,Im) and (Re,Im) blocks of  scale * G M  (M symmetric): one warp per point i
__global__ void __launch_bounds__(256) rowdot3_kernel(const double* __restrict__ G, const double* __restrict__ M, const int n, const int Np, const double scale, double* __restrict__ drr, double* __restrict__ dii, double* __restrict__ dri)
{
	const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (i >= Np)
	{
		return;
	}
	const double2* gr = reinterpret_cast<const double2*>(G + size_t(i) * n);
	const double2* gi = reinterpret_cast<const double2*>(G + size_t(Np + i) * n);
	const double2* mr = reinterpret_cast<const double2*>(M + size_t(i) * n);
	const double2* mi = reinterpret_cast<const double2*>(M + size_t(Np + i) * n);
	double a = 0.0, b = 0.0, c = 0.0;
	for (int k = lane; k < n / 2; k += 32)
	{
		const double2 g0 = gr[k], g1 = gi[k], m0 = mr[k], m1 = mi[k];
		a = fma(g0.x, m0.x, fma(g0.y, m0.y, a));
		b = fma(g1.x, m1.x, fma(g1.y, m1.y, b));
		c = fma(g0.x, m1.x, fma(g0.y, m1.y, c));
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		a += __shfl_xor_sync(0xffffffffu, a, o);
		b += __shfl_xor_sync(0xffffffffu, b, o);
		c += __shfl_xor_sync(0xffffffffu, c, o);
	}
	if (lane == 0)
	{
		drr[i] = scale * a;
		dii[i] = scale * b;
		dri[i] = scale * c;
	}
}

__global__ void scale_copy_kernel(const double* __restrict__ in, const int count, const double scale, double* __restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count)
	{
		out[i] = scale * in[i];
	}
}

/// complex_kernel.cpp:444-474 for all 8 parameters, plus the dot products of the purity gradient.
/// dd = [8][3][Np] (diagonals of dM's rr, ii, ri blocks), dw = [8][n], u = Kaux w, z = [8][n] (Kaux_d[p] w).
/// out[p] = d error; out[8 + p] = u . dw[p]; out[16 + p] = w . z[p]
__global__ void __launch_bounds__(1024) complex_grad_scalars_kernel(const double* __restrict__ w, const double* __restrict__ ss, const double* __restrict__ cross, const double* __restrict__ dd, const double* __restrict__ dw, const double* __restrict__ u, const double* __restrict__ z, const int N, const int Np, const int want_avg, double* __restrict__ out)
{
	__shared__ double scratch[24 * 32];
	double s[24];
#pragma unroll
	for (int i = 0; i < 24; i++)
	{
		s[i] = 0.0;
	}
	const int n = 2 * Np;
	for (int i = threadIdx.x; i < N; i += 1024)
	{
		const double mrr = ss[i], mii = ss[Np + i], mri = cross[i];
		const double P = 0.25 * (mrr + mii), qr = 0.25 * (mrr - mii), qi = -0.5 * mri;
		const double vr = 0.5 * w[i], vi = 0.5 * w[Np + i];
		const double sqd = P * P - (qr * qr + qi * qi);
		// diff = (P v - conj(Q v)) / sqd
		const double qvr = qr * vr - qi * vi, qvi = qr * vi + qi * vr;
		const double fr = (P * vr - qvr) / sqd, fi = (P * vi + qvi) / sqd;
		const double f2 = fr * fr + fi * fi;
#pragma unroll
		for (int p = 0; p < 8; p++)
		{
			const double* d = dd + size_t(p) * 3 * Np;
			const double drr = d[i], dii = d[Np + i], dri = d[2 * Np + i];
			const double dP = 0.25 * (drr + dii), dqr = 0.25 * (drr - dii), dqi = -0.5 * dri;
			const double dvr = 0.5 * dw[size_t(p) * n + i], dvi = 0.5 * dw[size_t(p) * n + Np + i];
			// inner = dP v + P dv - conj(dQ v + Q dv)
			const double ar = dqr * vr - dqi * vi + qr * dvr - qi * dvi;
			const double ai = dqr * vi + dqi * vr + qr * dvi + qi * dvr;
			const double ir = dP * vr + P * dvr - ar, ii = dP * vi + P * dvi + ai;
			// Re(conj(diff) * inner) + den, den = -2 |diff|^2 (P dP - Re(conj(Q) dQ))
			const double num = fr * ir + fi * ii;
			const double den = -2.0 * f2 * (P * dP - (qr * dqr + qi * dqi));
			s[p] += (num + den) / sqd;
			if (want_avg)
			{
				s[8 + p] = fma(u[i], dw[size_t(p) * n + i], fma(u[Np + i], dw[size_t(p) * n + Np + i], s[8 + p]));
				s[16 + p] = fma(w[i], z[size_t(p) * n + i], fma(w[Np + i], z[size_t(p) * n + Np + i], s[16 + p]));
			}
		}
	}
	block_reduce<24, 1024>(s, scratch);
	if (threadIdx.x == 0)
	{
		for (int i = 0; i < 24; i++)
		{
			out[i] = s[i];
		}
	}
}

/// complex_kernel.cpp:648-667 in composite form.  Per query m (rows 2m, 2m+1 = Re, Im part):
///   g_p = diff_r (Dc*_p[2m] . w + C*[2m] . dw_p) + diff_i (Dc*_p[2m+1] . w + C*[2m+1] . dw_p),  p = 0..7
/// with the test-vs-training blocks regenerated on the fly.  part[block][8]
struct ValGradComplexSpec
{
	GaussBlock kr, ki, kc; // sigma^2 included: the blocks of the covariance itself
	double s2;			   // sigma^2 (the reference's derivative blocks carry no sigma^2: quirk q2)
	double sr, si;		   // sub-magnitudes
	double cx[2][2];	   // [R|I][d]: (1 / l - l / l_C^2)   coefficient of K_C in D_ri
	double cd[2][2];	   // [R|I][d]: l / (2 l_C)           coefficient of dK_C/dl_C in D_ri
	double inv_mag;		   // 1 / sigma
};
__global__ void __launch_bounds__(256) valgrad_complex_kernel(const ValGradComplexSpec sp, const double2* __restrict__ Xq, const long long Q, const double2* __restrict__ yq, const double2* __restrict__ cut, const double rescale, const double2* __restrict__ Xt, const int N, const int Np, const double* __restrict__ w, const double* __restrict__ dw, double* __restrict__ part)
{
	__shared__ double scratch[8 * 8];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n = 2 * Np;
	double tot[8];
#pragma unroll
	for (int p = 0; p < 8; p++)
	{
		tot[p] = 0.0;
	}
	for (long long m = blockIdx.x * 8ll + warp; m < Q; m += 8ll * gridDim.x)
	{
		const double2 a = Xq[m];
		double sr_[8], si_[8]; // per parameter: real-row and imag-row sums
#pragma unroll
		for (int p = 0; p < 8; p++)
		{
			sr_[p] = si_[p] = 0.0;
		}
		for (int j = lane; j < N; j += 32)
		{
			const double2 b = Xt[j];
			const double wr = w[j], wi = w[Np + j];
			// sub-kernel values WITHOUT sigma^2 and their length factors
			const double rx = (a.x - b.x) * sp.kr.inv_lx, rp = (a.y - b.y) * sp.kr.inv_lp;
			const double ix = (a.x - b.x) * sp.ki.inv_lx, ip = (a.y - b.y) * sp.ki.inv_lp;
			const double cx = (a.x - b.x) * sp.kc.inv_lx, cp = (a.y - b.y) * sp.kc.inv_lp;
			const double kR = sp.kr.mag2 / sp.s2 * exp(-0.5 * (rx * rx + rp * rp));
			const double kI = sp.ki.mag2 / sp.s2 * exp(-0.5 * (ix * ix + ip * ip));
			const double kC = sp.kc.mag2 / sp.s2 * exp(-0.5 * (cx * cx + cp * cp));
			const bool same = (a.x == b.x && a.y == b.y);
			// covariance rows (with sigma^2 and the delta term of delta_kernel)
			const double crr = sp.s2 * kR + (same ? sp.kr.diag_add : 0.0), cii = sp.s2 * kI + (same ? sp.ki.diag_add : 0.0), cri = sp.s2 * kC;
			// p = 0: D = 2 C / sigma
			sr_[0] += 2.0 * sp.inv_mag * (crr * wr + cri * wi);
			si_[0] += 2.0 * sp.inv_mag * (cri * wr + cii * wi);
			// p = 1 (sigma_R): D_rr = 2 K_R / s_R, D_ri = K_C / s_R ; p = 4 (sigma_I): D_ii = 2 K_I / s_I, D_ri = K_C / s_I
			sr_[1] += (2.0 * kR * wr + kC * wi) / sp.sr;
			si_[1] += (kC * wr) / sp.sr;
			sr_[4] += (kC * wi) / sp.si;
			si_[4] += (kC * wr + 2.0 * kI * wi) / sp.si;
			// lengths: D_rr = K_R o f_d (R), D_ii = K_I o f_d (I), D_ri = cx K_C + cd K_C o f_d^C
			const double fR[2] = {rx * rx * sp.kr.inv_lx, rp * rp * sp.kr.inv_lp};
			const double fI[2] = {ix * ix * sp.ki.inv_lx, ip * ip * sp.ki.inv_lp};
			const double fC[2] = {cx * cx * sp.kc.inv_lx, cp * cp * sp.kc.inv_lp};
#pragma unroll
			for (int d = 0; d < 2; d++)
			{
				const double driR = sp.cx[0][d] * kC + sp.cd[0][d] * kC * fC[d];
				sr_[2 + d] += kR * fR[d] * wr + driR * wi;
				si_[2 + d] += driR * wr;
				const double driI = sp.cx[1][d] * kC + sp.cd[1][d] * kC * fC[d];
				sr_[5 + d] += driI * wi;
				si_[5 + d] += driI * wr + kI * fI[d] * wi;
			}
			// C* dw_p for every parameter (p = 7: D* = 0 for distinct buffers, only this term)
#pragma unroll
			for (int p = 0; p < 8; p++)
			{
				const double dwr = dw[size_t(p) * n + j], dwi = dw[size_t(p) * n + Np + j];
				sr_[p] += crr * dwr + cri * dwi;
				si_[p] += cri * dwr + cii * dwi;
			}
		}
		const double2 c = cut[m], y = yq[m];
		const double dr = rescale * (c.x - y.x), di = rescale * (c.y - y.y);
#pragma unroll
		for (int p = 0; p < 8; p++)
		{
			tot[p] += dr * sr_[p] + di * si_[p];
		}
	}
	block_reduce<8, 256>(tot, scratch);
	if (threadIdx.x == 0)
	{
		for (int p = 0; p < 8; p++)
		{
			part[blockIdx.x * 8 + p] = tot[p];
		}
	}
}
__global__ void sum8_kernel(const double* __restrict__ part, const int blocks, double* __restrict__ out)
{
	const int i = threadIdx.x;
	if (i < 8)
	{
		double s = 0.0;
		for (int b = 0; b < blocks; b++)
		{
			s += part[b * 8 + i];
		}
		out[i] = s;
	}
}

/// NLML scalars (test/gpr.cpp:515): out[0] = sum_I ln L_II = -sum ln W_II (padding rows have W_II = 1),
/// out[1] = y'^T v, out[2] = v^T v, out[3] = tr(M) over the real rows (M = full inverse, may be null).
__global__ void __launch_bounds__(1024) nlml_scalars_kernel(const double* __restrict__ W, const double* __restrict__ M, const double* __restrict__ label, const double* __restrict__ v, const int N, const int Np, const int n, double* __restrict__ out)
{
	__shared__ double scratch[4 * 32];
	double s[4] = {0.0, 0.0, 0.0, 0.0};
	for (int I = threadIdx.x; I < n; I += 1024)
	{
		const bool real_row = I % Np < N;
		s[0] -= log(W[size_t(I) * n + I]);
		s[1] = fma(label[I], v[I], s[1]);
		s[2] = fma(v[I], v[I], s[2]);
		s[3] += (real_row && M != nullptr) ? M[size_t(I) * n + I] : 0.0;
	}
	block_reduce<4, 1024>(s, scratch);
	if (threadIdx.x == 0)
	{
		out[0] = s[0];
		out[1] = s[1];
		out[2] = s[2];
		out[3] = s[3];
	}
}

/// tr[(M - w w^T) D] = sum_IJ (M_IJ - w_I w_J) D(I, J) for a composite matrix D generated on the fly (never stored):
/// the gradient of the NLML (test/gpr.cpp:487-493, :525).  One CTA per 128 x 128 tile; partials summed by sum_kernel-like pass.
__global__ void __launch_bounds__(256) trace_quad_kernel(const CompTerms ct, const double2* __restrict__ X, const int N, const int Np, const int n, const double* __restrict__ M, const double* __restrict__ w, double* __restrict__ part)
{
	__shared__ double2 xr[128], xc[128];
	__shared__ double wr[128], wc[128];
	__shared__ double scratch[8];
	const int I0 = blockIdx.y * 128, J0 = blockIdx.x * 128;
	const int rb = I0 / Np, cb = J0 / Np;
	const int i0 = I0 - rb * Np, j0 = J0 - cb * Np;
	if (threadIdx.x < 128)
	{
		xr[threadIdx.x] = X[i0 + threadIdx.x];
		wr[threadIdx.x] = w[I0 + threadIdx.x];
	}
	else
	{
		xc[threadIdx.x - 128] = X[j0 + threadIdx.x - 128];
		wc[threadIdx.x - 128] = w[J0 + threadIdx.x - 128];
	}
	__syncthreads();
	const BlockTerms& bt = ct.b[rb + cb];
	double s[1] = {0.0};
	if (bt.n > 0)
	{
		const int c = threadIdx.x & 127;
		if (j0 + c < N)
		{
			for (int r = threadIdx.x >> 7; r < 128 && i0 + r < N; r += 2)
			{
				const double a = M[size_t(I0 + r) * n + J0 + c] - wr[r] * wc[c];
				s[0] = fma(a, eval_terms(bt, xr[r], xc[c]), s[0]);
			}
		}
	}
	block_reduce<1, 256>(s, scratch);
	if (threadIdx.x == 0)
	{
		part[blockIdx.y * gridDim.x + blockIdx.x] = s[0];
	}
}

/// out[k] = sum of part[k][0 .. count)
__global__ void __launch_bounds__(1024) sum_rows_kernel(const double* __restrict__ part, const int count, double* __restrict__ out)
{
	__shared__ double scratch[32];
	double s[1] = {0.0};
	const double* p = part + size_t(blockIdx.x) * count;
	for (int i = threadIdx.x; i < count; i += 1024)
	{
		s[0] += p[i];
	}
	block_reduce<1, 1024>(s, scratch);
	if (threadIdx.x == 0)
	{
		out[blockIdx.x] = s[0];
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

namespace
{
struct CSub
{
	double sr, si, sc, lr[2], li[2], lc[2];
};
CSub csub(const double* th)
{
	CSub c{};
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
Term term(double mag, const double* l, double coef, int mode)
{
	return Term{mag * mag, 1.0 / l[0], 1.0 / l[1], coef, mode};
}
void push(BlockTerms& b, const Term& t)
{
	if (t.coef != 0.0)
	{
		b.t[b.n++] = t;
	}
}
/// Composite derivative of the covariance blocks over sub-kernel parameter p = 1..6 (sigma_R, l_Rx, l_Rp, sigma_I, l_Ix,
/// l_Ip), including the chain rule through the derived correlation kernel (complex_kernel.cpp:37-46, 96-126), times
/// `scale` (1: the reference's arrays, which lack sigma^2 -- quirk q2; sigma^2: the true derivative of C).
CompTerms sub_param_terms(const CSub& c, const int p, const double scale)
{
	const bool isR = p <= 3;
	const int d = isR ? p - 2 : p - 5; // -1: sub-magnitude
	const double smag = isR ? c.sr : c.si;
	const double* sl = isR ? c.lr : c.li;
	CompTerms ct{};
	BlockTerms& diag = ct.b[isR ? 0 : 2];
	if (d < 0)
	{
		push(diag, term(smag, sl, scale * 2.0 / smag, 0));
		push(ct.b[1], term(c.sc, c.lc, scale / smag, 0));
	}
	else
	{
		push(diag, term(smag, sl, scale, 1 + d));
		push(ct.b[1], term(c.sc, c.lc, scale * (1.0 / sl[d] - sl[d] / (c.lc[d] * c.lc[d])), 0));
		push(ct.b[1], term(c.sc, c.lc, scale * 0.5 * sl[d] / c.lc[d], 1 + d));
	}
	return ct;
}
/// purity auxiliary kernels (kernel.h:285-294, complex_kernel.cpp:206-219): magnitude and lengths
struct Aux
{
	double mag, l[2];
};
Aux aux_of(double mag, const double* l)
{
	return Aux{mag * mag * std::sqrt(l[0] * l[1]), {std::sqrt(2.0) * l[0], std::sqrt(2.0) * l[1]}};
}
Aux mixed_of(double ma, const double* la, double mb, const double* lb)
{
	Aux a{};
	double prod = 1.0;
	for (int d = 0; d < 2; d++)
	{
		prod *= 0.5 * (1.0 / (la[d] * la[d]) + 1.0 / (lb[d] * lb[d]));
		a.l[d] = std::sqrt(la[d] * la[d] + lb[d] * lb[d]);
	}
	a.mag = ma * mb / std::sqrt(std::sqrt(prod));
	return a;
}
} // namespace

void complex_derivatives(gple_ctx* ctx, gple_model* m, unsigned flags, const double* h_scal, gple_complex_scalars* r)
{
	(void)h_scal;
	const int N = int(m->N), Np = m->Np, n = m->n;
	const double* th = m->theta;
	const double sigma = th[0], noise = th[7];
	const CSub c = csub(th);
	const double2* X = reinterpret_cast<const double2*>(m->X);
	ensure_full_inverse(ctx, m); // M = C^-1
	if (m->dv == nullptr)
	{
		m->dv = static_cast<double*>(ctx->pool.alloc(size_t(8) * n * sizeof(double)));
	}
	double* dw = m->dv;
	double* dd = ctx->ws.get<double>("cderiv.dd", size_t(8) * 3 * Np);
	double* t = ctx->ws.get<double>("cderiv.t", size_t(n));
	double* u = ctx->ws.get<double>("cderiv.u", size_t(n));
	double* z = ctx->ws.get<double>("cderiv.z", size_t(8) * n);
	double* Dm = ctx->ws.get<double>("deriv.dK", size_t(n) * n);
	double* G = ctx->ws.get<double>("deriv.G", size_t(n) * n);
	GPLE_CUDA(cudaMemsetAsync(z, 0, size_t(8) * n * sizeof(double), ctx->stream));
	GPLE_CUDA(cudaMemsetAsync(u, 0, size_t(n) * sizeof(double), ctx->stream));
	const int rb = (n + 7) / 8, pb = (Np + 7) / 8;
	const double* ss = m->kinv_diag;
	const double* cross = m->kinv_diag + n;
	// p = 0 (global magnitude): Dc = 2 C / sigma  =>  dM = -2 M / sigma, dw = -2 w / sigma
	GPLE_LAUNCH(ctx, scale_copy_kernel, (n + 255) / 256, 256, 0, m->v, n, -2.0 / sigma, dw);
	GPLE_LAUNCH(ctx, scale_copy_kernel, (Np + 255) / 256, 256, 0, ss, Np, -2.0 / sigma, dd);
	GPLE_LAUNCH(ctx, scale_copy_kernel, (Np + 255) / 256, 256, 0, ss + Np, Np, -2.0 / sigma, dd + Np);
	GPLE_LAUNCH(ctx, scale_copy_kernel, (Np + 255) / 256, 256, 0, cross, Np, -2.0 / sigma, dd + 2 * Np);
	// p = 7 (noise): D = 2 sigma_n I (no sigma^2: quirk q2), Dt = 0  =>  Dc = sigma_n I  =>  dM = -sigma_n M M
	GPLE_LAUNCH(ctx, rowdot3_kernel, pb, 256, 0, m->Kinv, m->Kinv, n, Np, -noise, dd + size_t(7) * 3 * Np, dd + size_t(7) * 3 * Np + Np, dd + size_t(7) * 3 * Np + 2 * Np);
	GPLE_LAUNCH(ctx, matvec_kernel, rb, 256, 0, m->Kinv, n, m->v, -noise, dw + size_t(7) * n);
	// p = 1..6: sub-kernel parameters (complex_kernel.cpp:37-46, 96-126), blocks WITHOUT sigma^2 (quirk q2)
	for (int p = 1; p <= 6; p++)
	{
		const CompTerms ct = sub_param_terms(c, p, 1.0);
		GPLE_LAUNCH(ctx, build_comp_kernel, dim3(n / 128, n / 128), 256, 0, ct, X, N, Np, n, Dm);
		gemm::GemmArgs a{};
		a.A = m->Kinv;
		a.B = Dm; // symmetric
		a.C = G;
		a.lda = a.ldb = a.ldc = size_t(n);
		a.M = a.N = a.K = n;
		a.alpha = 1.0;
		a.beta = 0.0;
		gemm_nt(ctx, a);
		double* dp = dd + size_t(p) * 3 * Np;
		GPLE_LAUNCH(ctx, rowdot3_kernel, pb, 256, 0, G, m->Kinv, n, Np, -1.0, dp, dp + Np, dp + 2 * Np);
		GPLE_LAUNCH(ctx, matvec_kernel, rb, 256, 0, Dm, n, m->v, 1.0, t);
		GPLE_LAUNCH(ctx, matvec_kernel, rb, 256, 0, m->Kinv, n, t, -1.0, dw + size_t(p) * n);
	}
	const bool avg = (flags & GPLE_CALC_AVERAGE) != 0;
	if (avg)
	{
		// complex_kernel.cpp:475-590 in composite form: d purity_p = GF / s^2 [ w^T Kaux dw_p + 1/2 w^T Kaux_d[p] w ]
		const Aux Rp = aux_of(c.sr, c.lr), Ip = aux_of(c.si, c.li), Cp = aux_of(c.sc, c.lc);
		const Aux RC = mixed_of(c.sr, c.lr, c.sc, c.lc), IC = mixed_of(c.si, c.li, c.sc, c.lc);
		CompTerms ka{};
		push(ka.b[0], term(Rp.mag, Rp.l, 1.0, 0));
		push(ka.b[0], term(Cp.mag, Cp.l, 1.0, 0));
		push(ka.b[2], term(Ip.mag, Ip.l, 1.0, 0));
		push(ka.b[2], term(Cp.mag, Cp.l, 1.0, 0));
		push(ka.b[1], term(RC.mag, RC.l, 1.0, 0));
		push(ka.b[1], term(IC.mag, IC.l, 1.0, 0));
		GPLE_LAUNCH(ctx, comp_matvec_kernel, rb, 256, 0, ka, X, N, Np, m->v, u);
		const double sq2 = std::sqrt(2.0);
		for (int p = 1; p <= 6; p++)
		{
			const bool isR = p <= 3;
			const int d = isR ? p - 2 : p - 5;
			const double smag = isR ? c.sr : c.si;
			const double* sl = isR ? c.lr : c.li;
			const Aux& own = isR ? Rp : Ip;
			CompTerms kd{};
			BlockTerms& diag = kd.b[isR ? 0 : 2];
			BlockTerms& other = kd.b[isR ? 2 : 0];
			if (d < 0)
			{
				// complex_kernel.cpp:525-529 / 548-552
				push(diag, term(own.mag, own.l, 4.0 / smag, 0));
				push(diag, term(Cp.mag, Cp.l, 2.0 / smag, 0));
				push(other, term(Cp.mag, Cp.l, 2.0 / smag, 0));
				push(kd.b[1], term(RC.mag, RC.l, (isR ? 3.0 : 1.0) / smag, 0));
				push(kd.b[1], term(IC.mag, IC.l, (isR ? 1.0 : 3.0) / smag, 0));
			}
			else
			{
				// complex_kernel.cpp:534-541 / 558-564.  Quirk q10: the sub-kernel derivative entry used is [2 + d] of
				// [mag, l_x, l_p, noise]: the l_p derivative (mode 2) for d == 0, the zero noise derivative for d == 1.
				const int mode = (d == 0) ? 2 : -1;
				const double oc2 = sl[d] / (c.lc[d] * c.lc[d]), oc = sl[d] / c.lc[d];
				const double orc = sl[d] / RC.l[d], oic = sl[d] / IC.l[d];
				push(diag, term(own.mag, own.l, 1.0 / sl[d], 0));
				const double cC = 2.0 / sl[d] - 3.0 * oc2 / 2.0;
				push(diag, term(Cp.mag, Cp.l, cC, 0));
				push(other, term(Cp.mag, Cp.l, cC, 0));
				const double wRC = isR ? 1.5 * orc : 0.5 * orc, wIC = isR ? 0.5 * oic : 1.5 * oic;
				push(kd.b[1], term(RC.mag, RC.l, ((isR ? 2.0 : 1.0) / sl[d] - oc2 / 2.0) - wRC / RC.l[d], 0));
				push(kd.b[1], term(IC.mag, IC.l, ((isR ? 1.0 : 2.0) / sl[d] - oc2 / 2.0) - wIC / IC.l[d], 0));
				if (mode > 0)
				{
					push(diag, term(own.mag, own.l, sq2, mode));
					if (diag.n < 4)
					{
						push(diag, term(Cp.mag, Cp.l, oc / sq2, mode));
					}
					push(other, term(Cp.mag, Cp.l, oc / sq2, mode));
					push(kd.b[1], term(RC.mag, RC.l, wRC, mode));
					push(kd.b[1], term(IC.mag, IC.l, wIC, mode));
				}
			}
			GPLE_LAUNCH(ctx, comp_matvec_kernel, rb, 256, 0, kd, X, N, Np, m->v, z + size_t(p) * n);
		}
	}
	double* d_out = ctx->ws.get<double>("cderiv.out", 32);
	GPLE_LAUNCH(ctx, complex_grad_scalars_kernel, 1, 1024, 0, m->v, ss, cross, dd, dw, u, z, N, Np, int(avg), d_out);
	double h[24];
	read_back(ctx, d_out, 24, h);
	if (flags & GPLE_CALC_ERROR)
	{
		for (int p = 0; p < 8; p++)
		{
			r->d_error[p] = 2.0 * h[p];
		}
	}
	if (avg)
	{
		const double gf = (2.0 * M_PI) * 2.0 * M_PI, s = r->rescale;
		for (int p = 0; p < 8; p++)
		{
			// v^H K1 dv etc. in terms of w: 2 Re(..) + 2 Re(..) = w^T Kaux dw ; second part = 1/2 w^T Kaux_d w
			r->d_purity[p] = gf * (h[8 + p] + 0.5 * h[16 + p]) / (s * s);
		}
	}
}

void validation_gradient(gple_ctx* ctx, const gple_model* m, const double* d_Xq, size_t Q, const double* d_yq, const double* d_cut, double* h_grad)
{
	if (m->dv == nullptr)
	{
		for (int p = 0; p < m->nparam(); p++)
		{
			h_grad[p] = std::nan("");
		}
		return;
	}
	if (m->is_complex)
	{
		const double* th = m->theta;
		const CSub c = csub(th);
		const double s2 = th[0] * th[0], hn = 0.5 * s2 * th[7] * th[7];
		ValGradComplexSpec sp{};
		sp.kr = GaussBlock{s2 * c.sr * c.sr, 1.0 / c.lr[0], 1.0 / c.lr[1], hn};
		sp.ki = GaussBlock{s2 * c.si * c.si, 1.0 / c.li[0], 1.0 / c.li[1], hn};
		sp.kc = GaussBlock{s2 * c.sc * c.sc, 1.0 / c.lc[0], 1.0 / c.lc[1], 0.0};
		sp.s2 = s2;
		sp.sr = c.sr;
		sp.si = c.si;
		sp.inv_mag = 1.0 / th[0];
		for (int d = 0; d < 2; d++)
		{
			sp.cx[0][d] = 1.0 / c.lr[d] - c.lr[d] / (c.lc[d] * c.lc[d]);
			sp.cd[0][d] = 0.5 * c.lr[d] / c.lc[d];
			sp.cx[1][d] = 1.0 / c.li[d] - c.li[d] / (c.lc[d] * c.lc[d]);
			sp.cd[1][d] = 0.5 * c.li[d] / c.lc[d];
		}
		const int blocks = int(std::min<size_t>(592, (Q + 7) / 8));
		double* part = ctx->ws.get<double>("valgrad.part", size_t(blocks) * 8 + 8);
		GPLE_LAUNCH(ctx, valgrad_complex_kernel, blocks, 256, 0, sp, reinterpret_cast<const double2*>(d_Xq), (long long)Q, reinterpret_cast<const double2*>(d_yq), reinterpret_cast<const double2*>(d_cut), m->rescale, reinterpret_cast<const double2*>(m->X), int(m->N), m->Np, m->v, m->dv, part);
		double* out = part + size_t(blocks) * 8;
		GPLE_LAUNCH(ctx, sum8_kernel, 1, 32, 0, part, blocks, out);
		double h[8];
		read_back(ctx, out, 8, h);
		for (int p = 0; p < 8; p++)
		{
			h_grad[p] = 2.0 * h[p];
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

/// NLML / LLT objective of test/gpr.cpp:470-532 on a trained element model:
///   value = y'^T v / 2 + sum ln L_II,   grad_p = tr[(M - v v^T) dC/dtheta_p] / 2   (M = C^-1, true derivatives).
/// The magnitude and noise components are closed forms (tr(M C) = rows, dC/dsigma_n proportional to I); the length and
/// sub-kernel components generate dC on the fly.  Complex element: composite [Re f; Im f] process.
void nlml_device(gple_ctx* ctx, gple_model* m, double* value, double* grad)
{
	const int N = int(m->N), Np = m->Np, n = m->n;
	const double2* X = reinterpret_cast<const double2*>(m->X);
	double* out = ctx->ws.get<double>("nlml.out", 16);
	if (grad != nullptr)
	{
		ensure_full_inverse(ctx, m);
	}
	GPLE_LAUNCH(ctx, nlml_scalars_kernel, 1, 1024, 0, m->W, grad != nullptr ? m->Kinv : nullptr, m->label, m->v, N, Np, n, out);
	const int np = m->nparam();
	const int tiles = (n / 128) * (n / 128);
	int generated[6], ng = 0;
	if (grad != nullptr)
	{
		double* part = ctx->ws.get<double>("nlml.part", size_t(6) * tiles);
		if (m->is_complex)
		{
			const CSub c = csub(m->theta);
			for (int p = 1; p <= 6; p++)
			{
				const CompTerms ct = sub_param_terms(c, p, m->theta[0] * m->theta[0]);
				GPLE_LAUNCH(ctx, trace_quad_kernel, dim3(n / 128, n / 128), 256, 0, ct, X, N, Np, n, m->Kinv, m->v, part + size_t(ng) * tiles);
				generated[ng++] = p;
			}
		}
		else
		{
			const double l[2] = {m->theta[1], m->theta[2]};
			for (int d = 0; d < 2; d++)
			{
				CompTerms ct{};
				push(ct.b[0], term(m->theta[0], l, 1.0, 1 + d));
				GPLE_LAUNCH(ctx, trace_quad_kernel, dim3(n / 128, n / 128), 256, 0, ct, X, N, Np, n, m->Kinv, m->v, part + size_t(ng) * tiles);
				generated[ng++] = 1 + d;
			}
		}
		GPLE_LAUNCH(ctx, sum_rows_kernel, ng, 1024, 0, part, tiles, out + 4);
	}
	double h[16];
	read_back(ctx, out, 4 + ng, h);
	*value = 0.5 * h[1] + h[0];
	if (grad != nullptr)
	{
		const double mag = m->theta[0], noise = m->theta[np - 1];
		const double rows = double(m->is_complex ? 2 * N : N);
		grad[0] = (rows - h[1]) / mag;
		grad[np - 1] = (m->is_complex ? 0.5 : 1.0) * mag * mag * noise * (h[3] - h[2]);
		for (int k = 0; k < ng; k++)
		{
			grad[generated[k]] = 0.5 * h[4 + k];
		}
	}
}
} // namespace gple
