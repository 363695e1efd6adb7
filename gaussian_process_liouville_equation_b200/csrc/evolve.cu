// Per-point Trotter evolve step, Tully-model PES and Monte-Carlo-integral observables on the GPU.
//
// Reference path replaced: gple/pes.cpp:42-189, gple/evolve.cpp:53-478, gple/predict.cpp:43-244.
// The reference evolves each point inside a par_unseq loop and calls the GPR predictor through a
// std::function once per backward-propagated query (8 single-point PredictiveKernel constructions per
// evolved point, each streaming the whole N x N inverse).  Here one step is
//   evolve_pre   (elementwise)  : forward move + the 9 backward query points of every evolved point,
//                                  scattered into one query array per TARGET element;
//   predict_device (DMMA GEMMs) : one batched prediction per target element over all its queries;
//   evolve_post  (elementwise)  : phases, off-diagonal rotations and branch recombination.
// The geometry is recomputed in the post kernel (a few PES evaluations) instead of being stored.
#include "evolve.cuh"
#include "gpr.cuh"
#include "gpr_kernels.cuh"

namespace gple
{
namespace
{
struct Sym2
{
	double a00, a01, a11;
};

__device__ __forceinline__ int sgn(const double v)
{
	return (v > 0.0) - (v < 0.0);
}

/// gple/pes.cpp:42-63
__device__ __forceinline__ Sym2 diabatic_potential(const int model, const double x)
{
	Sym2 V{0.0, 0.0, 0.0};
	if (model == GPLE_SAC)
	{
		V.a00 = sgn(x) * 0.01 * (1.0 - exp(-sgn(x) * 1.6 * x));
		V.a11 = -V.a00;
		V.a01 = 0.005 * exp(-1.0 * (x * x));
	}
	else if (model == GPLE_DAC)
	{
		V.a11 = 0.05 - 0.10 * exp(-0.28 * (x * x));
		V.a01 = 0.015 * exp(-0.06 * (x * x));
	}
	else
	{
		V.a00 = 6e-4;
		V.a11 = -6e-4;
		V.a01 = 0.10 * (1 - sgn(x) * (exp(-sgn(x) * 0.90 * x) - 1));
	}
	return V;
}

/// gple/pes.cpp:69-88
__device__ __forceinline__ Sym2 diabatic_force(const int model, const double x)
{
	Sym2 F{0.0, 0.0, 0.0};
	if (model == GPLE_SAC)
	{
		F.a00 = -0.01 * 1.6 * exp(-sgn(x) * 1.6 * x);
		F.a11 = -F.a00;
		F.a01 = 2.0 * 0.005 * 1.0 * x * exp(-1.0 * (x * x));
	}
	else if (model == GPLE_DAC)
	{
		F.a11 = -2 * 0.10 * 0.28 * x * exp(-0.28 * (x * x));
		F.a01 = 2 * 0.015 * 0.06 * x * exp(-0.06 * (x * x));
	}
	else
	{
		F.a01 = -0.10 * 0.90 * exp(-sgn(x) * 0.90 * x);
	}
	return F;
}

/// All adiabatic quantities at one position from a single evaluation of the diabatic matrices:
/// E (pes.cpp:127-148), F_adia = C^T F C, lower triangle (pes.cpp:100-125,154-167), d_10 (pes.cpp:172-189)
struct Adiabatic
{
	double E0, E1, F00, F10, F11, d10;
};
__device__ __forceinline__ Adiabatic adiabatic(const int model, const double x)
{
	const Sym2 V = diabatic_potential(model, x), F = diabatic_force(model, x);
	const double diff = V.a00 - V.a11;
	const double s = sqrt(diff * diff + 4.0 * (V.a01 * V.a01));
	const double r00 = (-s + diff) / (2.0 * V.a01), r01 = (s + diff) / (2.0 * V.a01);
	const double n0 = sqrt(r00 * r00 + 1.0), n1 = sqrt(r01 * r01 + 1.0);
	const double c00 = r00 / n0, c01 = r01 / n1, c10 = 1.0 / n0, c11 = 1.0 / n1;
	Adiabatic a;
	a.E0 = (-s + (V.a00 + V.a11)) / 2.0;
	a.E1 = (s + (V.a00 + V.a11)) / 2.0;
	const double t00 = c00 * F.a00 + c10 * F.a01, t01 = c00 * F.a01 + c10 * F.a11;
	const double t10 = c01 * F.a00 + c11 * F.a01, t11 = c01 * F.a01 + c11 * F.a11;
	a.F00 = t00 * c00 + t01 * c10;
	a.F10 = t10 * c00 + t11 * c10;
	a.F11 = t10 * c01 + t11 * c11;
	a.d10 = a.F10 / (a.E1 - a.E0);
	return a;
}

/// gple/evolve.cpp:53-100 (CouplingCriterion == 0: true unless a NaN appears; quirks q8, q9)
__device__ __forceinline__ bool is_coupling(const Adiabatic& a, const double p, const double mass, const double dt)
{
	const double favg = (0.0 + a.F00 + a.F11) / 2.0;
	return fabs(-a.d10 * p / mass) * dt >= 0.0 || fabs(a.F10 / favg) >= 0.0;
}

/// gple/evolve.cpp:125-148
__device__ __forceinline__ void adiabatic_evolve(const int model, double& x, double& p, const double mass, const double dt, const int drc, const int row, const int col)
{
	x += drc * dt / 2.0 * (p / mass);
	const Adiabatic a = adiabatic(model, x);
	const double fr = row == 0 ? a.F00 : a.F11, fc = col == 0 ? a.F00 : a.F11;
	p += drc * dt / 2.0 * (fr + fc);
	x += drc * dt / 2.0 * (p / mass);
}

struct Geometry
{
	double x2, p1;
	double p2[3];
	double qx[3][3], qp[3][3]; // [target element (00, 10, 11)][branch]
};

/// gple/evolve.cpp:232-266
__device__ __forceinline__ Geometry backward_geometry(const int model, const double x0, const double p0, const double mass, const double dt, const int row, const int col)
{
	Geometry g;
	const bool couple = is_coupling(adiabatic(model, x0), p0, mass, dt);
	double x2 = x0, p1 = p0;
	adiabatic_evolve(model, x2, p1, mass, dt / 2.0, -1, row, col);
	g.x2 = x2;
	g.p1 = p1;
	const double f01 = adiabatic(model, x2).F10 * (couple ? 1.0 : 0.0);
#pragma unroll
	for (int b = 0; b < 3; b++)
	{
		const double nbr = double(b - 1);
		g.p2[b] = p1 + dt * -1.0 * nbr * f01;
		const double x3 = x2 + -1 * (dt / 4.0) * g.p2[b] / mass;
		const Adiabatic a = adiabatic(model, x3);
#pragma unroll
		for (int e = 0; e < 3; e++)
		{
			const double fi = (e == 0) ? a.F00 : a.F11, fj = (e == 2) ? a.F11 : a.F00; // (0,0), (1,0), (1,1)
			const double p3 = g.p2[b] + -1 * (dt / 2.0) / 2.0 * (fi + fj);
			g.qp[e][b] = p3;
			g.qx[e][b] = x3 + -1 * (dt / 4.0) * p3 / mass;
		}
	}
	return g;
}

struct Cplx
{
	double re, im;
};
__device__ __forceinline__ Cplx cmul(const Cplx a, const Cplx b)
{
	return Cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ Cplx cexp_i(const double phase)
{
	double s, c;
	sincos(phase, &s, &c);
	return Cplx{c, s};
}

/// gple/evolve.cpp:214-228
__device__ __forceinline__ void offdiagonal_rotation(const int model, Cplx (&rho)[3], const double x, const double p, const double mass, const double dt)
{
	const Adiabatic a = adiabatic(model, x);
	const double phi = p / mass * (-a.d10) * (is_coupling(a, p, mass, dt) ? 1.0 : 0.0);
	double s, c;
	sincos(2.0 * phi * dt, &s, &c);
	const Cplx o0 = rho[0], o1 = rho[1], o2 = rho[2];
	rho[0] = Cplx{(1.0 + c) / 2.0 * o0.re - s * o1.re + (1.0 - c) / 2.0 * o2.re, (1.0 + c) / 2.0 * o0.im + (1.0 - c) / 2.0 * o2.im};
	rho[1] = Cplx{s / 2.0 * o0.re + c * o1.re - s / 2.0 * o2.re, s / 2.0 * o0.im + o1.im - s / 2.0 * o2.im};
	rho[2] = Cplx{(1.0 - c) / 2.0 * o0.re + s * o1.re + (1.0 + c) / 2.0 * o2.re, (1.0 - c) / 2.0 * o0.im + (1.0 + c) / 2.0 * o2.im};
}

/// Where the queries of one source element land in the per-target query arrays.
struct QueryMap
{
	double2* q[3];		 // query coordinate arrays of the 3 target elements (nullptr: no predictor)
	long long offset[3]; // first query of this source element in each target array
	int nbranch[3];		 // 3, or 2 when target == source and the own density is known (branches -1, +1)
};

/// Forward move (evolve.cpp:399-407) and query generation.  `forward == 0`: new_point_predict (no move).
__global__ void __launch_bounds__(256) evolve_pre_kernel(const int model, double4* __restrict__ pts, const double2* __restrict__ r_in, const long long n, const int row, const int col, const double mass, const double dt, const int forward, const QueryMap qm)
{
	const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k >= n)
	{
		return;
	}
	double x, p;
	if (forward)
	{
		const double4 pt = pts[k];
		x = pt.x;
		p = pt.y;
		adiabatic_evolve(model, x, p, mass, dt / 2, 1, row, col);
		adiabatic_evolve(model, x, p, mass, dt / 2, 1, row, col);
		pts[k] = make_double4(x, p, pt.z, pt.w);
	}
	else
	{
		const double2 r = r_in[k];
		x = r.x;
		p = r.y;
	}
	const Geometry g = backward_geometry(model, x, p, mass, dt, row, col);
#pragma unroll
	for (int e = 0; e < 3; e++)
	{
		if (qm.q[e] == nullptr)
		{
			continue;
		}
		double2* dst = qm.q[e] + qm.offset[e] + k * qm.nbranch[e];
		if (qm.nbranch[e] == 3)
		{
			dst[0] = make_double2(g.qx[e][0], g.qp[e][0]);
			dst[1] = make_double2(g.qx[e][1], g.qp[e][1]);
			dst[2] = make_double2(g.qx[e][2], g.qp[e][2]);
		}
		else
		{
			dst[0] = make_double2(g.qx[e][0], g.qp[e][0]);
			dst[1] = make_double2(g.qx[e][2], g.qp[e][2]);
		}
	}
}

struct PredMap
{
	const double* c[3]; // cutoff predictions per target: real (1 double) for 00 / 11, complex (2 doubles) for 10
	long long offset[3];
	int nbranch[3];
};

/// evolve.cpp:269-365: gather the 9 densities, apply phases / rotations, recombine the branches.
__global__ void __launch_bounds__(256) evolve_post_kernel(const int model, double4* __restrict__ pts, const double2* __restrict__ r_in, double2* __restrict__ rho_out, const long long n, const int row, const int col, const double mass, const double dt, const int forward, const PredMap pm)
{
	const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (k >= n)
	{
		return;
	}
	double x0, p0;
	Cplx own{0.0, 0.0};
	if (forward)
	{
		const double4 pt = pts[k];
		x0 = pt.x;
		p0 = pt.y;
		own = Cplx{pt.z, pt.w};
	}
	else
	{
		const double2 r = r_in[k];
		x0 = r.x;
		p0 = r.y;
	}
	const Geometry g = backward_geometry(model, x0, p0, mass, dt, row, col);
	const int self = row + col; // lower-triangular index of (row, col): (0,0)->0, (1,0)->1, (1,1)->2
	Cplx rho[3][3];				// [element][branch]
#pragma unroll
	for (int e = 0; e < 3; e++)
	{
#pragma unroll
		for (int b = 0; b < 3; b++)
		{
			Cplx v{0.0, 0.0};
			if (forward && e == self && b == 1)
			{
				v = own;
			}
			else if (pm.c[e] != nullptr)
			{
				const int slot = (pm.nbranch[e] == 3) ? b : (b >> 1);
				const long long idx = pm.offset[e] + k * pm.nbranch[e] + slot;
				if (e == 1)
				{
					v = Cplx{pm.c[e][2 * idx], pm.c[e][2 * idx + 1]};
				}
				else
				{
					v = Cplx{pm.c[e][idx], 0.0};
				}
			}
			rho[e][b] = v;
		}
	}
	const Adiabatic a2 = adiabatic(model, g.x2);
	Cplx comb[3] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
	for (int b = 0; b < 3; b++)
	{
		// evolve.cpp:310-311: omega0(x2, x4[(1,0)][b], Forward, 0, 1)
		const Adiabatic a4 = adiabatic(model, g.qx[1][b]);
		const double omega = 1 * (a2.E0 - a2.E1 + a4.E0 - a4.E1) / 2.0 / 1.0;
		rho[1][b] = cmul(rho[1][b], cexp_i(omega * dt / 2));
		Cplx view[3] = {rho[0][b], rho[1][b], rho[2][b]};
		offdiagonal_rotation(model, view, g.x2, g.p2[b], mass, dt / 2.0);
		if (b == 0)
		{
			const Cplx value{(view[0].re + 2.0 * view[1].re + view[2].re) / 4.0, (view[0].im + view[2].im) / 4.0};
			comb[0].re += value.re;
			comb[0].im += value.im;
			comb[1].re += value.re;
			comb[1].im += value.im;
			comb[2].re += value.re;
			comb[2].im += value.im;
		}
		else if (b == 1)
		{
			const Cplx value{(view[0].re - view[2].re) / 2.0, (view[0].im - view[2].im) / 2.0};
			comb[0].re += value.re;
			comb[0].im += value.im;
			comb[1].im += view[1].im;
			comb[2].re -= value.re;
			comb[2].im -= value.im;
		}
		else
		{
			const Cplx value{(view[0].re - 2.0 * view[1].re + view[2].re) / 4.0, (view[0].im + view[2].im) / 4.0};
			comb[0].re += value.re;
			comb[0].im += value.im;
			comb[1].re -= value.re;
			comb[1].im -= value.im;
			comb[2].re += value.re;
			comb[2].im += value.im;
		}
	}
	offdiagonal_rotation(model, comb, g.x2, g.p1, mass, dt / 2.0);
	Cplx result = comb[self];
	if (row != col)
	{
		const Adiabatic a0 = adiabatic(model, x0);
		const double omega = 1 * (a0.E0 - a0.E1 + a2.E0 - a2.E1) / 2.0 / 1.0;
		result = cmul(result, cexp_i(omega * dt / 2.0));
	}
	if (forward)
	{
		pts[k] = make_double4(x0, p0, result.re, result.im);
	}
	else
	{
		// new_point_predict (evolve.cpp:434-442) returns 0 where is_coupling is false.  With CouplingCriterion == 0 that only happens
		// where the criterion itself is NaN (e.g. Tully's SAC beyond |x| ~ 27, where V01 underflows and the adiabatic forces are
		// 0 / 0): the reference reports an exact 0 there, which is_very_small counts as "small", not a NaN
		const bool couple = is_coupling(adiabatic(model, x0), p0, mass, dt);
		rho_out[k] = couple ? make_double2(result.re, result.im) : make_double2(0.0, 0.0);
	}
}

__global__ void pes_kernel(const int model, const double* __restrict__ x, const long long n, double* __restrict__ E, double* __restrict__ F, double* __restrict__ D)
{
	const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (i >= n)
	{
		return;
	}
	const Adiabatic a = adiabatic(model, x[i]);
	E[2 * i] = a.E0;
	E[2 * i + 1] = a.E1;
	F[3 * i] = a.F00;
	F[3 * i + 1] = a.F10;
	F[3 * i + 2] = a.F11;
	D[i] = a.d10;
}

constexpr int OBS_BLOCKS = 296;
/// gple/predict.cpp:43-244: nine running sums in one pass, two-stage deterministic reduction
__global__ void __launch_bounds__(256) observables_kernel(const int model, const double4* __restrict__ pts, const long long n, const double mass, const int pes_index, double* __restrict__ part)
{
	__shared__ double scratch[9 * 8];
	double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	for (long long k = blockIdx.x * 256ll + threadIdx.x; k < n; k += 256ll * gridDim.x)
	{
		const double4 pt = pts[k];
		const double x = pt.x, p = pt.y, w = pt.z;
		const Adiabatic a = adiabatic(model, x);
		s[0] += w;
		s[1] += x * w;
		s[2] += p * w;
		s[3] += x;
		s[4] += p;
		s[5] += x * x;
		s[6] += p * p;
		s[7] += ((p * p / mass) / 2.0 + (pes_index == 0 ? a.E0 : a.E1)) * w;
		s[8] += pt.z * pt.z + pt.w * pt.w;
	}
	block_reduce<9, 256>(s, scratch);
	if (threadIdx.x == 0)
	{
		for (int i = 0; i < 9; i++)
		{
			part[blockIdx.x * 9 + i] = s[i];
		}
	}
}
__global__ void observables_final_kernel(const double* __restrict__ part, const int blocks, double* __restrict__ out)
{
	const int i = threadIdx.x;
	if (i < 9)
	{
		double s = 0.0;
		for (int b = 0; b < blocks; b++)
		{
			s += part[b * 9 + i];
		}
		out[i] = s;
	}
}

int branches(const int source, const int target, const bool own_density)
{
	return (own_density && source == target) ? 2 : 3;
}

} // namespace

/// Shared by gple_evolve (forward = 1: three point sets in place) and gple_new_point_predict (forward = 0).
void evolve_device(gple_ctx* ctx, const int pes_model, const gple_model* const models[3], double* d_pts[3], const size_t counts[3], const double mass, const double dt)
{
	// totals per target element
	long long total[3] = {0, 0, 0}, offset[3][3];
	for (int t = 0; t < 3; t++)
	{
		for (int s = 0; s < 3; s++)
		{
			offset[s][t] = total[t];
			total[t] += (long long)counts[s] * branches(s, t, true);
		}
	}
	double2* qbuf[3] = {nullptr, nullptr, nullptr};
	double* cbuf[3] = {nullptr, nullptr, nullptr};
	const char* qn[3] = {"evolve.q0", "evolve.q1", "evolve.q2"};
	const char* cn[3] = {"evolve.c0", "evolve.c1", "evolve.c2"};
	for (int t = 0; t < 3; t++)
	{
		if (models[t] != nullptr && total[t] > 0)
		{
			qbuf[t] = ctx->ws.get<double2>(qn[t], size_t(total[t]));
			cbuf[t] = ctx->ws.get<double>(cn[t], size_t(total[t]) * (t == 1 ? 2 : 1));
		}
	}
	static const int rows[3] = {0, 1, 1}, cols[3] = {0, 0, 1};
	for (int s = 0; s < 3; s++)
	{
		if (counts[s] == 0)
		{
			continue;
		}
		QueryMap qm{};
		for (int t = 0; t < 3; t++)
		{
			qm.q[t] = qbuf[t];
			qm.offset[t] = offset[s][t];
			qm.nbranch[t] = branches(s, t, true);
		}
		GPLE_LAUNCH(ctx, evolve_pre_kernel, unsigned((counts[s] + 255) / 256), 256, 0, pes_model, reinterpret_cast<double4*>(d_pts[s]), nullptr, (long long)counts[s], rows[s], cols[s], mass, dt, 1, qm);
	}
	for (int t = 0; t < 3; t++)
	{
		if (qbuf[t] != nullptr)
		{
			predict_device(ctx, models[t], reinterpret_cast<const double*>(qbuf[t]), size_t(total[t]), nullptr, nullptr, nullptr, cbuf[t], nullptr);
		}
	}
	for (int s = 0; s < 3; s++)
	{
		if (counts[s] == 0)
		{
			continue;
		}
		PredMap pm{};
		for (int t = 0; t < 3; t++)
		{
			pm.c[t] = cbuf[t];
			pm.offset[t] = offset[s][t];
			pm.nbranch[t] = branches(s, t, true);
		}
		GPLE_LAUNCH(ctx, evolve_post_kernel, unsigned((counts[s] + 255) / 256), 256, 0, pes_model, reinterpret_cast<double4*>(d_pts[s]), nullptr, nullptr, (long long)counts[s], rows[s], cols[s], mass, dt, 1, pm);
	}
}

void new_point_predict_device(gple_ctx* ctx, const int pes_model, const gple_model* const models[3], const double* d_r, const size_t n, const int row, const int col, const double mass, const double dt, double* d_out)
{
	double2* qbuf[3] = {nullptr, nullptr, nullptr};
	double* cbuf[3] = {nullptr, nullptr, nullptr};
	const char* qn[3] = {"evolve.q0", "evolve.q1", "evolve.q2"};
	const char* cn[3] = {"evolve.c0", "evolve.c1", "evolve.c2"};
	QueryMap qm{};
	PredMap pm{};
	for (int t = 0; t < 3; t++)
	{
		if (models[t] != nullptr)
		{
			qbuf[t] = ctx->ws.get<double2>(qn[t], 3 * n);
			cbuf[t] = ctx->ws.get<double>(cn[t], 3 * n * (t == 1 ? 2 : 1));
		}
		qm.q[t] = qbuf[t];
		qm.offset[t] = 0;
		qm.nbranch[t] = 3;
	}
	GPLE_LAUNCH(ctx, evolve_pre_kernel, unsigned((n + 255) / 256), 256, 0, pes_model, nullptr, reinterpret_cast<const double2*>(d_r), (long long)n, row, col, mass, dt, 0, qm);
	for (int t = 0; t < 3; t++)
	{
		if (qbuf[t] != nullptr)
		{
			predict_device(ctx, models[t], reinterpret_cast<const double*>(qbuf[t]), 3 * n, nullptr, nullptr, nullptr, cbuf[t], nullptr);
		}
		pm.c[t] = cbuf[t];
		pm.offset[t] = 0;
		pm.nbranch[t] = 3;
	}
	GPLE_LAUNCH(ctx, evolve_post_kernel, unsigned((n + 255) / 256), 256, 0, pes_model, nullptr, reinterpret_cast<const double2*>(d_r), reinterpret_cast<double2*>(d_out), (long long)n, row, col, mass, dt, 0, pm);
}

void pes_device(gple_ctx* ctx, const int pes_model, const double* d_x, const size_t n, double* E, double* F, double* D)
{
	GPLE_LAUNCH(ctx, pes_kernel, unsigned((n + 255) / 256), 256, 0, pes_model, d_x, (long long)n, E, F, D);
}

void observables_device(gple_ctx* ctx, const int pes_model, const double* d_pts, const size_t n, const double mass, const int pes_index, double* d_out9)
{
	double* part = ctx->ws.get<double>("obs.part", OBS_BLOCKS * 9);
	const int blocks = int(std::min<size_t>(OBS_BLOCKS, (n + 255) / 256));
	GPLE_LAUNCH(ctx, observables_kernel, blocks, 256, 0, pes_model, reinterpret_cast<const double4*>(d_pts), (long long)n, mass, pes_index, part);
	GPLE_LAUNCH(ctx, observables_final_kernel, 1, 32, 0, part, blocks, d_out9);
}

} // namespace gple
