// TEST INFRASTRUCTURE ONLY -- CPU oracle, dynamics part: Tully-model PES (gple/pes.cpp), the per-point
// Trotter evolve step (gple/evolve.cpp), Monte-Carlo-integral observables (gple/predict.cpp:43-244) and
// the analytic initial distribution (gple/mc.cpp:30-50).  PARITY UNPINNED (see gple_oracle.hpp).
// The reference fixes the model at compile time (pes.h:38-41); here it is a run-time argument.
#pragma once
#include "gple_oracle_complex.hpp"

namespace orc
{
enum Model : int
{
	SAC = 0,
	DAC = 1,
	ECR = 2
};

/// gple/pes.h:14-17
inline int sgn(const double v)
{
	return (v > 0.0) - (v < 0.0);
}

struct Sym2
{
	double a00, a01, a11;
};

/// gple/pes.cpp:42-63 (constants pes.cpp:12-36)
inline Sym2 diabatic_potential(const Model m, const double x)
{
	Sym2 V{0.0, 0.0, 0.0};
	switch (m)
	{
	case SAC:
		V.a00 = sgn(x) * 0.01 * (1.0 - std::exp(-sgn(x) * 1.6 * x));
		V.a11 = -V.a00;
		V.a01 = 0.005 * std::exp(-1.0 * (x * x));
		break;
	case DAC:
		V.a11 = 0.05 - 0.10 * std::exp(-0.28 * (x * x));
		V.a01 = 0.015 * std::exp(-0.06 * (x * x));
		break;
	case ECR:
		V.a00 = 6e-4;
		V.a11 = -6e-4;
		V.a01 = 0.10 * (1 - sgn(x) * (std::exp(-sgn(x) * 0.90 * x) - 1));
		break;
	}
	return V;
}

/// gple/pes.cpp:69-88
inline Sym2 diabatic_force(const Model m, const double x)
{
	Sym2 F{0.0, 0.0, 0.0};
	switch (m)
	{
	case SAC:
		F.a00 = -0.01 * 1.6 * std::exp(-sgn(x) * 1.6 * x);
		F.a11 = -F.a00;
		F.a01 = 2.0 * 0.005 * 1.0 * x * std::exp(-1.0 * (x * x));
		break;
	case DAC:
		F.a11 = -2 * 0.10 * 0.28 * x * std::exp(-0.28 * (x * x));
		F.a01 = 2 * 0.015 * 0.06 * x * std::exp(-0.06 * (x * x));
		break;
	case ECR:
		F.a01 = -0.10 * 0.90 * std::exp(-sgn(x) * 0.90 * x);
		break;
	}
	return F;
}

/// gple/pes.cpp:100-123 (NumPES == 2 branch); returns C row-major {c00, c01, c10, c11}
inline std::array<double, 4> diabatic_to_adiabatic_matrix(const Model m, const double x)
{
	const Sym2 V = diabatic_potential(m, x);
	const double s = std::sqrt(sq(V.a00 - V.a11) + 4.0 * sq(V.a01));
	double r00 = -1.0 * s, r01 = 1.0 * s;
	r00 += V.a00 - V.a11;
	r01 += V.a00 - V.a11;
	r00 /= 2.0 * V.a01;
	r01 /= 2.0 * V.a01;
	const double r10 = 1.0, r11 = 1.0;
	const double n0 = std::sqrt(r00 * r00 + r10 * r10), n1 = std::sqrt(r01 * r01 + r11 * r11);
	return {r00 / n0, r01 / n1, r10 / n0, r11 / n1};
}

/// gple/pes.cpp:127-148
inline std::array<double, 2> adiabatic_potential(const Model m, const double x)
{
	const Sym2 V = diabatic_potential(m, x);
	const double s = std::sqrt(sq(V.a00 - V.a11) + sq(2.0 * V.a01));
	return {(-1.0 * s + (V.a00 + V.a11)) / 2.0, (1.0 * s + (V.a00 + V.a11)) / 2.0};
}

/// gple/pes.cpp:154-167; returns the symmetric matrix built from the LOWER triangle of C^T F C
inline Sym2 adiabatic_force(const Model m, const double x)
{
	const Sym2 F = diabatic_force(m, x);
	const auto C = diabatic_to_adiabatic_matrix(m, x);
	// T = C^T F (2x2), then T C
	const double t00 = C[0] * F.a00 + C[2] * F.a01, t01 = C[0] * F.a01 + C[2] * F.a11;
	const double t10 = C[1] * F.a00 + C[3] * F.a01, t11 = C[1] * F.a01 + C[3] * F.a11;
	Sym2 R;
	R.a00 = t00 * C[0] + t01 * C[2];
	R.a01 = t10 * C[0] + t11 * C[2]; // element (1,0); mirrored to (0,1) by selfadjointView<Lower>
	R.a11 = t10 * C[1] + t11 * C[3];
	return R;
}

/// gple/pes.cpp:172-189; returns d_{10} (d_{01} = -d_{10})
inline double adiabatic_coupling_10(const Model m, const double x)
{
	const auto E = adiabatic_potential(m, x);
	const Sym2 F = adiabatic_force(m, x);
	return F.a01 / (E[1] - E[0]);
}

using Distribution = std::function<cplx(const double x, const double p, std::size_t row, std::size_t col)>;

/// gple/mc.cpp:30-50
inline cplx initial_distribution(
	const std::array<double, 2>& r0,
	const std::array<double, 2>& sigma0,
	const double x,
	const double p,
	const std::size_t row,
	const std::size_t col,
	const std::array<double, 2>& init_pop,
	const std::array<double, 2>& init_phase
)
{
	const double gw = std::exp(-(sq((x - r0[0]) / sigma0[0]) + sq((p - r0[1]) / sigma0[1])) / 2.0) / (2.0 * std::numbers::pi * (sigma0[0] * sigma0[1]));
	const double sw = 0.0 + sq(init_pop[0]) + sq(init_pop[1]);
	return gw * init_pop[row] * init_pop[col] / sw * std::exp(1.0i * (init_phase[row] - init_phase[col]));
}

/// gple/evolve.cpp:53-100 with CouplingCriterion == 0, IsAdiabatic == false (quirks q8, q9)
inline bool is_coupling(const Model m, const double x, const double p, const double mass, const double dt)
{
	const Sym2 F = adiabatic_force(m, x);
	const double nac01 = -adiabatic_coupling_10(m, x);
	const double favg = (0.0 + F.a00 + F.a11) / 2.0;
	return std::abs(nac01 * p / mass) * dt >= 0.0 || std::abs(F.a01 / favg) >= 0.0;
}

/// gple/evolve.cpp:125-148
inline void adiabatic_evolve(const Model m, double& x, double& p, const double mass, const double dt, const int drc, const std::size_t row, const std::size_t col)
{
	x += drc * dt / 2.0 * (p / mass);
	const Sym2 F = adiabatic_force(m, x);
	const double fr = row == 0 ? F.a00 : F.a11, fc = col == 0 ? F.a00 : F.a11;
	p += drc * dt / 2.0 * (fr + fc);
	x += drc * dt / 2.0 * (p / mass);
}

/// gple/evolve.cpp:157-172
inline double calculate_omega0(const Model m, const double x0, const double x2, const int drc, const std::size_t row, const std::size_t col)
{
	if (row == col)
	{
		return 0.0;
	}
	const auto E0 = adiabatic_potential(m, x0), E2 = adiabatic_potential(m, x2);
	return drc * (E0[row] - E0[col] + E2[row] - E2[col]) / 2.0 / hbar;
}

inline std::size_t tri_index(const std::size_t row, const std::size_t col)
{
	return row * (row + 1) / 2 + col;
}

/// gple/evolve.cpp:214-228
inline void offdiagonal_rotation(const Model m, std::array<cplx, 3>& rho, const double x, const double p, const double mass, const double dt)
{
	const double phi = p / mass * (-adiabatic_coupling_10(m, x)) * static_cast<double>(is_coupling(m, x, p, mass, dt));
	const double c = std::cos(2.0 * phi * dt), s = std::sin(2.0 * phi * dt);
	const std::array<cplx, 3> o = rho;
	rho[0] = (1.0 + c) / 2.0 * o[0] - s * o[1].real() + (1.0 - c) / 2.0 * o[2];
	rho[1] = s / 2.0 * o[0] + c * o[1].real() + 1.0i * o[1].imag() - s / 2.0 * o[2];
	rho[2] = (1.0 - c) / 2.0 * o[0] + s * o[1].real() + (1.0 + c) / 2.0 * o[2];
}

/// The 9 backward-propagated query points of gple/evolve.cpp:232-266.  q[e][b] = (x4, p3) for
/// target element e (lower-triangular index) and branch b (n = -1, 0, +1); also returns x2, p1, p2[b].
struct BackwardGeometry
{
	double x2, p1;
	std::array<double, 3> p2;
	std::array<std::array<std::array<double, 2>, 3>, 3> q;
};

inline BackwardGeometry backward_geometry(const Model m, const double x0, const double p0, const double mass, const double dt, const std::size_t row, const std::size_t col)
{
	constexpr int drc = -1;
	BackwardGeometry g;
	const bool couple = is_coupling(m, x0, p0, mass, dt);
	double x2 = x0, p1 = p0;
	adiabatic_evolve(m, x2, p1, mass, dt / 2.0, drc, row, col);
	g.x2 = x2;
	g.p1 = p1;
	const double f01 = adiabatic_force(m, x2).a01 * static_cast<double>(couple);
	for (std::size_t b = 0; b < 3; b++)
	{
		const double n = static_cast<double>(static_cast<int>(b) - 1);
		g.p2[b] = p1 + dt * static_cast<double>(drc) * n * f01;
		const double x3 = x2 + drc * (dt / 4.0) * g.p2[b] / mass;
		const Sym2 F = adiabatic_force(m, x3);
		for (std::size_t i = 0; i < NumPES; i++)
		{
			for (std::size_t j = 0; j <= i; j++)
			{
				const double fi = i == 0 ? F.a00 : F.a11, fj = j == 0 ? F.a00 : F.a11;
				const double p3 = g.p2[b] + drc * (dt / 2.0) / 2.0 * (fi + fj);
				const double x4 = x3 + drc * (dt / 4.0) * p3 / mass;
				g.q[tri_index(i, j)][b] = {x4, p3};
			}
		}
	}
	return g;
}

/// gple/evolve.cpp:184-372
inline cplx non_adiabatic_evolve_predict(
	const Model m,
	const double x0,
	const double p0,
	const std::optional<cplx> density,
	const double mass,
	const double dt,
	const Distribution& distribution,
	const std::size_t row,
	const std::size_t col
)
{
	const BackwardGeometry g = backward_geometry(m, x0, p0, mass, dt, row, col);
	std::array<std::array<cplx, 3>, 3> rho; // [element][branch]
	for (std::size_t i = 0; i < NumPES; i++)
	{
		for (std::size_t j = 0; j <= i; j++)
		{
			const std::size_t e = tri_index(i, j);
			for (std::size_t b = 0; b < 3; b++)
			{
				if (i == row && j == col && b == 1 && density.has_value())
				{
					rho[e][b] = density.value();
				}
				else
				{
					rho[e][b] = distribution(g.q[e][b][0], g.q[e][b][1], i, j);
				}
			}
		}
	}
	std::array<cplx, 3> comb{0.0, 0.0, 0.0};
	for (std::size_t b = 0; b < 3; b++)
	{
		rho[1][b] *= std::exp(calculate_omega0(m, g.x2, g.q[1][b][0], 1, 0, 1) * dt / 2 * 1.0i);
		std::array<cplx, 3> view{rho[0][b], rho[1][b], rho[2][b]};
		offdiagonal_rotation(m, view, g.x2, g.p2[b], mass, dt / 2.0);
		switch (static_cast<int>(b) - 1)
		{
		case -1:
		{
			const cplx value = (view[0] + 2.0 * view[1].real() + view[2]) / 4.0;
			comb[0] += value;
			comb[1] += value;
			comb[2] += value;
			break;
		}
		case 0:
		{
			const cplx value = (view[0] - view[2]) / 2.0;
			comb[0] += value;
			comb[1] += 1.0i * view[1].imag();
			comb[2] -= value;
			break;
		}
		default:
		{
			const cplx value = (view[0] - 2.0 * view[1].real() + view[2]) / 4.0;
			comb[0] += value;
			comb[1] -= value;
			comb[2] += value;
			break;
		}
		}
	}
	offdiagonal_rotation(m, comb, g.x2, g.p1, mass, dt / 2.0);
	const cplx result = comb[tri_index(row, col)];
	if (row != col)
	{
		return result * std::exp(calculate_omega0(m, x0, g.x2, 1, 0, 1) * dt / 2.0 * 1.0i);
	}
	return result;
}

/// One phase-space point: 32-byte AoS exactly like gple/storage.h:232-297 {r: (x, p), rho: complex}
struct PhaseSpacePoint
{
	double x, p;
	cplx rho;
};
static_assert(sizeof(PhaseSpacePoint) == 32);

/// gple/evolve.cpp:377-423, one element
inline void evolve_element(
	const Model m,
	PhaseSpacePoint* pts,
	const std::size_t n,
	const double mass,
	const double dt,
	const Distribution& distribution,
	const std::size_t row,
	const std::size_t col
)
{
	parallel_for(
		n,
		[&](const std::size_t k)
		{
			PhaseSpacePoint& psp = pts[k];
			const double x0 = psp.x, p0 = psp.p;
			if (is_coupling(m, x0, p0, mass, dt))
			{
				double x = x0, p = p0;
				adiabatic_evolve(m, x, p, mass, dt / 2, 1, row, col);
				adiabatic_evolve(m, x, p, mass, dt / 2, 1, row, col);
				psp.x = x;
				psp.p = p;
				psp.rho = non_adiabatic_evolve_predict(m, x, p, psp.rho, mass, dt, distribution, row, col);
			}
			else
			{
				double x = x0, p = p0;
				adiabatic_evolve(m, x, p, mass, dt, 1, row, col);
				psp.rho = distribution(x0, p0, row, col) * std::exp(-calculate_omega0(m, x0, x, 1, row, col) * dt * 1.0i);
				psp.x = x;
				psp.p = p;
			}
		},
		1
	);
}

/// gple/evolve.cpp:425-443
inline cplx new_point_predict(const Model m, const double x, const double p, const double mass, const double dt, const Distribution& distribution, const std::size_t row, const std::size_t col)
{
	if (is_coupling(m, x, p, mass, dt))
	{
		return non_adiabatic_evolve_predict(m, x, p, std::nullopt, mass, dt, distribution, row, col);
	}
	return 0.0;
}

/// Sums of gple/predict.cpp:65-244 for one element: out = {sum Re rho, sum x Re rho, sum p Re rho,
/// sum x, sum p, sum x^2, sum p^2, sum (p^2/2m + E_pes(x)) Re rho, sum |rho|^2}
inline std::array<double, 9> observable_sums(const Model m, const PhaseSpacePoint* pts, const std::size_t n, const double mass, const std::size_t pes_index)
{
	std::array<double, 9> s{};
	for (std::size_t k = 0; k < n; k++)
	{
		const double x = pts[k].x, p = pts[k].p, w = pts[k].rho.real();
		s[0] += w;
		s[1] += x * w;
		s[2] += p * w;
		s[3] += x;
		s[4] += p;
		s[5] += x * x;
		s[6] += p * p;
		s[7] += ((p * p / mass) / 2.0 + adiabatic_potential(m, x)[pes_index]) * w;
		s[8] += std::norm(pts[k].rho);
	}
	return s;
}

} // namespace orc
