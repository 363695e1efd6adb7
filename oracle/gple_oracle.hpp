// TEST INFRASTRUCTURE ONLY -- CPU oracle: a restatement of the reference's per-time-step GPR hot
// path (kaigu1997/gaussian_process_liouville_equation, directory
// gaussian_process_liouville_equation/ = "gple/").  PARITY UNPINNED: the reference ships no golden
// vectors / KATs / unit tests and cannot be compiled here (Eigen, NLopt,
// xtensor, MKL, TBB absent), so this oracle is pinned only by the analytic known answers,
// finite-difference, brute-force LOOCV, numpy and mpmath cross-checks in tests/test_oracle_*.py.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Every function cites the reference file:line it follows.  The arithmetic deliberately keeps the
// reference's formulation (pivoted LDL^T, explicit inverse via solve(Identity), per-query
// k K^-1 k^T variance, materialised derivative matrices), including its quirks.
#pragma once
#include "linalg.hpp"

#include <array>
#include <memory>
#include <numbers>
#include <optional>
#include <tuple>

namespace orc
{
// gple/stdafx.h:107-125
constexpr double hbar = 1.0;
constexpr std::size_t NumPES = 2, Dim = 1, PhaseDim = 2, NumTriangularElements = 3;
constexpr double PurityFactor = 2.0 * std::numbers::pi * hbar; // power<Dim>(2 pi hbar)
constexpr double ConnectingPoint = 2.0;						   // gple/kernel.h:16
constexpr double RescaleMaximum = 10.0;						   // gple/kernel.h:37
constexpr std::size_t NumRealParams = 4;					   // gple/kernel.h:33
constexpr std::size_t NumComplexParams = 8;					   // gple/complex_kernel.h:22

template <typename T>
inline T sq(const T x)
{
	return x * x;
}

/// Phase-space coordinates, 2 x n column-major (gple/stdafx.h:153). `id` emulates the reference's
/// `LeftFeature.data() == RightFeature.data()` identity test (kernel.cpp:11,50,112,187,205).
struct Points
{
	const double* p = nullptr;
	std::size_t n = 0;
	double x(std::size_t i) const { return p[2 * i]; }
	double mom(std::size_t i) const { return p[2 * i + 1]; }
};

/// gple/kernel.h:41 KernelParameter = (magnitude, characteristic lengths, noise)
struct KParam
{
	double mag = 1.0;
	std::array<double, PhaseDim> l{1.0, 1.0};
	double noise = 0.0;
};

// ----------------------------------------------------------------------------------------------
// gple/kernel.cpp
// ----------------------------------------------------------------------------------------------

/// gple/kernel.cpp:8-31
inline Mat delta_kernel(const Points& L, const Points& R, const bool same)
{
	if (same)
	{
		return Mat::identity(L.n);
	}
	Mat result(L.n, R.n);
	parallel_for(
		R.n,
		[&](const std::size_t c)
		{
			for (std::size_t r = 0; r < L.n; r++)
			{
				result(r, c) = static_cast<double>(L.x(r) == R.x(c) && L.mom(r) == R.mom(c));
			}
		},
		64
	);
	return result;
}

/// gple/kernel.cpp:38-85.  As called from KernelBase (kernel.cpp:227) the feature arguments are the
/// object's own member copies, so the "training set" branch is never taken:
/// every element is evaluated, including the diagonal (exp(-0) = 1).
inline Mat gaussian_kernel(const std::array<double, PhaseDim>& l, const Points& L, const Points& R)
{
	Mat result(L.n, R.n);
	parallel_for(
		R.n,
		[&](const std::size_t c)
		{
			for (std::size_t r = 0; r < L.n; r++)
			{
				const double d0 = (L.x(r) - R.x(c)) / l[0], d1 = (L.mom(r) - R.mom(c)) / l[1];
				result(r, c) += std::exp(-(d0 * d0 + d1 * d1) / 2.0);
			}
		},
		16
	);
	return result;
}

/// gple/kernel.cpp:99-160
inline std::array<Mat, PhaseDim> gaussian_derivative_over_char_length(
	const std::array<double, PhaseDim>& l,
	const Points& L,
	const Points& R,
	const bool same,
	const Mat& G
)
{
	std::array<Mat, PhaseDim> result{G, G};
	const bool training = same && L.n == R.n;
	parallel_for(
		R.n,
		[&](const std::size_t c)
		{
			for (std::size_t r = training ? c + 1 : 0; r < L.n; r++)
			{
				const double d0 = (L.x(r) - R.x(c)) / l[0], d1 = (L.mom(r) - R.mom(c)) / l[1];
				result[0](r, c) *= d0 * d0 / l[0];
				result[1](r, c) *= d1 * d1 / l[1];
			}
		},
		16
	);
	if (training)
	{
		for (std::size_t d = 0; d < PhaseDim; d++)
		{
			for (std::size_t i = 0; i < L.n; i++)
			{
				result[d](i, i) = 0.0;
			}
			result[d] = selfadjoint_lower(result[d]);
		}
	}
	return result;
}

/// gple/kernel.cpp:217-242 (+ calculate_derivative, kernel.cpp:168-215)
struct KernelBase
{
	KParam prm;
	Points L, R;
	bool same;
	Mat K;
	std::optional<std::array<Mat, NumRealParams>> dK;

	KernelBase(const KParam& P, const Points& left, const Points& right, const bool same_buffer, const bool deriv):
		prm(P), L(left), R(right), same(same_buffer)
	{
		K = gaussian_kernel(prm.l, L, R);
		const Mat delta = delta_kernel(L, R, same);
		const double m2 = sq(prm.mag), n2 = sq(prm.noise);
		for (std::size_t i = 0; i < K.d.size(); i++)
		{
			K.d[i] = m2 * (K.d[i] + n2 * delta.d[i]);
		}
		if (deriv)
		{
			std::array<Mat, NumRealParams> D;
			// magnitude (kernel.cpp:181)
			D[0] = K;
			for (auto& x : D[0].d)
			{
				x *= 2.0 / prm.mag;
			}
			// characteristic lengths (kernel.cpp:184-202)
			Mat G = K;
			if (same)
			{
				const double nn = sq(prm.mag * prm.noise);
				for (std::size_t i = 0; i < L.n; i++)
				{
					G(i, i) -= nn;
				}
			}
			auto dl = gaussian_derivative_over_char_length(prm.l, L, R, same, G);
			D[1] = std::move(dl[0]);
			D[2] = std::move(dl[1]);
			// noise (kernel.cpp:205-212)
			D[3] = Mat(L.n, R.n);
			if (same)
			{
				for (std::size_t i = 0; i < L.n; i++)
				{
					D[3](i, i) = 2.0 * m2 * prm.noise;
				}
			}
			dK = std::move(D);
		}
	}
};

/// gple/kernel.h:285-294
inline KParam construct_purity_auxiliary_kernel_params(const KParam& o)
{
	KParam r;
	r.mag = sq(o.mag) * std::sqrt(o.l[0] * o.l[1]);
	r.l = {std::numbers::sqrt2 * o.l[0], std::numbers::sqrt2 * o.l[1]};
	r.noise = 0.0;
	return r;
}

/// gple/kernel.h:301-332
template <typename T>
inline Vec cutoff_factor(const std::vector<T>& pred, const Vec& var)
{
	Vec result(pred.size());
	for (std::size_t i = 0; i < pred.size(); i++)
	{
		const double ps = std::norm(pred[i]);
		if (ps >= sq(ConnectingPoint) * var[i])
		{
			result[i] = 1.0;
		}
		else if (ps <= var[i])
		{
			result[i] = 0.0;
		}
		else
		{
			const double a = std::abs(pred[i]) / std::sqrt(var[i]);
			result[i] = (3.0 * ConnectingPoint - 2.0 * a - 1.0) * sq(a - 1) / ((ConnectingPoint - 1) * (ConnectingPoint - 1) * (ConnectingPoint - 1));
		}
	}
	return result;
}

inline double dot(const Vec& a, const Vec& b)
{
	double s = 0;
	for (std::size_t i = 0; i < a.size(); i++)
	{
		s += a[i] * b[i];
	}
	return s;
}
inline double sum(const Vec& a)
{
	double s = 0;
	for (const double x : a)
	{
		s += x;
	}
	return s;
}

/// gple/kernel.cpp:244-479
struct TrainingKernel
{
	std::vector<double> coords; // owned copy of the training features (2 x N)
	Points X;
	KParam prm;
	std::array<double, NumRealParams> theta;
	std::unique_ptr<KernelBase> base;
	double rescale = 1.0;
	Vec label;
	Mat inverse;
	Vec v;
	std::optional<double> error, population, purity;
	std::optional<std::array<double, PhaseDim>> first_order;
	std::optional<std::array<Mat, NumRealParams>> dinv;
	std::optional<std::array<Vec, NumRealParams>> dv;
	std::optional<std::array<double, NumRealParams>> derror, dpopulation, dpurity;

	TrainingKernel(const double* th, const double* feat, const cplx* y, const std::size_t N, const bool is_err, const bool is_avg, const bool is_deriv):
		coords(feat, feat + 2 * N)
	{
		X = Points{coords.data(), N};
		for (std::size_t i = 0; i < NumRealParams; i++)
		{
			theta[i] = th[i];
		}
		prm.mag = th[0];
		prm.l = {th[1], th[2]};
		prm.noise = th[3];
		base = std::make_unique<KernelBase>(prm, X, X, true, is_deriv);
		// kernel.cpp:279-280 (imaginary part of label discarded: quirk q7)
		double mx = 0.0;
		for (std::size_t i = 0; i < N; i++)
		{
			mx = std::max(mx, std::abs(y[i].real()));
		}
		rescale = RescaleMaximum / mx;
		label.resize(N);
		for (std::size_t i = 0; i < N; i++)
		{
			label[i] = y[i].real() * rescale;
		}
		// kernel.cpp:281-283
		const LDLT<double> dec(base->K);
		inverse = dec.solve(Mat::identity(N));
		v = dec.solve(label);
		// kernel.cpp:285
		if (is_err)
		{
			double e = 0;
			for (std::size_t i = 0; i < N; i++)
			{
				e += sq(v[i] / inverse(i, i));
			}
			error = e;
		}
		const double TwoPi = 2.0 * std::numbers::pi;
		std::unique_ptr<KernelBase> aux;
		if (is_avg)
		{
			// kernel.cpp:286-312
			const double f = TwoPi * sq(prm.mag) * prm.l[0] * prm.l[1];
			population = f * sum(v) / rescale;
			std::array<double, PhaseDim> r{0.0, 0.0};
			for (std::size_t i = 0; i < N; i++)
			{
				r[0] += X.x(i) * v[i];
				r[1] += X.mom(i) * v[i];
			}
			first_order = std::array<double, PhaseDim>{f * r[0] / rescale, f * r[1] / rescale};
			// kernel.cpp:313-335
			aux = std::make_unique<KernelBase>(construct_purity_auxiliary_kernel_params(prm), X, X, true, is_deriv);
			const Vec K1v = matvec(aux->K, v);
			purity = PurityFactor * std::numbers::pi * dot(v, K1v) / sq(rescale);
		}
		if (is_deriv)
		{
			// kernel.cpp:337-364
			std::array<Mat, NumRealParams> DI;
			DI[0] = inverse;
			for (auto& x : DI[0].d)
			{
				x *= -2.0 / prm.mag;
			}
			for (std::size_t d = 0; d < PhaseDim; d++)
			{
				Mat t = matmul(matmul(inverse, (*base->dK)[1 + d]), inverse);
				for (auto& x : t.d)
				{
					x = -x;
				}
				DI[1 + d] = std::move(t);
			}
			{
				Mat scaled = inverse;
				const double c = -2 * sq(prm.mag) * prm.noise;
				for (auto& x : scaled.d)
				{
					x *= c;
				}
				DI[3] = matmul(inverse, scaled);
			}
			// kernel.cpp:365-379
			std::array<Vec, NumRealParams> DV;
			for (std::size_t p = 0; p < NumRealParams; p++)
			{
				DV[p] = matvec(DI[p], label);
			}
			// kernel.cpp:381-400
			if (is_err)
			{
				std::array<double, NumRealParams> de{};
				for (std::size_t p = 0; p < NumRealParams; p++)
				{
					double s = 0;
					for (std::size_t i = 0; i < N; i++)
					{
						const double invd = inverse(i, i), diff = v[i] / invd;
						s += diff / invd * (DV[p][i] - diff * DI[p](i, i));
					}
					de[p] = 2.0 * s;
				}
				derror = de;
			}
			if (is_avg)
			{
				// kernel.cpp:401-435
				const double f = TwoPi * sq(prm.mag) * prm.l[0] * prm.l[1];
				std::array<double, NumRealParams> dp{};
				dp[0] = 0.0;
				for (std::size_t d = 0; d < PhaseDim; d++)
				{
					dp[1 + d] = f * (sum(v) / prm.l[d] + sum(DV[1 + d]));
				}
				dp[3] = f * sum(DV[3]);
				for (auto& x : dp)
				{
					x /= rescale;
				}
				dpopulation = dp;
				// kernel.cpp:436-477
				const double gf = PurityFactor * std::numbers::pi;
				std::array<double, NumRealParams> du{};
				const Vec K1v = matvec(aux->K, v);
				du[0] = 0.0;
				for (std::size_t d = 0; d < PhaseDim; d++)
				{
					Mat comb = aux->K;
					const Mat& dK1 = (*aux->dK)[1 + d];
					for (std::size_t i = 0; i < comb.d.size(); i++)
					{
						comb.d[i] = comb.d[i] / prm.l[d] + std::numbers::sqrt2 * dK1.d[i];
					}
					du[1 + d] = (dot(v, matvec(comb, v)) + 2.0 * dot(DV[1 + d], K1v)) * gf;
				}
				du[3] = 2.0 * gf * dot(DV[3], K1v);
				for (auto& x : du)
				{
					x /= sq(rescale);
				}
				dpurity = du;
			}
			dinv = std::move(DI);
			dv = std::move(DV);
		}
	}

	/// gple/kernel.h:167-179
	double get_magnitude() const
	{
		const double w = dot(label, v) / static_cast<double>(label.size());
		return w < 0 ? std::sqrt(-w) : std::sqrt(w);
	}
};

/// gple/kernel.cpp:481-544
struct PredictiveKernel
{
	std::unique_ptr<KernelBase> base;
	double rescale;
	Vec prediction, variance, cutoff_prediction;
	std::optional<double> error;
	std::optional<std::array<double, NumRealParams>> derror;

	PredictiveKernel(const double* feat, const std::size_t M, const TrainingKernel& k, const bool is_deriv, const double* test_label):
		rescale(k.rescale)
	{
		const Points T{feat, M};
		base = std::make_unique<KernelBase>(k.prm, T, k.X, false, is_deriv);
		const Mat& Ks = base->K;
		prediction = matvec(Ks, k.v); // kernel.cpp:495
		// kernel.cpp:496-518; prior = 1x1 self kernel incl. noise (quirk q3)
		variance.resize(M);
		const std::size_t N = k.X.n;
		const double prior = sq(k.prm.mag) * (std::exp(-0.0) + sq(k.prm.noise) * 1.0);
		parallel_for(
			M,
			[&](const std::size_t m)
			{
				// row * Inverse -> row vector, then dot with the row
				double q = 0.0;
				for (std::size_t c = 0; c < N; c++)
				{
					const double* ic = k.inverse.col(c);
					double s = 0.0;
					for (std::size_t r = 0; r < N; r++)
					{
						s += Ks(m, r) * ic[r];
					}
					q += s * Ks(m, c);
				}
				variance[m] = prior - q;
			},
			4
		);
		const Vec cf = cutoff_factor(prediction, variance); // kernel.cpp:519
		cutoff_prediction.resize(M);
		for (std::size_t m = 0; m < M; m++)
		{
			cutoff_prediction[m] = prediction[m] * cf[m] / rescale;
		}
		if (test_label != nullptr)
		{
			Vec lbl(M);
			double e = 0;
			for (std::size_t m = 0; m < M; m++)
			{
				lbl[m] = test_label[m] * rescale;
				e += sq(prediction[m] - lbl[m]); // raw prediction: quirk q1
			}
			error = e;
			if (is_deriv)
			{
				// kernel.cpp:524-541 (cutoff prediction in the residual: quirk q1)
				Vec diff(M);
				for (std::size_t m = 0; m < M; m++)
				{
					diff[m] = cutoff_prediction[m] * rescale - lbl[m];
				}
				std::array<double, NumRealParams> de{};
				for (std::size_t p = 0; p < NumRealParams; p++)
				{
					const Vec a = matvec((*base->dK)[p], k.v), b = matvec(Ks, (*k.dv)[p]);
					double s = 0;
					for (std::size_t m = 0; m < M; m++)
					{
						s += diff[m] * (a[m] + b[m]);
					}
					de[p] = 2.0 * s;
				}
				derror = de;
			}
		}
	}
};

} // namespace orc
