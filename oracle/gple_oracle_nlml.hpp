// TEST INFRASTRUCTURE ONLY (see gple_oracle.hpp).  Negative log marginal likelihood with LLT, the objective the
// reference's test programme minimises (test/gpr.cpp:470-532; formula doc :475-496), restated for the gple/ element
// models so that it can serve as an alternative loss on the same kernels:
//     NLML = y'^T K^-1 y' / 2 + sum_i ln L_ii        (the constant n/2 ln 2 pi is dropped, test/gpr.cpp:481-483)
//     d NLML / d theta = tr[(K^-1 - b b^T) dK/dtheta] / 2,   b = K^-1 y'      (test/gpr.cpp:487-493, :525)
// y' are the model's rescaled labels (kernel.cpp:279-280 / complex_kernel.cpp:262-263).  The reference has no NLML
// for the complex element; there the likelihood is that of the real composite process [Re f; Im f] whose covariance
// is [[K_rr, K_ri], [K_ri, K_ii]] with K = K_rr + K_ii, Kt = K_rr - K_ii + 2i K_ri (complex_kernel.h:12-13).
// Derivatives are the TRUE ones: for the complex element the reference's derivative arrays lack a factor sigma^2
// on every parameter but the global magnitude (quirk q2, complex_kernel.cpp:37-51), which is restored here.
#pragma once
#include "gple_oracle_complex.hpp"

namespace orc
{
/// Unpivoted Cholesky (Eigen::LLT of test/gpr.cpp:511); returns false if not positive definite.
inline bool llt_lower(Mat& A)
{
	const std::size_t n = A.rows;
	for (std::size_t j = 0; j < n; j++)
	{
		double d = A(j, j);
		for (std::size_t k = 0; k < j; k++)
		{
			d -= sq(A(j, k));
		}
		if (!(d > 0.0))
		{
			return false;
		}
		const double ljj = std::sqrt(d);
		A(j, j) = ljj;
		for (std::size_t i = j + 1; i < n; i++)
		{
			double s = A(i, j);
			for (std::size_t k = 0; k < j; k++)
			{
				s -= A(i, k) * A(j, k);
			}
			A(i, j) = s / ljj;
		}
	}
	return true;
}

/// Solve (L L^T) x = b in place.
inline void llt_solve(const Mat& L, Vec& b)
{
	const std::size_t n = L.rows;
	for (std::size_t i = 0; i < n; i++)
	{
		double s = b[i];
		for (std::size_t k = 0; k < i; k++)
		{
			s -= L(i, k) * b[k];
		}
		b[i] = s / L(i, i);
	}
	for (std::size_t ii = n; ii-- > 0;)
	{
		double s = b[ii];
		for (std::size_t k = ii + 1; k < n; k++)
		{
			s -= L(k, ii) * b[k];
		}
		b[ii] = s / L(ii, ii);
	}
}

/// value and (optionally) gradient for a symmetric covariance C, labels y and derivative matrices dC[p]
inline double nlml_core(const Mat& Cov, const Vec& y, const std::vector<Mat>& dC, double* grad)
{
	const std::size_t n = Cov.rows;
	Mat L = Cov;
	if (!llt_lower(L))
	{
		return std::numeric_limits<double>::quiet_NaN();
	}
	Vec b = y;
	llt_solve(L, b);
	double value = 0.5 * dot(y, b);
	for (std::size_t i = 0; i < n; i++)
	{
		value += std::log(std::abs(L(i, i)));
	}
	if (grad != nullptr)
	{
		// K^-1 column by column
		Mat Inv(n, n);
		for (std::size_t j = 0; j < n; j++)
		{
			Vec e(n, 0.0);
			e[j] = 1.0;
			llt_solve(L, e);
			for (std::size_t i = 0; i < n; i++)
			{
				Inv(i, j) = e[i];
			}
		}
		for (std::size_t p = 0; p < dC.size(); p++)
		{
			double t = 0.0;
			for (std::size_t j = 0; j < n; j++)
			{
				for (std::size_t i = 0; i < n; i++)
				{
					t += (Inv(i, j) - b[i] * b[j]) * dC[p](j, i);
				}
			}
			grad[p] = 0.5 * t;
		}
	}
	return value;
}

/// Real element: theta = (sigma_f, l_x, l_p, sigma_n); k must have been built with derivatives if grad != nullptr.
inline double nlml_real(const TrainingKernel& k, double* grad)
{
	std::vector<Mat> dC;
	if (grad != nullptr)
	{
		for (std::size_t p = 0; p < NumRealParams; p++)
		{
			dC.push_back((*k.base->dK)[p]);
		}
	}
	return nlml_core(k.base->K, k.label, dC, grad);
}

/// Complex element as the composite real process; 8 parameters in the order of complex_kernel.cpp:230-256.
inline double nlml_complex(const TrainingComplexKernel& k, double* grad)
{
	const std::size_t N = k.label.size();
	const Mat& K = k.base->K;
	const CMat& Kt = k.base->Kt;
	auto compose = [N](const Mat& D, const CMat& Dt, const double scale)
	{
		Mat Cc(2 * N, 2 * N);
		for (std::size_t j = 0; j < N; j++)
		{
			for (std::size_t i = 0; i < N; i++)
			{
				const double d = D(i, j), tr = Dt(i, j).real(), ti = Dt(i, j).imag();
				Cc(i, j) = scale * 0.5 * (d + tr);
				Cc(i + N, j + N) = scale * 0.5 * (d - tr);
				Cc(i, j + N) = scale * 0.5 * ti;
				Cc(i + N, j) = scale * 0.5 * ti;
			}
		}
		return Cc;
	};
	const Mat Cov = compose(K, Kt, 1.0);
	Vec y(2 * N);
	for (std::size_t i = 0; i < N; i++)
	{
		y[i] = k.label[i].real();
		y[i + N] = k.label[i].imag();
	}
	std::vector<Mat> dC;
	if (grad != nullptr)
	{
		const double s2 = sq(k.prm.mag);
		for (std::size_t p = 0; p < NumComplexParams; p++)
		{
			dC.push_back(compose((*k.base->dK)[p], (*k.base->dKt)[p], p == 0 ? 1.0 : s2));
		}
	}
	return nlml_core(Cov, y, dC, grad);
}
} // namespace orc
