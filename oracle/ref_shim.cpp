// TEST INFRASTRUCTURE ONLY -- oracle/ref_shim.cpp
//
// C entry points over the reference's OWN classes and functions, so that Python tests can run the
// unmodified reference translation units (compiled where they lie under /root/reference by
// oracle/Makefile.ref, against the stand-in headers of oracle/refstub/) next to the oracle
// restatement and the CUDA path.  One library per PES model (pes.h:41 fixes `TestModel` at compile
// time): oracle/_ref/libgple_ref_{sac,dac,ecr}.so.  The entry points mirror oracle/gple_oracle_c.cpp's
// (`orc_*` -> `ref_*`), which lets oracle/ref.py reuse oracle/oracle.py's wrappers.
// Nothing here is product code; nothing in the product path links it.
#include "stdafx.h"

#include "complex_kernel.h"
#include "evolve.h"
#include "kernel.h"
#include "mc.h"
#include "pes.h"
#include "predict.h"
#include "storage.h"

#include <cstring>
#include <limits>
#include <memory>

namespace
{
using cplx = std::complex<double>;
constexpr double NaN = std::numeric_limits<double>::quiet_NaN();

PhasePoints make_points(const double* X, std::size_t n)
{
	PhasePoints p(PhaseDim, n);
	std::memcpy(p.data(), X, n * PhaseDim * sizeof(double)); // column-major 2 x n == interleaved (x, p)
	return p;
}
ElementTrainingSet make_set(const double* X, const double* y_c, std::size_t n)
{
	Eigen::VectorXcd y(n);
	std::memcpy(static_cast<void*>(y.data()), y_c, n * sizeof(cplx));
	return std::make_tuple(make_points(X, n), y);
}
template <typename M>
void copy_out(const M& m, double* out)
{
	// column-major scan (the layout of Eigen::MatrixXd / MatrixXcd)
	using S = typename M::Scalar;
	S* o = reinterpret_cast<S*>(out);
	for (Eigen::Index j = 0; j < m.cols(); j++)
	{
		for (Eigen::Index i = 0; i < m.rows(); i++)
		{
			*o++ = m(i, j);
		}
	}
}
template <std::size_t N>
void put(double* out, const std::array<double, N>& a)
{
	for (std::size_t i = 0; i < N; i++)
	{
		out[i] = a[i];
	}
}
template <std::size_t N>
void put_nan(double* out)
{
	for (std::size_t i = 0; i < N; i++)
	{
		out[i] = NaN;
	}
}

struct RealModel
{
	std::unique_ptr<TrainingKernel> k;
	bool err, avg, deriv;
};
struct ComplexModel
{
	std::unique_ptr<TrainingComplexKernel> k;
	bool err, avg, deriv;
};

ElementPoints make_element_points(const double* pts, std::size_t n)
{
	ElementPoints e;
	e.reserve(n);
	for (std::size_t i = 0; i < n; i++)
	{
		ClassicalPhaseVector r;
		r << pts[4 * i], pts[4 * i + 1];
		e.emplace_back(r, cplx(pts[4 * i + 2], pts[4 * i + 3]));
	}
	return e;
}
void store_element_points(const ElementPoints& e, double* pts)
{
	for (std::size_t i = 0; i < e.size(); i++)
	{
		const auto& [r, rho] = e[i];
		pts[4 * i] = r[0];
		pts[4 * i + 1] = r[1];
		pts[4 * i + 2] = rho.real();
		pts[4 * i + 3] = rho.imag();
	}
}

/// The reference's `predict_distribution` (gple/main.cpp:75-101) over individually held element models
/// instead of the `all_kernels` aggregate: a single-point PredictiveKernel / PredictiveComplexKernel per call.
DistributionFunction make_distribution(const RealModel* m00, const ComplexModel* m10, const RealModel* m11)
{
	return [m00, m10, m11](const ClassicalPhaseVector& r, const std::size_t RowIndex, const std::size_t ColIndex) -> cplx
	{
		if (RowIndex == ColIndex)
		{
			const RealModel* m = RowIndex == 0 ? m00 : m11;
			if (m != nullptr)
			{
				return PredictiveKernel(r, *m->k, false).get_cutoff_prediction().value();
			}
			return 0.0;
		}
		if (m10 != nullptr)
		{
			return PredictiveComplexKernel(r, *m10->k, false).get_cutoff_prediction().value();
		}
		return 0.0;
	};
}
DistributionFunction make_analytic(const double* a)
{
	ClassicalPhaseVector r0, s0;
	r0 << a[0], a[1];
	s0 << a[2], a[3];
	const std::array<double, NumPES> pop{a[4], a[5]}, ph{a[6], a[7]};
	return [r0, s0, pop, ph](const ClassicalPhaseVector& r, const std::size_t RowIndex, const std::size_t ColIndex) -> cplx
	{
		return initial_distribution(r0, s0, r, RowIndex, ColIndex, pop, ph);
	};
}
} // namespace

extern "C"
{
	/// 0 SAC, 1 DAC, 2 ECR: the model this library was compiled for (pes.h:31-41)
	int ref_model(void)
	{
		return static_cast<int>(TestModel);
	}
	int ref_num_threads(void)
	{
		return static_cast<int>(orc::num_threads());
	}
	int ref_have_blas(void)
	{
		return refstub::have_blas() ? 1 : 0;
	}

	// ---- kernel.cpp ------------------------------------------------------------------------------
	void ref_kernel_real(const double* XL, std::size_t nL, const double* XR, std::size_t nR, const double* th, int same, int deriv, double* K, double* dK)
	{
		KernelBase::KernelParameter p;
		auto& [mag, len, noise] = p;
		mag = th[0];
		len << th[1], th[2];
		noise = th[3];
		const PhasePoints L = make_points(XL, nL), R = make_points(XR, nR);
		// "same" = the caller passes one object for both features, which is what the reference's pointer tests see
		const KernelBase kb(p, L, same != 0 ? L : R, deriv != 0);
		copy_out(kb.get_kernel(), K);
		if (deriv != 0 && dK != nullptr)
		{
			for (std::size_t i = 0; i < KernelBase::NumTotalParameters; i++)
			{
				copy_out(kb.get_derivative()[i], dK + i * nL * nR);
			}
		}
	}
	void* ref_train_real(const double* th, const double* X, const double* y_c, std::size_t N, int err, int avg, int deriv)
	{
		const ElementTrainingSet set = make_set(X, y_c, N);
		auto* m = new RealModel{nullptr, err != 0, avg != 0, deriv != 0};
		m->k = std::make_unique<TrainingKernel>(ParameterVector(th, th + KernelBase::NumTotalParameters), set, m->err, m->avg, m->deriv);
		return m;
	}
	void ref_free_real(void* h)
	{
		delete static_cast<RealModel*>(h);
	}
	/// out[19] = rescale, error, population, <x>, <p>, purity, magnitude, derr[4], dpop[4], dpur[4]
	void ref_train_real_scalars(const void* h, double* out)
	{
		const auto* m = static_cast<const RealModel*>(h);
		const TrainingKernel& k = *m->k;
		out[0] = k.get_rescale_factor();
		out[1] = m->err ? k.get_error() : NaN;
		out[2] = m->avg ? k.get_population() : NaN;
		out[3] = m->avg ? k.get_1st_order_average()[0] : NaN;
		out[4] = m->avg ? k.get_1st_order_average()[1] : NaN;
		out[5] = m->avg ? k.get_purity() : NaN;
		out[6] = k.get_magnitude();
		if (m->err && m->deriv)
		{
			put(out + 7, k.get_error_derivative());
		}
		else
		{
			put_nan<4>(out + 7);
		}
		if (m->avg && m->deriv)
		{
			put(out + 11, k.get_population_derivative());
			put(out + 15, k.get_purity_derivative());
		}
		else
		{
			put_nan<8>(out + 11);
		}
	}
	/// which: 0 K, 1 K^-1, 2 v, 4+p dv[p], 8+p dK[p]  (the label and dK^-1 have no getter in kernel.h)
	int ref_train_real_get(const void* h, int which, double* out)
	{
		const auto* m = static_cast<const RealModel*>(h);
		const TrainingKernel& k = *m->k;
		if (which == 0)
		{
			copy_out(k.get_kernel(), out);
		}
		else if (which == 1)
		{
			copy_out(k.get_inverse(), out);
		}
		else if (which == 2)
		{
			copy_out(k.get_inverse_times_label(), out);
		}
		else if (which >= 4 && which < 8 && m->deriv)
		{
			copy_out(k.get_inverse_times_label_derivative()[which - 4], out);
		}
		else if (which >= 8 && which < 12 && m->deriv)
		{
			copy_out(k.get_derivative()[which - 8], out);
		}
		else
		{
			return 1;
		}
		return 0;
	}
	void ref_predict_real(const void* h, const double* Xq, std::size_t Q, const double* yq, int deriv, double* pred, double* var, double* cutoff, double* err, double* derr)
	{
		const auto* m = static_cast<const RealModel*>(h);
		std::optional<Eigen::VectorXd> label = std::nullopt;
		if (yq != nullptr)
		{
			Eigen::VectorXd y(Q);
			std::memcpy(y.data(), yq, Q * sizeof(double));
			label = y;
		}
		const PredictiveKernel pk(make_points(Xq, Q), *m->k, deriv != 0, label);
		if (pred != nullptr)
		{
			// `Prediction` is private (kernel.h:391); it is K* v with the public pieces (kernel.cpp:495)
			copy_out(Eigen::VectorXd(pk.get_kernel() * m->k->get_inverse_times_label()), pred);
		}
		if (var != nullptr)
		{
			copy_out(pk.get_variance(), var);
		}
		if (cutoff != nullptr)
		{
			copy_out(pk.get_cutoff_prediction(), cutoff);
		}
		if (err != nullptr)
		{
			*err = yq != nullptr ? pk.get_error() : NaN;
		}
		if (derr != nullptr)
		{
			if (yq != nullptr && deriv != 0)
			{
				put(derr, pk.get_error_derivative());
			}
			else
			{
				put_nan<4>(derr);
			}
		}
	}

	// ---- complex_kernel.cpp ----------------------------------------------------------------------
	void ref_kernel_complex(const double* XL, std::size_t nL, const double* XR, std::size_t nR, const double* th, int same, int deriv, double* K, double* Kt, double* dK, double* dKt)
	{
		ComplexKernelBase::KernelParameter p;
		auto& [mag, sub, noise] = p;
		mag = th[0];
		std::get<0>(sub[0]) = th[1];
		std::get<1>(sub[0]) << th[2], th[3];
		std::get<0>(sub[1]) = th[4];
		std::get<1>(sub[1]) << th[5], th[6];
		noise = th[7];
		const PhasePoints L = make_points(XL, nL), R = make_points(XR, nR);
		const ComplexKernelBase kb(p, L, same != 0 ? L : R, deriv != 0);
		copy_out(kb.get_kernel(), K);
		copy_out(kb.get_pseudo_kernel(), Kt);
		if (deriv != 0)
		{
			for (std::size_t i = 0; i < ComplexKernelBase::NumTotalParameters; i++)
			{
				if (dK != nullptr)
				{
					copy_out(kb.get_derivative()[i], dK + i * nL * nR);
				}
				if (dKt != nullptr)
				{
					copy_out(kb.get_pseudo_derivative()[i], dKt + 2 * i * nL * nR);
				}
			}
		}
	}
	void* ref_train_complex(const double* th, const double* X, const double* y_c, std::size_t N, int err, int avg, int deriv)
	{
		const ElementTrainingSet set = make_set(X, y_c, N);
		auto* m = new ComplexModel{nullptr, err != 0, avg != 0, deriv != 0};
		m->k = std::make_unique<TrainingComplexKernel>(ParameterVector(th, th + ComplexKernelBase::NumTotalParameters), set, m->err, m->avg, m->deriv);
		return m;
	}
	void ref_free_complex(void* h)
	{
		delete static_cast<ComplexModel*>(h);
	}
	/// out[20] = rescale, error, purity, magnitude, derr[8], dpur[8]
	void ref_train_complex_scalars(const void* h, double* out)
	{
		const auto* m = static_cast<const ComplexModel*>(h);
		const TrainingComplexKernel& k = *m->k;
		out[0] = k.get_rescale_factor();
		out[1] = m->err ? k.get_error() : NaN;
		out[2] = m->avg ? k.get_purity() : NaN;
		out[3] = k.get_magnitude();
		if (m->err && m->deriv)
		{
			put(out + 4, k.get_error_derivative());
		}
		else
		{
			put_nan<8>(out + 4);
		}
		if (m->avg && m->deriv)
		{
			put(out + 12, k.get_purity_derivative());
		}
		else
		{
			put_nan<8>(out + 12);
		}
	}
	/// which: 0 K (real), 1 Kt, 2 P, 3 Q (complex N*N), 4 v (complex N), 8+p dv[p]
	int ref_train_complex_get(const void* h, int which, double* out)
	{
		const auto* m = static_cast<const ComplexModel*>(h);
		const TrainingComplexKernel& k = *m->k;
		switch (which)
		{
		case 0:
			copy_out(k.get_kernel(), out);
			return 0;
		case 1:
			copy_out(k.get_pseudo_kernel(), out);
			return 0;
		case 2:
			copy_out(k.get_upper_left_block_of_augmented_inverse(), out);
			return 0;
		case 3:
			copy_out(k.get_lower_left_block_of_augmented_inverse(), out);
			return 0;
		case 4:
			copy_out(k.get_upper_part_of_augmented_inverse_times_label(), out);
			return 0;
		default:
			if (which >= 8 && which < 16 && m->deriv)
			{
				copy_out(k.get_upper_part_of_augmented_inverse_times_label_derivative()[which - 8], out);
				return 0;
			}
			return 1;
		}
	}
	void ref_predict_complex(const void* h, const double* Xq, std::size_t Q, const double* yq_c, int deriv, double* pred, double* var, double* cutoff, double* err, double* derr)
	{
		const auto* m = static_cast<const ComplexModel*>(h);
		std::optional<Eigen::VectorXcd> label = std::nullopt;
		if (yq_c != nullptr)
		{
			Eigen::VectorXcd y(Q);
			std::memcpy(static_cast<void*>(y.data()), yq_c, Q * sizeof(cplx));
			label = y;
		}
		const PredictiveComplexKernel pk(make_points(Xq, Q), *m->k, deriv != 0, label);
		if (pred != nullptr)
		{
			// `Prediction` is private (complex_kernel.h:380); K* v + K~* conj(v) with the public pieces (complex_kernel.cpp:608)
			const Eigen::VectorXcd& v = m->k->get_upper_part_of_augmented_inverse_times_label();
			copy_out(Eigen::VectorXcd(pk.get_kernel() * v + pk.get_pseudo_kernel() * v.conjugate()), pred);
		}
		if (var != nullptr)
		{
			copy_out(pk.get_variance(), var);
		}
		if (cutoff != nullptr)
		{
			copy_out(pk.get_cutoff_prediction(), cutoff);
		}
		if (err != nullptr)
		{
			*err = yq_c != nullptr ? pk.get_error() : NaN;
		}
		if (derr != nullptr)
		{
			if (yq_c != nullptr && deriv != 0)
			{
				put(derr, pk.get_error_derivative());
			}
			else
			{
				put_nan<8>(derr);
			}
		}
	}

	// ---- pes.cpp ---------------------------------------------------------------------------------
	/// E: 2n; F: 3n (F00, F10, F11 adiabatic); D: n (d_10).  `model` must be the one this library was built for.
	int ref_pes(int model, const double* x, std::size_t n, double* E, double* F, double* D)
	{
		if (model != static_cast<int>(TestModel))
		{
			return 1;
		}
		for (std::size_t i = 0; i < n; i++)
		{
			ClassicalVector<double> xi;
			xi << x[i];
			const QuantumVector<double> e = adiabatic_potential(xi);
			const Tensor3d f = adiabatic_force(xi);
			const Tensor3d d = adiabatic_coupling(xi);
			E[2 * i] = e[0];
			E[2 * i + 1] = e[1];
			F[3 * i] = f(0, 0, 0);
			F[3 * i + 1] = f(0, 1, 0);
			F[3 * i + 2] = f(0, 1, 1);
			D[i] = d(0, 1, 0);
		}
		return 0;
	}

	// ---- evolve.cpp ------------------------------------------------------------------------------
	int ref_evolve(int model, double* pts00, std::size_t n00, double* pts10, std::size_t n10, double* pts11, std::size_t n11, double mass, double dt, const void* h00, const void* h10, const void* h11, const double* analytic)
	{
		if (model != static_cast<int>(TestModel))
		{
			return 1;
		}
		AllPoints density;
		density(0, 0) = make_element_points(pts00, n00);
		density(1, 0) = make_element_points(pts10, n10);
		density(1, 1) = make_element_points(pts11, n11);
		const DistributionFunction dist = analytic != nullptr
			? make_analytic(analytic)
			: make_distribution(static_cast<const RealModel*>(h00), static_cast<const ComplexModel*>(h10), static_cast<const RealModel*>(h11));
		ClassicalVector<double> m;
		m << mass;
		evolve(density, m, dt, dist);
		store_element_points(density(0, 0), pts00);
		store_element_points(density(1, 0), pts10);
		store_element_points(density(1, 1), pts11);
		return 0;
	}
	int ref_new_point_predict(int model, const double* r, std::size_t n, double mass, double dt, int row, int col, const void* h00, const void* h10, const void* h11, double* out_c)
	{
		if (model != static_cast<int>(TestModel))
		{
			return 1;
		}
		const DistributionFunction dist = make_distribution(static_cast<const RealModel*>(h00), static_cast<const ComplexModel*>(h10), static_cast<const RealModel*>(h11));
		ClassicalVector<double> m;
		m << mass;
		orc::parallel_for(
			n,
			[&](const std::size_t k)
			{
				ClassicalPhaseVector rk;
				rk << r[2 * k], r[2 * k + 1];
				const cplx v = new_point_predict(rk, m, dt, dist, static_cast<std::size_t>(row), static_cast<std::size_t>(col));
				out_c[2 * k] = v.real();
				out_c[2 * k + 1] = v.imag();
			}
		);
		return 0;
	}
	/// is_very_small (evolve.cpp:444-478): out[3] = flags of (0,0), (1,0), (1,1)
	int ref_is_very_small(int model, const double* pts00, std::size_t n00, const double* pts10, std::size_t n10, const double* pts11, std::size_t n11, double mass, double dt, const void* h00, const void* h10, const void* h11, int* out)
	{
		if (model != static_cast<int>(TestModel))
		{
			return 1;
		}
		AllPoints density;
		density(0, 0) = make_element_points(pts00, n00);
		density(1, 0) = make_element_points(pts10, n10);
		density(1, 1) = make_element_points(pts11, n11);
		const DistributionFunction dist = make_distribution(static_cast<const RealModel*>(h00), static_cast<const ComplexModel*>(h10), static_cast<const RealModel*>(h11));
		ClassicalVector<double> m;
		m << mass;
		const QuantumStorage<bool> s = is_very_small(density, m, dt, dist);
		out[0] = s(0, 0) ? 1 : 0;
		out[1] = s(1, 0) ? 1 : 0;
		out[2] = s(1, 1) ? 1 : 0;
		return 0;
	}

	// ---- predict.cpp -----------------------------------------------------------------------------
	/// MC-integral observables of one diagonal element (predict.cpp:65-244):
	/// out[6] = <x>, <p>, std x, std p, <E> on surface pes_index, sum |rho|^2
	int ref_observables(int model, const double* pts, std::size_t n, double mass, int pes_index, double* out)
	{
		if (model != static_cast<int>(TestModel))
		{
			return 1;
		}
		const ElementPoints e = make_element_points(pts, n);
		ClassicalVector<double> m;
		m << mass;
		const ClassicalPhaseVector avg = calculate_1st_order_average_one_surface(e), sd = calculate_standard_deviation_one_surface(e);
		out[0] = avg[0];
		out[1] = avg[1];
		out[2] = sd[0];
		out[3] = sd[1];
		out[4] = calculate_total_energy_average_one_surface(e, m, static_cast<std::size_t>(pes_index));
		AllPoints all;
		all(static_cast<std::size_t>(pes_index), static_cast<std::size_t>(pes_index)) = e;
		out[5] = calculate_purity_each_element(all)(pes_index, pes_index);
		return 0;
	}
	/// TrainingKernels aggregate (predict.cpp:362-463) over three point sets with err = avg = true:
	/// theta: 4 + 8 + 4 doubles in the order (0,0), (1,0), (1,1); out[5] = population, <x>, <p>, energy(E0, E1 given), purity
	void ref_training_kernels(const double* theta, const double* pts00, std::size_t n00, const double* pts10, std::size_t n10, const double* pts11, std::size_t n11, const double* energies, double* out)
	{
		AllPoints density;
		density(0, 0) = make_element_points(pts00, n00);
		density(1, 0) = make_element_points(pts10, n10);
		density(1, 1) = make_element_points(pts11, n11);
		QuantumStorage<ParameterVector> params;
		params(0, 0) = ParameterVector(theta, theta + 4);
		params(1, 0) = ParameterVector(theta + 4, theta + 12);
		params(1, 1) = ParameterVector(theta + 12, theta + 16);
		const TrainingKernels ks(params, density);
		QuantumVector<double> E;
		E << energies[0], energies[1];
		out[0] = ks.calculate_population();
		const ClassicalPhaseVector r = ks.calculate_1st_order_average();
		out[1] = r[0];
		out[2] = r[1];
		out[3] = ks.calculate_total_energy_average(E);
		out[4] = ks.calculate_purity();
	}

	// ---- mc.cpp ----------------------------------------------------------------------------------
	void ref_initial_distribution(const double* analytic, const double* r, std::size_t n, int row, int col, double* out_c)
	{
		const DistributionFunction dist = make_analytic(analytic);
		for (std::size_t k = 0; k < n; k++)
		{
			ClassicalPhaseVector rk;
			rk << r[2 * k], r[2 * k + 1];
			const cplx v = dist(rk, static_cast<std::size_t>(row), static_cast<std::size_t>(col));
			out_c[2 * k] = v.real();
			out_c[2 * k + 1] = v.imag();
		}
	}
}
