// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle for ctypes (tests/, smoke(), and the
// cpu_baseline / --impl reference legs of bench.py).  The product library never links this file.
#include "gple_oracle_dynamics.hpp"
#include "gple_oracle_mc.hpp"
#include "gple_oracle_nlml.hpp"

#include <chrono>
#include <cstring>

using namespace orc;

namespace
{
constexpr double NaN = std::numeric_limits<double>::quiet_NaN();

template <std::size_t N>
void put(double* out, const std::optional<std::array<double, N>>& a)
{
	for (std::size_t i = 0; i < N; i++)
	{
		out[i] = a.has_value() ? (*a)[i] : NaN;
	}
}

void copy_mat(const Mat& m, double* out)
{
	std::memcpy(out, m.d.data(), m.d.size() * sizeof(double));
}
void copy_cmat(const CMat& m, double* out)
{
	std::memcpy(out, m.d.data(), m.d.size() * sizeof(cplx));
}

struct Predictors
{
	const TrainingKernel* diag[2] = {nullptr, nullptr};
	const TrainingComplexKernel* off = nullptr;
};

/// gple/main.cpp:75-101 predict_distribution
Distribution make_distribution(const Predictors& pr)
{
	return [pr](const double x, const double p, const std::size_t row, const std::size_t col) -> cplx
	{
		const double r[2] = {x, p};
		if (row == col)
		{
			if (pr.diag[row] != nullptr)
			{
				return PredictiveKernel(r, 1, *pr.diag[row], false, nullptr).cutoff_prediction[0];
			}
			return 0.0;
		}
		if (pr.off != nullptr)
		{
			return PredictiveComplexKernel(r, 1, *pr.off, false, nullptr).cutoff_prediction[0];
		}
		return 0.0;
	};
}
} // namespace

extern "C"
{
	int orc_num_threads(void)
	{
		return static_cast<int>(num_threads());
	}

	// ---- real kernel -------------------------------------------------------------------------
	void orc_kernel_real(const double* XL, std::size_t nL, const double* XR, std::size_t nR, const double* th, int same, int deriv, double* K, double* dK)
	{
		KParam p;
		p.mag = th[0];
		p.l = {th[1], th[2]};
		p.noise = th[3];
		const KernelBase kb(p, Points{XL, nL}, Points{XR, nR}, same != 0, deriv != 0);
		copy_mat(kb.K, K);
		if (deriv != 0 && dK != nullptr)
		{
			for (std::size_t i = 0; i < NumRealParams; i++)
			{
				copy_mat((*kb.dK)[i], dK + i * nL * nR);
			}
		}
	}

	void* orc_train_real(const double* th, const double* X, const double* y_c, std::size_t N, int err, int avg, int deriv)
	{
		return new TrainingKernel(th, X, reinterpret_cast<const cplx*>(y_c), N, err != 0, avg != 0, deriv != 0);
	}
	void orc_free_real(void* h)
	{
		delete static_cast<TrainingKernel*>(h);
	}
	/// out[19] = rescale, error, population, <x>, <p>, purity, magnitude, derr[4], dpop[4], dpur[4]
	void orc_train_real_scalars(const void* h, double* out)
	{
		const auto* k = static_cast<const TrainingKernel*>(h);
		out[0] = k->rescale;
		out[1] = k->error.value_or(NaN);
		out[2] = k->population.value_or(NaN);
		out[3] = k->first_order.has_value() ? (*k->first_order)[0] : NaN;
		out[4] = k->first_order.has_value() ? (*k->first_order)[1] : NaN;
		out[5] = k->purity.value_or(NaN);
		out[6] = k->get_magnitude();
		put(out + 7, k->derror);
		put(out + 11, k->dpopulation);
		put(out + 15, k->dpurity);
	}
	/// which: 0 K, 1 K^-1 (N*N), 2 v (N), 3 label (N), 4+p dv[p] (N), 8+p dK[p], 12+p dKinv[p]
	int orc_train_real_get(const void* h, int which, double* out)
	{
		const auto* k = static_cast<const TrainingKernel*>(h);
		const std::size_t N = k->X.n;
		if (which == 0)
		{
			copy_mat(k->base->K, out);
		}
		else if (which == 1)
		{
			copy_mat(k->inverse, out);
		}
		else if (which == 2)
		{
			std::memcpy(out, k->v.data(), N * sizeof(double));
		}
		else if (which == 3)
		{
			std::memcpy(out, k->label.data(), N * sizeof(double));
		}
		else if (which >= 4 && which < 8 && k->dv.has_value())
		{
			std::memcpy(out, (*k->dv)[which - 4].data(), N * sizeof(double));
		}
		else if (which >= 8 && which < 12 && k->base->dK.has_value())
		{
			copy_mat((*k->base->dK)[which - 8], out);
		}
		else if (which >= 12 && which < 16 && k->dinv.has_value())
		{
			copy_mat((*k->dinv)[which - 12], out);
		}
		else
		{
			return 1;
		}
		return 0;
	}
	/// err[1], derr[4] written only when yq != NULL (derr only if deriv)
	void orc_predict_real(const void* h, const double* Xq, std::size_t Q, const double* yq, int deriv, double* pred, double* var, double* cutoff, double* err, double* derr)
	{
		const auto* k = static_cast<const TrainingKernel*>(h);
		const PredictiveKernel pk(Xq, Q, *k, deriv != 0, yq);
		if (pred != nullptr)
		{
			std::memcpy(pred, pk.prediction.data(), Q * sizeof(double));
		}
		if (var != nullptr)
		{
			std::memcpy(var, pk.variance.data(), Q * sizeof(double));
		}
		if (cutoff != nullptr)
		{
			std::memcpy(cutoff, pk.cutoff_prediction.data(), Q * sizeof(double));
		}
		if (err != nullptr)
		{
			*err = pk.error.value_or(NaN);
		}
		if (derr != nullptr)
		{
			put(derr, pk.derror);
		}
	}

	// ---- complex kernel ----------------------------------------------------------------------
	/// K: nL*nR doubles; Kt: nL*nR complex (interleaved); dK: 8 real matrices; dKt: 8 complex matrices
	void orc_kernel_complex(const double* XL, std::size_t nL, const double* XR, std::size_t nR, const double* th, int same, int deriv, double* K, double* Kt, double* dK, double* dKt)
	{
		const ComplexKernelBase kb(unpack_complex(th), Points{XL, nL}, Points{XR, nR}, same != 0, deriv != 0);
		copy_mat(kb.K, K);
		copy_cmat(kb.Kt, Kt);
		if (deriv != 0)
		{
			for (std::size_t i = 0; i < NumComplexParams; i++)
			{
				if (dK != nullptr)
				{
					copy_mat((*kb.dK)[i], dK + i * nL * nR);
				}
				if (dKt != nullptr)
				{
					copy_cmat((*kb.dKt)[i], dKt + 2 * i * nL * nR);
				}
			}
		}
	}
	void* orc_train_complex(const double* th, const double* X, const double* y_c, std::size_t N, int err, int avg, int deriv)
	{
		return new TrainingComplexKernel(th, X, reinterpret_cast<const cplx*>(y_c), N, err != 0, avg != 0, deriv != 0);
	}
	void orc_free_complex(void* h)
	{
		delete static_cast<TrainingComplexKernel*>(h);
	}
	/// out[20] = rescale, error, purity, magnitude, derr[8], dpur[8]
	void orc_train_complex_scalars(const void* h, double* out)
	{
		const auto* k = static_cast<const TrainingComplexKernel*>(h);
		out[0] = k->rescale;
		out[1] = k->error.value_or(NaN);
		out[2] = k->purity.value_or(NaN);
		out[3] = k->get_magnitude();
		put(out + 4, k->derror);
		put(out + 12, k->dpurity);
	}
	/// which: 0 K (real N*N), 1 Kt, 2 P, 3 Q (complex N*N), 4 v, 5 label (complex N), 8+p dv[p] (complex N)
	int orc_train_complex_get(const void* h, int which, double* out)
	{
		const auto* k = static_cast<const TrainingComplexKernel*>(h);
		const std::size_t N = k->X.n;
		switch (which)
		{
		case 0:
			copy_mat(k->base->K, out);
			return 0;
		case 1:
			copy_cmat(k->base->Kt, out);
			return 0;
		case 2:
			copy_cmat(k->P, out);
			return 0;
		case 3:
			copy_cmat(k->Q, out);
			return 0;
		case 4:
			std::memcpy(out, k->v.data(), N * sizeof(cplx));
			return 0;
		case 5:
			std::memcpy(out, k->label.data(), N * sizeof(cplx));
			return 0;
		default:
			if (which >= 8 && which < 16 && k->dv.has_value())
			{
				std::memcpy(out, (*k->dv)[which - 8].data(), N * sizeof(cplx));
				return 0;
			}
			return 1;
		}
	}
	/// pred, cutoff: complex Q (interleaved); var: Q doubles; yq complex or NULL
	void orc_predict_complex(const void* h, const double* Xq, std::size_t Q, const double* yq_c, int deriv, double* pred, double* var, double* cutoff, double* err, double* derr)
	{
		const auto* k = static_cast<const TrainingComplexKernel*>(h);
		const PredictiveComplexKernel pk(Xq, Q, *k, deriv != 0, reinterpret_cast<const cplx*>(yq_c));
		if (pred != nullptr)
		{
			std::memcpy(pred, pk.prediction.data(), Q * sizeof(cplx));
		}
		if (var != nullptr)
		{
			std::memcpy(var, pk.variance.data(), Q * sizeof(double));
		}
		if (cutoff != nullptr)
		{
			std::memcpy(cutoff, pk.cutoff_prediction.data(), Q * sizeof(cplx));
		}
		if (err != nullptr)
		{
			*err = pk.error.value_or(NaN);
		}
		if (derr != nullptr)
		{
			put(derr, pk.derror);
		}
	}

	// ---- loose function (gple/opt.cpp:441-482) ------------------------------------------------
	/// nparam 4 (real) or 8 (complex); grad may be NULL ("grad.empty()"); returns the loss after make_normal
	double orc_loose_function(const double* x, int nparam, double* grad, const double* X, const double* y_c, std::size_t N, const double* Xe, const double* ye_c, std::size_t M)
	{
		auto make_normal = [](double& d)
		{
			if (std::isnan(d) || std::isinf(d))
			{
				d = std::numeric_limits<double>::max();
			}
		};
		double result = 0.0;
		const bool g = grad != nullptr;
		if (nparam == 4)
		{
			const TrainingKernel k(x, X, reinterpret_cast<const cplx*>(y_c), N, true, false, g);
			std::vector<double> ye(M);
			for (std::size_t i = 0; i < M; i++)
			{
				ye[i] = ye_c[2 * i];
			}
			const PredictiveKernel pk(Xe, M, k, g, ye.data());
			result = *k.error + *pk.error;
			if (g)
			{
				for (std::size_t i = 0; i < 4; i++)
				{
					grad[i] = (*k.derror)[i] + (*pk.derror)[i];
				}
			}
		}
		else
		{
			const TrainingComplexKernel k(x, X, reinterpret_cast<const cplx*>(y_c), N, true, false, g);
			const PredictiveComplexKernel pk(Xe, M, k, g, reinterpret_cast<const cplx*>(ye_c));
			result = *k.error + *pk.error;
			if (g)
			{
				for (std::size_t i = 0; i < 8; i++)
				{
					grad[i] = (*k.derror)[i] + (*pk.derror)[i];
				}
			}
		}
		make_normal(result);
		if (g)
		{
			for (int i = 0; i < nparam; i++)
			{
				make_normal(grad[i]);
			}
		}
		return result;
	}

	// ---- PES ---------------------------------------------------------------------------------
	/// E: 2n (E0,E1 per point); F: 3n (F00, F10, F11 adiabatic); D: n (d_10)
	void orc_pes(int model, const double* x, std::size_t n, double* E, double* F, double* D)
	{
		const Model m = static_cast<Model>(model);
		for (std::size_t i = 0; i < n; i++)
		{
			const auto e = adiabatic_potential(m, x[i]);
			const Sym2 f = adiabatic_force(m, x[i]);
			E[2 * i] = e[0];
			E[2 * i + 1] = e[1];
			F[3 * i] = f.a00;
			F[3 * i + 1] = f.a01;
			F[3 * i + 2] = f.a11;
			D[i] = adiabatic_coupling_10(m, x[i]);
		}
	}

	// ---- evolve --------------------------------------------------------------------------------
	/// Evolves the three elements (order rho00, rho10, rho11 = lower-triangular index) in place.
	/// pts[e]: 4 doubles per point (x, p, Re rho, Im rho).  Predictors: handles from orc_train_real for
	/// rho00 / rho11 and orc_train_complex for rho10 (NULL = element absent, predicts 0: main.cpp:85-99).
	/// If analytic != NULL it replaces the predictors by gple/mc.cpp:30-50:
	/// analytic[8] = x0, p0, sigma_x, sigma_p, pop0, pop1, phase0, phase1.
	void orc_evolve(int model, double* pts00, std::size_t n00, double* pts10, std::size_t n10, double* pts11, std::size_t n11, double mass, double dt, const void* h00, const void* h10, const void* h11, const double* analytic)
	{
		const Model m = static_cast<Model>(model);
		Predictors pr;
		pr.diag[0] = static_cast<const TrainingKernel*>(h00);
		pr.diag[1] = static_cast<const TrainingKernel*>(h11);
		pr.off = static_cast<const TrainingComplexKernel*>(h10);
		Distribution dist = make_distribution(pr);
		if (analytic != nullptr)
		{
			const std::array<double, 2> r0{analytic[0], analytic[1]}, s0{analytic[2], analytic[3]}, pop{analytic[4], analytic[5]}, ph{analytic[6], analytic[7]};
			dist = [=](const double x, const double p, const std::size_t row, const std::size_t col) -> cplx
			{
				return initial_distribution(r0, s0, x, p, row, col, pop, ph);
			};
		}
		// same visiting order as evolve.cpp:386-389: (0,0), (1,0), (1,1)
		evolve_element(m, reinterpret_cast<PhaseSpacePoint*>(pts00), n00, mass, dt, dist, 0, 0);
		evolve_element(m, reinterpret_cast<PhaseSpacePoint*>(pts10), n10, mass, dt, dist, 1, 0);
		evolve_element(m, reinterpret_cast<PhaseSpacePoint*>(pts11), n11, mass, dt, dist, 1, 1);
	}

	/// The 9 query points of one evolved point: out[18] = for e in (00,10,11), b in (-1,0,1): (x4, p3).
	/// (x, p) is the point AFTER the forward move (i.e. the `r` passed to non_adiabatic_evolve_predict).
	void orc_backward_queries(int model, double x, double p, double mass, double dt, int row, int col, double* out)
	{
		const BackwardGeometry g = backward_geometry(static_cast<Model>(model), x, p, mass, dt, row, col);
		for (std::size_t e = 0; e < 3; e++)
		{
			for (std::size_t b = 0; b < 3; b++)
			{
				out[(e * 3 + b) * 2] = g.q[e][b][0];
				out[(e * 3 + b) * 2 + 1] = g.q[e][b][1];
			}
		}
	}

	/// gple/evolve.cpp:425-443 for n points: out = complex n
	void orc_new_point_predict(int model, const double* r, std::size_t n, double mass, double dt, int row, int col, const void* h00, const void* h10, const void* h11, double* out_c)
	{
		Predictors pr;
		pr.diag[0] = static_cast<const TrainingKernel*>(h00);
		pr.diag[1] = static_cast<const TrainingKernel*>(h11);
		pr.off = static_cast<const TrainingComplexKernel*>(h10);
		const Distribution dist = make_distribution(pr);
		parallel_for(
			n,
			[&](const std::size_t k)
			{
				const cplx v = new_point_predict(static_cast<Model>(model), r[2 * k], r[2 * k + 1], mass, dt, dist, row, col);
				out_c[2 * k] = v.real();
				out_c[2 * k + 1] = v.imag();
			}
		);
	}

	/// out[9]: see orc::observable_sums
	void orc_observable_sums(int model, const double* pts, std::size_t n, double mass, int pes_index, double* out)
	{
		const auto s = observable_sums(static_cast<Model>(model), reinterpret_cast<const PhaseSpacePoint*>(pts), n, mass, static_cast<std::size_t>(pes_index));
		for (std::size_t i = 0; i < 9; i++)
		{
			out[i] = s[i];
		}
	}

	/// gple/mc.cpp:30-50
	void orc_initial_distribution(const double* analytic, const double* r, std::size_t n, int row, int col, double* out_c)
	{
		const std::array<double, 2> r0{analytic[0], analytic[1]}, s0{analytic[2], analytic[3]}, pop{analytic[4], analytic[5]}, ph{analytic[6], analytic[7]};
		for (std::size_t k = 0; k < n; k++)
		{
			const cplx v = initial_distribution(r0, s0, r[2 * k], r[2 * k + 1], row, col, pop, ph);
			out_c[2 * k] = v.real();
			out_c[2 * k + 1] = v.imag();
		}
	}
	// ---- NLML / LLT objective (test/gpr.cpp:470-532) on the element models ------------------------------
	/// grad: 4 doubles or NULL (the model must have been trained with deriv = 1 when grad is requested)
	double orc_nlml_real(const void* h, double* grad)
	{
		return nlml_real(*static_cast<const TrainingKernel*>(h), grad);
	}
	/// grad: 8 doubles or NULL
	double orc_nlml_complex(const void* h, double* grad)
	{
		return nlml_complex(*static_cast<const TrainingComplexKernel*>(h), grad);
	}
	// ---- Metropolis sampling (gple/mc.cpp:143-243) -------------------------------------------------------
	/// kind 0: analytic initial distribution (analytic[8] as in orc_evolve); 1: predict_distribution of the models
	/// (main.cpp:75-101); 2: new_point_predict over the models (evolve.cpp:425-443, needs model / mass / dt).
	/// pts: n x 4 (x, p, re, im) in/out; accept: n doubles or NULL; chain_out: n x (num_steps + 1) x 2 or NULL.
	/// Point k walks chain chain0 + k (a block of a sharded point set keeps the streams it would have in the whole set).
	void orc_markov_chains(int kind, const double* analytic, const void* h00, const void* h10, const void* h11, int model, double mass, double dt, int row, int col, double* pts, std::size_t n, std::size_t num_steps, double max_displacement, std::uint64_t seed, std::uint64_t stream, std::uint64_t chain0, double* accept, double* chain_out)
	{
		Predictors pr;
		pr.diag[0] = static_cast<const TrainingKernel*>(h00);
		pr.diag[1] = static_cast<const TrainingKernel*>(h11);
		pr.off = static_cast<const TrainingComplexKernel*>(h10);
		const Distribution predict = make_distribution(pr);
		Distribution dist = predict;
		if (kind == 0)
		{
			const std::array<double, 2> r0{analytic[0], analytic[1]}, s0{analytic[2], analytic[3]}, pop{analytic[4], analytic[5]}, ph{analytic[6], analytic[7]};
			dist = [=](const double x, const double p, const std::size_t r, const std::size_t c) -> cplx
			{
				return initial_distribution(r0, s0, x, p, r, c, pop, ph);
			};
		}
		else if (kind == 2)
		{
			dist = [=](const double x, const double p, const std::size_t r, const std::size_t c) -> cplx
			{
				return new_point_predict(static_cast<Model>(model), x, p, mass, dt, predict, r, c);
			};
		}
		parallel_for(
			n,
			[&](const std::size_t k)
			{
				double x = pts[4 * k], p = pts[4 * k + 1];
				cplx rho;
				const double a = markov_chain(dist, std::size_t(row), std::size_t(col), x, p, rho, num_steps, max_displacement, seed, stream, chain0 + k, chain_out != nullptr ? chain_out + k * 2 * (num_steps + 1) : nullptr);
				pts[4 * k] = x;
				pts[4 * k + 1] = p;
				pts[4 * k + 2] = rho.real();
				pts[4 * k + 3] = rho.imag();
				if (accept != nullptr)
				{
					accept[k] = a;
				}
			}
		);
	}
	/// mean over the n chains of chain_autocorrelation: out[len / 2], chains laid out as orc_markov_chains writes them
	void orc_chain_autocorrelation(const double* chains, std::size_t n, std::size_t len, double* out)
	{
		std::vector<double> one(len / 2);
		std::fill(out, out + len / 2, 0.0);
		for (std::size_t k = 0; k < n; k++)
		{
			chain_autocorrelation(chains + k * 2 * len, len, one.data());
			for (std::size_t j = 0; j < len / 2; j++)
			{
				out[j] += one[j] / double(n);
			}
		}
	}
	/// the three uniforms of (seed, stream, chain, step): lets the tests pin the RNG itself
	void orc_philox_draws(std::uint64_t seed, std::uint64_t stream, std::uint64_t chain, std::uint32_t step, double* out3)
	{
		const auto u = Philox::draws(seed, stream, chain, step);
		out3[0] = u[0];
		out3[1] = u[1];
		out3[2] = u[2];
	}
}
