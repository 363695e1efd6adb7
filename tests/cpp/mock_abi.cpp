// TEST INFRASTRUCTURE ONLY.  A CPU stand-in for libgple_b200.so: the C-ABI of include/gple_b200.h answered by the oracle
// (oracle/gple_oracle_c.cpp), so that the host-side C++ (host/gple_host.hpp, gple_opt.hpp, gple_mc.hpp, examples/mqcle_run.cpp)
// can be EXECUTED by the CPU test tier, not only compiled.  It is built into tests/cpp/_build/mock/ by the tests, never into
// the package, and nothing in the product links or loads it: the product library has no CPU path.
#include "../../include/gple_b200.h"

#include "../../oracle/gple_oracle_c.cpp"

#include <string>

struct gple_ctx
{
	std::string last_error;
	unsigned long long calls = 0;
};
struct gple_model
{
	int is_complex = 0;
	std::size_t N = 0;
	void* h = nullptr; // orc::TrainingKernel / orc::TrainingComplexKernel
	bool deriv = false;
};

namespace
{
const void* handle(const gple_model* m)
{
	return m != nullptr ? m->h : nullptr;
}
int fail(gple_ctx* ctx, const char* what)
{
	if (ctx != nullptr)
	{
		ctx->last_error = what;
	}
	return GPLE_ERR_ARG;
}
} // namespace

extern "C"
{
	int gple_ctx_create(int, gple_ctx** ctx)
	{
		*ctx = new gple_ctx();
		return GPLE_OK;
	}
	int gple_ctx_destroy(gple_ctx* ctx)
	{
		delete ctx;
		return GPLE_OK;
	}
	int gple_ctx_set_stream(gple_ctx*, void*) { return GPLE_OK; }
	int gple_ctx_sync(gple_ctx*) { return GPLE_OK; }
	int gple_ctx_set_option(gple_ctx*, int, int) { return GPLE_OK; }
	int gple_gate_statistics(gple_ctx*, unsigned long long out[4])
	{
		out[0] = out[1] = out[2] = out[3] = 0;
		return GPLE_OK;
	}
	const char* gple_last_error(const gple_ctx* ctx) { return ctx != nullptr ? ctx->last_error.c_str() : ""; }
	unsigned long long gple_launch_count(const gple_ctx* ctx) { return ctx != nullptr ? ctx->calls : 0; }
	const char* gple_version(void) { return "mock (CPU oracle) -- tests only"; }

	int gple_kernel_real(gple_ctx*, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[4], int same_set, double* K_out, double* dK_out)
	{
		orc_kernel_real(XL, nL, XR, nR, theta, same_set, dK_out != nullptr, K_out, dK_out);
		return GPLE_OK;
	}
	int gple_kernel_complex(gple_ctx*, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[8], int same_set, double* K_out, double* Kt_out)
	{
		orc_kernel_complex(XL, nL, XR, nR, theta, same_set, 0, K_out, Kt_out, nullptr, nullptr);
		return GPLE_OK;
	}
	int gple_kernel_complex_derivatives(gple_ctx*, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[8], int same_set, double* dK_out, double* dKt_out)
	{
		std::vector<double> K(nL * nR), Kt(2 * nL * nR);
		orc_kernel_complex(XL, nL, XR, nR, theta, same_set, 1, K.data(), Kt.data(), dK_out, dKt_out);
		return GPLE_OK;
	}
	int gple_train_real(gple_ctx* ctx, const double* X, const double* y, size_t N, const double theta[4], unsigned flags, gple_model** model, gple_real_scalars* out)
	{
		if (X == nullptr || y == nullptr || N == 0 || model == nullptr)
		{
			return fail(ctx, "gple_train_real: null argument");
		}
		auto* m = new gple_model();
		m->N = N;
		m->deriv = (flags & GPLE_CALC_DERIVATIVE) != 0;
		m->h = orc_train_real(theta, X, y, N, (flags & GPLE_CALC_ERROR) != 0, (flags & GPLE_CALC_AVERAGE) != 0, m->deriv);
		if (out != nullptr)
		{
			static_assert(sizeof(gple_real_scalars) == 19 * sizeof(double), "scalar block layout");
			orc_train_real_scalars(m->h, reinterpret_cast<double*>(out));
		}
		*model = m;
		return GPLE_OK;
	}
	int gple_train_complex(gple_ctx* ctx, const double* X, const double* y, size_t N, const double theta[8], unsigned flags, gple_model** model, gple_complex_scalars* out)
	{
		if (X == nullptr || y == nullptr || N == 0 || model == nullptr)
		{
			return fail(ctx, "gple_train_complex: null argument");
		}
		auto* m = new gple_model();
		m->is_complex = 1;
		m->N = N;
		m->deriv = (flags & GPLE_CALC_DERIVATIVE) != 0;
		m->h = orc_train_complex(theta, X, y, N, (flags & GPLE_CALC_ERROR) != 0, (flags & GPLE_CALC_AVERAGE) != 0, m->deriv);
		if (out != nullptr)
		{
			static_assert(sizeof(gple_complex_scalars) == 20 * sizeof(double), "scalar block layout");
			orc_train_complex_scalars(m->h, reinterpret_cast<double*>(out));
		}
		*model = m;
		return GPLE_OK;
	}
	int gple_model_get(gple_ctx* ctx, const gple_model*, int, double*) { return fail(ctx, "gple_model_get: not provided by the mock"); }
	int gple_model_is_complex(const gple_model* m) { return m != nullptr ? m->is_complex : -1; }
	size_t gple_model_size(const gple_model* m) { return m != nullptr ? m->N : 0; }
	int gple_model_destroy(gple_ctx*, gple_model* m)
	{
		if (m != nullptr)
		{
			m->is_complex ? orc_free_complex(m->h) : orc_free_real(m->h);
			delete m;
		}
		return GPLE_OK;
	}
	int gple_model_nlml(gple_ctx*, gple_model* m, double* value, double* grad)
	{
		*value = m->is_complex ? orc_nlml_complex(m->h, grad) : orc_nlml_real(m->h, grad);
		return GPLE_OK;
	}

	int gple_predict_real(gple_ctx*, const gple_model* m, const double* Xq, size_t Q, const double* yq, double* pred, double* var, double* cut, double* err, double* derr)
	{
		orc_predict_real(m->h, Xq, Q, yq, derr != nullptr, pred, var, cut, err, derr);
		return GPLE_OK;
	}
	int gple_predict_complex(gple_ctx*, const gple_model* m, const double* Xq, size_t Q, const double* yq, double* pred, double* var, double* cut, double* err, double* derr)
	{
		orc_predict_complex(m->h, Xq, Q, yq, derr != nullptr, pred, var, cut, err, derr);
		return GPLE_OK;
	}
	int gple_validation_error(gple_ctx* ctx, const gple_model* m, const double* Xe, const double* ye, size_t M, double* error, double* grad)
	{
		if (grad != nullptr && !m->deriv)
		{
			return fail(ctx, "gple_validation_error: the gradient needs a model trained with GPLE_CALC_DERIVATIVE");
		}
		if (m->is_complex)
		{
			orc_predict_complex(m->h, Xe, M, ye, grad != nullptr, nullptr, nullptr, nullptr, error, grad);
		}
		else
		{
			std::vector<double> re(M); // opt.cpp:451: the real kernel sees ExtraTrainingLabel.real()
			for (std::size_t i = 0; i < M; i++)
			{
				re[i] = ye[2 * i];
			}
			orc_predict_real(m->h, Xe, M, re.data(), grad != nullptr, nullptr, nullptr, nullptr, error, grad);
		}
		return GPLE_OK;
	}
	int gple_loose_function(gple_ctx*, const double* x, int nparam, double* grad, const double* X, const double* y, size_t N, const double* Xe, const double* ye, size_t M, double* value)
	{
		*value = orc_loose_function(x, nparam, grad, X, y, N, Xe, ye, M);
		return GPLE_OK;
	}

	int gple_pes(gple_ctx*, int pes_model, const double* x, size_t n, double* E, double* F, double* D)
	{
		orc_pes(pes_model, x, n, E, F, D);
		return GPLE_OK;
	}
	int gple_evolve(gple_ctx*, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, double* pts00, size_t n00, double* pts10, size_t n10, double* pts11, size_t n11, double mass, double dt)
	{
		orc_evolve(pes_model, pts00, n00, pts10, n10, pts11, n11, mass, dt, handle(m00), handle(m10), handle(m11), nullptr);
		return GPLE_OK;
	}
	// single-process mock of the multi-GPU entry points: a context without a communicator (rank 0 of 1)
	int gple_comm_unique_id(unsigned char*) { return GPLE_ERR_COMM; }
	int gple_ctx_comm_init(gple_ctx*, int, int, const unsigned char*) { return GPLE_ERR_COMM; }
	int gple_ctx_comm_info(const gple_ctx*, int* rank, int* nranks)
	{
		if (rank != nullptr)
		{
			*rank = 0;
		}
		if (nranks != nullptr)
		{
			*nranks = 1;
		}
		return GPLE_OK;
	}
	int gple_allreduce_sum(gple_ctx*, double*, size_t) { return GPLE_OK; }
	int gple_evolve_sharded(gple_ctx* ctx, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, double* pts00, size_t n00, double* pts10, size_t n10, double* pts11, size_t n11, double mass, double dt)
	{
		return gple_evolve(ctx, pes_model, m00, m10, m11, pts00, n00, pts10, n10, pts11, n11, mass, dt);
	}
	int gple_new_point_predict(gple_ctx*, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, const double* r, size_t n, int row, int col, double mass, double dt, double* out)
	{
		orc_new_point_predict(pes_model, r, n, mass, dt, row, col, handle(m00), handle(m10), handle(m11), out);
		return GPLE_OK;
	}
	int gple_observables(gple_ctx*, int pes_model, const double* pts, size_t n, double mass, int pes_index, double out[9])
	{
		orc_observable_sums(pes_model, pts, n, mass, pes_index, out);
		return GPLE_OK;
	}
	int gple_markov_chains(gple_ctx* ctx, const gple_mc_source* s, double* pts, size_t n, size_t num_steps, double max_displacement, unsigned long long seed, unsigned long long stream, unsigned long long chain0, double* accept_ratio, double* chains)
	{
		(void)ctx;
		orc_markov_chains(s->kind, s->analytic, handle(s->m00), handle(s->m10), handle(s->m11), s->pes_model, s->mass, s->dt, s->row, s->col, pts, n, num_steps, max_displacement, seed, stream, chain0, accept_ratio, chains);
		return GPLE_OK;
	}
	int gple_chain_autocorrelation(gple_ctx*, const double* chains, size_t n, size_t len, double* out)
	{
		orc_chain_autocorrelation(chains, n, len, out);
		return GPLE_OK;
	}
}
