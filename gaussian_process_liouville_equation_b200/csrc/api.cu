// extern "C" boundary of libgple_b200.so (declared in include/gple_b200.h).
#include "chol.cuh"
#include "comm.cuh"
#include "evolve.cuh"
#include "gpr.cuh"
#include "mc.cuh"

#include <exception>

using namespace gple;

namespace
{
template <typename F>
int guarded(gple_ctx* ctx, F&& f)
{
	if (ctx == nullptr)
	{
		return GPLE_ERR_ARG;
	}
	try
	{
		GPLE_CUDA(cudaSetDevice(ctx->device));
		return f();
	}
	catch (const CudaError& e)
	{
		char buf[512];
		std::snprintf(buf, sizeof(buf), "%s failed at %s:%d: %s", e.what, e.file, e.line, cudaGetErrorString(e.code));
		ctx->last_error = buf;
		cudaGetLastError();
		return GPLE_ERR_CUDA;
	}
	catch (const ArgError& e)
	{
		ctx->last_error = e.what;
		return GPLE_ERR_ARG;
	}
	catch (const CommError& e)
	{
		ctx->last_error = e.what;
		return GPLE_ERR_COMM;
	}
	catch (const std::exception& e)
	{
		ctx->last_error = e.what();
		return GPLE_ERR_CUDA;
	}
}

void require(bool ok, const char* what)
{
	if (!ok)
	{
		throw ArgError{what};
	}
}

void sync(gple_ctx* ctx)
{
	GPLE_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- FP64 peak probes: register-resident dependent chains, enough independent chains to fill the pipes ----
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, const int iters)
{
	double acc[16][2];
#pragma unroll
	for (int i = 0; i < 16; i++)
	{
		acc[i][0] = acc[i][1] = 0.0;
	}
	const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int i = 0; i < 16; i++)
		{
			gemm::dmma884(acc[i], a, b);
		}
	}
	double s = 0.0;
#pragma unroll
	for (int i = 0; i < 16; i++)
	{
		s += acc[i][0] + acc[i][1];
	}
	if (s == 12345.678)
	{
		out[0] = s;
	}
}
/// Same DMMA stream as the GEMM inner loop (8 x 4 register tile: 32 accumulators, 8 distinct A and 4 distinct B
/// operands per k-step, no memory traffic at all): the ceiling of a register-tiled mma.sync FP64 kernel.
__global__ void __launch_bounds__(256, 1) dmma_tile_peak_kernel(double* out, const int iters)
{
	double acc[8][4][2];
#pragma unroll
	for (int i = 0; i < 8; i++)
	{
#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			acc[i][j][0] = acc[i][j][1] = 0.0;
		}
	}
	double a[8], b[4];
#pragma unroll
	for (int i = 0; i < 8; i++)
	{
		a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
	}
#pragma unroll
	for (int j = 0; j < 4; j++)
	{
		b[j] = 1.0 - (threadIdx.x + j) * 1e-9;
	}
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int i = 0; i < 8; i++)
		{
#pragma unroll
			for (int j = 0; j < 4; j++)
			{
				gemm::dmma884(acc[i][j], a[i], b[j]);
			}
		}
#pragma unroll
		for (int i = 0; i < 8; i++)
		{
			a[i] += 1e-12; // operands change every k-step, as in a real GEMM
		}
	}
	double s = 0.0;
#pragma unroll
	for (int i = 0; i < 8; i++)
	{
#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			s += acc[i][j][0] + acc[i][j][1];
		}
	}
	if (s == 12345.678)
	{
		out[0] = s;
	}
}
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, const int iters)
{
	double acc[16];
#pragma unroll
	for (int i = 0; i < 16; i++)
	{
		acc[i] = threadIdx.x * 1e-3 + i;
	}
	const double a = 1.0 + 1e-12, b = 1e-9;
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int i = 0; i < 16; i++)
		{
			acc[i] = fma(acc[i], a, b);
		}
	}
	double s = 0.0;
#pragma unroll
	for (int i = 0; i < 16; i++)
	{
		s += acc[i];
	}
	if (s == 12345.678)
	{
		out[0] = s;
	}
}
} // namespace

namespace
{
/// Validation half of loose_function (opt.cpp:455-470): squared error of the raw prediction on the extra set and, optionally,
/// its gradient (kernel.cpp:519-541 / complex_kernel.cpp:645-667).  Xe / ye: host or device; ye complex interleaved.
void validation_error(gple_ctx* ctx, const gple_model* m, const double* Xe, const double* ye, const size_t M, double* error, double* grad)
{
	const size_t w = m->is_complex ? 2 : 1;
	DeviceArray<double> xe(ctx, Xe, 2 * M, false), yc(ctx, ye, 2 * M, false);
	// opt.cpp:451: the real kernel sees ExtraTrainingLabel.real()
	double* yq = yc.dev;
	if (!m->is_complex)
	{
		yq = ctx->ws.get<double>("loose.yre", M);
		GPLE_CUDA(cudaMemcpy2DAsync(yq, sizeof(double), yc.dev, 2 * sizeof(double), sizeof(double), M, cudaMemcpyDeviceToDevice, ctx->stream));
	}
	double* d_err = ctx->ws.get<double>("pred.err", 8);
	double* d_cut = grad != nullptr ? ctx->ws.get<double>("pred.cut_tmp", w * M) : nullptr;
	predict_device(ctx, m, xe.dev, M, yq, nullptr, nullptr, d_cut, d_err);
	GPLE_CUDA(cudaMemcpyAsync(ctx->h_pinned + 128, d_err, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	sync(ctx);
	*error = ctx->h_pinned[128];
	if (grad != nullptr)
	{
		validation_gradient(ctx, m, xe.dev, M, yq, d_cut, grad);
	}
}
} // namespace

extern "C"
{
	const char* gple_version(void)
	{
		return "gple_b200 0.1 (sm_100a)";
	}

	int gple_ctx_create(int device, gple_ctx** out)
	{
		if (out == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		*out = nullptr;
		int count = 0;
		if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
		{
			cudaGetLastError();
			return GPLE_ERR_CUDA; // no CUDA device: there is deliberately no CPU fallback
		}
		gple_ctx* ctx = new gple_ctx();
		ctx->device = device;
		const int rc = guarded(
			ctx,
			[&]() -> int
			{
				GPLE_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
				GPLE_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
				GPLE_CUDA(cudaEventCreateWithFlags(&ctx->ev_panel, cudaEventDisableTiming));
				GPLE_CUDA(cudaEventCreateWithFlags(&ctx->ev_bulk, cudaEventDisableTiming));
				GPLE_CUDA(cudaEventCreateWithFlags(&ctx->ev_graph, cudaEventDisableTiming));
				ctx->stream = ctx->own_stream;
				ctx->h_pinned_count = 512;
				GPLE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_pinned), ctx->h_pinned_count * sizeof(double)));
				// staging buffers of host-pointer arguments: a memory pool of this context's own, so that freed blocks are kept
				// (the default pool returns them to the driver at every synchronisation) and are never recycled into another
				// context's stream (which would chain the streams of concurrently working contexts together)
				{
					cudaMemPoolProps pp{};
					pp.allocType = cudaMemAllocationTypePinned;
					pp.handleTypes = cudaMemHandleTypeNone;
					pp.location.type = cudaMemLocationTypeDevice;
					pp.location.id = device;
					GPLE_CUDA(cudaMemPoolCreate(&ctx->mempool, &pp));
					unsigned long long threshold = ~0ull;
					GPLE_CUDA(cudaMemPoolSetAttribute(ctx->mempool, cudaMemPoolAttrReleaseThreshold, &threshold));
				}
				cudaDeviceProp prop{};
				GPLE_CUDA(cudaGetDeviceProperties(&prop, device));
				ctx->num_sms = prop.multiProcessorCount;
				gpr_setup_attributes();
				return GPLE_OK;
			}
		);
		if (rc != GPLE_OK)
		{
			std::fprintf(stderr, "gple_ctx_create: %s\n", ctx->last_error.c_str());
			delete ctx;
			return rc;
		}
		*out = ctx;
		return GPLE_OK;
	}

	int gple_ctx_destroy(gple_ctx* ctx)
	{
		if (ctx == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		cudaSetDevice(ctx->device);
		cudaStreamSynchronize(ctx->stream);
		comm_destroy(ctx);
		ctx->ws.release();
		ctx->pool.release();
		if (ctx->h_pinned != nullptr)
		{
			cudaFreeHost(ctx->h_pinned);
		}
		if (ctx->mempool != nullptr)
		{
			cudaMemPoolDestroy(ctx->mempool);
		}
		if (ctx->aux_stream != nullptr)
		{
			cudaStreamDestroy(ctx->aux_stream);
			cudaEventDestroy(ctx->ev_panel);
			cudaEventDestroy(ctx->ev_bulk);
		}
		for (auto& g : ctx->factorise_graph_cache)
		{
			cudaGraphExecDestroy(g.exec);
		}
		if (ctx->ev_graph != nullptr)
		{
			cudaEventDestroy(ctx->ev_graph);
		}
		if (ctx->own_stream != nullptr)
		{
			cudaStreamDestroy(ctx->own_stream);
		}
		delete ctx;
		return GPLE_OK;
	}

	int gple_ctx_set_stream(gple_ctx* ctx, void* stream)
	{
		if (ctx == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		ctx->stream = stream != nullptr ? static_cast<cudaStream_t>(stream) : ctx->own_stream;
		return GPLE_OK;
	}

	int gple_ctx_set_option(gple_ctx* ctx, int option, int value)
	{
		if (ctx == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		if (option == GPLE_OPT_GATED_VARIANCE)
		{
			ctx->gated_variance = value != 0;
			return GPLE_OK;
		}
		if (option == GPLE_OPT_REFINE_SOLUTION)
		{
			ctx->refine_solution = value != 0;
			return GPLE_OK;
		}
		if (option == GPLE_OPT_FACTORISE_GRAPHS)
		{
			ctx->factorise_graphs = value != 0;
			return GPLE_OK;
		}
		if (option == GPLE_OPT_GATE_STAGE_TILES && value >= -1)
		{
			ctx->gate_stage_tiles = value;
			return GPLE_OK;
		}
		if (option == GPLE_OPT_GATE_STAGE2_TILES && value >= -1)
		{
			ctx->gate_stage2_tiles = value;
			return GPLE_OK;
		}
		if (option == GPLE_OPT_GATE_STAGE_TILES_IM && value >= -1)
		{
			ctx->gate_stage_tiles_im = value;
			return GPLE_OK;
		}
		return GPLE_ERR_ARG;
	}

	int gple_ctx_set_gate_schedule(gple_ctx* ctx, int complex_element, int stages, const int* re_end, const int* im_end)
	{
		if (ctx == nullptr || stages < 0 || stages > 15 || (stages > 0 && re_end == nullptr))
		{
			return GPLE_ERR_ARG;
		}
		std::vector<int> flat;
		for (int k = 0; k < stages; k++)
		{
			const int im = im_end != nullptr ? im_end[k] : 0;
			if (re_end[k] < 0 || im < 0)
			{
				return GPLE_ERR_ARG;
			}
			flat.push_back(re_end[k]);
			flat.push_back(im);
		}
		ctx->gate_schedule[complex_element != 0 ? 1 : 0] = flat;
		return GPLE_OK;
	}

	int gple_gate_schedule_automatic(int complex_element, int blocks, int* re_end, int* im_end, int capacity)
	{
		if (blocks < 0 || capacity < 0 || (capacity > 0 && (re_end == nullptr || im_end == nullptr)))
		{
			return -int(GPLE_ERR_ARG);
		}
		return gple::gate_schedule_automatic_host(complex_element, blocks, re_end, im_end, capacity);
	}

	int gple_gate_statistics(gple_ctx* ctx, unsigned long long out[4])
	{
		if (ctx == nullptr || out == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		out[0] = ctx->gate_rows_total;
		out[1] = ctx->gate_rows_variance;
		out[2] = ctx->gate_rows_zero;
		out[3] = ctx->gate_rows_stage_b;
		ctx->gate_rows_total = ctx->gate_rows_variance = ctx->gate_rows_zero = ctx->gate_rows_stage_b = 0;
		return GPLE_OK;
	}

	int gple_ctx_sync(gple_ctx* ctx)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	const char* gple_last_error(const gple_ctx* ctx)
	{
		return ctx != nullptr ? ctx->last_error.c_str() : "null context";
	}

	unsigned long long gple_launch_count(const gple_ctx* ctx)
	{
		return ctx != nullptr ? ctx->launches : 0ull;
	}

	int gple_kernel_real(gple_ctx* ctx, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[4], int same_set, double* K_out, double* dK_out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(XL != nullptr && XR != nullptr && theta != nullptr && K_out != nullptr && nL > 0 && nR > 0, "gple_kernel_real: null argument or empty set");
				DeviceArray<double> l(ctx, XL, 2 * nL, false), r(ctx, XR, 2 * nR, false), k(ctx, K_out, nL * nR, true), dk(ctx, dK_out, 4 * nL * nR, true);
				kernel_real_device(ctx, l.dev, int(nL), r.dev, int(nR), theta, same_set, k.dev, dk.dev);
				k.finish();
				dk.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_kernel_complex(gple_ctx* ctx, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[8], int same_set, double* K_out, double* Kt_out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(XL != nullptr && XR != nullptr && theta != nullptr && K_out != nullptr && Kt_out != nullptr && nL > 0 && nR > 0, "gple_kernel_complex: null argument or empty set");
				DeviceArray<double> l(ctx, XL, 2 * nL, false), r(ctx, XR, 2 * nR, false), k(ctx, K_out, nL * nR, true), kt(ctx, Kt_out, 2 * nL * nR, true);
				kernel_complex_device(ctx, l.dev, int(nL), r.dev, int(nR), theta, same_set, k.dev, kt.dev);
				k.finish();
				kt.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_kernel_complex_derivatives(gple_ctx* ctx, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[8], int same_set, double* dK_out, double* dKt_out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(XL != nullptr && XR != nullptr && theta != nullptr && (dK_out != nullptr || dKt_out != nullptr) && nL > 0 && nR > 0, "gple_kernel_complex_derivatives: null argument or empty set");
				DeviceArray<double> l(ctx, XL, 2 * nL, false), r(ctx, XR, 2 * nR, false), k(ctx, dK_out, 8 * nL * nR, true), kt(ctx, dKt_out, 16 * nL * nR, true);
				kernel_complex_derivatives_device(ctx, l.dev, int(nL), r.dev, int(nR), theta, same_set, k.dev, kt.dev);
				k.finish();
				kt.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_train_real(gple_ctx* ctx, const double* X, const double* y, size_t N, const double theta[4], unsigned flags, gple_model** model, gple_real_scalars* out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(X != nullptr && y != nullptr && theta != nullptr && model != nullptr && N > 0, "gple_train_real: null argument or empty training set");
				*model = nullptr;
				return train_real(ctx, X, y, N, theta, flags, model, out);
			}
		);
	}

	int gple_train_complex(gple_ctx* ctx, const double* X, const double* y, size_t N, const double theta[8], unsigned flags, gple_model** model, gple_complex_scalars* out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(X != nullptr && y != nullptr && theta != nullptr && model != nullptr && N > 0, "gple_train_complex: null argument or empty training set");
				*model = nullptr;
				return train_complex(ctx, X, y, N, theta, flags, model, out);
			}
		);
	}

	int gple_model_is_complex(const gple_model* m)
	{
		return m != nullptr ? m->is_complex : -1;
	}
	int gple_model_nlml(gple_ctx* ctx, gple_model* m, double* value, double* grad)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(m != nullptr && value != nullptr, "gple_model_nlml: null argument");
				nlml_device(ctx, m, value, grad);
				return GPLE_OK;
			}
		);
	}
	size_t gple_model_size(const gple_model* m)
	{
		return m != nullptr ? m->N : 0;
	}
	int gple_model_destroy(gple_ctx* ctx, gple_model* m)
	{
		if (ctx == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		cudaSetDevice(ctx->device);
		free_model(ctx, m); // buffers go back to the context's pool; reuse is ordered on the context's stream
		return GPLE_OK;
	}

	int gple_model_get(gple_ctx* ctx, const gple_model* cm, int which, double* out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(cm != nullptr && out != nullptr, "gple_model_get: null argument");
				gple_model* m = const_cast<gple_model*>(cm);
				const size_t N = m->N, Np = size_t(m->Np), n = size_t(m->n);
				std::vector<double> h;
				auto fetch = [&](const double* d, size_t count)
				{
					h.resize(count);
					GPLE_CUDA(cudaMemcpyAsync(h.data(), d, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
					sync(ctx);
				};
				std::vector<double> res;
				if (!m->is_complex)
				{
					if (which == GPLE_FIELD_INVERSE)
					{
						ensure_full_inverse(ctx, m);
						fetch(m->Kinv, n * n);
						res.resize(N * N);
						for (size_t c = 0; c < N; c++)
						{
							for (size_t r = 0; r < N; r++)
							{
								res[c * N + r] = h[r * n + c];
							}
						}
					}
					else if (which == GPLE_FIELD_INV_LABEL || which == GPLE_FIELD_LABEL)
					{
						fetch(which == GPLE_FIELD_LABEL ? m->label : m->v, n);
						res.assign(h.begin(), h.begin() + N);
					}
					else if (which >= GPLE_FIELD_INV_LABEL_DERIV && which < GPLE_FIELD_INV_LABEL_DERIV + 4 && m->dv != nullptr)
					{
						fetch(m->dv + size_t(which - GPLE_FIELD_INV_LABEL_DERIV) * n, n);
						res.assign(h.begin(), h.begin() + N);
					}
					else
					{
						return GPLE_ERR_STATE;
					}
				}
				else
				{
					if (which == GPLE_FIELD_INV_LABEL || which == GPLE_FIELD_LABEL)
					{
						// v = (wr + i wi) / 2 ; label = yr + i yi
						fetch(which == GPLE_FIELD_LABEL ? m->label : m->v, n);
						const double f = which == GPLE_FIELD_LABEL ? 1.0 : 0.5;
						res.resize(2 * N);
						for (size_t i = 0; i < N; i++)
						{
							res[2 * i] = f * h[i];
							res[2 * i + 1] = f * h[Np + i];
						}
					}
					else if (which == GPLE_FIELD_UPPER_LEFT || which == GPLE_FIELD_LOWER_LEFT)
					{
						// P = (Mrr + Mii + i (Mir - Mri)) / 4 ; Q = (Mrr - Mii - i (Mir + Mri)) / 4 with M = C^-1
						ensure_full_inverse(ctx, m);
						fetch(m->Kinv, n * n);
						res.resize(2 * N * N);
						const double sg = which == GPLE_FIELD_UPPER_LEFT ? 1.0 : -1.0;
						for (size_t c = 0; c < N; c++)
						{
							for (size_t r = 0; r < N; r++)
							{
								const double mrr = h[r * n + c], mii = h[(Np + r) * n + Np + c], mri = h[r * n + Np + c], mir = h[(Np + r) * n + c];
								res[2 * (c * N + r)] = 0.25 * (mrr + sg * mii);
								res[2 * (c * N + r) + 1] = 0.25 * (sg * mir - mri);
							}
						}
					}
					else
					{
						return GPLE_ERR_STATE;
					}
				}
				// `out` may be a device pointer as well
				DeviceArray<double> o(ctx, out, res.size(), true);
				if (o.owned)
				{
					std::memcpy(out, res.data(), res.size() * sizeof(double));
				}
				else
				{
					GPLE_CUDA(cudaMemcpyAsync(o.dev, res.data(), res.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
					sync(ctx);
				}
				return GPLE_OK;
			}
		);
	}

	static int predict_any(gple_ctx* ctx, const gple_model* m, int want_complex, const double* Xq, size_t Q, const double* yq, double* pred_out, double* var_out, double* cutoff_out, double* err_out, double* derr_out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(m != nullptr && Xq != nullptr && Q > 0, "gple_predict: null argument or no query point");
				require(m->is_complex == want_complex, "gple_predict: model kind does not match the entry point");
				const size_t w = m->is_complex ? 2 : 1;
				DeviceArray<double> xq(ctx, Xq, 2 * Q, false), y(ctx, yq, w * Q, false);
				DeviceArray<double> pred(ctx, pred_out, w * Q, true), var(ctx, var_out, Q, true), cut(ctx, cutoff_out, w * Q, true);
				const bool want_grad = derr_out != nullptr && yq != nullptr;
				double* d_cut = cut.dev;
				if (want_grad && d_cut == nullptr)
				{
					d_cut = ctx->ws.get<double>("pred.cut_tmp", w * Q);
				}
				double* d_err = (err_out != nullptr && yq != nullptr) ? ctx->ws.get<double>("pred.err", 8) : nullptr;
				predict_device(ctx, m, xq.dev, Q, y.dev, pred.dev, var.dev, d_cut, d_err);
				pred.finish();
				var.finish();
				cut.finish();
				if (d_err != nullptr)
				{
					GPLE_CUDA(cudaMemcpyAsync(ctx->h_pinned + 128, d_err, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
				}
				sync(ctx);
				if (d_err != nullptr)
				{
					*err_out = ctx->h_pinned[128];
				}
				if (want_grad)
				{
					validation_gradient(ctx, m, xq.dev, Q, y.dev, d_cut, derr_out);
				}
				return GPLE_OK;
			}
		);
	}

	int gple_predict_real(gple_ctx* ctx, const gple_model* m, const double* Xq, size_t Q, const double* yq, double* pred_out, double* var_out, double* cutoff_out, double* err_out, double* derr_out)
	{
		return predict_any(ctx, m, 0, Xq, Q, yq, pred_out, var_out, cutoff_out, err_out, derr_out);
	}
	int gple_predict_complex(gple_ctx* ctx, const gple_model* m, const double* Xq, size_t Q, const double* yq, double* pred_out, double* var_out, double* cutoff_out, double* err_out, double* derr_out)
	{
		return predict_any(ctx, m, 1, Xq, Q, yq, pred_out, var_out, cutoff_out, err_out, derr_out);
	}

	int gple_validation_error(gple_ctx* ctx, const gple_model* m, const double* Xe, const double* ye, size_t M, double* error, double* grad)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(m != nullptr && Xe != nullptr && ye != nullptr && error != nullptr && M > 0, "gple_validation_error: null argument");
				require(grad == nullptr || m->dv != nullptr, "gple_validation_error: the gradient needs a model trained with GPLE_CALC_DERIVATIVE");
				validation_error(ctx, m, Xe, ye, M, error, grad);
				return GPLE_OK;
			}
		);
	}

	int gple_loose_function(gple_ctx* ctx, const double* x, int nparam, double* grad, const double* X, const double* y, size_t N, const double* Xe, const double* ye, size_t M, double* value)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(x != nullptr && X != nullptr && y != nullptr && Xe != nullptr && ye != nullptr && value != nullptr && N > 0 && M > 0, "gple_loose_function: null argument");
				require(nparam == 4 || nparam == 8, "gple_loose_function: nparam must be 4 (real) or 8 (complex)");
				const unsigned flags = GPLE_CALC_ERROR | (grad != nullptr ? unsigned(GPLE_CALC_DERIVATIVE) : 0u);
				gple_model* m = nullptr;
				double trn_err = 0.0, trn_grad[8];
				int rc;
				if (nparam == 4)
				{
					gple_real_scalars s{};
					rc = train_real(ctx, X, y, N, x, flags, &m, &s);
					trn_err = s.error;
					std::memcpy(trn_grad, s.d_error, 4 * sizeof(double));
				}
				else
				{
					gple_complex_scalars s{};
					rc = train_complex(ctx, X, y, N, x, flags, &m, &s);
					trn_err = s.error;
					std::memcpy(trn_grad, s.d_error, 8 * sizeof(double));
				}
				if (rc != GPLE_OK)
				{
					free_model(ctx, m);
					*value = std::nan("");
					return rc;
				}
				double val_err = 0.0, vg[8];
				validation_error(ctx, m, Xe, ye, M, &val_err, grad != nullptr ? vg : nullptr);
				*value = trn_err + val_err;
				if (grad != nullptr)
				{
					for (int p = 0; p < nparam; p++)
					{
						grad[p] = trn_grad[p] + vg[p];
					}
				}
				free_model(ctx, m);
				return GPLE_OK;
			}
		);
	}

	int gple_pes(gple_ctx* ctx, int pes_model, const double* x, size_t n, double* E, double* F, double* D)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(x != nullptr && E != nullptr && F != nullptr && D != nullptr && n > 0, "gple_pes: null argument");
				require(pes_model >= GPLE_SAC && pes_model <= GPLE_ECR, "gple_pes: unknown model");
				DeviceArray<double> dx(ctx, x, n, false), e(ctx, E, 2 * n, true), f(ctx, F, 3 * n, true), d(ctx, D, n, true);
				pes_device(ctx, pes_model, dx.dev, n, e.dev, f.dev, d.dev);
				e.finish();
				f.finish();
				d.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_evolve(gple_ctx* ctx, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, double* pts00, size_t n00, double* pts10, size_t n10, double* pts11, size_t n11, double mass, double dt)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(pes_model >= GPLE_SAC && pes_model <= GPLE_ECR, "gple_evolve: unknown model");
				require((m00 == nullptr || !m00->is_complex) && (m11 == nullptr || !m11->is_complex) && (m10 == nullptr || m10->is_complex), "gple_evolve: diagonal elements take real models, rho10 a complex model");
				require((n00 == 0 || pts00 != nullptr) && (n10 == 0 || pts10 != nullptr) && (n11 == 0 || pts11 != nullptr), "gple_evolve: null point set");
				DeviceArray<double> a(ctx, pts00, 4 * n00, false), b(ctx, pts10, 4 * n10, false), c(ctx, pts11, 4 * n11, false);
				a.write_back = b.write_back = c.write_back = true;
				const gple_model* models[3] = {m00, m10, m11};
				double* d_pts[3] = {a.dev, b.dev, c.dev};
				const size_t counts[3] = {n00, n10, n11};
				evolve_device(ctx, pes_model, models, d_pts, counts, mass, dt);
				a.finish();
				b.finish();
				c.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	// ---- multi-GPU ----------------------------------------------------------------------------------
	int gple_comm_unique_id(unsigned char id[GPLE_COMM_ID_BYTES])
	{
		if (id == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		try
		{
			comm_unique_id(id);
			return GPLE_OK;
		}
		catch (const CommError& e)
		{
			std::fprintf(stderr, "gple_comm_unique_id: %s\n", e.what);
			return GPLE_ERR_COMM;
		}
	}
	int gple_ctx_comm_init(gple_ctx* ctx, int rank, int nranks, const unsigned char id[GPLE_COMM_ID_BYTES])
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(id != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, "gple_ctx_comm_init: bad rank / size / id");
				comm_init(ctx, rank, nranks, id);
				return GPLE_OK;
			}
		);
	}
	int gple_ctx_comm_info(const gple_ctx* ctx, int* rank, int* nranks)
	{
		if (ctx == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		if (rank != nullptr)
		{
			*rank = ctx->comm_rank;
		}
		if (nranks != nullptr)
		{
			*nranks = ctx->comm_size;
		}
		return GPLE_OK;
	}
	int gple_partition(size_t total, int rank, int nranks, size_t* lo, size_t* hi)
	{
		if (lo == nullptr || hi == nullptr || nranks < 1 || rank < 0 || rank >= nranks)
		{
			return GPLE_ERR_ARG;
		}
		partition(total, rank, nranks, *lo, *hi);
		return GPLE_OK;
	}
	int gple_allgather_points(gple_ctx* ctx, double* pts, size_t total)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(pts != nullptr || total == 0, "gple_allgather_points: null point set");
				DeviceArray<double> p(ctx, pts, 4 * total, false);
				p.write_back = true;
				allgather_blocks(ctx, p.dev, total, 4);
				p.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}
	int gple_allreduce_sum(gple_ctx* ctx, double* values, size_t count)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(values != nullptr && count > 0, "gple_allreduce_sum: null argument");
				DeviceArray<double> v(ctx, values, count, false);
				v.write_back = true;
				allreduce_sum(ctx, v.dev, count);
				v.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}
	int gple_model_bcast(gple_ctx* ctx, gple_model** model, int root)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(model != nullptr && root >= 0 && root < ctx->comm_size, "gple_model_bcast: null model slot or bad root");
				model_bcast(ctx, model, root);
				sync(ctx);
				return GPLE_OK;
			}
		);
	}
	int gple_evolve_sharded(gple_ctx* ctx, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, double* pts00, size_t n00, double* pts10, size_t n10, double* pts11, size_t n11, double mass, double dt)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(pes_model >= GPLE_SAC && pes_model <= GPLE_ECR, "gple_evolve_sharded: unknown model");
				require((m00 == nullptr || !m00->is_complex) && (m11 == nullptr || !m11->is_complex) && (m10 == nullptr || m10->is_complex), "gple_evolve_sharded: diagonal elements take real models, rho10 a complex model");
				require((n00 == 0 || pts00 != nullptr) && (n10 == 0 || pts10 != nullptr) && (n11 == 0 || pts11 != nullptr), "gple_evolve_sharded: null point set");
				DeviceArray<double> a(ctx, pts00, 4 * n00, false), b(ctx, pts10, 4 * n10, false), c(ctx, pts11, 4 * n11, false);
				a.write_back = b.write_back = c.write_back = true;
				const gple_model* models[3] = {m00, m10, m11};
				double* d_pts[3] = {a.dev, b.dev, c.dev};
				const size_t totals[3] = {n00, n10, n11};
				evolve_sharded_device(ctx, pes_model, models, d_pts, totals, mass, dt);
				a.finish();
				b.finish();
				c.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_new_point_predict(gple_ctx* ctx, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, const double* r, size_t n, int row, int col, double mass, double dt, double* out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(r != nullptr && out != nullptr && n > 0, "gple_new_point_predict: null argument");
				require(row >= 0 && row < 2 && col >= 0 && col <= row, "gple_new_point_predict: (row, col) must be a lower-triangular index");
				DeviceArray<double> dr(ctx, r, 2 * n, false), o(ctx, out, 2 * n, true);
				const gple_model* models[3] = {m00, m10, m11};
				new_point_predict_device(ctx, pes_model, models, dr.dev, n, row, col, mass, dt, o.dev);
				o.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_markov_chains(gple_ctx* ctx, const gple_mc_source* src, double* pts, size_t n, size_t num_steps, double max_displacement, unsigned long long seed, unsigned long long stream, unsigned long long chain0, double* accept_ratio, double* chains)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(src != nullptr && pts != nullptr && n > 0, "gple_markov_chains: null argument");
				require(src->kind >= GPLE_MC_ANALYTIC && src->kind <= GPLE_MC_NEW_POINT, "gple_markov_chains: unknown source kind");
				require(src->row >= 0 && src->row < 2 && src->col >= 0 && src->col <= src->row, "gple_markov_chains: (row, col) must be a lower-triangular index");
				require(max_displacement > 0.0 && num_steps < 0xffffffffull, "gple_markov_chains: bad displacement or step count");
				DeviceArray<double> p(ctx, pts, 4 * n, true), a(ctx, accept_ratio, n, true), c(ctx, chains, 2 * n * (num_steps + 1), true);
				if (p.host != nullptr)
				{
					GPLE_CUDA(cudaMemcpyAsync(p.dev, p.host, 4 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
				}
				markov_chains_device(ctx, *src, p.dev, n, num_steps, max_displacement, seed, stream, chain0, a.dev, c.dev);
				p.finish();
				a.finish();
				c.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_chain_autocorrelation(gple_ctx* ctx, const double* chains, size_t n, size_t len, double* out)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(chains != nullptr && out != nullptr && n > 0 && len >= 2, "gple_chain_autocorrelation: null argument");
				DeviceArray<double> c(ctx, chains, 2 * n * len, false), o(ctx, out, len / 2, true);
				chain_autocorrelation_device(ctx, c.dev, n, len, o.dev);
				o.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_observables(gple_ctx* ctx, int pes_model, const double* pts, size_t n, double mass, int pes_index, double out[9])
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(pts != nullptr && out != nullptr && n > 0, "gple_observables: null argument");
				DeviceArray<double> p(ctx, pts, 4 * n, false), o(ctx, out, 9, true);
				observables_device(ctx, pes_model, p.dev, n, mass, pes_index, o.dev);
				o.finish();
				sync(ctx);
				return GPLE_OK;
			}
		);
	}

	int gple_tune_variance_gemm(gple_ctx* ctx, int variant, int rows, int n, int iters, double* ms_per_launch)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(rows > 0 && rows % 128 == 0 && n > 0 && n % 128 == 0 && iters > 0 && ms_per_launch != nullptr, "gple_tune_variance_gemm: rows and n must be multiples of 128");
				*ms_per_launch = bench_variance_gemm(ctx, variant, rows, n, iters);
				return GPLE_OK;
			}
		);
	}

	int gple_set_potrf_flat(int n)
	{
		return set_potrf_flat(n < 128 ? 128 : n);
	}
	int gple_set_variance_gemm_variant(int variant)
	{
		return set_variance_gemm_variant(variant) == 0 ? GPLE_OK : GPLE_ERR_ARG;
	}

	int gple_profile_enable(gple_ctx* ctx, int on)
	{
		if (ctx == nullptr)
		{
			return GPLE_ERR_ARG;
		}
		ctx->prof_on = on != 0;
		return GPLE_OK;
	}

	int gple_profile_read(gple_ctx* ctx, int slot, double* total_ms, unsigned long long* launches, double* work)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(slot >= 0 && slot < 4, "gple_profile_read: unknown slot");
				sync(ctx);
				auto& p = ctx->prof[slot];
				double tot = 0.0;
				for (auto& e : p.ev)
				{
					float ms = 0.f;
					GPLE_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
					tot += ms;
					cudaEventDestroy(e.first);
					cudaEventDestroy(e.second);
				}
				p.ev.clear();
				if (total_ms != nullptr)
				{
					*total_ms = tot;
				}
				if (launches != nullptr)
				{
					*launches = p.launches;
				}
				if (work != nullptr)
				{
					*work = p.work;
				}
				p.launches = 0;
				p.work = 0.0;
				return GPLE_OK;
			}
		);
	}

	int gple_measure_dmma_tile_peak(gple_ctx* ctx, double* tflops)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				require(tflops != nullptr, "gple_measure_dmma_tile_peak: null argument");
				double* d = ctx->ws.get<double>("peak.out", 8);
				const int iters = 2048, blocks = ctx->num_sms;
				cudaEvent_t e0, e1;
				GPLE_CUDA(cudaEventCreate(&e0));
				GPLE_CUDA(cudaEventCreate(&e1));
				double best = 0.0;
				for (int rep = 0; rep < 4; rep++)
				{
					float ms = 0.f;
					GPLE_CUDA(cudaEventRecord(e0, ctx->stream));
					GPLE_LAUNCH(ctx, dmma_tile_peak_kernel, blocks, 256, 0, d, iters);
					GPLE_CUDA(cudaEventRecord(e1, ctx->stream));
					GPLE_CUDA(cudaEventSynchronize(e1));
					GPLE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
					best = std::max(best, double(blocks) * 8 * 32 * iters * 512.0 / (ms * 1e-3) / 1e12);
				}
				cudaEventDestroy(e0);
				cudaEventDestroy(e1);
				*tflops = best;
				return GPLE_OK;
			}
		);
	}

	int gple_measure_fp64_peak(gple_ctx* ctx, double* dmma_tflops, double* dfma_tflops)
	{
		return guarded(
			ctx,
			[&]() -> int
			{
				double* d = ctx->ws.get<double>("peak.out", 8);
				const int iters = 4096, blocks = ctx->num_sms * 4;
				cudaEvent_t e0, e1;
				GPLE_CUDA(cudaEventCreate(&e0));
				GPLE_CUDA(cudaEventCreate(&e1));
				float ms = 0.f;
				double best_mma = 0.0, best_fma = 0.0;
				for (int rep = 0; rep < 4; rep++)
				{
					GPLE_CUDA(cudaEventRecord(e0, ctx->stream));
					GPLE_LAUNCH(ctx, dmma_peak_kernel, blocks, 256, 0, d, iters);
					GPLE_CUDA(cudaEventRecord(e1, ctx->stream));
					GPLE_CUDA(cudaEventSynchronize(e1));
					GPLE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
					best_mma = std::max(best_mma, double(blocks) * 8 * 16 * iters * 512.0 / (ms * 1e-3) / 1e12);
					GPLE_CUDA(cudaEventRecord(e0, ctx->stream));
					GPLE_LAUNCH(ctx, dfma_peak_kernel, blocks, 256, 0, d, iters * 4);
					GPLE_CUDA(cudaEventRecord(e1, ctx->stream));
					GPLE_CUDA(cudaEventSynchronize(e1));
					GPLE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
					best_fma = std::max(best_fma, double(blocks) * 256 * 16 * (iters * 4.0) * 2.0 / (ms * 1e-3) / 1e12);
				}
				cudaEventDestroy(e0);
				cudaEventDestroy(e1);
				if (dmma_tflops != nullptr)
				{
					*dmma_tflops = best_mma;
				}
				if (dfma_tflops != nullptr)
				{
					*dfma_tflops = best_fma;
				}
				return GPLE_OK;
			}
		);
	}
}
