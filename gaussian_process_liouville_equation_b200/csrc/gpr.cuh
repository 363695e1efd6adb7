// Host-callable entry points of the GPR part (gpr.cu, gpr_deriv.cu).
#pragma once
#include "common.cuh"
#include "model.cuh"

namespace gple
{
void gpr_setup_attributes();
int train_real(gple_ctx* ctx, const double* X, const double* y, size_t N, const double* theta, unsigned flags, gple_model** model, gple_real_scalars* out);
int train_complex(gple_ctx* ctx, const double* X, const double* y, size_t N, const double* theta, unsigned flags, gple_model** model, gple_complex_scalars* out);
/// All pointers are device pointers.  d_pred / d_cut: 1 (real) or 2 (complex, interleaved) doubles per point;
/// d_err accumulates the squared validation error against d_yq (same layout as d_pred); any output may be null.
void predict_device(gple_ctx* ctx, const gple_model* m, const double* d_Xq, size_t Q, const double* d_yq, double* d_pred, double* d_var, double* d_cut, double* d_err);
int set_variance_gemm_variant(int variant);
double bench_variance_gemm(gple_ctx* ctx, int variant, int rows, int n, int iters);
void kernel_real_device(gple_ctx* ctx, const double* XL, int nL, const double* XR, int nR, const double* theta, int same, double* K, double* dK);
void kernel_complex_derivatives_device(gple_ctx* ctx, const double* XL, int nL, const double* XR, int nR, const double* theta, int same, double* dK, double* dKt);
void kernel_complex_device(gple_ctx* ctx, const double* XL, int nL, const double* XR, int nR, const double* theta, int same, double* K, double* Kt);

/// gpr_deriv.cu: parameter gradients (kernel.cpp:337-477, 524-541; complex_kernel.cpp:379-590, 648-667)
void real_derivatives(gple_ctx* ctx, gple_model* m, unsigned flags, const double* h_scal, gple_real_scalars* r);
void complex_derivatives(gple_ctx* ctx, gple_model* m, unsigned flags, const double* h_scal, gple_complex_scalars* r);
/// d(validation error)/d theta for a model trained with GPLE_CALC_DERIVATIVE; d_cut / d_yq as in predict_device
void validation_gradient(gple_ctx* ctx, const gple_model* m, const double* d_Xq, size_t Q, const double* d_yq, const double* d_cut, double* h_grad);
/// NLML / LLT objective (test/gpr.cpp:470-532) of a trained model; grad (nparam doubles, host) may be null
void nlml_device(gple_ctx* ctx, gple_model* m, double* value, double* grad);
/// the automatic schedule of the staged bound for a model of `blocks` 128-blocks of training points (host only); returns the number
/// of stages before the last one and writes at most `capacity` of them
int gate_schedule_automatic_host(int complex_element, int blocks, int* re_end, int* im_end, int capacity);
} // namespace gple
