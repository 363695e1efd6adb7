// Host-callable entry points of the dynamics part (evolve.cu).  All pointers are device pointers.
#pragma once
#include "common.cuh"
#include "model.cuh"

namespace gple
{
/// models / d_pts / counts in lower-triangular element order (rho00, rho10, rho11)
void evolve_device(gple_ctx* ctx, int pes_model, const gple_model* const models[3], double* d_pts[3], const size_t counts[3], double mass, double dt);
void new_point_predict_device(gple_ctx* ctx, int pes_model, const gple_model* const models[3], const double* d_r, size_t n, int row, int col, double mass, double dt, double* d_out);
void pes_device(gple_ctx* ctx, int pes_model, const double* d_x, size_t n, double* E, double* F, double* D);
void observables_device(gple_ctx* ctx, int pes_model, const double* d_pts, size_t n, double mass, int pes_index, double* d_out9);
} // namespace gple
