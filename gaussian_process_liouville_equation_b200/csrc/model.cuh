// Device-resident state of one trained density-matrix element.
#pragma once
#include "common.cuh"

struct gple_model
{
	gple_ctx* owner = nullptr; // context whose buffer pool the device buffers came from
	int is_complex = 0;
	size_t N = 0;	// training points
	int Np = 0;		// N padded to a multiple of 128
	int n = 0;		// order of the factorised matrix: Np (real kernel) or 2 Np (complex: composite [Re; Im] process)
	double theta[8] = {0};
	unsigned flags = 0;
	double rescale = 1.0;
	double prior = 0.0; // k**: prior variance of a single point including noise
	// device buffers (owned)
	double* X = nullptr;	 // 2 * Np: training coordinates, padded with zeros
	double* W = nullptr;	 // n x n: L^-1, row-major lower triangle (K^-1 = W^T W)
	double* v = nullptr;	 // n: K^-1 y'  (complex: w = C^-1 [Re y'; Im y'])
	double* label = nullptr; // n: rescaled labels y'
	double* kinv_diag = nullptr; // real: Np diag(K^-1); complex: 3 * Np (M_rr, M_ii, M_ri diagonals)
	double* Kinv = nullptr;	 // n x n full inverse, built on demand (derivatives, getters)
	double* dv = nullptr;	 // nparam * n: d v / d theta (only with GPLE_CALC_DERIVATIVE)
	int nparam() const { return is_complex ? 8 : 4; }
};

namespace gple
{
void free_model(gple_ctx* ctx, gple_model* m);
/// Kinv = W^T W (full symmetric), cached in the model
void ensure_full_inverse(gple_ctx* ctx, gple_model* m);
} // namespace gple
