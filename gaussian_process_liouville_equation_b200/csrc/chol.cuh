// Blocked Cholesky / triangular inverse / GEMM entry points (see chol.cu).
#pragma once
#include "common.cuh"
#include "dmma_gemm.cuh"

namespace gple
{
void chol_setup_attributes();
/// C = beta C + alpha A B^T (row-major, dims multiples of 128 / 16)
void gemm_nt(gple_ctx* ctx, const gemm::GemmArgs& a);
/// C = beta C + alpha A B
void gemm_nn(gple_ctx* ctx, const gemm::GemmArgs& a);
/// In place A (n x n, row-major, lower) <- L with A = L L^T, strictly-upper part zeroed; if W != nullptr,
/// W <- L^-1 (lower, upper part zero).  *d_info (device) = 0, or 1 + index of the first non-positive pivot.
void potrf_trtri(gple_ctx* ctx, double* A, double* W, int n, int* d_info);
/// Block size at or below which potrf runs the right-looking 128-block sweep instead of recursing; returns the old value
int set_potrf_flat(int n);
} // namespace gple
