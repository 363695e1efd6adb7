// Host-callable entry points of the Metropolis sampler (mc.cu).  All data pointers are device pointers.
#pragma once
#include "common.cuh"
#include "model.cuh"

namespace gple
{
/// n chains in lock-step: d_pts (n x 4: x, p, Re rho, Im rho) in/out; d_accept (n) and d_chain (n x (num_steps + 1) x 2) optional.
/// Chain k draws from the Philox stream (seed, stream, chain0 + k).
void markov_chains_device(gple_ctx* ctx, const gple_mc_source& src, double* d_pts, size_t n, size_t num_steps, double max_displacement, unsigned long long seed, unsigned long long stream, unsigned long long chain0, double* d_accept, double* d_chain);
/// per-device function attributes of the sampler kernels (called from gpr_setup_attributes)
void mc_setup_attributes();
/// d_out[len / 2] = mean autocorrelation of the n chains of length len (gple/mc.cpp:230-243)
void chain_autocorrelation_device(gple_ctx* ctx, const double* d_chains, size_t n, size_t len, double* d_out);
} // namespace gple
