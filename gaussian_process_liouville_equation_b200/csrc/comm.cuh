// Host-callable multi-GPU entry points (comm.cu): NCCL communicator of a context, block partition, in-place all-gather.
#pragma once
#include "common.cuh"
#include "model.cuh"

namespace gple
{
struct CommError
{
	const char* what;
};

/// Contiguous block [lo, hi) of `total` records owned by `rank`; blocks differ by at most one record.
inline void partition(const size_t total, const int rank, const int nranks, size_t& lo, size_t& hi)
{
	lo = total * size_t(rank) / size_t(nranks);
	hi = total * size_t(rank + 1) / size_t(nranks);
}

void comm_unique_id(unsigned char* id);
void comm_init(gple_ctx* ctx, int rank, int nranks, const unsigned char* id);
void comm_destroy(gple_ctx* ctx);
void allgather_blocks(gple_ctx* ctx, double* d_all, size_t total, size_t width);
void allreduce_sum(gple_ctx* ctx, double* d_values, size_t count);
void model_bcast(gple_ctx* ctx, gple_model** model, int root);
void evolve_sharded_device(gple_ctx* ctx, int pes_model, const gple_model* const models[3], double* d_pts[3], const size_t totals[3], double mass, double dt);
} // namespace gple
