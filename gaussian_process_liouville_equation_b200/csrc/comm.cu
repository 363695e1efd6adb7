// Multi-GPU part of the C-ABI (include/gple_b200.h, "multi-GPU"): one process per GPU, one NCCL communicator per context.
// The path shards over the evolved points; its only per-step exchange is the all-gather of the evolved
// (r, rho) sets, after which every rank rebuilds its element models from the full sets (gple/main.cpp:140-141, 176;
// predict.cpp:246-280).  NCCL is bound at run time (dlopen of libnccl.so.2 -- in a torch process that is the copy torch
// already loaded), so that the library also loads on a box without NCCL and single-GPU users never touch it.
#include "comm.cuh"
#include "evolve.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

namespace gple
{
namespace
{
struct NcclApi
{
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
};

const NcclApi& nccl()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(
		once,
		[]()
		{
			// the copy already in the process first (torch's), then the loader's search path
			void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
			if (h == nullptr)
			{
				h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
			}
			if (h == nullptr)
			{
				return;
			}
			api.handle = h;
#define GPLE_NCCL_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name))
			GPLE_NCCL_SYM(GetUniqueId);
			GPLE_NCCL_SYM(CommInitRank);
			GPLE_NCCL_SYM(CommDestroy);
			GPLE_NCCL_SYM(AllGather);
			GPLE_NCCL_SYM(AllReduce);
			GPLE_NCCL_SYM(Broadcast);
			GPLE_NCCL_SYM(GroupStart);
			GPLE_NCCL_SYM(GroupEnd);
			GPLE_NCCL_SYM(GetErrorString);
#undef GPLE_NCCL_SYM
			api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce && api.Broadcast && api.GroupStart && api.GroupEnd && api.GetErrorString;
		}
	);
	return api;
}

void check_nccl(const ncclResult_t r, const char* what)
{
	if (r != ncclSuccess)
	{
		static thread_local char buf[256];
		std::snprintf(buf, sizeof(buf), "%s: %s", what, nccl().GetErrorString(r));
		throw CommError{buf};
	}
}
void require_api()
{
	if (!nccl().ok)
	{
		throw CommError{"libnccl.so.2 could not be loaded (multi-GPU entry points need NCCL)"};
	}
}
static_assert(sizeof(ncclUniqueId) == GPLE_COMM_ID_BYTES, "gple_b200.h: GPLE_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
} // namespace

void comm_unique_id(unsigned char* id)
{
	require_api();
	ncclUniqueId u;
	check_nccl(nccl().GetUniqueId(&u), "ncclGetUniqueId");
	std::memcpy(id, &u, sizeof(u));
}

void comm_init(gple_ctx* ctx, const int rank, const int nranks, const unsigned char* id)
{
	require_api();
	if (ctx->comm != nullptr)
	{
		throw ArgError{"gple_ctx_comm_init: this context already has a communicator"};
	}
	ncclUniqueId u;
	std::memcpy(&u, id, sizeof(u));
	ncclComm_t c = nullptr;
	check_nccl(nccl().CommInitRank(&c, nranks, u, rank), "ncclCommInitRank");
	ctx->comm = c;
	ctx->comm_rank = rank;
	ctx->comm_size = nranks;
}

void comm_destroy(gple_ctx* ctx)
{
	if (ctx->comm != nullptr && nccl().ok)
	{
		nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
	}
	ctx->comm = nullptr;
	ctx->comm_rank = 0;
	ctx->comm_size = 1;
}

/// In-place all-gather of a block-partitioned array of `total` records of `width` doubles (rank r owns partition(total, r)).
void allgather_blocks(gple_ctx* ctx, double* d_all, const size_t total, const size_t width)
{
	if (ctx->comm_size <= 1 || total == 0)
	{
		return;
	}
	const NcclApi& n = nccl();
	const ncclComm_t c = static_cast<ncclComm_t>(ctx->comm);
	const int G = ctx->comm_size;
	if (total % size_t(G) == 0)
	{
		const size_t count = total / size_t(G) * width;
		check_nccl(n.AllGather(d_all + size_t(ctx->comm_rank) * count, d_all, count, ncclDouble, c, ctx->stream), "ncclAllGather");
		return;
	}
	// uneven blocks: every rank broadcasts its own block to its place (one fused group)
	check_nccl(n.GroupStart(), "ncclGroupStart");
	for (int r = 0; r < G; r++)
	{
		size_t lo, hi;
		partition(total, r, G, lo, hi);
		if (hi > lo)
		{
			check_nccl(n.Broadcast(d_all + lo * width, d_all + lo * width, (hi - lo) * width, ncclDouble, r, c, ctx->stream), "ncclBroadcast");
		}
	}
	check_nccl(n.GroupEnd(), "ncclGroupEnd");
}

void allreduce_sum(gple_ctx* ctx, double* d_values, const size_t count)
{
	if (ctx->comm_size <= 1 || count == 0)
	{
		return;
	}
	check_nccl(nccl().AllReduce(d_values, d_values, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->comm), ctx->stream), "ncclAllReduce");
}

/// One element model, trained on `root`, replicated on every rank: the owner sends a small header (sizes, parameters,
/// scalars) and then its device buffers (X, W = L^-1, v, labels, diag(K^-1)); the other ranks allocate the same buffers from
/// their own pools.  On entry *model is the trained model on the root and ignored elsewhere (NULL on every rank = the element
/// is not populated: nothing is sent).  Kinv / dv (on-demand, derivative-only) are not replicated.
void model_bcast(gple_ctx* ctx, gple_model** model, const int root)
{
	if (ctx->comm_size <= 1)
	{
		return;
	}
	const NcclApi& n = nccl();
	const ncclComm_t c = static_cast<ncclComm_t>(ctx->comm);
	struct Header
	{
		double present, is_complex, N, Np, n, flags, rescale, prior, theta[8];
	};
	static_assert(sizeof(Header) == 16 * sizeof(double), "header is 16 doubles");
	Header h{};
	const bool mine = ctx->comm_rank == root;
	if (mine && *model != nullptr)
	{
		const gple_model* m = *model;
		h.present = 1.0;
		h.is_complex = m->is_complex;
		h.N = double(m->N);
		h.Np = m->Np;
		h.n = m->n;
		h.flags = m->flags;
		h.rescale = m->rescale;
		h.prior = m->prior;
		std::memcpy(h.theta, m->theta, sizeof(h.theta));
	}
	double* d_h = ctx->ws.get<double>("comm.header", 16);
	if (mine)
	{
		GPLE_CUDA(cudaMemcpyAsync(d_h, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
	}
	check_nccl(n.Broadcast(d_h, d_h, 16, ncclDouble, root, c, ctx->stream), "ncclBroadcast(header)");
	GPLE_CUDA(cudaMemcpyAsync(&h, d_h, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	GPLE_CUDA(cudaStreamSynchronize(ctx->stream));
	if (h.present == 0.0)
	{
		if (!mine)
		{
			*model = nullptr;
		}
		return;
	}
	gple_model* m = mine ? *model : new gple_model();
	const size_t nn = size_t(h.n), Np = size_t(h.Np);
	if (!mine)
	{
		m->owner = ctx;
		m->is_complex = int(h.is_complex);
		m->N = size_t(h.N);
		m->Np = int(h.Np);
		m->n = int(h.n);
		m->flags = unsigned(h.flags);
		m->rescale = h.rescale;
		m->prior = h.prior;
		std::memcpy(m->theta, h.theta, sizeof(h.theta));
		m->X = static_cast<double*>(ctx->pool.alloc(2 * Np * sizeof(double)));
		m->W = static_cast<double*>(ctx->pool.alloc(nn * nn * sizeof(double)));
		m->v = static_cast<double*>(ctx->pool.alloc(nn * sizeof(double)));
		m->label = static_cast<double*>(ctx->pool.alloc(nn * sizeof(double)));
		m->kinv_diag = static_cast<double*>(ctx->pool.alloc(3 * nn * sizeof(double)));
	}
	check_nccl(n.GroupStart(), "ncclGroupStart");
	check_nccl(n.Broadcast(m->X, m->X, 2 * Np, ncclDouble, root, c, ctx->stream), "ncclBroadcast(X)");
	check_nccl(n.Broadcast(m->W, m->W, nn * nn, ncclDouble, root, c, ctx->stream), "ncclBroadcast(W)");
	check_nccl(n.Broadcast(m->v, m->v, nn, ncclDouble, root, c, ctx->stream), "ncclBroadcast(v)");
	check_nccl(n.Broadcast(m->label, m->label, nn, ncclDouble, root, c, ctx->stream), "ncclBroadcast(label)");
	check_nccl(n.Broadcast(m->kinv_diag, m->kinv_diag, 3 * nn, ncclDouble, root, c, ctx->stream), "ncclBroadcast(kinv_diag)");
	check_nccl(n.GroupEnd(), "ncclGroupEnd");
	*model = m;
}

/// evolve() over point sets that are block-partitioned over the ranks: every rank moves its own block of each element,
/// then the blocks are all-gathered in place, so that every rank leaves with the full evolved sets.
void evolve_sharded_device(gple_ctx* ctx, int pes_model, const gple_model* const models[3], double* d_pts[3], const size_t totals[3], double mass, double dt)
{
	double* local[3];
	size_t counts[3];
	for (int e = 0; e < 3; e++)
	{
		size_t lo, hi;
		partition(totals[e], ctx->comm_rank, ctx->comm_size, lo, hi);
		local[e] = d_pts[e] != nullptr ? d_pts[e] + 4 * lo : nullptr;
		counts[e] = hi - lo;
	}
	evolve_device(ctx, pes_model, models, local, counts, mass, dt);
	for (int e = 0; e < 3; e++)
	{
		allgather_blocks(ctx, d_pts[e], totals[e], 4);
	}
}
} // namespace gple
