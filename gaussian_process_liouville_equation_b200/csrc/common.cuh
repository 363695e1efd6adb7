// Shared declarations of the B200 (sm_100a) GPR-MQCLE library: context, device buffers, status plumbing.
#pragma once
#include <mutex>
#include "../../include/gple_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace gple
{
constexpr int TILE = 128; // all device matrices are padded to a multiple of this

inline size_t round_up(size_t n, size_t m)
{
	return (n + m - 1) / m * m;
}

struct CudaError
{
	cudaError_t code;
	const char* what;
	const char* file;
	int line;
};

#define GPLE_CUDA(expr)                                                         \
	do                                                                          \
	{                                                                           \
		const cudaError_t e__ = (expr);                                         \
		if (e__ != cudaSuccess)                                                 \
		{                                                                       \
			throw ::gple::CudaError{e__, #expr, __FILE__, __LINE__};            \
		}                                                                       \
	} while (0)

struct ArgError
{
	const char* what;
};

/// Grow-only named scratch buffers: no cudaMalloc on the steady-state path.
struct Workspace
{
	std::map<std::string, std::pair<void*, size_t>> bufs;
	void* get(const std::string& name, size_t bytes)
	{
		auto& e = bufs[name];
		if (e.second < bytes)
		{
			if (e.first != nullptr)
			{
				GPLE_CUDA(cudaFree(e.first));
				e.first = nullptr;
				e.second = 0;
			}
			GPLE_CUDA(cudaMalloc(&e.first, bytes));
			e.second = bytes;
		}
		return e.first;
	}
	template <typename T>
	T* get(const std::string& name, size_t count)
	{
		return static_cast<T*>(get(name, count * sizeof(T)));
	}
	void release()
	{
		for (auto& kv : bufs)
		{
			if (kv.second.first != nullptr)
			{
				cudaFree(kv.second.first);
			}
		}
		bufs.clear();
	}
};
/// Size-bucketed free list for the per-model device buffers: a model created and destroyed every time step (or every
/// optimiser evaluation) must not pay cudaMalloc / cudaFree (device-wide syncs, and peer mapping once NCCL is up).
struct BlockPool
{
	std::multimap<size_t, void*> free_blocks;
	std::map<void*, size_t> sizes;
	std::mutex lock; // a model is freed by whichever thread drops the last reference to it
	void* alloc(size_t bytes)
	{
		const std::lock_guard<std::mutex> guard(lock);
		bytes = round_up(bytes == 0 ? 1 : bytes, 512);
		auto it = free_blocks.lower_bound(bytes);
		if (it != free_blocks.end() && it->first <= bytes + bytes / 4)
		{
			void* p = it->second;
			free_blocks.erase(it);
			return p;
		}
		void* p = nullptr;
		GPLE_CUDA(cudaMalloc(&p, bytes));
		sizes[p] = bytes;
		return p;
	}
	void free(void* p)
	{
		if (p != nullptr)
		{
			const std::lock_guard<std::mutex> guard(lock);
			free_blocks.emplace(sizes.at(p), p);
		}
	}
	void release()
	{
		for (auto& kv : sizes)
		{
			cudaFree(kv.first);
		}
		sizes.clear();
		free_blocks.clear();
	}
};
} // namespace gple

struct gple_ctx
{
	int device = 0;
	cudaStream_t own_stream = nullptr;
	// look-ahead of the factorisation: the bulk of a trailing update runs here while the next leaf runs on `stream`
	cudaMemPool_t mempool = nullptr; // staging buffers of host-pointer arguments (DeviceArray)
	cudaStream_t aux_stream = nullptr;
	cudaEvent_t ev_panel = nullptr, ev_bulk = nullptr;
	cudaStream_t stream = nullptr;
	unsigned long long launches = 0;
	std::string last_error;
	gple::Workspace ws;
	gple::BlockPool pool; // device buffers of the element models
	// multi-GPU: NCCL communicator of this context (ncclComm_t; NULL = single GPU), see comm.cu
	void* comm = nullptr;
	int comm_rank = 0, comm_size = 1;
	double* h_pinned = nullptr; // small pinned staging area for scalar read-backs
	size_t h_pinned_count = 0;
	int num_sms = 148;
	// bound-gated variance (GPLE_OPT_GATED_VARIANCE) and its statistics
	bool gated_variance = true;
	bool refine_solution = true; // GPLE_OPT_REFINE_SOLUTION
	// CUDA graphs of the factorisation (potrf + trtri: ~100 short dependent launches on two streams), one per (size, buffers)
	bool factorise_graphs = true; // GPLE_OPT_FACTORISE_GRAPHS
	struct FactoriseGraph
	{
		int n = 0;
		const void *A = nullptr, *W = nullptr, *info = nullptr, *dinv = nullptr, *T = nullptr;
		int potrf_flat = 0;
		cudaGraphExec_t exec = nullptr;
		unsigned long long launches = 0, uses = 0;
	};
	std::vector<FactoriseGraph> factorise_graph_cache;
	cudaEvent_t ev_graph = nullptr;
	unsigned long long graph_replays = 0, graph_captures = 0;
	int gate_stage_tiles = -1;	  // GPLE_OPT_GATE_STAGE_TILES (-1: automatic)
	int gate_stage_tiles_im = -1; // GPLE_OPT_GATE_STAGE_TILES_IM (-1: automatic)
	std::vector<int> gate_schedule[2]; // gple_ctx_set_gate_schedule: (re_end, im_end) pairs, [0] real / [1] complex element; empty: automatic
	int gate_stage2_tiles = -1;	  // GPLE_OPT_GATE_STAGE2_TILES (with an explicit stage: -1 = five eighths of the blocks)
	unsigned long long gate_rows_total = 0, gate_rows_variance = 0, gate_rows_zero = 0, gate_rows_stage_b = 0;
	// optional per-kernel event timing (gple_profile_*)
	bool prof_on = false;
	struct ProfSlot
	{
		std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
		double work = 0.0;
		unsigned long long launches = 0;
	} prof[4];
};

namespace gple
{
/// RAII view of a caller array on the device: copies a host array in (and optionally back out).
template <typename T>
struct DeviceArray
{
	gple_ctx* ctx;
	T* dev = nullptr;
	T* host = nullptr; // non-null when the caller passed a host pointer
	size_t count = 0;
	bool owned = false;
	bool write_back = false;

	DeviceArray(gple_ctx* c, const T* ptr, size_t n, bool is_output): ctx(c), count(n), write_back(is_output)
	{
		if (ptr == nullptr || n == 0)
		{
			return;
		}
		cudaPointerAttributes at{};
		const cudaError_t e = cudaPointerGetAttributes(&at, ptr);
		if (e != cudaSuccess)
		{
			cudaGetLastError();
		}
		if (e == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged))
		{
			dev = const_cast<T*>(ptr);
			return;
		}
		host = const_cast<T*>(ptr);
		if (ctx->mempool != nullptr)
		{
			GPLE_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&dev), n * sizeof(T), ctx->mempool, ctx->stream));
		}
		else
		{
			GPLE_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dev), n * sizeof(T), ctx->stream));
		}
		owned = true;
		if (!is_output)
		{
			GPLE_CUDA(cudaMemcpyAsync(dev, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
		}
	}
	/// copy results back to a host caller (stream-ordered; caller syncs)
	void finish()
	{
		if (owned && write_back && host != nullptr)
		{
			GPLE_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
		}
	}
	~DeviceArray()
	{
		if (owned && dev != nullptr)
		{
			cudaFreeAsync(dev, ctx->stream);
		}
	}
	DeviceArray(const DeviceArray&) = delete;
	DeviceArray& operator=(const DeviceArray&) = delete;
	explicit operator bool() const { return dev != nullptr; }
};

#define GPLE_LAUNCH(ctx, kernel, grid, block, smem, ...)                        \
	do                                                                          \
	{                                                                           \
		kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);        \
		(ctx)->launches++;                                                      \
		GPLE_CUDA(cudaGetLastError());                                          \
	} while (0)

/// RAII event bracket around one or more launches of a profiled kernel
struct ProfScope
{
	gple_ctx* ctx;
	int slot;
	cudaEvent_t e1 = nullptr;
	ProfScope(gple_ctx* c, int s, double work, unsigned long long launches): ctx(c), slot(s)
	{
		if (!ctx->prof_on)
		{
			return;
		}
		cudaEvent_t e0;
		GPLE_CUDA(cudaEventCreate(&e0));
		GPLE_CUDA(cudaEventCreate(&e1));
		GPLE_CUDA(cudaEventRecord(e0, ctx->stream));
		ctx->prof[slot].ev.emplace_back(e0, e1);
		ctx->prof[slot].work += work;
		ctx->prof[slot].launches += launches;
	}
	~ProfScope()
	{
		if (e1 != nullptr)
		{
			cudaEventRecord(e1, ctx->stream);
		}
	}
};

// ---- parameter blocks passed by value to kernels -------------------------------------------------

/// Gaussian ARD kernel sigma_f^2 * (exp(-1/2 sum ((dx)/l)^2) + sigma_n^2 delta): gple/kernel.h:25-28
struct GaussParam
{
	double mag2;   // sigma_f^2
	double inv_lx; // 1 / l_x
	double inv_lp; // 1 / l_p
	double noise2; // sigma_n^2
};

inline GaussParam make_gauss(double mag, double lx, double lp, double noise)
{
	return GaussParam{mag * mag, 1.0 / lx, 1.0 / lp, noise * noise};
}

} // namespace gple
