// comm_host.cuh -- the one exchange step of the path, inside the library: per-chain partials of the shards are combined over
// NCCL (NVLink / NVSwitch), stream-ordered behind the reduction kernel that produced them.
//
// The exchange is an ALL-GATHER of every shard's small partial block followed by a local combination in rank order
// (comm_combine_kernel): one collective per batched call, and -- unlike an all-reduce, whose summation order is NCCL's choice
// of algorithm and channel count -- bit-identical results on every rank, from run to run and for any NCCL version.
// Sizes: PopPK 3 C doubles per rank (1.5 KB at 64 chains), cell_population C (2 T + 1) doubles (13 KB at 16 chains x 50
// timepoints): latency-bound, so what matters is that it is enqueued on the same stream right behind the producer.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): a host that never shards over several GPUs does not need it, and a
// process that already carries an NCCL (PyTorch bundles one under the same soname) shares that copy instead of loading a second.
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <vector>

#include "common_host.cuh"

namespace bcm3b200 {

struct NcclApi {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	std::string why; // why loading failed

	static NcclApi& get()
	{
		static NcclApi api;
		static std::once_flag once;
		std::call_once(once, [] { api.load(); });
		return api;
	}
	bool ok() const { return lib != nullptr; }

private:
	void load()
	{
		// Order: an explicit override; the copy this process already carries (PyTorch bundles one under the same soname, and the
		// dynamic loader resolves a later DT_NEEDED "libnccl.so.2" to whatever object of that soname is loaded first -- loading
		// the system library ahead of `import torch` would hand torch a libnccl it was not built against); the system library.
		if (const char* over = getenv("BCM3B200_NCCL_LIB")) {
			if (*over) {
				lib = dlopen(over, RTLD_NOW | RTLD_GLOBAL);
				if (!lib) why = dlerror();
			}
		}
		if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
		for (const char* n : { "libnccl.so.2", "libnccl.so" }) {
			if (lib) break;
			lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
			if (!lib) why = dlerror();
		}
		if (!lib) return;
		bool all = true;
		auto sym = [&](const char* name) {
			void* p = dlsym(lib, name);
			if (!p) {
				all = false;
				why = std::string("missing symbol ") + name;
			}
			return p;
		};
		GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
		CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
		CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
		CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
		AllGather = (decltype(AllGather))sym("ncclAllGather");
		GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
		GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
		GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
		if (!all) lib = nullptr;
	}
};

#define NCCL_TRY(expr)                                                                                                              \
	do {                                                                                                                            \
		ncclResult_t r_ = (expr);                                                                                                   \
		if (r_ != ncclSuccess) return bcm3b200::fail(BCM3B200_ERR_CUDA, "%s failed: %s", #expr, NcclApi::get().GetErrorString(r_)); \
	} while (0)

// gathered [W][n] -> out [n]: entries [0, n_sum) are added in rank order, entries [n_sum, n) take the minimum over the ranks
__global__ void comm_combine_kernel(const double* __restrict__ gathered, int W, int n, int n_sum, double* __restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double v = gathered[i];
	for (int r = 1; r < W; r++) {
		const double x = gathered[(long long)r * n + i];
		v = (i < n_sum) ? v + x : fmin(v, x);
	}
	out[i] = v;
}

// One rank's end of a communicator (one process per GPU), or one device's end inside a single process that drives several.
struct CommEnd {
	ncclComm_t comm = nullptr;
	int world = 1, rank = 0;
	DevBuf<double> gathered;
	~CommEnd()
	{
		if (comm && NcclApi::get().ok()) NcclApi::get().CommDestroy(comm);
	}
	bool active() const { return comm != nullptr && world > 1; }
	// d_local [n] of every rank -> d_out [n] on every rank, on `stream`; call between GroupStart/GroupEnd when one thread drives several ends
	int gather(const double* d_local, size_t n, cudaStream_t stream)
	{
		CUDA_TRY(gathered.ensure((size_t)world * n));
		NCCL_TRY(NcclApi::get().AllGather(d_local, gathered.p, n, ncclDouble, comm, stream));
		return BCM3B200_OK;
	}
	int combine(size_t n, size_t n_sum, double* d_out, cudaStream_t stream)
	{
		comm_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(gathered.p, world, (int)n, (int)n_sum, d_out);
		CUDA_TRY(cudaGetLastError());
		return BCM3B200_OK;
	}
	int gather_combine(const double* d_local, size_t n, size_t n_sum, double* d_out, cudaStream_t stream)
	{
		int rc = gather(d_local, n, stream);
		if (rc != BCM3B200_OK) return rc;
		return combine(n, n_sum, d_out, stream);
	}
};

} // namespace bcm3b200
