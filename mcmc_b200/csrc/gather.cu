// Assembly of draws_out across the GPUs of one box (BASELINE north_star: "chains shard embarrassingly across
// the 8 GPUs with an NCCL/NVLink all-gather only to assemble draws_out") behind the C ABI, so that C/C++ callers — not only
// the Python plumbing — get it: every rank passes its chain-major block [count_r][n_keep][n_dim] and receives the full
// [n_chains_total][n_keep][n_dim] array in rank order.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the host process already loaded — e.g. PyTorch's — is the one
// that is used), so libmcmc_b200.so itself has no NCCL link dependency and single-GPU users need no NCCL at all.
// Equal shards -> one ncclAllGather; unequal shards -> one grouped ncclBroadcast per rank (no padding, no second copy).
// There is nothing to fuse it with: sampling has no exchange step and a persistent kernel's rows are final only when it
// ends; the gather is NVLink-bandwidth-bound (each rank receives (N-1)/N of the whole array) and is reported separately.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "engine.h"

namespace mcmcb200
{

// the subset of nccl.h that is used (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // 0 = ncclSuccess
constexpr int ncclDouble = 8;

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (!api.h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.h, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.h, "ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.h, "ncclAllGather"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(api.h, "ncclBroadcast"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.h, "ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.h, "ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.h, "ncclGetErrorString"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.Broadcast || !api.GroupStart || !api.GroupEnd) {
            dlclose(api.h);
            api.h = nullptr;
        }
    });
    if (!api.h) {
        set_error("NCCL is not available (dlopen libnccl.so.2 failed): multi-GPU assembly of draws_out needs it");
        return nullptr;
    }
    return &api;
}

struct Comm {
    ncclComm_t comm;
    int world, rank, device;
};

#define MCMCB200_NCCL_TRY(api, expr)                                                                             \
    do {                                                                                                         \
        ncclResult_t _r = (expr);                                                                                \
        if (_r != 0) {                                                                                           \
            set_error("%s failed: %s", #expr, (api)->GetErrorString ? (api)->GetErrorString(_r) : "NCCL error"); \
            return MCMCB200_ERR_CUDA;                                                                            \
        }                                                                                                        \
    } while (0)

}  // namespace mcmcb200

using namespace mcmcb200;

extern "C" {

int mcmcb200_comm_unique_id(void* id_out, size_t id_bytes)
{
    NcclApi* api = nccl_api();
    if (!api) return MCMCB200_ERR_UNSUPPORTED;
    if (!id_out || id_bytes < sizeof(ncclUniqueId)) { set_error("comm_unique_id: need a %zu-byte buffer", sizeof(ncclUniqueId)); return MCMCB200_ERR_INVALID_ARG; }
    ncclUniqueId id;
    MCMCB200_NCCL_TRY(api, api->GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof(id));
    return MCMCB200_OK;
}

int mcmcb200_comm_init(const void* id, size_t id_bytes, int32_t world_size, int32_t rank, int32_t device, void** comm_out)
{
    NcclApi* api = nccl_api();
    if (!api) return MCMCB200_ERR_UNSUPPORTED;
    if (!id || id_bytes < sizeof(ncclUniqueId) || !comm_out || world_size < 1 || rank < 0 || rank >= world_size) {
        set_error("comm_init: bad arguments");
        return MCMCB200_ERR_INVALID_ARG;
    }
    int prev = 0;
    MCMCB200_CUDA_TRY(cudaGetDevice(&prev));
    if (device >= 0) MCMCB200_CUDA_TRY(cudaSetDevice(device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    const ncclResult_t r = api->CommInitRank(&c, world_size, uid, rank);
    if (device >= 0) cudaSetDevice(prev);
    if (r != 0) { set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "NCCL error"); return MCMCB200_ERR_CUDA; }
    *comm_out = new Comm{c, world_size, rank, device >= 0 ? device : prev};
    return MCMCB200_OK;
}

int mcmcb200_comm_destroy(void* comm)
{
    NcclApi* api = nccl_api();
    if (!api || !comm) return MCMCB200_OK;
    Comm* c = static_cast<Comm*>(comm);
    api->CommDestroy(c->comm);
    delete c;
    return MCMCB200_OK;
}

int mcmcb200_allgather_draws(void* comm, const double* local_dev, const int64_t* chains_per_rank, int64_t n_keep, int32_t n_dim,
                             double* full_dev, void* stream)
{
    NcclApi* api = nccl_api();
    if (!api) return MCMCB200_ERR_UNSUPPORTED;
    Comm* c = static_cast<Comm*>(comm);
    if (!c || !local_dev || !chains_per_rank || !full_dev || n_keep < 0 || n_dim <= 0) { set_error("allgather_draws: bad arguments"); return MCMCB200_ERR_INVALID_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t row = (size_t)n_keep * (size_t)n_dim;
    bool equal = true;
    for (int r = 1; r < c->world; ++r) equal = equal && chains_per_rank[r] == chains_per_rank[0];
    if (equal) {
        MCMCB200_NCCL_TRY(api, api->AllGather(local_dev, full_dev, (size_t)chains_per_rank[0] * row, ncclDouble, c->comm, st));
        return MCMCB200_OK;
    }
    MCMCB200_NCCL_TRY(api, api->GroupStart());
    size_t off = 0;
    for (int r = 0; r < c->world; ++r) {
        const size_t n = (size_t)chains_per_rank[r] * row;
        const ncclResult_t rr = api->Broadcast(r == c->rank ? local_dev : full_dev + off, full_dev + off, n, ncclDouble, r, c->comm, st);
        if (rr != 0) { api->GroupEnd(); set_error("ncclBroadcast failed: %s", api->GetErrorString ? api->GetErrorString(rr) : "NCCL error"); return MCMCB200_ERR_CUDA; }
        off += n;
    }
    MCMCB200_NCCL_TRY(api, api->GroupEnd());
    return MCMCB200_OK;
}

}  // extern "C"
