// The one collective of the path: an NCCL all-reduce of the 64-byte batch-statistics vector, for hosts that
// have no NCCL binding of their own.  libnccl.so.2 is resolved at run time so that libddp.so has no link-time dependency on it
// (inside a PyTorch process the already loaded NCCL is picked up).
#include <dlfcn.h>
#include <cstring>
#include <nccl.h>
#include <mutex>
#include "ddp_common.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    });
    return api;
}

int nccl_fail(ddp_handle_s* h, ncclResult_t r, const char* what) {
    if (h) h->err = std::string(what) + ": " + nccl().GetErrorString(r);
    return DDP_ERR_CUDA;
}

static_assert(sizeof(ncclUniqueId) == 128, "ddp_comm_unique_id hands out 128 bytes");

}  // namespace

extern "C" {

int ddp_comm_unique_id(void* id128) {
    if (!id128) return DDP_ERR_INVALID;
    if (!nccl().ok) return DDP_ERR_UNSUPPORTED;
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess) return DDP_ERR_CUDA;
    memcpy(id128, &id, sizeof(id));
    return DDP_OK;
}

int ddp_comm_init(ddp_handle_t h, int32_t nranks, int32_t rank, const void* id128) {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return DDP_ERR_INVALID;
    if (!nccl().ok) { h->err = "ddp_comm_init: libnccl.so.2 not found"; return DDP_ERR_UNSUPPORTED; }
    if (h->comm) ddp_comm_destroy(h);
    if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return DDP_ERR_CUDA; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    ncclResult_t r = nccl().CommInitRank(&c, nranks, id, rank);
    if (r != ncclSuccess) return nccl_fail(h, r, "ncclCommInitRank");
    h->comm = c;
    return DDP_OK;
}

int ddp_comm_allreduce_stats_f64(ddp_handle_t h, double* stats8) {
    if (!h || !stats8) return DDP_ERR_INVALID;
    if (!h->comm) { h->err = "ddp_comm_allreduce_stats_f64: call ddp_comm_init first"; return DDP_ERR_INVALID; }
    ncclResult_t r = nccl().AllReduce(stats8, stats8, 8, ncclDouble, ncclSum, (ncclComm_t)h->comm, h->stream);
    if (r != ncclSuccess) return nccl_fail(h, r, "ncclAllReduce");
    h->launches++;
    return DDP_OK;
}

int ddp_comm_destroy(ddp_handle_t h) {
    if (!h) return DDP_ERR_INVALID;
    if (h->comm && nccl().ok) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        nccl().CommDestroy((ncclComm_t)h->comm);
    }
    h->comm = nullptr;
    return DDP_OK;
}

}  // extern "C"
