// NCCL plumbing for the row-block-sharded genome-wide ICE: one in-stream allreduce of the
// marginal vector per iteration over NVLink/NVSwitch.  libnccl.so.2 is resolved at run time with
// dlopen (the copy torch already loaded in the process, or the system one), so the library has
// no link-time NCCL dependency and loads on machines without NCCL.
#include <dlfcn.h>
#include <string.h>
#include "hc_common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat64 = 8;   // ncclDataType_t::ncclFloat64
constexpr int kNcclSum = 0;       // ncclRedOp_t::ncclSum

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    if (a.handle || a.ok) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
    if (!a.handle) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
    return a;
}

int nccl_fail(const char* what, ncclResult_t r) {
    hc_set_error("%s: %s", what, api().GetErrorString ? api().GetErrorString(r) : "NCCL error");
    return HC_ERR_NCCL;
}

}  // namespace

int hc_nccl_allreduce_f64(void* comm, const double* send, double* recv, size_t count, cudaStream_t s) {
    NcclApi& a = api();
    if (!a.ok) { hc_set_error("libnccl.so.2 not available"); return HC_ERR_NCCL; }
    ncclResult_t r = a.AllReduce(send, recv, count, kNcclFloat64, kNcclSum, (ncclComm_t)comm, s);
    return r == 0 ? HC_OK : nccl_fail("ncclAllReduce", r);
}

extern "C" int hc_nccl_available(void) { return api().ok ? 1 : 0; }

extern "C" int hc_nccl_unique_id(void* h_id128) {
    NcclApi& a = api();
    if (!a.ok) { hc_set_error("libnccl.so.2 not available"); return HC_ERR_NCCL; }
    ncclUniqueId id;
    ncclResult_t r = a.GetUniqueId(&id);
    if (r != 0) return nccl_fail("ncclGetUniqueId", r);
    memcpy(h_id128, &id, sizeof(id));
    return HC_OK;
}

extern "C" int hc_nccl_comm_init(const void* h_id128, int32_t nranks, int32_t rank, void** h_comm) {
    NcclApi& a = api();
    if (!a.ok) { hc_set_error("libnccl.so.2 not available"); return HC_ERR_NCCL; }
    HC_REQUIRE(h_comm != nullptr && nranks > 0 && rank >= 0 && rank < nranks, "rank / nranks");
    ncclUniqueId id;
    memcpy(&id, h_id128, sizeof(id));
    ncclComm_t c = nullptr;
    ncclResult_t r = a.CommInitRank(&c, nranks, id, rank);
    if (r != 0) return nccl_fail("ncclCommInitRank", r);
    *h_comm = (void*)c;
    return HC_OK;
}

extern "C" int hc_nccl_comm_destroy(void* comm) {
    NcclApi& a = api();
    if (!a.ok || !comm) return HC_OK;
    ncclResult_t r = a.CommDestroy((ncclComm_t)comm);
    return r == 0 ? HC_OK : nccl_fail("ncclCommDestroy", r);
}

extern "C" int hc_nccl_allreduce_sum_f64(void* comm, double* buf, int64_t count, void* stream) {
    return hc_nccl_allreduce_f64(comm, buf, buf, (size_t)count, (cudaStream_t)stream);
}
