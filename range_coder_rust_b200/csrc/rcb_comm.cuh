// rcb_comm.cuh -- the hot path's only exchange step, behind the C ABI: one all-reduce (sum) of the
// K-entry u64 symbol-count table over NCCL (NVLink 5 / NVSwitch) when a static model is shared by
// chunks that live on several GPUs (the reference's caller owns the histogram loop,
// examples/sample_impl.rs:77-81, and would sum its per-GPU tables exactly here).  No payload byte ever
// crosses GPUs.  Included by rcb_api.cu (same translation unit: it needs rcb_ctx).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 on first use) instead of at link time: a host
// process that already carries an NCCL -- torch ships its own copy under the same soname -- keeps
// exactly one copy in the process, and a host that never shards over GPUs needs no NCCL at all.
// Only the types and enums come from <nccl.h>.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    const char* names[] = {getenv("RCB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return nullptr;
#define RCB_NCCL_SYM(field, sym)                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
    if (!api.field) return nullptr;
    RCB_NCCL_SYM(GetVersion, "ncclGetVersion")
    RCB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    RCB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    RCB_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    RCB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    RCB_NCCL_SYM(AllReduce, "ncclAllReduce")
    RCB_NCCL_SYM(GroupStart, "ncclGroupStart")
    RCB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    RCB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef RCB_NCCL_SYM
    api.ok = true;
    return &api;
}

}  // namespace

struct rcb_comm {
    ncclComm_t comm = nullptr;
    int n_ranks = 0, rank = 0, device = 0;
    ncclResult_t last = ncclSuccess;
};

static_assert(RCB_UNIQUE_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");

extern "C" int rcb_comm_unique_id(uint8_t* id) {
    if (!id) return RCB_ERR_INVALID_ARGUMENT;
    NcclApi* n = nccl_api();
    if (!n) return RCB_ERR_NCCL;
    ncclUniqueId u;
    if (n->GetUniqueId(&u) != ncclSuccess) return RCB_ERR_NCCL;
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return RCB_OK;
}

extern "C" int rcb_comm_init_rank(rcb_ctx* c, const uint8_t* id, int n_ranks, int rank, rcb_comm** out) {
    if (!c || !id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return RCB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    NcclApi* n = nccl_api();
    if (!n) return RCB_ERR_NCCL;
    ON_DEVICE(c);
    rcb_comm* k = new (std::nothrow) rcb_comm();
    if (!k) return RCB_ERR_INVALID_ARGUMENT;
    ncclUniqueId u;
    memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
    k->n_ranks = n_ranks;
    k->rank = rank;
    k->device = c->device;
    k->last = n->CommInitRank(&k->comm, n_ranks, u, rank);
    if (k->last != ncclSuccess) {
        delete k;
        return RCB_ERR_NCCL;
    }
    *out = k;
    return RCB_OK;
}

extern "C" int rcb_comm_init_all(rcb_ctx* const* ctxs, int n_ctx, rcb_comm** out) {
    if (!ctxs || !out || n_ctx < 1 || n_ctx > 64) return RCB_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_ctx; i++) {
        out[i] = nullptr;
        if (!ctxs[i]) return RCB_ERR_INVALID_ARGUMENT;
        for (int j = 0; j < i; j++)
            if (ctxs[j]->device == ctxs[i]->device) return RCB_ERR_INVALID_ARGUMENT;  // one rank per GPU
    }
    NcclApi* n = nccl_api();
    if (!n) return RCB_ERR_NCCL;
    int devs[64];
    ncclComm_t comms[64];
    for (int i = 0; i < n_ctx; i++) devs[i] = ctxs[i]->device;
    int prev = -1;
    cudaGetDevice(&prev);
    const ncclResult_t r = n->CommInitAll(comms, n_ctx, devs);
    if (prev >= 0) cudaSetDevice(prev);
    if (r != ncclSuccess) return RCB_ERR_NCCL;
    for (int i = 0; i < n_ctx; i++) {
        rcb_comm* k = new (std::nothrow) rcb_comm();
        if (!k) return RCB_ERR_INVALID_ARGUMENT;
        k->comm = comms[i];
        k->n_ranks = n_ctx;
        k->rank = i;
        k->device = devs[i];
        out[i] = k;
    }
    return RCB_OK;
}

extern "C" int rcb_comm_destroy(rcb_comm* k) {
    if (!k) return RCB_OK;
    NcclApi* n = nccl_api();
    int rc = RCB_OK;
    if (n && k->comm) {
        DeviceGuard dg(k->device);
        if (n->CommDestroy(k->comm) != ncclSuccess) rc = RCB_ERR_NCCL;
    }
    delete k;
    return rc;
}

extern "C" int rcb_comm_info(const rcb_comm* k, int* n_ranks, int* rank, int* nccl_version) {
    if (!k) return RCB_ERR_INVALID_ARGUMENT;
    if (n_ranks) *n_ranks = k->n_ranks;
    if (rank) *rank = k->rank;
    if (nccl_version) {
        NcclApi* n = nccl_api();
        *nccl_version = 0;
        if (n) n->GetVersion(nccl_version);
    }
    return RCB_OK;
}

extern "C" const char* rcb_comm_last_error(const rcb_comm* k) {
    NcclApi* n = nccl_api();
    if (!n) return "NCCL library not found (libnccl.so.2; set RCB_NCCL_LIB)";
    return n->GetErrorString(k ? k->last : ncclSuccess);
}

// Sum of the ranks' count tables, in place, on the ctx's stream (no synchronisation: the model build
// that follows is ordered after it on the same stream).
extern "C" int rcb_allreduce_counts(rcb_ctx* c, rcb_comm* k, void* d_counts, uint32_t K) {
    if (!c || !k || !k->comm || !d_counts || K == 0 || K > MAX_K) return RCB_ERR_INVALID_ARGUMENT;
    if (k->device != c->device) return RCB_ERR_INVALID_ARGUMENT;
    NcclApi* n = nccl_api();
    if (!n) return RCB_ERR_NCCL;
    ON_DEVICE(c);
    k->last = n->AllReduce(d_counts, d_counts, (size_t)K, ncclUint64, ncclSum, k->comm, c->stream);
    if (k->last != ncclSuccess) return RCB_ERR_NCCL;
    c->launches++;  // NCCL's kernel, issued by this library on the path
    return RCB_OK;
}

// Single-process hosts (one thread driving every GPU, ncclCommInitAll): the per-rank calls must be
// grouped or the first one would wait for the others forever.
extern "C" int rcb_allreduce_counts_multi(rcb_ctx* const* ctxs, rcb_comm* const* comms, void* const* d_counts,
                                          uint32_t K, int n_ctx) {
    if (!ctxs || !comms || !d_counts || n_ctx < 1 || K == 0 || K > MAX_K) return RCB_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_ctx; i++)
        if (!ctxs[i] || !comms[i] || !comms[i]->comm || !d_counts[i] || comms[i]->device != ctxs[i]->device)
            return RCB_ERR_INVALID_ARGUMENT;
    NcclApi* n = nccl_api();
    if (!n) return RCB_ERR_NCCL;
    int prev = -1;
    cudaGetDevice(&prev);
    int rc = RCB_OK;
    if (n->GroupStart() != ncclSuccess) return RCB_ERR_NCCL;
    for (int i = 0; i < n_ctx; i++) {
        comms[i]->last = n->AllReduce(d_counts[i], d_counts[i], (size_t)K, ncclUint64, ncclSum, comms[i]->comm,
                                      ctxs[i]->stream);
        if (comms[i]->last != ncclSuccess) rc = RCB_ERR_NCCL;
        ctxs[i]->launches++;
    }
    if (n->GroupEnd() != ncclSuccess) rc = RCB_ERR_NCCL;
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}
