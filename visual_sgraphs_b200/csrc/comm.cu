// comm.cu — vsg_comm: one NCCL communicator per process (one process per GPU) for the two matcher paths that have a real
// exchange step: train-sharded brute-force kNN-2 (all-gather of the per-rank top-2 lists) and map-sharded
// SearchByProjection (claim-state token + all-gather of the per-rank assignments).  Extraction shards by frame and never
// touches this file.
//
// NCCL is resolved with dlopen at vsg_comm_create time ("libnccl.so.2": the copy already loaded into the process, e.g.
// PyTorch's, or the system one), so libvsg_cuda.so has no link-time dependency on it and single-GPU users never load it.
// The reference has no counterpart: it is a single-process CPU program.
#include <dlfcn.h>

#include <cstring>

#include "comm_internal.h"
#include "vsg_internal.cuh"

namespace {

// the part of nccl.h this file needs (ABI-stable since NCCL 2.0)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

bool load_nccl(NcclApi *api) {
    static NcclApi cached;
    static bool tried = false, ok = false;
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
            if ((cached.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (cached.handle) {
            auto sym = [&](const char *s) { return dlsym(cached.handle, s); };
            cached.GetUniqueId = (decltype(cached.GetUniqueId))sym("ncclGetUniqueId");
            cached.CommInitRank = (decltype(cached.CommInitRank))sym("ncclCommInitRank");
            cached.CommDestroy = (decltype(cached.CommDestroy))sym("ncclCommDestroy");
            cached.AllGather = (decltype(cached.AllGather))sym("ncclAllGather");
            cached.Send = (decltype(cached.Send))sym("ncclSend");
            cached.Recv = (decltype(cached.Recv))sym("ncclRecv");
            cached.GetErrorString = (decltype(cached.GetErrorString))sym("ncclGetErrorString");
            cached.GetVersion = (decltype(cached.GetVersion))sym("ncclGetVersion");
            ok = cached.GetUniqueId && cached.CommInitRank && cached.CommDestroy && cached.AllGather && cached.Send && cached.Recv &&
                 cached.GetErrorString;
        }
    }
    *api = cached;
    return ok;
}

}  // namespace

struct vsg_comm {
    NcclApi api;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
};

namespace vsg {

static vsg_status nccl_check(vsg_comm *c, ncclResult_t r, const char *what) {
    if (r == 0) return VSG_OK;
    set_error("%s: NCCL error %d (%s)", what, r, c->api.GetErrorString ? c->api.GetErrorString(r) : "?");
    return VSG_ERR_CUDA;
}

int comm_rank(const vsg_comm *c) { return c->rank; }
int comm_size(const vsg_comm *c) { return c->nranks; }

vsg_status comm_all_gather(vsg_comm *c, const void *send_dev, void *recv_dev, size_t bytes_per_rank, cudaStream_t stream) {
    return nccl_check(c, c->api.AllGather(send_dev, recv_dev, bytes_per_rank, ncclInt8, c->comm, stream), "ncclAllGather");
}
vsg_status comm_send(vsg_comm *c, const void *buf_dev, size_t bytes, int peer, cudaStream_t stream) {
    return nccl_check(c, c->api.Send(buf_dev, bytes, ncclInt8, peer, c->comm, stream), "ncclSend");
}
vsg_status comm_recv(vsg_comm *c, void *buf_dev, size_t bytes, int peer, cudaStream_t stream) {
    return nccl_check(c, c->api.Recv(buf_dev, bytes, ncclInt8, peer, c->comm, stream), "ncclRecv");
}

}  // namespace vsg

using namespace vsg;

extern "C" {

vsg_status vsg_comm_unique_id(uint8_t id_out[VSG_COMM_ID_BYTES]) {
    static_assert(VSG_COMM_ID_BYTES == sizeof(ncclUniqueId), "ncclUniqueId is 128 bytes");
    if (!id_out) return VSG_ERR_INVALID;
    NcclApi api;
    if (!load_nccl(&api)) { set_error("vsg_comm_unique_id: libnccl.so.2 not found (%s)", dlerror() ? dlerror() : "missing symbols"); return VSG_ERR_CUDA; }
    ncclUniqueId id;
    const ncclResult_t r = api.GetUniqueId(&id);
    if (r != 0) { set_error("ncclGetUniqueId: NCCL error %d (%s)", r, api.GetErrorString(r)); return VSG_ERR_CUDA; }
    memcpy(id_out, &id, sizeof(id));
    return VSG_OK;
}

vsg_status vsg_comm_create(const uint8_t id[VSG_COMM_ID_BYTES], int nranks, int rank, int device, vsg_comm **out) {
    if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) return VSG_ERR_INVALID;
    *out = nullptr;
    NcclApi api;
    if (!load_nccl(&api)) { set_error("vsg_comm_create: libnccl.so.2 not found (%s)", dlerror() ? dlerror() : "missing symbols"); return VSG_ERR_CUDA; }
    CK(cudaSetDevice(device));
    vsg_comm *c = new vsg_comm;
    c->api = api;
    c->rank = rank; c->nranks = nranks; c->device = device;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    const vsg_status st = nccl_check(c, api.CommInitRank(&c->comm, nranks, uid, rank), "ncclCommInitRank");
    if (st != VSG_OK) { delete c; return st; }
    *out = c;
    return VSG_OK;
}

void vsg_comm_destroy(vsg_comm *c) {
    if (!c) return;
    if (c->comm) c->api.CommDestroy(c->comm);
    delete c;
}

int vsg_comm_rank(const vsg_comm *c) { return c ? c->rank : -1; }
int vsg_comm_size(const vsg_comm *c) { return c ? c->nranks : 0; }

int vsg_comm_nccl_version(void) {
    NcclApi api;
    int v = 0;
    if (load_nccl(&api) && api.GetVersion) api.GetVersion(&v);
    return v;
}

}  // extern "C"
