// C ABI: the communicator of the multi-GPU paths (ising_comm_*).  One process per GPU; the
// collectives the hot path needs - the all-gather of replica energies per tempering swap step,
// the halo send/recv of a strip decomposition - are NCCL calls on the context's stream, made by
// this library, so that any host language that can call C reaches them.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that has already loaded a NCCL
// (PyTorch bundles its own) shares that copy instead of mapping a second one, and single-GPU
// users need no NCCL at all.  Only the unique id has to travel between the ranks; the host layer
// broadcasts it with whatever it has (torch.distributed, MPI, a file).
#include <dlfcn.h>
#include <nccl.h>

#include "api_internal.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            const char* why = dlerror();   // (a second call would return NULL: the message is consumed)
            api.err = std::string("cannot load NCCL: ") + (why ? why : "libnccl.so.2 not found");
            return;
        }
        bool ok = true;
        auto sym = [&](const char* n) {
            void* p = dlsym(api.handle, n);
            if (!p) { ok = false; api.err = std::string("NCCL symbol missing: ") + n; }
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return api;
}

}  // namespace

struct ising_comm {
    ising_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

#define NCCL_TRY(ctx, call)                                                                    \
    do {                                                                                       \
        ncclResult_t _r = (call);                                                              \
        if (_r != ncclSuccess)                                                                 \
            return fail((ctx), ISING_E_CUDA, "%s failed: %s", #call, nccl_api().GetErrorString(_r)); \
    } while (0)

extern "C" int ising_comm_unique_id(uint8_t* out, uint64_t capacity) {
    if (!out || capacity < ISING_COMM_ID_BYTES) return fail(nullptr, ISING_E_INVALID, "id buffer too small");
    NcclApi& n = nccl_api();
    if (!n.handle) return fail(nullptr, ISING_E_UNSUPPORTED, "%s", n.err.c_str());
    static_assert(sizeof(ncclUniqueId) <= ISING_COMM_ID_BYTES, "unique id size");
    ncclUniqueId id;
    NCCL_TRY(nullptr, n.GetUniqueId(&id));
    memset(out, 0, ISING_COMM_ID_BYTES);
    memcpy(out, &id, sizeof id);
    return ISING_OK;
}

extern "C" int ising_comm_create(ising_ctx* ctx, const uint8_t* id_bytes, int rank, int world,
                                 ising_comm** out) {
    CtxLock _lk(ctx);
    if (!ctx || !id_bytes || !out) return fail(ctx, ISING_E_INVALID, "ctx/id/out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, ISING_E_INVALID, "need 0 <= rank < world");
    NcclApi& n = nccl_api();
    if (!n.handle) return fail(ctx, ISING_E_UNSUPPORTED, "%s", n.err.c_str());
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    std::unique_ptr<ising_comm> c(new ising_comm);
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    NCCL_TRY(ctx, n.CommInitRank(&c->comm, world, id, rank));
    *out = c.release();
    ctx_retain(ctx);
    return ISING_OK;
}

extern "C" void ising_comm_destroy(ising_comm* c) {
    if (!c) return;
    struct Release { ising_ctx* c; ~Release() { ctx_release(c); } } _rel{c->ctx};   // after the lock is gone
    CtxLock _lk(c->ctx);
    if (c->comm) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
        nccl_api().CommDestroy(c->comm);
    }
    delete c;
}

extern "C" int ising_comm_info(const ising_comm* c, int* rank, int* world) {
    if (!c) return fail(nullptr, ISING_E_INVALID, "comm is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return ISING_OK;
}

// ---- internal helpers used by api_pt.cu / api_strip.cu ---------------------------------------
int comm_rank(const ising_comm* c) { return c ? c->rank : 0; }
int comm_world(const ising_comm* c) { return c ? c->world : 1; }

// recv[r * bytes .. (r+1) * bytes) = rank r's send[0 .. bytes)
int comm_allgather_bytes(ising_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    NCCL_TRY(c->ctx, nccl_api().AllGather(send, recv, bytes, ncclUint8, c->comm, st));
    return ISING_OK;
}

int comm_allreduce_sum_u64(ising_comm* c, const void* send, void* recv, size_t count, cudaStream_t st) {
    NCCL_TRY(c->ctx, nccl_api().AllReduce(send, recv, count, ncclUint64, ncclSum, c->comm, st));
    return ISING_OK;
}

// Periodic ring: send_up -> rank-1, send_down -> rank+1, recv_up <- rank-1, recv_down <- rank+1
// (bytes each), one NCCL group.  With two ranks both neighbours are the same peer: NCCL matches
// sends and receives between a pair in issue order, so the peer's "down" message (sent second)
// must be received second.
int comm_ring_exchange(ising_comm* c, const void* send_up, const void* send_down, void* recv_up,
                       void* recv_down, size_t bytes, cudaStream_t st) {
    NcclApi& n = nccl_api();
    const int up = (c->rank + c->world - 1) % c->world, down = (c->rank + 1) % c->world;
    NCCL_TRY(c->ctx, n.GroupStart());
    ncclResult_t r1 = n.Send(send_up, bytes, ncclUint8, up, c->comm, st);
    ncclResult_t r2 = n.Send(send_down, bytes, ncclUint8, down, c->comm, st);
    // the peer's first send goes "up" (towards lower rank): it is what arrives from `down`
    ncclResult_t r3 = n.Recv(recv_down, bytes, ncclUint8, down, c->comm, st);
    ncclResult_t r4 = n.Recv(recv_up, bytes, ncclUint8, up, c->comm, st);
    ncclResult_t r5 = n.GroupEnd();
    for (ncclResult_t r : {r1, r2, r3, r4, r5})
        if (r != ncclSuccess) return fail(c->ctx, ISING_E_CUDA, "NCCL ring exchange failed: %s", n.GetErrorString(r));
    return ISING_OK;
}
