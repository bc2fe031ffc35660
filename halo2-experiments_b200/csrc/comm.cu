// Comm implementations (comm.hpp) and their C ABI: b200zk_comm_* (one process per GPU, NCCL) and
// b200zk_group_* (one process driving several GPUs, the ranks are threads).
#include "context.hpp"
#include "comm.hpp"
#include <nccl.h>
#include <dlfcn.h>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

namespace b200zk {

// ------------------------------------------------------------------ ranks = threads of one process
struct LocalState {
    int world = 0;
    std::mutex mu;
    std::condition_variable cv;
    int count = 0;
    uint64_t gen = 0;
    bool aborted = false;
    std::vector<char*> win_base;
    std::vector<size_t> win_bytes;
    std::vector<cudaEvent_t> ev_ready, ev_done;          // per rank, on that rank's device
    std::vector<std::vector<uint8_t>> host;              // per rank allgather_host slot
    // false when a rank aborted (now or earlier): the caller gives up instead of waiting for ever
    bool barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return false;
        const uint64_t g = gen;
        if (++count == world) { count = 0; ++gen; cv.notify_all(); return true; }
        cv.wait(lk, [&] { return gen != g || aborted; });
        return gen != g;
    }
    void abort_all() {
        std::lock_guard<std::mutex> lk(mu);
        aborted = true;
        cv.notify_all();
    }
    void reset() {
        std::lock_guard<std::mutex> lk(mu);
        aborted = false; count = 0;
    }
};

struct LocalComm : Comm {
    std::shared_ptr<LocalState> s;
    int32_t set_window(b200zk_ctx*, void* base, size_t bytes) override {
        s->win_base[rank] = (char*)base; s->win_bytes[rank] = bytes;
        return B200ZK_OK;
    }
    int32_t share(b200zk_ctx* ctx, const CommPiece* pieces, size_t count, cudaStream_t st) override {
        if (world == 1) return B200ZK_OK;
        ZK_CUDA(ctx, cudaEventRecord(s->ev_ready[rank], st));
        if (!s->barrier()) return fail(ctx, B200ZK_ECUDA, "comm", "a peer rank failed");
        char* mine = s->win_base[rank];
        std::vector<char> waited(world, 0);
        for (size_t i = 0; i < count; ++i) {
            const CommPiece& p = pieces[i];
            if (p.owner == rank || p.bytes == 0) continue;
            size_t off = (char*)p.ptr - mine;
            if ((char*)p.ptr < mine || off + p.bytes > s->win_bytes[rank] || off + p.bytes > s->win_bytes[p.owner])
                return fail(ctx, B200ZK_EINVAL, "comm", "shared buffer outside the registered window");
            if (!waited[p.owner]) { ZK_CUDA(ctx, cudaStreamWaitEvent(st, s->ev_ready[p.owner], 0)); waited[p.owner] = 1; }
            ZK_CUDA(ctx, cudaMemcpyAsync(p.ptr, s->win_base[p.owner] + off, p.bytes, cudaMemcpyDefault, st));
        }
        // nobody overwrites what it shared before every reader's copy has run
        ZK_CUDA(ctx, cudaEventRecord(s->ev_done[rank], st));
        if (!s->barrier()) return fail(ctx, B200ZK_ECUDA, "comm", "a peer rank failed");
        for (int r = 0; r < world; ++r) if (r != rank) ZK_CUDA(ctx, cudaStreamWaitEvent(st, s->ev_done[r], 0));
        return B200ZK_OK;
    }
    int32_t allgather_host(b200zk_ctx* ctx, const void* mine, size_t bytes, void* all, cudaStream_t st) override {
        if (bytes > COMM_HOST_MAX) return fail(ctx, B200ZK_EINVAL, "comm", "allgather_host payload too large");
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
        if (world == 1) { memcpy(all, mine, bytes); return B200ZK_OK; }
        memcpy(s->host[rank].data(), mine, bytes);
        if (!s->barrier()) return fail(ctx, B200ZK_ECUDA, "comm", "a peer rank failed");
        for (int r = 0; r < world; ++r) memcpy((char*)all + (size_t)r * bytes, s->host[r].data(), bytes);
        if (!s->barrier()) return fail(ctx, B200ZK_ECUDA, "comm", "a peer rank failed");
        return B200ZK_OK;
    }
    void abort() override { s->abort_all(); }
};

// ------------------------------------------------------------------ one process per GPU: NCCL
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommAbort) CommAbort = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

// The NCCL already in the process (torch's bundled copy has the same soname) or the system one.
static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.handle = h;
#define ZK_SYM(name) api.name = (decltype(api.name))dlsym(h, "nccl" #name)
        ZK_SYM(GetUniqueId); ZK_SYM(CommInitRank); ZK_SYM(CommDestroy); ZK_SYM(CommAbort); ZK_SYM(Broadcast); ZK_SYM(AllGather);
        ZK_SYM(GroupStart); ZK_SYM(GroupEnd); ZK_SYM(GetErrorString); ZK_SYM(GetVersion);
#undef ZK_SYM
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.AllGather || !api.GroupStart || !api.GroupEnd)
            api.handle = nullptr;
    });
    return api.handle ? &api : nullptr;
}

#define ZK_NCCL(ctx, expr)                                                                          \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) return fail((ctx), B200ZK_ECUDA, #expr, api->GetErrorString ? api->GetErrorString(_r) : "nccl error"); \
    } while (0)

struct NcclComm : Comm {
    NcclApi* api = nullptr;
    ncclComm_t comm = nullptr;
    char* d_stage = nullptr;            // world * COMM_HOST_MAX
    char* h_stage = nullptr;            // pinned, same size
    ~NcclComm() override {
        if (comm && api) api->CommDestroy(comm);
        if (d_stage) cudaFree(d_stage);
        if (h_stage) cudaFreeHost(h_stage);
    }
    int32_t set_window(b200zk_ctx*, void*, size_t) override { return B200ZK_OK; }
    int32_t share(b200zk_ctx* ctx, const CommPiece* pieces, size_t count, cudaStream_t st) override {
        if (world == 1 || count == 0) return B200ZK_OK;
        ZK_NCCL(ctx, api->GroupStart());
        for (size_t i = 0; i < count; ++i) {
            if (pieces[i].bytes == 0) continue;
            ncclResult_t r = api->Broadcast(pieces[i].ptr, pieces[i].ptr, pieces[i].bytes, ncclUint8, pieces[i].owner, comm, st);
            if (r != ncclSuccess) { api->GroupEnd(); return fail(ctx, B200ZK_ECUDA, "ncclBroadcast", api->GetErrorString ? api->GetErrorString(r) : ""); }
        }
        ZK_NCCL(ctx, api->GroupEnd());
        return B200ZK_OK;
    }
    int32_t allgather_host(b200zk_ctx* ctx, const void* mine, size_t bytes, void* all, cudaStream_t st) override {
        if (bytes > COMM_HOST_MAX) return fail(ctx, B200ZK_EINVAL, "comm", "allgather_host payload too large");
        if (world == 1) { ZK_CUDA(ctx, cudaStreamSynchronize(st)); memcpy(all, mine, bytes); return B200ZK_OK; }
        const size_t slot = (bytes + 15) / 16 * 16;
        memcpy(h_stage + (size_t)rank * slot, mine, bytes);
        ZK_CUDA(ctx, cudaMemcpyAsync(d_stage + (size_t)rank * slot, h_stage + (size_t)rank * slot, slot, cudaMemcpyHostToDevice, st));
        ZK_NCCL(ctx, api->AllGather(d_stage + (size_t)rank * slot, d_stage, slot, ncclUint8, comm, st));
        ZK_CUDA(ctx, cudaMemcpyAsync(h_stage, d_stage, slot * world, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
        for (int r = 0; r < world; ++r) memcpy((char*)all + (size_t)r * bytes, h_stage + (size_t)r * slot, bytes);
        return B200ZK_OK;
    }
    void abort() override {
        if (comm && api && api->CommAbort) { api->CommAbort(comm); comm = nullptr; }
    }
};

}  // namespace b200zk

using namespace b200zk;

struct b200zk_group {
    std::vector<b200zk_ctx*> ctxs;
    std::shared_ptr<LocalState> state;
};

extern "C" {

// ---- one process per GPU ---------------------------------------------------------------------
int32_t b200zk_comm_unique_id(void* id_out128) {
    if (!id_out128) return B200ZK_EINVAL;
    NcclApi* api = nccl_api();
    if (!api) return B200ZK_ENODEV;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (api->GetUniqueId(&id) != ncclSuccess) return B200ZK_ECUDA;
    memcpy(id_out128, &id, 128);
    return B200ZK_OK;
}

int32_t b200zk_ctx_comm_init(b200zk_ctx* ctx, uint32_t world, uint32_t rank, const void* id128) {
    if (!ctx || !id128 || world == 0 || rank >= world) return B200ZK_EINVAL;
    if (ctx->comm) return fail(ctx, B200ZK_EINVAL, "comm_init", "this ctx already has a communicator");
    NcclApi* api = nccl_api();
    if (!api) return fail(ctx, B200ZK_ENODEV, "comm_init", "libnccl.so.2 not found");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    NcclComm* c = new (std::nothrow) NcclComm();
    if (!c) return B200ZK_ENOMEM;
    c->api = api; c->rank = (int)rank; c->world = (int)world;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = api->CommInitRank(&c->comm, (int)world, id, (int)rank);
    if (r != ncclSuccess) { c->comm = nullptr; delete c; return fail(ctx, B200ZK_ECUDA, "ncclCommInitRank", api->GetErrorString ? api->GetErrorString(r) : ""); }
    if (cudaMalloc(&c->d_stage, (size_t)world * COMM_HOST_MAX) != cudaSuccess || cudaHostAlloc(&c->h_stage, (size_t)world * COMM_HOST_MAX, cudaHostAllocDefault) != cudaSuccess) {
        delete c;
        return fail(ctx, B200ZK_ENOMEM, "comm_init", "staging buffers");
    }
    ctx->comm = c;
    ctx->comm_owned = true;
    return B200ZK_OK;
}

int32_t b200zk_ctx_comm_destroy(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_EINVAL;
    if (ctx->comm && ctx->comm_owned) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); delete ctx->comm; }
    ctx->comm = nullptr; ctx->comm_owned = false;
    return B200ZK_OK;
}

uint32_t b200zk_ctx_comm_world(const b200zk_ctx* ctx) { return ctx && ctx->comm ? (uint32_t)ctx->comm->world : 1u; }
uint32_t b200zk_ctx_comm_rank(const b200zk_ctx* ctx) { return ctx && ctx->comm ? (uint32_t)ctx->comm->rank : 0u; }

// ---- one process, several GPUs -----------------------------------------------------------------
void b200zk_group_destroy(b200zk_group* g) {
    if (!g) return;
    for (size_t r = 0; r < g->ctxs.size(); ++r) {
        b200zk_ctx* ctx = g->ctxs[r];
        if (!ctx) continue;
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        delete ctx->comm; ctx->comm = nullptr;
        if (g->state) {
            if (g->state->ev_ready[r]) cudaEventDestroy(g->state->ev_ready[r]);
            if (g->state->ev_done[r]) cudaEventDestroy(g->state->ev_done[r]);
        }
        b200zk_ctx_destroy(ctx);
    }
    delete g;
}

int32_t b200zk_group_create(const int32_t* devices, uint32_t n, b200zk_group** out) {
    if (!devices || !out || n == 0 || n > 64) return B200ZK_EINVAL;
    *out = nullptr;
    b200zk_group* g = new (std::nothrow) b200zk_group();
    if (!g) return B200ZK_ENOMEM;
    g->state = std::make_shared<LocalState>();
    LocalState& s = *g->state;
    s.world = (int)n;
    s.win_base.assign(n, nullptr); s.win_bytes.assign(n, 0);
    s.ev_ready.assign(n, nullptr); s.ev_done.assign(n, nullptr);
    s.host.assign(n, std::vector<uint8_t>(COMM_HOST_MAX));
    g->ctxs.assign(n, nullptr);
    for (uint32_t r = 0; r < n; ++r) {
        int32_t rc = b200zk_ctx_create(devices[r], &g->ctxs[r]);
        if (rc != B200ZK_OK) { b200zk_group_destroy(g); return rc; }
        if (cudaEventCreateWithFlags(&s.ev_ready[r], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.ev_done[r], cudaEventDisableTiming) != cudaSuccess) { b200zk_group_destroy(g); return B200ZK_ECUDA; }
        LocalComm* c = new (std::nothrow) LocalComm();
        if (!c) { b200zk_group_destroy(g); return B200ZK_ENOMEM; }
        c->rank = (int)r; c->world = (int)n; c->s = g->state;
        g->ctxs[r]->comm = c;
    }
    // direct NVLink copies between the devices of the group
    for (uint32_t a = 0; a < n; ++a)
        for (uint32_t b = 0; b < n; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can) {
                cudaSetDevice(devices[a]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e != cudaSuccess) cudaGetLastError();          // already enabled: fine
            }
        }
    *out = g;
    return B200ZK_OK;
}

uint32_t b200zk_group_size(const b200zk_group* g) { return g ? (uint32_t)g->ctxs.size() : 0; }
b200zk_ctx* b200zk_group_ctx(b200zk_group* g, uint32_t rank) { return g && rank < g->ctxs.size() ? g->ctxs[rank] : nullptr; }
// after a failed collective call: clears the abort flag so that the group can be used again
int32_t b200zk_group_reset(b200zk_group* g) {
    if (!g) return B200ZK_EINVAL;
    g->state->reset();
    return B200ZK_OK;
}

}  // extern "C"
