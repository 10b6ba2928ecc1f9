// comm.cu - the one exchange step of the tile-sharded path: an NCCL all-gather of every rank's
// result rows over NVLink / NVSwitch, issued by the library on the handle's stream.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2 - the copy the process already has, e.g. PyTorch's, wins), so the
// library has no link-time dependency on it and single-GPU use never touches it.  The communicator is created from a
// 128-byte unique id that the host program distributes by whatever means it has (y3_comm_unique_id on rank 0, then a
// broadcast); one process per GPU, one communicator per handle.
#include "comm.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace y3 {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
constexpr ncclDataType_t kNcclInt64 = 4, kNcclFloat64 = 8;      // nccl.h: ncclInt64 = 4, ncclFloat64 = ncclDouble = 8

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.so) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* so = nullptr;
    for (const char* n : names) { so = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (so) break; }   // already in the process
    for (const char* n : names) { if (so) break; so = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
    Y3_CHECK(so, Y3_ERR_UNSUPPORTED, "NCCL is not available (dlopen libnccl.so.2: %s)", dlerror());
    auto sym = [&](const char* name) {
        void* p = dlsym(so, name);
        Y3_CHECK(p, Y3_ERR_UNSUPPORTED, "libnccl has no symbol %s", name);
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.so = so;
    return api;
}

#define Y3_NCCL(expr)                                                                                   \
    do {                                                                                                \
        const ncclResult_t r_ = (expr);                                                                 \
        if (r_ != 0) fail(Y3_ERR_CUDA, "%s failed: %s", #expr, nccl().GetErrorString(r_));              \
    } while (0)
}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    std::vector<double> share;
};

void comm_unique_id(uint8_t* id) {
    ncclUniqueId u;
    Y3_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, u.internal, 128);
}

void comm_init(y3_context* ctx, int rank, int nranks, const uint8_t* id) {
    Y3_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, Y3_ERR_INVALID, "rank %d outside 0..%d", rank, nranks - 1);
    comm_destroy(ctx);
    Comm* c = new Comm();
    c->rank = rank; c->nranks = nranks;
    c->share.assign((size_t)nranks, 1.0 / nranks);
    if (nranks > 1) {
        ncclUniqueId u;
        memcpy(u.internal, id, 128);
        const ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, u, rank);
        if (r != 0) { delete c; fail(Y3_ERR_CUDA, "ncclCommInitRank failed: %s", nccl().GetErrorString(r)); }
    }
    ctx->comm = c;
}

void comm_destroy(y3_context* ctx) {
    Comm* c = static_cast<Comm*>(ctx->comm);
    if (!c) return;
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
    ctx->comm = nullptr;
}

int comm_rank(const y3_context* ctx) { return ctx->comm ? static_cast<Comm*>(ctx->comm)->rank : 0; }
int comm_size(const y3_context* ctx) { return ctx->comm ? static_cast<Comm*>(ctx->comm)->nranks : 1; }
const double* comm_shares(y3_context* ctx) {
    Comm* c = static_cast<Comm*>(ctx->comm);
    return c ? c->share.data() : nullptr;
}

void comm_update_shares(y3_context* ctx, const long long* tiles, const long long* micros) {
    Comm* c = static_cast<Comm*>(ctx->comm);
    // Off by default.  Measured on 4 B200s (20000^2 image, 702 tiles per rank): a rank's time for the same tiles moves by +-5 %
    // from step to step (70-88 ms; power management, not a persistent property of the GPU), so shares that follow the last
    // measurement chase noise and push ranks over a tile-batch boundary: 4030 Mpix/s against 4868 with equal shares.
    static const bool adaptive = getenv("Y3_ADAPTIVE_SHARDS") != nullptr;
    if (!c || c->nranks == 1 || !adaptive) return;
    std::vector<double> thr((size_t)c->nranks);
    double sum = 0;
    for (int r = 0; r < c->nranks; ++r) {
        if (tiles[r] <= 0 || micros[r] <= 0) return;             // a rank without work / timing: keep the current shares
        thr[r] = (double)tiles[r] / (double)micros[r];
        sum += thr[r];
    }
    double norm = 0;
    for (int r = 0; r < c->nranks; ++r) {
        const double target = thr[r] / sum;
        c->share[r] = std::max(0.5 * c->share[r] + 0.5 * target, 0.25 / c->nranks);      // damped, never starving a rank
        norm += c->share[r];
    }
    for (int r = 0; r < c->nranks; ++r) c->share[r] /= norm;
    static const bool dbg = getenv("Y3_DEBUG_TIMING") != nullptr;
    if (dbg && c->rank == 0) {
        fprintf(stderr, "y3: shard update:");
        for (int r = 0; r < c->nranks; ++r) fprintf(stderr, " [%lld tiles %.2f ms -> %.4f]", tiles[r], micros[r] / 1000.0, c->share[r]);
        fprintf(stderr, "\n");
    }
}

int comm_nccl_version() { int v = 0; nccl().GetVersion(&v); return v; }

void comm_all_gather_i64(y3_context* ctx, const long long* send_dev, long long* recv_dev, size_t count) {
    Comm* c = static_cast<Comm*>(ctx->comm);
    if (!c || c->nranks == 1) {
        Y3_CUDA(cudaMemcpyAsync(recv_dev, send_dev, count * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    Y3_NCCL(nccl().AllGather(send_dev, recv_dev, count, kNcclInt64, c->comm, ctx->stream));
}

void comm_all_gather_f64(y3_context* ctx, const double* send_dev, double* recv_dev, size_t count) {
    Comm* c = static_cast<Comm*>(ctx->comm);
    if (!c || c->nranks == 1) {
        Y3_CUDA(cudaMemcpyAsync(recv_dev, send_dev, count * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    Y3_NCCL(nccl().AllGather(send_dev, recv_dev, count, kNcclFloat64, c->comm, ctx->stream));
}

}  // namespace y3
