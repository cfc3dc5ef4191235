// gk_compositor.cpp — include/gknext_compositor.h: the frame-end exchange of the multi-GPU compositor driven from C++
// over NCCL (one process per GPU).  It only sequences calls of the CUDA backend's C ABI with stream barriers between
// them; the data path is the backend's peer-to-peer store kernels over NVLink (gk_exchange.cu).
#include "../../include/gknext_compositor.h"
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_err;
const char* gkc_last_error(void) { return g_err.c_str(); }

struct GkCompositor {
    GkContext* ctx = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    int rank = 0, world = 1, device = 0;
    int* dToken = nullptr;          // the 4 bytes every barrier all-reduces
    unsigned char* dScratch = nullptr; // handle exchange: world x GK_EXCHANGE_IPC_BYTES
    bool peersOpen = false, shardOpen = false;
};

#define GKC_NCCL(call)                                                                         \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess) {                                                               \
            g_err = std::string(#call) + ": " + ncclGetErrorString(r_);                        \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)
#define GKC_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)
#define GKC_GK(call)                                                                           \
    do {                                                                                       \
        GkStatus s_ = (call);                                                                  \
        if (s_ != GK_OK) {                                                                     \
            g_err = std::string(#call) + ": " + gk_last_error();                               \
            return (int)s_;                                                                    \
        }                                                                                      \
    } while (0)

static_assert(sizeof(ncclUniqueId) <= GKC_UNIQUE_ID_BYTES, "ncclUniqueId must fit the id buffer");

extern "C" {

int gkc_get_unique_id(void* id, size_t bytes)
{
    if (!id || bytes < GKC_UNIQUE_ID_BYTES) {
        g_err = "gkc_get_unique_id: buffer of GKC_UNIQUE_ID_BYTES expected";
        return -1;
    }
    ncclUniqueId u;
    GKC_NCCL(ncclGetUniqueId(&u));
    memset(id, 0, bytes);
    memcpy(id, &u, sizeof(u));
    return 0;
}

// all-gathers `bytes` bytes per rank through the device scratch buffer; host in, host out (setup only)
static int allGatherHost(GkCompositor* c, const void* mine, size_t bytes, std::vector<unsigned char>& all)
{
    GKC_CUDA(cudaMemcpyAsync(c->dScratch + (size_t)c->rank * bytes, mine, bytes, cudaMemcpyHostToDevice, c->stream));
    GKC_NCCL(ncclAllGather(c->dScratch + (size_t)c->rank * bytes, c->dScratch, bytes, ncclUint8, c->comm, c->stream));
    all.resize(bytes * (size_t)c->world);
    GKC_CUDA(cudaMemcpyAsync(all.data(), c->dScratch, all.size(), cudaMemcpyDeviceToHost, c->stream));
    GKC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int gkc_barrier(GkCompositor* c)
{
    if (!c) return -1;
    GKC_NCCL(ncclAllReduce(c->dToken, c->dToken, 1, ncclInt32, ncclSum, c->comm, c->stream));
    return 0;
}

int gkc_create(GkContext* ctx, int rank, int world, const void* id, size_t bytes, GkCompositor** out)
{
    if (!ctx || !out || !id || bytes < sizeof(ncclUniqueId) || world < 2 || rank < 0 || rank >= world) {
        g_err = "gkc_create: invalid argument (world >= 2, 0 <= rank < world, id from gkc_get_unique_id)";
        return -1;
    }
    GkCompositor* c = new GkCompositor();
    c->ctx = ctx, c->rank = rank, c->world = world;
    c->stream = (cudaStream_t)gk_stream(ctx);
    auto fail = [&](int code) {
        if (c->comm) ncclCommAbort(c->comm);
        if (c->dToken) cudaFree(c->dToken);
        if (c->dScratch) cudaFree(c->dScratch);
        delete c;
        return code;
    };
    // the context made its device current when it was created; the communicator lives on that device
    if (cudaGetDevice(&c->device) != cudaSuccess) {
        g_err = "gkc_create: no current CUDA device";
        return fail(-2);
    }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = ncclCommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
        g_err = std::string("ncclCommInitRank: ") + ncclGetErrorString(r);
        c->comm = nullptr;
        return fail(-2);
    }
    if (cudaMalloc(&c->dToken, sizeof(int)) != cudaSuccess || cudaMalloc(&c->dScratch, (size_t)world * GK_EXCHANGE_IPC_BYTES) != cudaSuccess ||
        cudaMemsetAsync(c->dToken, 0, sizeof(int), c->stream) != cudaSuccess) {
        g_err = "gkc_create: device allocation failed";
        return fail(-3);
    }
    // exchange planes of every rank -> CUDA IPC handles -> all-gather -> map
    unsigned char mine[GK_EXCHANGE_IPC_BYTES] = {0};
    std::vector<unsigned char> all;
    GkStatus s = gk_exchange_ipc_handles(ctx, mine, sizeof(mine));
    int ok = s == GK_OK ? 1 : 0;
    if (!ok) g_err = std::string("gk_exchange_ipc_handles: ") + gk_last_error();
    // a rank that failed still takes part in the collectives below, so that the others do not wait for it forever
    if (allGatherHost(c, mine, sizeof(mine), all) != 0) return fail(-2);
    if (ok) {
        s = gk_exchange_open_peers(ctx, all.data(), (uint32_t)world);
        if (s != GK_OK) ok = 0, g_err = std::string("gk_exchange_open_peers: ") + gk_last_error();
    }
    // every rank must come to the same verdict: min over ranks of `ok`
    int* dOk = nullptr;
    int okAll = 0;
    if (cudaMalloc(&dOk, sizeof(int)) != cudaSuccess) return fail(-3);
    cudaMemcpyAsync(dOk, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream);
    r = ncclAllReduce(dOk, dOk, 1, ncclInt32, ncclMin, c->comm, c->stream);
    cudaMemcpyAsync(&okAll, dOk, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(dOk);
    if (r != ncclSuccess || !okAll) {
        if (ok) g_err = "gkc_create: a peer could not map the exchange planes (no peer access between the devices?)";
        if (ok) gk_exchange_close_peers(ctx);
        return fail(-5);
    }
    c->peersOpen = true;
    *out = c;
    return 0;
}

int gkc_enable_frame_sharding(GkCompositor* c)
{
    if (!c || !c->peersOpen) {
        g_err = "gkc_enable_frame_sharding: compositor not ready";
        return -4;
    }
    unsigned char mine[64];
    std::vector<unsigned char> all;
    GKC_GK(gk_frame_shard_handle(c->ctx, mine, sizeof(mine)));
    if (allGatherHost(c, mine, sizeof(mine), all) != 0) return -2;
    GKC_GK(gk_frame_shard_open(c->ctx, all.data(), (uint32_t)c->world));
    if (gkc_barrier(c) != 0) return -2;
    GKC_CUDA(cudaStreamSynchronize(c->stream));
    c->shardOpen = true;
    return 0;
}

int gkc_composite_frame(GkCompositor* c)
{
    if (!c || !c->peersOpen) {
        g_err = "gkc_composite_frame: compositor not ready";
        return -4;
    }
    GKC_GK(gk_readback_wait(c->ctx)); // peers are about to overwrite planes an asynchronous read-back may still be reading
    if (gkc_barrier(c) != 0) return -2; // every rank has finished reading last frame's planes
    GKC_GK(gk_exchange_push(c->ctx));
    return gkc_barrier(c);              // all rows have landed
}

int gkc_composite_final(GkCompositor* c, int dst_rank)
{
    if (!c || !c->peersOpen) {
        g_err = "gkc_composite_final: compositor not ready";
        return -4;
    }
    GKC_GK(gk_readback_wait(c->ctx));
    if (gkc_barrier(c) != 0) return -2;
    GKC_GK(gk_exchange_push_final(c->ctx, dst_rank));
    return gkc_barrier(c);
}

int gkc_composite_frame_shard(GkCompositor* c, int dst_rank)
{
    if (!c || !c->shardOpen) {
        g_err = "gkc_composite_frame_shard: gkc_enable_frame_sharding first";
        return -4;
    }
    GKC_GK(gk_readback_wait(c->ctx));
    if (gkc_barrier(c) != 0) return -2; // every rank has consumed the gather buffers / rtDenoised of the last super-step
    GKC_GK(gk_frame_shard_push(c->ctx));
    if (gkc_barrier(c) != 0) return -2; // all rows have landed
    GKC_GK(gk_frame_shard_accumulate(c->ctx));
    GKC_GK(gk_exchange_push_final(c->ctx, dst_rank));
    return gkc_barrier(c);              // the presenting rank holds the whole image
}

int gkc_world(const GkCompositor* c) { return c ? c->world : 0; }
int gkc_rank(const GkCompositor* c) { return c ? c->rank : -1; }

void gkc_destroy(GkCompositor* c)
{
    if (!c) return;
    // collective: nobody frees planes another rank still has mapped
    if (c->peersOpen) gk_exchange_close_peers(c->ctx);
    if (c->comm) {
        ncclAllReduce(c->dToken, c->dToken, 1, ncclInt32, ncclSum, c->comm, c->stream);
        cudaStreamSynchronize(c->stream);
        ncclCommDestroy(c->comm);
    }
    if (c->dToken) cudaFree(c->dToken);
    if (c->dScratch) cudaFree(c->dScratch);
    delete c;
}

} // extern "C"
