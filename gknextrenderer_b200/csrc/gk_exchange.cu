// gk_exchange.cu — frame-end exchange helpers of the multi-GPU compositor.
//
// The reference is single-GPU; this is the one step the tile-partitioned frame adds.  Rank r owns
// the row tiles with (row / tileRows) % tileCount == r.  At frame end
//   pack   : the owned rows of the six integrator planes the filters read (diffuse, specular,
//            albedo, normal, object id, motion = 44 B/pixel) are gathered into one contiguous
//            staging buffer                                   (this file, one launch)
//   gather : ncclAllGather over NVLink on the staging buffers (host side: torch.distributed)
//   unpack : every rank scatters all ranks' rows back into its full planes (one launch)
// after which each rank filters the whole image locally (halos need no further traffic).
// Both kernels are pure streaming copies in 4-byte words, coalesced along rows.
#include "gk_context.h"
#include <cuda_fp16.h>

namespace gk {

namespace {

constexpr int kExchangePlanes = 6;
const GkPlane kPlaneIds[kExchangePlanes] = {GK_PLANE_OUTPUT_DIFFUSE, GK_PLANE_OUTPUT_SPECULAR, GK_PLANE_ALBEDO, GK_PLANE_NORMAL, GK_PLANE_OBJECT_ID0, GK_PLANE_MOTION};
const uint32_t kPlaneWords[kExchangePlanes] = {2, 2, 2, 2, 1, 2}; // 4-byte words per pixel

struct ExchangeArgs {
    uint32_t* plane[kExchangePlanes];
    uint32_t words[kExchangePlanes];    // words per pixel
    uint64_t offset[kExchangePlanes];   // word offset of the plane inside one rank's staging block
    uint32_t width, height, tileRows, tileCount, blocksPerRank;
    uint64_t rankWords; // words per rank block
};

// One thread per word of the staging buffer of `ranks` ranks (pack: ranks == 1 and rank0 = own rank).
template <bool kPack>
__global__ void __launch_bounds__(256) k_exchange(ExchangeArgs A, uint32_t* __restrict__ staging, uint32_t rank0, uint32_t ranks)
{
    const uint64_t total = A.rankWords * ranks;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t rk = (uint32_t)(i / A.rankWords) + rank0;
        uint64_t w = i % A.rankWords;
        int p = kExchangePlanes - 1;
#pragma unroll
        for (int q = kExchangePlanes - 1; q > 0; --q)
            if (w < A.offset[q]) p = q - 1;
        w -= A.offset[p];
        const uint32_t rowWords = A.width * A.words[p];
        const uint32_t lrow = (uint32_t)(w / rowWords), col = (uint32_t)(w % rowWords);
        const uint32_t blk = lrow / A.tileRows, within = lrow % A.tileRows;
        const uint32_t row = (blk * A.tileCount + rk) * A.tileRows + within;
        if (row >= A.height) continue; // padding rows of the last block
        uint32_t* px = A.plane[p] + (uint64_t)row * rowWords + col;
        if (kPack) staging[i] = *px;
        else *px = staging[i];
    }
}

ExchangeArgs makeArgs(const Context& c)
{
    ExchangeArgs A;
    A.width = c.width, A.height = c.height, A.tileRows = c.tileRows, A.tileCount = c.tileCount;
    A.blocksPerRank = (c.height + c.tileRows * c.tileCount - 1) / (c.tileRows * c.tileCount);
    uint64_t off = 0;
    for (int p = 0; p < kExchangePlanes; ++p) {
        A.plane[p] = (uint32_t*)c.planes.p[kPlaneIds[p]];
        A.words[p] = kPlaneWords[p];
        A.offset[p] = off;
        off += (uint64_t)A.blocksPerRank * c.tileRows * c.width * kPlaneWords[p];
    }
    A.rankWords = off;
    return A;
}

// Peer-to-peer variant: one launch stores the rows this rank owns straight into the planes of every
// other rank over NVLink (no staging buffer, no unpack).  grid = (chunks of a row, owned rows, planes).
struct PushArgs {
    const uint32_t* src[kExchangePlanes];
    uint32_t* dst[kMaxPeers][kExchangePlanes];
    uint32_t words[kExchangePlanes];
    uint32_t width, height, tileRows, tileCount, tileIndex, peers;
    uint32_t peerMask; // destination ranks
};

__global__ void __launch_bounds__(256) k_exchange_push(PushArgs A)
{
    const uint32_t p = blockIdx.z, lrow = blockIdx.y;
    const uint32_t row = ((lrow / A.tileRows) * A.tileCount + A.tileIndex) * A.tileRows + lrow % A.tileRows;
    if (row >= A.height) return;
    const uint32_t rowWords = A.width * A.words[p];
    const uint64_t base = (uint64_t)row * rowWords;
    if ((rowWords & 3u) == 0) { // 128-bit path (every plane base is 256-byte aligned)
        const uint4* s = reinterpret_cast<const uint4*>(A.src[p] + base);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rowWords / 4; i += gridDim.x * blockDim.x) {
            const uint4 v = s[i];
            for (uint32_t r = 0; r < A.peers; ++r)
                if ((A.peerMask >> r) & 1u) reinterpret_cast<uint4*>(A.dst[r][p] + base)[i] = v;
        }
    } else {
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rowWords; i += gridDim.x * blockDim.x) {
            const uint32_t v = A.src[p][base + i];
            for (uint32_t r = 0; r < A.peers; ++r)
                if ((A.peerMask >> r) & 1u) A.dst[r][p][base + i] = v;
        }
    }
}

// ---- frame-sharded progressive rendering -------------------------------------------------------
// Rank s has traced the whole frame f0 + s.  k_shard_push sends every row of its three source planes
// (diffuse, specular, albedo) to the rank that owns the row: slot s of that rank's gather buffer.
struct ShardPushArgs {
    const uint2* src[3];
    uint2* dst[kMaxPeers]; // gather buffer of every rank (own one included)
    uint64_t slotPixels;   // pixels of one plane of one source slot
    uint32_t width, height, tileRows, world, rank;
};

__global__ void __launch_bounds__(256) k_shard_push(ShardPushArgs A)
{
    const uint32_t y = blockIdx.y, p = blockIdx.z;
    const uint32_t tile = y / A.tileRows;
    const uint32_t owner = tile % A.world;
    const uint32_t lrow = (tile / A.world) * A.tileRows + y % A.tileRows; // row inside the owner's share
    const uint2* s = A.src[p] + (uint64_t)y * A.width;
    uint2* d = A.dst[owner] + ((uint64_t)A.rank * 3 + p) * A.slotPixels + (uint64_t)lrow * A.width;
    if ((A.width & 1u) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(s);
        uint4* d4 = reinterpret_cast<uint4*>(d);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.width / 2; i += gridDim.x * blockDim.x) d4[i] = s4[i];
    } else {
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.width; i += gridDim.x * blockDim.x) d[i] = s[i];
    }
}

struct ShardAccArgs {
    const uint2* gather;
    const uint2* hist[3];
    uint2* out[3];
    uint64_t slotPixels;
    uint32_t width, height, tileRows, world, rank;
    float keep; // 1 / TemporalFrames, clamped (ReProject:76-82)
};

__device__ __forceinline__ float3 unpackHalf3(uint2 v)
{
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    return make_float3(a.x, a.y, b.x);
}
__device__ __forceinline__ uint2 packHalf4(float3 c, float a)
{
    const __half2 lo = __floats2half2_rn(c.x, c.y), hi = __floats2half2_rn(c.z, a);
    uint2 v;
    v.x = *reinterpret_cast<const uint32_t*>(&lo), v.y = *reinterpret_cast<const uint32_t*>(&hi);
    return v;
}

// One thread per owned pixel: history <- lerp(history, src_s, keep) for the frames s = 0..world-1 in order,
// rounded to RGBA16F after every step like `world` consecutive single-GPU frames (lerp(a,b,t) = a*(1-t) + b*t).
__global__ void __launch_bounds__(256) k_shard_accumulate(ShardAccArgs A)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, lrow = blockIdx.y;
    const uint32_t y = ((lrow / A.tileRows) * A.world + A.rank) * A.tileRows + lrow % A.tileRows;
    if (x >= A.width || y >= A.height) return;
    const uint64_t pi = (uint64_t)y * A.width + x, li = (uint64_t)lrow * A.width + x;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        uint2 packed = A.hist[ch][pi];
        for (uint32_t s = 0; s < A.world; ++s) {
            const float3 h = unpackHalf3(packed);
            const float3 c = unpackHalf3(A.gather[((uint64_t)s * 3 + ch) * A.slotPixels + li]);
            const float k = A.keep;
            packed = packHalf4(make_float3(h.x * (1.0f - k) + c.x * k, h.y * (1.0f - k) + c.y * k, h.z * (1.0f - k) + c.z * k), 1.0f);
        }
        A.out[ch][pi] = packed;
    }
}

} // namespace

static uint32_t shardRowsPadded(const Context& c) { return ((c.height + c.tileRows * c.tileCount - 1) / (c.tileRows * c.tileCount)) * c.tileRows; }

void frameShardClosePeers(Context& c)
{
    if (c.shardOpen)
        for (uint32_t r = 0; r < c.tileCount && r < (uint32_t)kMaxPeers; ++r)
            if (r != c.tileIndex && c.shardPeer[r]) cudaIpcCloseMemHandle(c.shardPeer[r]);
    for (int r = 0; r < kMaxPeers; ++r) c.shardPeer[r] = nullptr;
    c.shardOpen = false;
}

void frameShardRelease(Context& c)
{
    frameShardClosePeers(c);
    if (c.shardGather) cudaFree(c.shardGather);
    c.shardGather = nullptr;
}

GkStatus frameShardHandle(Context& c, void* out, size_t bytes)
{
    if (!out || bytes < sizeof(cudaIpcMemHandle_t)) {
        setLastError("gk_frame_shard_handle: buffer too small");
        return GK_ERR_INVALID_ARGUMENT;
    }
    if (!(c.flags & GK_CFG_TRACE_ALL_ROWS) || c.tileCount > (uint32_t)kMaxPeers) {
        setLastError("gk_frame_shard_handle: the context must be created with GK_CFG_TRACE_ALL_ROWS (and <= 16 ranks)");
        return GK_ERR_UNSUPPORTED;
    }
    if (!c.shardGather) {
        c.shardSlotPixels = (size_t)shardRowsPadded(c) * c.width;
        GK_CUDA(cudaMalloc(&c.shardGather, sizeof(uint2) * c.shardSlotPixels * 3 * c.tileCount));
        GK_CUDA(cudaMemsetAsync(c.shardGather, 0, sizeof(uint2) * c.shardSlotPixels * 3 * c.tileCount, c.stream));
        GK_CUDA(cudaStreamSynchronize(c.stream));
    }
    GK_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(out), c.shardGather));
    return GK_OK;
}

GkStatus frameShardOpen(Context& c, const void* handlesAll, uint32_t world)
{
    if (!handlesAll || world != c.tileCount || !c.shardGather) {
        setLastError("gk_frame_shard_open: call gk_frame_shard_handle first; world must equal GkConfig.tileCount");
        return GK_ERR_INVALID_ARGUMENT;
    }
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handlesAll);
    for (uint32_t r = 0; r < world; ++r) {
        if (r == c.tileIndex) {
            c.shardPeer[r] = c.shardGather;
            continue;
        }
        const cudaError_t e = cudaIpcOpenMemHandle(&c.shardPeer[r], h[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            setLastError(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            cudaGetLastError();
            return GK_ERR_CUDA;
        }
    }
    c.shardOpen = true;
    return GK_OK;
}

GkStatus frameShardPush(Context& c)
{
    if (!c.shardOpen) {
        setLastError("gk_frame_shard_push: gather buffers not opened");
        return GK_ERR_NOT_READY;
    }
    ShardPushArgs A;
    A.src[0] = (const uint2*)c.planes.p[GK_PLANE_OUTPUT_DIFFUSE], A.src[1] = (const uint2*)c.planes.p[GK_PLANE_OUTPUT_SPECULAR], A.src[2] = (const uint2*)c.planes.p[GK_PLANE_ALBEDO];
    for (int r = 0; r < kMaxPeers; ++r) A.dst[r] = (uint2*)c.shardPeer[r];
    A.slotPixels = c.shardSlotPixels;
    A.width = c.width, A.height = c.height, A.tileRows = c.tileRows, A.world = c.tileCount, A.rank = c.tileIndex;
    const dim3 grid((c.width / 2 + 255) / 256, c.height, 3);
    k_shard_push<<<grid, 256, 0, c.stream>>>(A);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return GK_OK;
}

GkStatus frameShardAccumulate(Context& c)
{
    if (!c.shardOpen || !c.haveUbo) {
        setLastError("gk_frame_shard_accumulate: gather buffers and UBO must be set");
        return GK_ERR_NOT_READY;
    }
    if (!(c.ubo.ProgressiveRender != 0 && c.ubo.BFSize == 0)) {
        setLastError("gk_frame_shard_accumulate: needs ProgressiveRender != 0 and BFSize == 0");
        return GK_ERR_UNSUPPORTED;
    }
    applyPendingHistorySwap(c); // the accumulated planes of the last super-step become the history
    GK_CUDA(cudaMemcpyAsync(c.dUbo, &c.ubo, sizeof(GkUniformBufferObject), cudaMemcpyHostToDevice, c.stream));
    void** P = c.planes.p;
    ShardAccArgs A;
    A.gather = c.shardGather;
    A.hist[0] = (const uint2*)P[GK_PLANE_HISTORY_DIFFUSE], A.hist[1] = (const uint2*)P[GK_PLANE_HISTORY_SPECULAR], A.hist[2] = (const uint2*)P[GK_PLANE_HISTORY_ALBEDO];
    A.out[0] = (uint2*)P[GK_PLANE_ACCUM_DIFFUSE], A.out[1] = (uint2*)P[GK_PLANE_ACCUM_SPECULAR], A.out[2] = (uint2*)P[GK_PLANE_ACCUM_ALBEDO];
    A.slotPixels = c.shardSlotPixels;
    A.width = c.width, A.height = c.height, A.tileRows = c.tileRows, A.world = c.tileCount, A.rank = c.tileIndex;
    const float t = 1.0f / float(c.ubo.TemporalFrames);
    A.keep = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
    {
        const void* bufs[] = {P[GK_PLANE_ACCUM_DIFFUSE], P[GK_PLANE_ACCUM_SPECULAR], P[GK_PLANE_ACCUM_ALBEDO]};
        waitAsyncCopyBeforeWriting(c, bufs, 3);
    }
    const dim3 grid((c.width + 255) / 256, shardRowsPadded(c), 1);
    k_shard_accumulate<<<grid, 256, 0, c.stream>>>(A);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return composeOwnedRows(c);
}

GkStatus exchangeIpcHandles(Context& c, void* out, size_t bytes)
{
    if (!out || bytes < kPeerBuffers * sizeof(cudaIpcMemHandle_t)) {
        setLastError("gk_exchange_ipc_handles: buffer too small");
        return GK_ERR_INVALID_ARGUMENT;
    }
    applyPendingHistorySwap(c);
    cudaIpcMemHandle_t* h = static_cast<cudaIpcMemHandle_t*>(out);
    for (int p = 0; p < kExchangePlanes; ++p) GK_CUDA(cudaIpcGetMemHandle(&h[p], c.planes.p[kPlaneIds[p]]));
    GK_CUDA(cudaIpcGetMemHandle(&h[kExchangePlanes], c.planes.p[GK_PLANE_OBJECT_ID1]));
    GK_CUDA(cudaIpcGetMemHandle(&h[kExchangePlanes + 1], c.planes.p[GK_PLANE_DENOISED]));
    c.peers.myId0 = c.planes.p[GK_PLANE_OBJECT_ID0], c.peers.myId1 = c.planes.p[GK_PLANE_OBJECT_ID1];
    return GK_OK;
}

void exchangeClosePeers(Context& c)
{
    if (!c.peers.open) return;
    for (uint32_t r = 0; r < c.peers.world; ++r)
        for (int b = 0; b < kPeerBuffers; ++b)
            if (c.peers.base[r][b]) cudaIpcCloseMemHandle(c.peers.base[r][b]), c.peers.base[r][b] = nullptr;
    c.peers.open = false;
}

GkStatus exchangeOpenPeers(Context& c, const void* handlesAll, uint32_t world)
{
    if (!handlesAll || world != c.tileCount || world > (uint32_t)kMaxPeers) {
        setLastError("gk_exchange_open_peers: world must equal GkConfig.tileCount (<= 16)");
        return GK_ERR_INVALID_ARGUMENT;
    }
    if (!c.peers.myId0) {
        setLastError("gk_exchange_open_peers: call gk_exchange_ipc_handles first");
        return GK_ERR_NOT_READY;
    }
    exchangeClosePeers(c);
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handlesAll);
    c.peers.world = world;
    for (uint32_t r = 0; r < world; ++r) {
        if (r == c.tileIndex) continue;
        for (int b = 0; b < kPeerBuffers; ++b) {
            const cudaError_t e = cudaIpcOpenMemHandle(&c.peers.base[r][b], h[r * kPeerBuffers + b], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                c.peers.open = true; // so that the handles opened so far are closed
                exchangeClosePeers(c);
                setLastError(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
                cudaGetLastError();
                return GK_ERR_CUDA;
            }
        }
    }
    c.peers.open = true;
    return GK_OK;
}

GkStatus exchangePush(Context& c)
{
    if (!c.peers.open) {
        setLastError("gk_exchange_push: peers not opened");
        return GK_ERR_NOT_READY;
    }
    PushArgs A;
    A.width = c.width, A.height = c.height, A.tileRows = c.tileRows, A.tileCount = c.tileCount, A.tileIndex = c.tileIndex, A.peers = c.peers.world;
    A.peerMask = ((1u << c.peers.world) - 1u) & ~(1u << c.tileIndex);
    // The two object-id buffers trade places every frame on every rank alike: address the peer's
    // buffer that plays the role of ObjectId0 this frame.
    const bool idSwapped = c.planes.p[GK_PLANE_OBJECT_ID0] != c.peers.myId0;
    for (int p = 0; p < kExchangePlanes; ++p) {
        A.src[p] = (const uint32_t*)c.planes.p[kPlaneIds[p]];
        A.words[p] = kPlaneWords[p];
        const int slot = (kPlaneIds[p] == GK_PLANE_OBJECT_ID0 && idSwapped) ? kExchangePlanes : p;
        for (uint32_t r = 0; r < c.peers.world; ++r) A.dst[r][p] = (uint32_t*)c.peers.base[r][slot];
    }
    const uint32_t blocksPerRank = (c.height + c.tileRows * c.tileCount - 1) / (c.tileRows * c.tileCount);
    const dim3 grid((c.width * 2 / 4 + 255) / 256, blocksPerRank * c.tileRows, kExchangePlanes);
    k_exchange_push<<<grid, 256, 0, c.stream>>>(A);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return GK_OK;
}

GkStatus exchangePushFinal(Context& c, int dstRank)
{
    if (!c.peers.open) {
        setLastError("gk_exchange_push_final: peers not opened");
        return GK_ERR_NOT_READY;
    }
    if (dstRank >= (int)c.peers.world) {
        setLastError("gk_exchange_push_final: dst_rank out of range");
        return GK_ERR_INVALID_ARGUMENT;
    }
    PushArgs A{};
    A.width = c.width, A.height = c.height, A.tileRows = c.tileRows, A.tileCount = c.tileCount, A.tileIndex = c.tileIndex, A.peers = c.peers.world;
    A.peerMask = dstRank < 0 ? ((1u << c.peers.world) - 1u) : (1u << dstRank);
    A.peerMask &= ~(1u << c.tileIndex);
    if (A.peerMask == 0) return GK_OK; // the destination is this rank: its rows are already in place
    A.src[0] = (const uint32_t*)c.planes.p[GK_PLANE_DENOISED];
    A.words[0] = 2;
    for (uint32_t r = 0; r < c.peers.world; ++r) A.dst[r][0] = (uint32_t*)c.peers.base[r][kExchangePlanes + 1];
    const uint32_t blocksPerRank = (c.height + c.tileRows * c.tileCount - 1) / (c.tileRows * c.tileCount);
    const dim3 grid((c.width * 2 / 4 + 255) / 256, blocksPerRank * c.tileRows, 1);
    k_exchange_push<<<grid, 256, 0, c.stream>>>(A);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return GK_OK;
}

size_t exchangeBytesPerRank(const Context& c) { return (size_t)makeArgs(c).rankWords * 4; }

GkStatus exchangePack(Context& c, void* dStaging)
{
    const ExchangeArgs A = makeArgs(c);
    const unsigned grid = (unsigned)std::min<uint64_t>((A.rankWords + 255) / 256, 148u * 16u);
    k_exchange<true><<<grid, 256, 0, c.stream>>>(A, (uint32_t*)dStaging, c.tileIndex, 1);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return GK_OK;
}

GkStatus exchangeUnpack(Context& c, const void* dAll)
{
    {   // the six exchange planes are about to be overwritten: an asynchronous read-back may still be reading one of them
        const void* bufs[] = {c.planes.p[GK_PLANE_OUTPUT_DIFFUSE], c.planes.p[GK_PLANE_OUTPUT_SPECULAR], c.planes.p[GK_PLANE_ALBEDO], c.planes.p[GK_PLANE_NORMAL],
                              c.planes.p[GK_PLANE_OBJECT_ID0], c.planes.p[GK_PLANE_MOTION]};
        waitAsyncCopyBeforeWriting(c, bufs, 6);
    }
    const ExchangeArgs A = makeArgs(c);
    const unsigned grid = (unsigned)std::min<uint64_t>((A.rankWords * c.tileCount + 255) / 256, 148u * 16u);
    k_exchange<false><<<grid, 256, 0, c.stream>>>(A, (uint32_t*)dAll, 0, c.tileCount);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return GK_OK;
}

} // namespace gk
