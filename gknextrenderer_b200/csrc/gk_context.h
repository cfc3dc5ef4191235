// gk_context.h — host-side state behind a GkContext handle.
#pragma once
#include "../../include/gknext_cuda.h"
#include "gk_bvh.cuh"
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace gk {

void setLastError(const std::string& s);

#define GK_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            gk::setLastError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return GK_ERR_CUDA;                                                                               \
        }                                                                                                     \
    } while (0)

// growable device buffer
template <class T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0; // elements
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr, o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept
    {
        if (this != &o) {
            release();
            p = o.p, cap = o.cap, o.p = nullptr, o.cap = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); } // temporaries in entry points are freed on every return path
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr, cap = 0;
        size_t want = n + n / 8 + 16;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr, cap = 0;
    }
    size_t bytes() const { return cap * sizeof(T); }
};

// Timing events that are destroyed on every return path.
struct ScopedEvents {
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    explicit ScopedEvents(int n)
    {
        for (int i = 0; i < n && i < 4; ++i) cudaEventCreate(&e[i]);
    }
    ~ScopedEvents()
    {
        for (cudaEvent_t x : e)
            if (x) cudaEventDestroy(x);
    }
    ScopedEvents(const ScopedEvents&) = delete;
    ScopedEvents& operator=(const ScopedEvents&) = delete;
};

// Binary radix tree (Karras 2012) over `n` primitives, kept for refit.
struct Lbvh {
    uint32_t n = 0;
    DevBuf<unsigned long long> keys, keysAlt;
    DevBuf<uint32_t> order, orderAlt; // sorted position -> primitive
    DevBuf<uint32_t> left, right;     // per internal node (n-1)
    DevBuf<uint32_t> parentI, parentL; // parent of internal node / of leaf (sorted position)
    DevBuf<uint32_t> first, last;      // key range of each internal node
    DevBuf<float4> ilo, ihi;           // internal-node boxes
    DevBuf<float4> llo, lhi;           // leaf boxes in sorted order
    DevBuf<float4> plo, phi;           // primitive boxes in primitive order
    DevBuf<uint32_t> group;            // primitive -> group (model) id
    DevBuf<int> flags;
    DevBuf<float> cost;                // collapse cost table, 8 floats per internal node (build only)
    bool costValid = false;
    float primCost = 1.f;
    void release();
};

struct ModelInfo { // per Assets::Model
    uint32_t vertexOffset, vertexCount;
    uint32_t indexOffset, indexCount; // into the concatenated index buffer
    uint32_t triOffset, triCount;     // into the concatenated (unsorted) triangle list
    uint32_t blasRoot;                // reference into blasNodes (leaf reference possible)
    uint32_t pad;
    float bmin[4], bmax[4];           // BLAS root box (BVH::aabbMin/aabbMax)
};

// Peer-to-peer frame exchange (multi-GPU, one process per GPU): the exchange planes of the other ranks,
// opened through CUDA IPC, so that one kernel can store this rank's rows straight into every peer.
constexpr int kMaxPeers = 16;
constexpr int kPeerBuffers = 8; // the six exchange planes, the second object-id buffer (the two swap every frame), rtDenoised
struct PeerExchange {
    bool open = false;
    uint32_t world = 0;
    void* base[kMaxPeers][kPeerBuffers] = {};
    void* myId0 = nullptr; // this rank's object-id buffers as they were exported
    void* myId1 = nullptr;
};

struct Planes {
    void* p[GK_PLANE_COUNT] = {};
    size_t bytes[GK_PLANE_COUNT] = {};
};

// SoA wavefront state, one slot per owned pixel ("path")
struct PathState {
    uint4* rng;
    float4* posMat;      // current vertex position, w = material index (bits)
    float4* nrmFlags;    // current vertex normal, w = packed bounce/sample/state flags (bits)
    float4* dirT;        // current direction
    float4* throughput;  // rayColor rgb, w unused
    float4* primPosMat;  // primary vertex position, w = material index
    float4* primNrm;     // primary vertex normal, w = |pixelOffset|
    float4* accDiffuse;  // FinalColor accumulator, w = direct-light shadow term
    float4* accSpec;     // FinalReflection accumulator
    uint32_t* pixel;     // image pixel index of the path
    uint32_t* rays;      // rays traced so far
    float4* dofVertex;   // scratch for the depth-of-field re-trace (initial vertex position, w = raw material slot)
    float4* dofNormal;   // initial vertex normal, w = node index bits
};

struct RayQueue {
    float4* o_tmin; // origin.xyz, tmin
    float4* d_tmax; // direction.xyz, tmax
    uint32_t* path; // owning path
    float4* hit_tuvp; // results: t, u, v, prim(bits)
    uint32_t* hit_inst;
    uint32_t* count; // device counter
};

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t width = 0, height = 0;
    uint32_t tileIndex = 0, tileCount = 1, tileRows = 16;
    uint32_t traceTileIndex = 0, traceTileCount = 1; // rows the path tracer covers (== tileIndex/tileCount unless GK_CFG_TRACE_ALL_ROWS)
    uint32_t ownedRows = 0, pathCount = 0;
    uint32_t flags = 0;

    // scene
    bool haveScene = false, haveInstances = false, haveUbo = false;
    std::vector<ModelInfo> models;
    DevBuf<ModelInfo> dModels;
    DevBuf<GkGPUVertex> dGpuVerts;
    DevBuf<uint32_t> dIndices;
    DevBuf<GkMaterial> dMaterials;
    uint32_t materialCount = 0;
    DevBuf<GkLightObject> dLights;
    DevBuf<float4> dFaceNormals; // per triangle (model order): FCPUBLASVertInfo::normal
    DevBuf<GkNodeProxy> dNodes;
    DevBuf<GkNodeProxy> dSparseNodes; // staging of gk_update_instances_sparse
    DevBuf<uint32_t> dSparseIdx;
    uint32_t nodeCount = 0;
    DevBuf<GkAmbientCube> dCubes;
    DevBuf<GkVoxelData> dVoxels;
    DevBuf<GkAmbientCube> dCubesPrev; // probe state before a gk_bake_probes call (what its gathers read)
    DevBuf<GkVoxelData> dVoxelsPrev;
    bool haveProbes = false;
    uint64_t totalTris = 0, instancedTris = 0;

    // acceleration structures
    Lbvh blasTree, tlasTree;
    DevBuf<TriRecord> dTris;
    DevBuf<WideNode> dBlasNodes, dTlasNodes;
    DevBuf<uint32_t> dBlasSrc, dTlasSrc; // 8 binary-tree references per wide node (refit)
    uint32_t blasNodeCount = 0, tlasNodeCount = 0;
    DevBuf<InstRecord> dInst;
    uint32_t tlasRoot = 0;
    DevBuf<uint32_t> dTaskA, dTaskB, dCounters;
    DevBuf<unsigned char> dSortTemp;
    DevBuf<float4> dGroupLo, dGroupHi;
    DevBuf<uint32_t> dGroupRoot;
    float msBlasBuild = 0, msTlasBuild = 0, msRefit = 0;

    // frame
    GkUniformBufferObject ubo{};
    GkUniformBufferObject* dUbo = nullptr;
    Planes planes;
    PathState paths{};
    std::vector<void*> pathAllocs;
    RayQueue extendQ[2]{}, shadowQ[2]{};
    std::vector<void*> queueAllocs;
    uint32_t* hCounts = nullptr; // pinned
    TraversalStats* dTravStats = nullptr;
    unsigned long long* dTailCounters = nullptr;
    float tailFraction = 0.25f;     // ... and this fraction of the frame's paths
    uint32_t tailThreshold = 0;     // finish the frame in one launch once this few paths are alive (0: never)
    bool travStats = false;
    uint32_t blasLeafMax = 4;       // triangles per BLAS leaf (<= kBlasLeafMax)
    uint32_t coopThreshold = 262144; // waves smaller than this use the 8-lanes-per-ray traversal (variant 0: measured optimum 65536, set by gk_create)
    GkFrameStats stats{};
    cudaEvent_t evA = nullptr, evB = nullptr;
    std::vector<cudaEvent_t> evPool;
    int captureWave = -1;
    PeerExchange peers;
    // frame-sharded progressive rendering: sources of the `world` frames of a super-step for the rows this rank owns
    uint2* shardGather = nullptr; // [source rank][3 planes][owned row][x]
    size_t shardSlotPixels = 0;   // pixels of one plane of one source
    void* shardPeer[kMaxPeers] = {};
    bool shardOpen = false;
    // asynchronous read-back (gk_readback_async): copy stream + the buffer it is still reading
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evCopyReady = nullptr, evCopyDone = nullptr;
    const void* asyncCopySrc = nullptr; // non-null while a copy may be in flight
    uint32_t coopDivisor = 4, tailDivisor = 8; // swept on 1/1 .. 1/8 frame shares (2 / 4 gain 3-5 % on a room share, lose 25 % on Cornell); the cooperative kernel / the tail never take more than paths/divisor (options "coop_divisor", "tail_divisor")
    int microTiles = 2; // path order: 0 rows, 1 = 8 x 4 pixel blocks per warp, 2 = additionally 16 x 16 squares per thread block (option "micro_tiles")
    uint32_t refitRejectedInARow = 0, refitBackoffLeft = 0; // see updateInstancesOnDevice
    float tlasAreaAtBuild = 0.f;    // summed internal-node area of the TLAS when it was last built
    uint32_t refitRejected = 0;     // refits that degraded the tree too much and became rebuilds
    DevBuf<uint32_t> dRootRef;
    bool concurrentShadow = true;   // shadow kernel of a wave on a second stream, next to the extend kernel
    cudaStream_t stream2 = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    unsigned laneBlock = 256;       // threads (= rays) per block of the one-ray-per-lane trace kernel
    int shadeMinBlocks = 3;         // launch bound of k_shade (tuning hook)
    bool blasPloc = false;          // PLOC topology (+ depth-first leaf renumbering) for the BLAS forest
    int blasPlocRadius = 16;
    DevBuf<uint32_t> dPlocGrp[2], dPlocLeafGrp, dPlocGroupBase;
    bool tlasPloc = true;           // PLOC topology for the TLAS (false: Karras radix tree)
    bool tightInstanceBounds = true;     // world boxes of rotated instances from their transformed vertices (not the 8 corners of the BLAS box)
    uint32_t tightBoundsMaxTris = 32768; // ... for models up to this many triangles
    float costTri = 0.3f;                // collapse cost model: one triangle test relative to one node visit
    int tlasPlocRadius = 64;             // search window of the nearest-neighbour pass
    uint32_t tlasPlocMax = 65536;        // ... up to this many instances (no gain measured on 200 k lattice bricks, 37 ms build)
    uint32_t tlasPlocMaxRebuild = 32768; // guard-forced (per-frame) rebuilds above this many instances keep the radix tree
    DevBuf<uint32_t> dPlocRef[2], dPlocNn, dPlocValid, dPlocPos;
    DevBuf<float4> dPlocLo[2], dPlocHi[2];
    int tlasSizeBits = 2;           // extended Morton code of the TLAS: box-size bits woven into the key (0 = plain Morton)
    bool sahCollapse = true;        // cost-driven wide collapse (false: greedy by surface area)
    // scheduled traversal kernel (gk_trace_sched.cuh)
    int traceVariant = 1;            // 0: while-while lane kernel / cooperative kernel, 1: persistent vote-scheduled kernel
    uint32_t schedRefillMin = 6;     // refill a warp's idle lanes once this many rays have finished
    uint32_t schedBiasN = 0;         // vote bias towards the node step (lanes)
    uint32_t schedKeepN = 12;        // a node phase goes on while this many lanes hold a node (33: one step per vote)
    uint32_t schedKeepT = 4;         // ditto for the triangle phase
    bool primaryLaneKernel = true;   // wave 0 (coherent camera rays) on the while-while lane kernel
    bool tailCoop = true;            // ... with eight lanes per path (k_tail_coop) instead of one
    uint32_t streamTailPaths = 262144; // streamed wave loop: finish the frame in one k_tail launch once at most this many rays are in flight
    uint32_t schedMinRays = 1000000; // waves of fewer live paths than this run on the one-ray-per-lane kernel (swept 0 .. 1 M on 1/1 .. 1/8 frame shares: -0.2 % .. -7 %)
    int schedBlocksPerSm = 0, smCount = 0;
    static constexpr uint32_t kCursorCount = 1024;
    uint32_t* dCursors = nullptr;    // one zeroed queue cursor per launch of a frame
    uint32_t cursorNext = 0;
    struct SchedStats* dSchedStats = nullptr;
    uint32_t* dOverflow = nullptr;   // set by a traversal that had to drop a stack entry
    uint32_t waveLookahead = 2;      // waves the host runs ahead of the device's published queue sizes (streamed wave loop)
    static constexpr uint32_t kWaveLimit = 4096;
    volatile uint32_t* hWave = nullptr; // mapped pinned memory: 4 words per wave {next extend count, next shadow count, frame tag, -}
    uint32_t* dWave = nullptr;          // the same memory as the device sees it
    uint32_t frameTag = 0;
    DevBuf<float4> dCapture;
    uint32_t capturedCount = 0;
    uint64_t frameIndex = 0;
    bool pendingHistorySwap = false, tracedSinceFilter = false;

    SceneView view() const
    {
        SceneView v;
        v.tlasNodes = dTlasNodes.p, v.blasNodes = dBlasNodes.p, v.tris = dTris.p, v.inst = dInst.p;
        v.tlasRoot = tlasRoot, v.instanceCount = nodeCount, v.overflowFlag = dOverflow;
        return v;
    }
};

// gk_scene.cu
GkStatus uploadScene(Context& c, const GkSceneDesc& d);
// gk_bvh_build.cu
GkStatus buildBlasForest(Context& c);
GkStatus updateInstances(Context& c, const GkNodeProxy* nodes, uint32_t count, bool refit);
GkStatus updateInstancesSparse(Context& c, const uint32_t* indices, const GkNodeProxy* proxies, uint32_t changed, bool refit);
// gk_integrator.cu
GkStatus allocFrameResources(Context& c);
void freeFrameResources(Context& c);
GkStatus traceFrame(Context& c);
GkStatus intersectDevice(Context& c, const float4* rays, uint32_t n, float* tuv, uint32_t* ids, bool anyHit);
GkStatus checkTraversalOverflow(Context& c); // synchronises; GK_ERR_UNSUPPORTED if a traversal dropped a stack entry
GkStatus raycastBatch(Context& c, const float* originDir, uint32_t n, GkRayCastResult* out);
// gk_probes.cu
GkStatus bakeProbes(Context& c, uint32_t first, uint32_t count);
GkStatus getProbes(Context& c, GkAmbientCube* cubes, GkVoxelData* voxels, size_t count);
// gk_filters.cu
GkStatus filterFrame(Context& c);
void applyPendingHistorySwap(Context& c);
// gk_exchange.cu
size_t exchangeBytesPerRank(const Context& c);
GkStatus exchangePack(Context& c, void* dStaging);
GkStatus exchangeUnpack(Context& c, const void* dAll);
GkStatus exchangeIpcHandles(Context& c, void* out, size_t bytes);
GkStatus exchangeOpenPeers(Context& c, const void* handlesAll, uint32_t world);
void waitAsyncCopyBeforeWriting(Context& c, const void* const* buffers, int count); // orders the stream after an in-flight read-back of any of them
GkStatus exchangePush(Context& c);
GkStatus exchangePushFinal(Context& c, int dstRank);
GkStatus frameShardHandle(Context& c, void* out, size_t bytes);
GkStatus frameShardOpen(Context& c, const void* handlesAll, uint32_t world);
GkStatus frameShardPush(Context& c);
GkStatus frameShardAccumulate(Context& c);
void frameShardRelease(Context& c);
void frameShardClosePeers(Context& c);
GkStatus composeOwnedRows(Context& c); // gk_filters.cu: k_denoise_jbf on the owned rows + history hand-over
GkStatus filterFrameOwnedRows(Context& c);
void exchangeClosePeers(Context& c);

} // namespace gk

struct GkContext {
    gk::Context c;
};
