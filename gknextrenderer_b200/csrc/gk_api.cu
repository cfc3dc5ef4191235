// gk_api.cu — the extern "C" entry points declared in include/gknext_cuda.h.
// Each function's header comment there cites the reference interface it replaces.
#include "gk_context.h"
#include "gk_shading.cuh"
#include <algorithm>
#include <cstring>
#include <new>

namespace gk {

static thread_local std::string g_lastError;
void setLastError(const std::string& s) { g_lastError = s; }

__global__ void k_pack_rays(const float* __restrict__ originDir, uint32_t n, float4* rays)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rays[2 * i] = make_float4(originDir[6 * i], originDir[6 * i + 1], originDir[6 * i + 2], 0.0f);
    rays[2 * i + 1] = make_float4(originDir[6 * i + 3], originDir[6 * i + 4], originDir[6 * i + 5], kPrimaryTMax);
}

// RayCastInCPU's result record (CPUAccelerationStructure.cpp:283-307) from a closest hit
__global__ void k_raycast_results(const float* __restrict__ originDir, const float* __restrict__ tuv, const uint32_t* __restrict__ ids, uint32_t n,
                                  const GkNodeProxy* __restrict__ nodes, const ModelInfo* __restrict__ models, const float4* __restrict__ faceNormals,
                                  GkRayCastResult* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GkRayCastResult R;
    memset(&R, 0, sizeof(R));
    const uint32_t node = ids[2 * i + 1], prim = ids[2 * i];
    if (node != kInvalid) {
        const float t = tuv[3 * i];
        const GkNodeProxy& p = nodes[node];
        const float4 fn = faceNormals[models[p.modelId / 10].triOffset + prim];
        R.HitPoint[0] = originDir[6 * i] + originDir[6 * i + 3] * t;
        R.HitPoint[1] = originDir[6 * i + 1] + originDir[6 * i + 4] * t;
        R.HitPoint[2] = originDir[6 * i + 2] + originDir[6 * i + 5] * t;
        const float* W = p.worldTS;
        R.Normal[0] = (W[0] * fn.x + W[4] * fn.y) + (W[8] * fn.z + W[12] * 0.0f);
        R.Normal[1] = (W[1] * fn.x + W[5] * fn.y) + (W[9] * fn.z + W[13] * 0.0f);
        R.Normal[2] = (W[2] * fn.x + W[6] * fn.y) + (W[10] * fn.z + W[14] * 0.0f);
        R.Normal[3] = (W[3] * fn.x + W[7] * fn.y) + (W[11] * fn.z + W[15] * 0.0f);
        R.T = t;
        R.InstanceId = p.instanceId;
        R.Hitted = 1;
    }
    out[i] = R;
}

// Read-bandwidth probe: every thread streams 128-bit loads (L1 bypassed) over the buffer `reps` times.
__global__ void __launch_bounds__(256) k_stream_read(const uint4* __restrict__ buf, size_t n16, int reps, uint4* sink)
{
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc.x ^= v.x, acc.y ^= v.y, acc.z ^= v.z, acc.w ^= v.w;
        }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) *sink = acc; // keeps the loads alive
}

// Task.RayCast.comp.slang:31-55 — rays of the in-place records
__global__ void k_task_pack(const GkRayCastIO* __restrict__ io, uint32_t n, float4* rays)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GkRayCastIn& c = io[i].Context;
    rays[2 * i] = make_float4(c.Origin[0], c.Origin[1], c.Origin[2], kEps);          // ray.TMin = EPS (Shading.slang:712)
    rays[2 * i + 1] = make_float4(c.Direction[0], c.Direction[1], c.Direction[2], 10000.0f); // TraceRay(..., 10000, ...)
}

// ... and the result half of the records (a miss only clears Hitted)
__global__ void k_task_results(GkRayCastIO* __restrict__ io, const float* __restrict__ tuv, const uint32_t* __restrict__ ids, uint32_t n, ShadeScene SS)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GkRayCastIO& R = io[i];
    const uint32_t node = ids[2 * i + 1], prim = ids[2 * i];
    if (node == kInvalid) {
        R.Result.Hitted = 0;
        return;
    }
    const f3 ro = mk3(R.Context.Origin[0], R.Context.Origin[1], R.Context.Origin[2]), rd = mk3(R.Context.Direction[0], R.Context.Direction[1], R.Context.Direction[2]);
    Vtx v;
    resolveHit(SS, ro, rd, tuv[3 * i], tuv[3 * i + 1], tuv[3 * i + 2], prim, node, v);
    R.Result.HitPoint[0] = v.Position.x, R.Result.HitPoint[1] = v.Position.y, R.Result.HitPoint[2] = v.Position.z, R.Result.HitPoint[3] = 1.0f;
    R.Result.Normal[0] = v.Normal.x, R.Result.Normal[1] = v.Normal.y, R.Result.Normal[2] = v.Normal.z, R.Result.Normal[3] = 0.0f;
    R.Result.Hitted = 1;
    R.Result.T = length3(v.Position - ro);
    R.Result.InstanceId = SS.nodes[node].instanceId;
    R.Result.MaterialId = v.MaterialIndex;
}

static GkStatus selectDevice(Context& c)
{
    GK_CUDA(cudaSetDevice(c.device));
    return GK_OK;
}

static GkPlane resolvePlane(const Context& c, GkPlane plane)
{
    // after a frame the reference's history images hold copies of the accumulated ones
    if (!c.pendingHistorySwap) return plane;
    switch (plane) {
    case GK_PLANE_HISTORY_DIFFUSE: return GK_PLANE_ACCUM_DIFFUSE;
    case GK_PLANE_HISTORY_SPECULAR: return GK_PLANE_ACCUM_SPECULAR;
    case GK_PLANE_HISTORY_ALBEDO: return GK_PLANE_ACCUM_ALBEDO;
    case GK_PLANE_OBJECT_ID1: return GK_PLANE_OBJECT_ID0;
    default: return plane;
    }
}

void waitAsyncCopyBeforeWriting(Context& c, const void* const* buffers, int count)
{
    if (!c.asyncCopySrc) return;
    for (int i = 0; i < count; ++i)
        if (buffers[i] == c.asyncCopySrc) {
            // device-side ordering only: the copy stays "in flight" for the host until gk_readback_wait
            // (or the next gk_readback_async) has synchronised with it
            cudaStreamWaitEvent(c.stream, c.evCopyDone, 0);
            return;
        }
}

} // namespace gk

using namespace gk;

#define GK_CHECK_CTX(ctx)                                  \
    if (!(ctx)) {                                          \
        gk::setLastError("null context");                  \
        return GK_ERR_INVALID_ARGUMENT;                    \
    }                                                      \
    Context& c = (ctx)->c;                                 \
    {                                                      \
        GkStatus _s = selectDevice(c);                     \
        if (_s != GK_OK) return _s;                        \
    }

extern "C" {

int gk_abi_version(void) { return GK_ABI_VERSION; }
const char* gk_last_error(void) { return g_lastError.c_str(); }

GkStatus gk_create(const GkConfig* cfg, GkContext** out)
{
    if (!cfg || !out || cfg->width == 0 || cfg->height == 0) {
        setLastError("gk_create: bad configuration");
        return GK_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        setLastError(std::string("gk_create: no CUDA device (") + cudaGetErrorString(e) + "); this backend has no CPU fallback");
        return GK_ERR_CUDA;
    }
    int dev = cfg->device;
    if (dev < 0) GK_CUDA(cudaGetDevice(&dev));
    if (dev >= count) {
        setLastError("gk_create: device ordinal out of range");
        return GK_ERR_INVALID_ARGUMENT;
    }
    cudaDeviceProp prop;
    GK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        setLastError(std::string("gk_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                     "; this library carries sm_100a code only");
        return GK_ERR_UNSUPPORTED;
    }
    GkContext* h = new (std::nothrow) GkContext();
    if (!h) return GK_ERR_OUT_OF_MEMORY;
    Context& c = h->c;
    c.device = dev;
    c.width = cfg->width, c.height = cfg->height;
    c.tileCount = cfg->tileCount ? cfg->tileCount : 1;
    c.tileIndex = cfg->tileIndex;
    c.tileRows = cfg->tileRows ? cfg->tileRows : 16;
    c.flags = cfg->flags;
    c.traceTileIndex = (c.flags & GK_CFG_TRACE_ALL_ROWS) ? 0u : c.tileIndex;
    c.traceTileCount = (c.flags & GK_CFG_TRACE_ALL_ROWS) ? 1u : c.tileCount;
    // measured on the 1080p room (sweep 16 K .. 4 M paths): the one-launch tail wins below ~2.5-3.5 K paths per SM (1 and 2 GPUs)
    c.tailThreshold = 3584u * (uint32_t)prop.multiProcessorCount;
    if (const char* e = getenv("GK_BLAS_LEAF")) c.blasLeafMax = (uint32_t)std::min(8, std::max(1, atoi(e)));
    if (const char* e = getenv("GK_CONCURRENT_SHADOW")) c.concurrentShadow = atoi(e) != 0;
    if (const char* e = getenv("GK_TRACE_BLOCK")) c.laneBlock = (unsigned)std::min(256, std::max(32, atoi(e) / 32 * 32));
    if (const char* e = getenv("GK_SHADE_BLOCKS")) c.shadeMinBlocks = atoi(e);
    if (const char* e = getenv("GK_TIGHT_BOUNDS")) c.tightInstanceBounds = atoi(e) != 0;
    if (const char* e = getenv("GK_COST_TRI")) c.costTri = (float)atof(e);
    if (const char* e = getenv("GK_BLAS_PLOC")) c.blasPloc = atoi(e) != 0;
    if (const char* e = getenv("GK_BLAS_PLOC_RADIUS")) c.blasPlocRadius = std::min(256, std::max(1, atoi(e)));
    if (const char* e = getenv("GK_TLAS_PLOC")) c.tlasPloc = atoi(e) != 0;
    if (const char* e = getenv("GK_TLAS_PLOC_RADIUS")) c.tlasPlocRadius = std::min(256, std::max(1, atoi(e)));
    if (const char* e = getenv("GK_TLAS_SIZE_BITS")) c.tlasSizeBits = std::min(7, std::max(0, atoi(e)));
    if (const char* e = getenv("GK_SAH_COLLAPSE")) c.sahCollapse = atoi(e) != 0;
    if (const char* e = getenv("GK_TAIL_FRACTION")) c.tailFraction = (float)atof(e);
    if (const char* e = getenv("GK_TAIL_THRESHOLD")) c.tailThreshold = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("GK_COOP_THRESHOLD")) c.coopThreshold = (uint32_t)strtoul(e, nullptr, 10); // tuning / test hook
    if (const char* e = getenv("GK_TRACE_VARIANT")) c.traceVariant = atoi(e);
    if (c.traceVariant != 1 && !getenv("GK_COOP_THRESHOLD")) c.coopThreshold = 65536u;
    if (const char* e = getenv("GK_SCHED_REFILL_MIN")) c.schedRefillMin = (uint32_t)std::min(32, std::max(1, atoi(e)));
    if (const char* e = getenv("GK_SCHED_BIAS_NODE")) c.schedBiasN = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("GK_SCHED_MIN_RAYS")) c.schedMinRays = (uint32_t)strtoul(e, nullptr, 10);
    if (c.tileIndex >= c.tileCount) {
        delete h;
        setLastError("gk_create: tileIndex >= tileCount");
        return GK_ERR_INVALID_ARGUMENT;
    }
    {
        cudaError_t ce = cudaSetDevice(dev);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking);
        if (ce != cudaSuccess) {
            setLastError(std::string("gk_create: ") + cudaGetErrorString(ce));
            c.stream = nullptr;
            delete h;
            return GK_ERR_CUDA;
        }
    }
    GkStatus s = allocFrameResources(c);
    if (s != GK_OK) {
        gk_destroy(h);
        return s;
    }
    *out = h;
    return GK_OK;
}

void gk_destroy(GkContext* ctx)
{
    if (!ctx) return;
    Context& c = ctx->c;
    cudaSetDevice(c.device);
    if (c.stream) cudaStreamSynchronize(c.stream);
    freeFrameResources(c);
    c.blasTree.release(), c.tlasTree.release();
    c.dModels.release(), c.dGpuVerts.release(), c.dIndices.release(), c.dMaterials.release(), c.dLights.release(), c.dFaceNormals.release();
    c.dNodes.release(), c.dSparseNodes.release(), c.dSparseIdx.release(), c.dCubes.release(), c.dVoxels.release(), c.dCubesPrev.release(), c.dVoxelsPrev.release(), c.dTris.release(), c.dBlasNodes.release(), c.dTlasNodes.release(), c.dBlasSrc.release(), c.dTlasSrc.release(), c.dInst.release();
    c.dTaskA.release(), c.dTaskB.release(), c.dCounters.release(), c.dSortTemp.release(), c.dGroupLo.release(), c.dGroupHi.release(), c.dGroupRoot.release(), c.dRootRef.release();
    for (int k = 0; k < 2; ++k) c.dPlocRef[k].release(), c.dPlocLo[k].release(), c.dPlocHi[k].release();
    c.dPlocNn.release(), c.dPlocValid.release(), c.dPlocPos.release();
    c.dPlocGrp[0].release(), c.dPlocGrp[1].release(), c.dPlocLeafGrp.release(), c.dPlocGroupBase.release();
    c.dCapture.release();
    for (cudaEvent_t e : c.evPool) cudaEventDestroy(e);
    if (c.evFork) cudaEventDestroy(c.evFork);
    if (c.evJoin) cudaEventDestroy(c.evJoin);
    if (c.stream2) cudaStreamDestroy(c.stream2);
    if (c.evCopyReady) cudaEventDestroy(c.evCopyReady);
    if (c.evCopyDone) cudaEventDestroy(c.evCopyDone);
    if (c.copyStream) cudaStreamDestroy(c.copyStream);
    if (c.stream) cudaStreamDestroy(c.stream);
    delete ctx;
}

GkStatus gk_resize(GkContext* ctx, uint32_t width, uint32_t height)
{
    GK_CHECK_CTX(ctx);
    if (width == 0 || height == 0) {
        setLastError("gk_resize: empty extent");
        return GK_ERR_INVALID_ARGUMENT;
    }
    GK_CUDA(cudaStreamSynchronize(c.stream));
    c.width = width, c.height = height;
    c.pendingHistorySwap = false;
    return allocFrameResources(c);
}

GkStatus gk_upload_scene(GkContext* ctx, const GkSceneDesc* scene)
{
    GK_CHECK_CTX(ctx);
    if (!scene) {
        setLastError("gk_upload_scene: null scene");
        return GK_ERR_INVALID_ARGUMENT;
    }
    return uploadScene(c, *scene);
}

GkStatus gk_update_materials(GkContext* ctx, const GkMaterial* materials, uint32_t count)
{
    GK_CHECK_CTX(ctx);
    if (!materials || count == 0) {
        setLastError("gk_update_materials: empty material list");
        return GK_ERR_INVALID_ARGUMENT;
    }
    GK_CUDA(cudaStreamSynchronize(c.stream));
    GK_CUDA(c.dMaterials.reserve(count));
    GK_CUDA(cudaMemcpyAsync(c.dMaterials.p, materials, sizeof(GkMaterial) * count, cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    c.materialCount = count;
    return GK_OK;
}

GkStatus gk_update_instances(GkContext* ctx, const GkNodeProxy* nodes, uint32_t count, int refit)
{
    GK_CHECK_CTX(ctx);
    return updateInstances(c, nodes, count, refit != 0);
}

GkStatus gk_update_instances_sparse(GkContext* ctx, const uint32_t* indices, const GkNodeProxy* proxies, uint32_t changed, int refit)
{
    GK_CHECK_CTX(ctx);
    return updateInstancesSparse(c, indices, proxies, changed, refit != 0);
}

GkStatus gk_set_probes(GkContext* ctx, const GkAmbientCube* cubes, const GkVoxelData* voxels, size_t count)
{
    GK_CHECK_CTX(ctx);
    if (!cubes || !voxels || count == 0) {
        c.haveProbes = false;
        return GK_OK;
    }
    if (count != (size_t)GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z) {
        setLastError("gk_set_probes: the probe grid is 192 x 48 x 192");
        return GK_ERR_INVALID_ARGUMENT;
    }
    GK_CUDA(c.dCubes.reserve(count));
    GK_CUDA(c.dVoxels.reserve(count));
    GK_CUDA(cudaMemcpyAsync(c.dCubes.p, cubes, sizeof(GkAmbientCube) * count, cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaMemcpyAsync(c.dVoxels.p, voxels, sizeof(GkVoxelData) * count, cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    c.haveProbes = true;
    return GK_OK;
}

GkStatus gk_bake_probes(GkContext* ctx, uint32_t first_probe, uint32_t count)
{
    GK_CHECK_CTX(ctx);
    return bakeProbes(c, first_probe, count);
}

GkStatus gk_get_probes(GkContext* ctx, GkAmbientCube* cubes, GkVoxelData* voxels, size_t count)
{
    GK_CHECK_CTX(ctx);
    return getProbes(c, cubes, voxels, count);
}

GkStatus gk_set_ubo(GkContext* ctx, const GkUniformBufferObject* ubo)
{
    GK_CHECK_CTX(ctx);
    if (!ubo) {
        setLastError("gk_set_ubo: null UBO");
        return GK_ERR_INVALID_ARGUMENT;
    }
    if (ubo->NumberOfSamples > 65535u || ubo->NumberOfBounces > 127u || ubo->MaxNumberOfBounces > 127u) {
        setLastError("gk_set_ubo: samples <= 65535 and bounces <= 127 are supported");
        return GK_ERR_UNSUPPORTED;
    }
    c.ubo = *ubo;
    c.haveUbo = true;
    return GK_OK;
}

GkStatus gk_trace_frame(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    return traceFrame(c);
}

GkStatus gk_filter_frame(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    return filterFrame(c);
}

GkStatus gk_render_frame(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    GkStatus s = traceFrame(c);
    if (s != GK_OK) return s;
    const float traceMs = c.stats.msTotal;
    s = filterFrame(c);
    c.stats.msTotal = traceMs + c.stats.msReproject + c.stats.msDenoise;
    c.frameIndex++;
    return s;
}

GkStatus gk_intersect_device(GkContext* ctx, const void* d_rays, uint32_t count, void* d_out_tuv, void* d_out_ids, int anyHit)
{
    GK_CHECK_CTX(ctx);
    return intersectDevice(c, (const float4*)d_rays, count, (float*)d_out_tuv, (uint32_t*)d_out_ids, anyHit != 0);
}

GkStatus gk_intersect(GkContext* ctx, const float* rays, uint32_t count, float* out_tuv, uint32_t* out_ids)
{
    GK_CHECK_CTX(ctx);
    if (count == 0) return GK_OK;
    if (!rays) {
        setLastError("gk_intersect: null rays");
        return GK_ERR_INVALID_ARGUMENT;
    }
    DevBuf<float4> dr;
    DevBuf<float> dt;
    DevBuf<uint32_t> di;
    GK_CUDA(dr.reserve(2 * (size_t)count));
    GK_CUDA(dt.reserve(3 * (size_t)count));
    GK_CUDA(di.reserve(2 * (size_t)count));
    GK_CUDA(cudaMemcpyAsync(dr.p, rays, 32 * (size_t)count, cudaMemcpyHostToDevice, c.stream));
    GkStatus s = intersectDevice(c, dr.p, count, dt.p, di.p, false);
    if (s == GK_OK) {
        if (out_tuv) GK_CUDA(cudaMemcpyAsync(out_tuv, dt.p, 12 * (size_t)count, cudaMemcpyDeviceToHost, c.stream));
        if (out_ids) GK_CUDA(cudaMemcpyAsync(out_ids, di.p, 8 * (size_t)count, cudaMemcpyDeviceToHost, c.stream));
        s = checkTraversalOverflow(c);
    }
    dr.release(), dt.release(), di.release();
    return s;
}

GkStatus gk_raycast(GkContext* ctx, const float* origin_dir, uint32_t count, GkRayCastResult* out)
{
    GK_CHECK_CTX(ctx);
    if (count == 0) return GK_OK;
    if (!origin_dir || !out) {
        setLastError("gk_raycast: null argument");
        return GK_ERR_INVALID_ARGUMENT;
    }
    if (!c.haveScene || !c.haveInstances) {
        setLastError("gk_raycast: scene and instances must be set first");
        return GK_ERR_NOT_READY;
    }
    DevBuf<float> dod, dt;
    DevBuf<float4> dr;
    DevBuf<uint32_t> di;
    DevBuf<GkRayCastResult> dres;
    GK_CUDA(dod.reserve(6 * (size_t)count));
    GK_CUDA(dr.reserve(2 * (size_t)count));
    GK_CUDA(dt.reserve(3 * (size_t)count));
    GK_CUDA(di.reserve(2 * (size_t)count));
    GK_CUDA(dres.reserve(count));
    GK_CUDA(cudaMemcpyAsync(dod.p, origin_dir, 24 * (size_t)count, cudaMemcpyHostToDevice, c.stream));
    k_pack_rays<<<(count + 255) / 256, 256, 0, c.stream>>>(dod.p, count, dr.p);
    GkStatus s = intersectDevice(c, dr.p, count, dt.p, di.p, false);
    if (s == GK_OK) {
        k_raycast_results<<<(count + 255) / 256, 256, 0, c.stream>>>(dod.p, dt.p, di.p, count, c.dNodes.p, c.dModels.p, c.dFaceNormals.p, dres.p);
        GK_CUDA(cudaGetLastError());
        GK_CUDA(cudaMemcpyAsync(out, dres.p, sizeof(GkRayCastResult) * (size_t)count, cudaMemcpyDeviceToHost, c.stream));
        s = checkTraversalOverflow(c);
    }
    dod.release(), dr.release(), dt.release(), di.release(), dres.release();
    return s;
}

GkStatus gk_raycast_task(GkContext* ctx, GkRayCastIO* io, uint32_t count)
{
    GK_CHECK_CTX(ctx);
    if (count == 0) return GK_OK;
    if (!io) {
        setLastError("gk_raycast_task: null argument");
        return GK_ERR_INVALID_ARGUMENT;
    }
    if (!c.haveScene || !c.haveInstances) {
        setLastError("gk_raycast_task: scene and instances must be set first");
        return GK_ERR_NOT_READY;
    }
    DevBuf<GkRayCastIO> dio;
    DevBuf<float4> dr;
    DevBuf<float> dt;
    DevBuf<uint32_t> di;
    GK_CUDA(dio.reserve(count));
    GK_CUDA(dr.reserve(2 * (size_t)count));
    GK_CUDA(dt.reserve(3 * (size_t)count));
    GK_CUDA(di.reserve(2 * (size_t)count));
    GK_CUDA(cudaMemcpyAsync(dio.p, io, sizeof(GkRayCastIO) * (size_t)count, cudaMemcpyHostToDevice, c.stream));
    k_task_pack<<<(count + 255) / 256, 256, 0, c.stream>>>(dio.p, count, dr.p);
    GkStatus s = intersectDevice(c, dr.p, count, dt.p, di.p, false);
    if (s != GK_OK) return s;
    ShadeScene SS;
    SS.verts = c.dGpuVerts.p, SS.indices = c.dIndices.p, SS.models = c.dModels.p, SS.materials = c.dMaterials.p, SS.nodes = c.dNodes.p, SS.inst = c.dInst.p;
    SS.cubes = nullptr, SS.voxels = nullptr, SS.materialCount = c.materialCount;
    k_task_results<<<(count + 255) / 256, 256, 0, c.stream>>>(dio.p, dt.p, di.p, count, SS);
    GK_CUDA(cudaGetLastError());
    GK_CUDA(cudaMemcpyAsync(io, dio.p, sizeof(GkRayCastIO) * (size_t)count, cudaMemcpyDeviceToHost, c.stream));
    return checkTraversalOverflow(c);
}

size_t gk_plane_bytes(const GkContext* ctx, GkPlane plane)
{
    if (!ctx || plane < 0 || plane >= GK_PLANE_COUNT) return 0;
    return ctx->c.planes.bytes[plane];
}

GkStatus gk_readback(GkContext* ctx, GkPlane plane, void* dst, size_t bytes)
{
    GK_CHECK_CTX(ctx);
    if (plane < 0 || plane >= GK_PLANE_COUNT || !dst || bytes != c.planes.bytes[plane]) {
        setLastError("gk_readback: bad plane or size");
        return GK_ERR_INVALID_ARGUMENT;
    }
    GK_CUDA(cudaMemcpyAsync(dst, c.planes.p[resolvePlane(c, plane)], bytes, cudaMemcpyDeviceToHost, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return GK_OK;
}

GkStatus gk_readback_async(GkContext* ctx, GkPlane plane, void* dst, size_t bytes)
{
    GK_CHECK_CTX(ctx);
    if (plane < 0 || plane >= GK_PLANE_COUNT || !dst || bytes != c.planes.bytes[plane]) {
        setLastError("gk_readback_async: bad plane or size");
        return GK_ERR_INVALID_ARGUMENT;
    }
    if (!c.copyStream) {
        GK_CUDA(cudaStreamCreateWithFlags(&c.copyStream, cudaStreamNonBlocking));
        GK_CUDA(cudaEventCreateWithFlags(&c.evCopyReady, cudaEventDisableTiming));
        GK_CUDA(cudaEventCreateWithFlags(&c.evCopyDone, cudaEventDisableTiming));
    }
    if (c.asyncCopySrc) GK_CUDA(cudaEventSynchronize(c.evCopyDone)); // one copy in flight
    const void* src = c.planes.p[resolvePlane(c, plane)];
    GK_CUDA(cudaEventRecord(c.evCopyReady, c.stream));
    GK_CUDA(cudaStreamWaitEvent(c.copyStream, c.evCopyReady, 0));
    GK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.copyStream));
    GK_CUDA(cudaEventRecord(c.evCopyDone, c.copyStream));
    c.asyncCopySrc = src;
    return GK_OK;
}

GkStatus gk_readback_wait(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    if (c.asyncCopySrc) {
        GK_CUDA(cudaEventSynchronize(c.evCopyDone));
        c.asyncCopySrc = nullptr;
    }
    return GK_OK;
}

GkStatus gk_upload_plane(GkContext* ctx, GkPlane plane, const void* src, size_t bytes)
{
    GK_CHECK_CTX(ctx);
    if (plane < 0 || plane >= GK_PLANE_COUNT || !src || bytes != c.planes.bytes[plane]) {
        setLastError("gk_upload_plane: bad plane or size");
        return GK_ERR_INVALID_ARGUMENT;
    }
    applyPendingHistorySwap(c);
    {
        const void* bufs[] = {c.planes.p[plane]};
        waitAsyncCopyBeforeWriting(c, bufs, 1); // an asynchronous read-back may still be reading this plane
    }
    GK_CUDA(cudaMemcpyAsync(c.planes.p[plane], src, bytes, cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return GK_OK;
}

void* gk_plane_device(GkContext* ctx, GkPlane plane)
{
    if (!ctx || plane < 0 || plane >= GK_PLANE_COUNT) return nullptr;
    return ctx->c.planes.p[resolvePlane(ctx->c, plane)];
}

size_t gk_exchange_bytes(const GkContext* ctx) { return ctx ? exchangeBytesPerRank(ctx->c) : 0; }

GkStatus gk_exchange_pack(GkContext* ctx, void* d_staging)
{
    GK_CHECK_CTX(ctx);
    if (!d_staging) return GK_ERR_INVALID_ARGUMENT;
    return exchangePack(c, d_staging);
}

GkStatus gk_exchange_unpack(GkContext* ctx, const void* d_all)
{
    GK_CHECK_CTX(ctx);
    if (!d_all) return GK_ERR_INVALID_ARGUMENT;
    return exchangeUnpack(c, d_all);
}

GkStatus gk_exchange_ipc_handles(GkContext* ctx, void* out, size_t bytes)
{
    GK_CHECK_CTX(ctx);
    return exchangeIpcHandles(c, out, bytes);
}

GkStatus gk_exchange_open_peers(GkContext* ctx, const void* handles_all, uint32_t world)
{
    GK_CHECK_CTX(ctx);
    return exchangeOpenPeers(c, handles_all, world);
}

GkStatus gk_filter_frame_owned(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    return filterFrameOwnedRows(c);
}

GkStatus gk_exchange_push_final(GkContext* ctx, int dst_rank)
{
    GK_CHECK_CTX(ctx);
    return exchangePushFinal(c, dst_rank);
}

GkStatus gk_frame_shard_handle(GkContext* ctx, void* out, size_t bytes)
{
    GK_CHECK_CTX(ctx);
    return frameShardHandle(c, out, bytes);
}

GkStatus gk_frame_shard_open(GkContext* ctx, const void* handles_all, uint32_t world)
{
    GK_CHECK_CTX(ctx);
    return frameShardOpen(c, handles_all, world);
}

GkStatus gk_frame_shard_push(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    return frameShardPush(c);
}

GkStatus gk_frame_shard_accumulate(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    return frameShardAccumulate(c);
}

GkStatus gk_exchange_close_peers(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    GK_CUDA(cudaStreamSynchronize(c.stream));
    exchangeClosePeers(c);
    frameShardClosePeers(c);
    return GK_OK;
}

GkStatus gk_exchange_push(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    return exchangePush(c);
}

void* gk_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void gk_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

GkStatus gk_synchronize(GkContext* ctx)
{
    GK_CHECK_CTX(ctx);
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return GK_OK;
}

GkStatus gk_get_stats(GkContext* ctx, GkFrameStats* out)
{
    GK_CHECK_CTX(ctx);
    if (!out) return GK_ERR_INVALID_ARGUMENT;
    GK_CUDA(cudaStreamSynchronize(c.stream));
    if (c.travStats) {
        TraversalStats h;
        GK_CUDA(cudaMemcpy(&h, c.dTravStats, sizeof(h), cudaMemcpyDeviceToHost));
        c.stats.nodeVisits = h.nodeVisits, c.stats.triTests = h.triTests, c.stats.tlasVisits = h.tlasVisits, c.stats.instanceEntries = h.instanceEntries, c.stats.maxStack = (uint32_t)h.maxStack;
    }
    *out = c.stats;
    return GK_OK;
}

GkStatus gk_get_bvh_info(GkContext* ctx, GkBvhInfo* out)
{
    GK_CHECK_CTX(ctx);
    if (!out) return GK_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof(*out));
    out->blasCount = (uint32_t)c.models.size();
    out->instanceCount = c.nodeCount;
    out->triangleCount = c.totalTris;
    out->instancedTriangles = c.instancedTris;
    out->blasNodes2 = c.blasTree.n ? c.blasTree.n - 1 : 0;
    out->blasNodes8 = c.blasNodeCount;
    out->tlasNodes2 = c.tlasTree.n ? c.tlasTree.n - 1 : 0;
    out->tlasNodes8 = c.tlasNodeCount;
    out->bytesGeometry = c.totalTris * sizeof(TriRecord) + c.dGpuVerts.bytes() + c.dIndices.bytes();
    out->bytesBvh = (uint64_t)(c.blasNodeCount + c.tlasNodeCount) * sizeof(WideNode) + (uint64_t)c.nodeCount * sizeof(InstRecord);
    out->msBlasBuild = c.msBlasBuild, out->msTlasBuild = c.msTlasBuild, out->msRefit = c.msRefit;
    out->refitsRejected = c.refitRejected, out->tlasAreaAtBuild = c.tlasAreaAtBuild;
    return GK_OK;
}

GkStatus gk_set_option(GkContext* ctx, const char* name, double value)
{
    GK_CHECK_CTX(ctx);
    if (!name) return GK_ERR_INVALID_ARGUMENT;
    const std::string n(name);
    if (n == "trace_variant") {
        c.traceVariant = (int)value;
        c.coopThreshold = c.traceVariant == 1 ? 262144u : 65536u; // measured optimum of each wave loop
    }
    else if (n == "sched_refill_min") c.schedRefillMin = (uint32_t)std::min(32.0, std::max(1.0, value));
    else if (n == "sched_bias_node") c.schedBiasN = (uint32_t)std::max(0.0, value);
    else if (n == "coop_divisor") c.coopDivisor = (uint32_t)std::max(1.0, value);
    else if (n == "tail_divisor") c.tailDivisor = (uint32_t)std::max(1.0, value);
    else if (n == "micro_tiles") c.microTiles = (int)value;
    else if (n == "sched_min_rays") c.schedMinRays = (uint32_t)std::max(0.0, value);
    else if (n == "sched_keep_node") c.schedKeepN = (uint32_t)std::min(33.0, std::max(1.0, value));
    else if (n == "sched_keep_tri") c.schedKeepT = (uint32_t)std::min(33.0, std::max(1.0, value));
    else if (n == "coop_threshold") c.coopThreshold = (uint32_t)std::max(0.0, value);
    else if (n == "primary_lane_kernel") c.primaryLaneKernel = value != 0;
    else if (n == "stream_tail_paths") c.streamTailPaths = (uint32_t)std::max(0.0, value);
    else if (n == "tail_coop") c.tailCoop = value != 0;
    else if (n == "tail_threshold") c.tailThreshold = (uint32_t)std::max(0.0, value);
    else if (n == "tail_fraction") c.tailFraction = (float)value;
    else if (n == "concurrent_shadow") c.concurrentShadow = value != 0;
    else if (n == "wave_lookahead") c.waveLookahead = (uint32_t)std::max(0.0, value);
    else {
        setLastError("gk_set_option: unknown option '" + n + "'");
        return GK_ERR_INVALID_ARGUMENT;
    }
    return GK_OK;
}

GkStatus gk_measure_read_bandwidth(GkContext* ctx, size_t bytes, int reps, float* out_gbps)
{
    GK_CHECK_CTX(ctx);
    if (!out_gbps || bytes < (1u << 20) || reps < 1) {
        setLastError("gk_measure_read_bandwidth: bytes >= 1 MiB, reps >= 1");
        return GK_ERR_INVALID_ARGUMENT;
    }
    DevBuf<uint4> buf;
    const size_t n16 = bytes / 16;
    GK_CUDA(buf.reserve(n16 + 1));
    GK_CUDA(cudaMemsetAsync(buf.p, 0x5a, n16 * 16, c.stream));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
    const unsigned grid = (unsigned)sms * 8u;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k_stream_read<<<grid, 256, 0, c.stream>>>(buf.p, n16, 2, buf.p + n16); // warm-up: brings the buffer into L2 when it fits
    float best = 0.f;
    for (int k = 0; k < 5; ++k) {
        cudaEventRecord(e0, c.stream);
        k_stream_read<<<grid, 256, 0, c.stream>>>(buf.p, n16, reps, buf.p + n16);
        cudaEventRecord(e1, c.stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0), cudaEventDestroy(e1), buf.release();
            setLastError(std::string("gk_measure_read_bandwidth: ") + cudaGetErrorString(e));
            return GK_ERR_CUDA;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms > 0) best = std::max(best, (float)((double)n16 * 16.0 * reps / (ms * 1e-3) / 1e9));
    }
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    buf.release();
    *out_gbps = best;
    return GK_OK;
}

GkStatus gk_set_traversal_stats(GkContext* ctx, int enabled)
{
    GK_CHECK_CTX(ctx);
    c.travStats = enabled != 0;
    GK_CUDA(cudaMemsetAsync(c.dTravStats, 0, sizeof(TraversalStats), c.stream));
    return GK_OK;
}

void* gk_stream(GkContext* ctx) { return ctx ? (void*)ctx->c.stream : nullptr; }

GkStatus gk_set_ray_capture(GkContext* ctx, int wave)
{
    GK_CHECK_CTX(ctx);
    c.captureWave = wave;
    return GK_OK;
}

GkStatus gk_get_captured_rays(GkContext* ctx, float* rays, uint32_t capacity, uint32_t* count)
{
    GK_CHECK_CTX(ctx);
    if (!count) return GK_ERR_INVALID_ARGUMENT;
    const uint32_t n = c.capturedCount < capacity ? c.capturedCount : capacity;
    *count = c.capturedCount;
    if (rays && n) {
        GK_CUDA(cudaMemcpyAsync(rays, c.dCapture.p, 32 * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
        GK_CUDA(cudaStreamSynchronize(c.stream));
    }
    return GK_OK;
}

} // extern "C"
