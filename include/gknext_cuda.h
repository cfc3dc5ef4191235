/*
 * gknext_cuda.h — C ABI of the B200 (sm_100a) CUDA backend for gkNextRenderer's
 * path-tracing hot path.  This is the drop-in boundary: a logic renderer registered behind
 * the reference's renderer switch forwards its five virtuals to these entry points
 * (INTEGRATION.md shows the shim).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every call returns GK_OK (0) or a negative GkStatus; gk_last_error() gives the text.
 *     Nothing throws across this boundary (the reference throws C++ exceptions,
 *     src/Utilities/Exception.hpp:10-16; the shim converts).
 *   - one caller thread per context (the reference drives its renderer from the main thread,
 *     src/Rendering/VulkanBaseRenderer.cpp:889-1030).  Calls enqueue work on the context's
 *     CUDA stream; gk_readback / gk_synchronize / gk_get_stats wait for it.
 *   - the caller keeps ownership of every host pointer; the backend copies during the call
 *     (the reference frees its CPU vertex arrays right after upload, src/Assets/Scene.cpp:195).
 *   - there is no CPU fallback: every entry point fails with GK_ERR_CUDA if no sm_100 device
 *     is usable.
 */
#ifndef GKNEXT_CUDA_H_
#define GKNEXT_CUDA_H_

#include "gknext_types.h"
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GK_ABI_VERSION 1
#define GK_TRAVERSAL_STACK 48 /* entries of the per-ray traversal stack */

typedef enum GkStatus {
    GK_OK = 0,
    GK_ERR_INVALID_ARGUMENT = -1,
    GK_ERR_CUDA = -2,
    GK_ERR_OUT_OF_MEMORY = -3,
    GK_ERR_NOT_READY = -4, /* e.g. render before a scene/instances/UBO were supplied */
    GK_ERR_UNSUPPORTED = -5
} GkStatus;

typedef struct GkContext GkContext;

typedef struct GkConfig {
    int32_t device;          /* CUDA ordinal, -1 = current device */
    uint32_t width, height;  /* render extent (SwapChain().RenderExtent()) */
    /* Multi-GPU frame partition.  With tileCount > 1 this context traces only
     * the image rows r with (r / tileRows) % tileCount == tileIndex; the other rows of the
     * path-tracer output planes are left untouched for the frame-end exchange. */
    uint32_t tileIndex, tileCount, tileRows;
    uint32_t flags;          /* GkConfigFlags */
    uint32_t reserved[6];
} GkConfig;

enum GkConfigFlags {
    GK_CFG_DEFAULT = 0,
    /* The path tracer covers the whole image although the context owns only its tile rows for the filters and the
     * exchange: frame-sharded progressive rendering (gk_frame_shard_*), where every rank traces a different frame. */
    GK_CFG_TRACE_ALL_ROWS = 1u << 0
};

/* Image planes, named after the reference's render targets
 * (src/Rendering/VulkanBaseRenderer.cpp:488-543, bindings :571-584). */
typedef enum GkPlane {
    GK_PLANE_OUTPUT_DIFFUSE = 0,   /* rtOutputDiffuse     RGBA16F  8 B/px */
    GK_PLANE_OUTPUT_SPECULAR = 1,  /* rtOutputSpecular    RGBA16F */
    GK_PLANE_ALBEDO = 2,           /* rtAlbedo_           RGBA16F */
    GK_PLANE_NORMAL = 3,           /* rtNormal_           RGBA16F */
    GK_PLANE_OBJECT_ID0 = 4,       /* rtObject0           R32_UINT 4 B/px */
    GK_PLANE_OBJECT_ID1 = 5,       /* rtObject1           R32_UINT (previous frame) */
    GK_PLANE_MOTION = 6,           /* rtMotionVector_     RG32F    8 B/px */
    GK_PLANE_DEPTH = 7,            /* rtPrevDepth         R32F */
    GK_PLANE_ACCUM_DIFFUSE = 8,    /* rtAccumlatedDiffuse RGBA16F */
    GK_PLANE_ACCUM_SPECULAR = 9,   /* rtAccumlatedSpecular */
    GK_PLANE_ACCUM_ALBEDO = 10,    /* rtAccumlatedAlbedo_ */
    GK_PLANE_HISTORY_DIFFUSE = 11, /* rtPingPong0 */
    GK_PLANE_HISTORY_SPECULAR = 12,/* rtPingPong1 */
    GK_PLANE_HISTORY_ALBEDO = 13,  /* rtPingPong3 */
    GK_PLANE_DENOISED = 14,        /* rtDenoised          RGBA16F (tonemapped final image) */
    /* un-quantised integrator outputs kept for parity checks */
    GK_PLANE_RADIANCE_DIFFUSE_F32 = 15, /* 4 x f32 */
    GK_PLANE_RADIANCE_SPECULAR_F32 = 16,/* 4 x f32 */
    GK_PLANE_PRIMARY_IDS = 17,     /* 2 x u32: {triangle, node-proxy index}, 0xffffffff = miss */
    GK_PLANE_PRIMARY_T = 18,       /* f32 hit distance of the primary ray */
    GK_PLANE_RAY_COUNT = 19,       /* u32 rays traced for the pixel this frame */
    GK_PLANE_COUNT = 20
} GkPlane;

typedef struct GkFrameStats {
    uint64_t primaryRays;   /* camera rays traced */
    uint64_t extensionRays; /* closest-hit bounce rays */
    uint64_t shadowRays;    /* any-hit rays */
    uint32_t waves;         /* extend/shade iterations of the wavefront loop */
    uint32_t launches;      /* kernels of this library launched for the frame */
    float msTotal;          /* whole gk_render_frame on the device */
    float msBvh;            /* instance update + TLAS build/refit, if it ran this frame */
    float msGenerate, msExtend, msShade, msShadow, msAccumulate; /* integrator */
    float msReproject, msDenoise;
    uint64_t nodeVisits, triTests; /* only when traversal statistics are enabled */
    uint64_t tlasVisits, instanceEntries; /* ditto: node visits in the TLAS, ray -> instance transitions */
    float msTail;       /* the single launch that finishes the last paths of the frame */
    uint32_t tailPaths; /* paths alive when that launch started */
    uint64_t tailExtensionRays, tailShadowRays; /* part of extensionRays / shadowRays traced by that launch */
    uint32_t maxStack;  /* traversal statistics: deepest stack seen; GK_TRAVERSAL_STACK + 1 means an entry was dropped */
    float msTrace;      /* extend + shadow kernels of all waves as one span per wave (the two run concurrently) */
    /* traversal statistics of the scheduled kernel (gk_set_traversal_stats): steps the warps ran per class
     * (0 node visit, 1 triangle test, 2 instance entry) and the lanes that took part; lanes / (32 * iters) is the
     * SIMD occupancy of that step */
    uint64_t schedIters[3], schedLanes[3];
    uint64_t schedRefills, schedRefillLanes; /* queue refills per warp and rays fetched by them */
    uint64_t schedPopIters, schedPopLanes;   /* votes taken by the warps, lanes alive at those votes */
} GkFrameStats;

typedef struct GkBvhInfo {
    uint32_t blasCount, instanceCount;
    uint64_t triangleCount;       /* unique triangles over all BLAS */
    uint64_t instancedTriangles;  /* sum over ray-visible instances */
    uint32_t blasNodes2, blasNodes8, tlasNodes2, tlasNodes8;
    uint64_t bytesGeometry, bytesBvh;
    float msBlasBuild, msTlasBuild, msRefit;
    uint32_t refitsRejected; /* refits that loosened the TLAS beyond the growth limit and were rebuilt instead */
    float tlasAreaAtBuild;   /* summed internal-node surface area of the TLAS at its last build */
} GkBvhInfo;

int gk_abi_version(void);
const char* gk_last_error(void);

/* Replaces LogicRendererBase::OnDeviceSet + CreateSwapChain(extent)
 * (src/Rendering/VulkanBaseRenderer.hpp:245-246; PathTracingRenderer.cpp:66-77). */
GkStatus gk_create(const GkConfig* cfg, GkContext** out);
/* Replaces DeleteSwapChain + destructor (PathTracingRenderer.cpp:79-95). */
void gk_destroy(GkContext* ctx);
/* Swap-chain resize without dropping the scene. */
GkStatus gk_resize(GkContext* ctx, uint32_t width, uint32_t height);

/* Replaces Scene::RebuildMeshBuffer's device upload (src/Assets/Scene.cpp:101-270) and
 * RayTraceBaseRenderer::CreateBottomLevelStructures (RayTraceBaseRenderer.cpp:298-362):
 * converts vertices to the fp16 GPUVertex layout, de-indexes fp32 triangles and builds one
 * BLAS per model on the GPU. */
GkStatus gk_upload_scene(GkContext* ctx, const GkSceneDesc* scene);
/* Replaces Scene::UpdateMaterial (Scene.cpp:335-350). */
GkStatus gk_update_materials(GkContext* ctx, const GkMaterial* materials, uint32_t count);
/* Replaces Scene::UpdateNodesGpuDriven's upload (Scene.cpp:464-511) plus
 * RayTraceBaseRenderer::AfterUpdateScene + the per-frame TLAS rebuild
 * (RayTraceBaseRenderer.cpp:176-228).  `refit` != 0 keeps the TLAS topology and only refits
 * boxes (valid when the instance count is unchanged); 0 rebuilds it. */
GkStatus gk_update_instances(GkContext* ctx, const GkNodeProxy* nodes, uint32_t count, int refit);
/* The same for a frame in which few nodes changed: proxies[k] replaces the record at indices[k] of the array a previous
 * gk_update_instances left on the device (the instance count stays), then the instance records, world boxes and the TLAS are
 * updated as above.  The reference rewrites and re-uploads every proxy on a dirty frame (Scene.cpp:464-511); a MagicaLego
 * frame moves a handful of bricks (MagicaLegoGameInstance.cpp:746-807): 1 % of 200 000 proxies are 0.4 MB instead of 41 MB. */
GkStatus gk_update_instances_sparse(GkContext* ctx, const uint32_t* indices, const GkNodeProxy* proxies, uint32_t changed, int refit);
/* Optional probe grid for the path terminator (Scene.cpp:258-259 buffers; AmbientCube.slang).
 * NULL (the default) means un-baked probes, i.e. all zero. */
GkStatus gk_set_probes(GkContext* ctx, const GkAmbientCube* cubes, const GkVoxelData* voxels, size_t count);

/* Replaces the probe-bake dispatch of RayTraceBaseRenderer::PostRender (RayTraceBaseRenderer.cpp:244-291: every frame a slice
 * of groupPerFrame x 64 probes starting at offsetInCubes) with Bake.HwAmbientCube.comp.slang:29-46 /
 * FGpuProbeGenerator::Render (common/AmbientCube.slang:574-629): classifies probes [first_probe, first_probe + count) of the
 * 192 x 48 x 192 grid against the scene (six axis rays, eight diagonal rays) and gathers direct / bounced light into the
 * faces of those near a surface (6 x (16 rays + light segment + sun ray)), blending with weight 1/8 into the stored RGB10A2
 * values.  Works on the probe buffers of gk_set_probes (all zero when none were uploaded); uses the UBO of gk_set_ubo (sky,
 * sun, LightCount).  Gathers read the probe state as it was before the call.  gk_get_probes copies the grid back. */
GkStatus gk_bake_probes(GkContext* ctx, uint32_t first_probe, uint32_t count);
GkStatus gk_get_probes(GkContext* ctx, GkAmbientCube* cubes, GkVoxelData* voxels, size_t count);

/* Replaces UniformBuffer::SetValue for the frame (VulkanBaseRenderer.cpp:1370-1374). */
GkStatus gk_set_ubo(GkContext* ctx, const GkUniformBufferObject* ubo);

/* Replaces PathTracingRenderer::Render (PathTracingRenderer.cpp:97-231): rt pass, the three
 * reproject passes, the compose (JBF) pass and the history copy, then ObjectId0 -> ObjectId1
 * (VulkanBaseRenderer.cpp:1287-1303). */
GkStatus gk_render_frame(GkContext* ctx);
/* Individual stages, for the multi-GPU compositor (trace on every rank, exchange, filter)
 * and for the filter-only benchmark. */
GkStatus gk_trace_frame(GkContext* ctx);   /* Core.PathTracing only */
GkStatus gk_filter_frame(GkContext* ctx);  /* ReProject x3 + DenoiseJBF + history/id copy */

/* Replaces NextEngine::RayCastGPU -> FCPUAccelerationStructure::RayCastInCPU
 * (src/Runtime/Engine.cpp:647-653; CPUAccelerationStructure.cpp:283-307), batched as in
 * Task.RayCast.comp.slang.  origin_dir: 6 floats per ray. */
GkStatus gk_raycast(GkContext* ctx, const float* origin_dir, uint32_t count, GkRayCastResult* out);
/* Replaces the GPU ray-cast task: NextEngine::RayCastGPU queues requests (src/Runtime/Engine.cpp:647-653), RayCastBuffer holds
 * them as RayCastIO records (src/Assets/UniformBuffer.hpp:81-114) and Task.RayCast.comp.slang:31-55 answers them in place, one
 * thread per record: FHardwareRayTracer::TraceRay(Origin, Direction, 10000) (Shading.slang:708-750: tmin EPS = 1e-3), then
 * HitPoint = Origin + Direction * t (w = 1), Normal = the INTERPOLATED shading normal taken to world space (w = 0),
 * T = |HitPoint - Origin|, InstanceId = the node's instance id, MaterialId = the node's material for the triangle's slot,
 * Hitted = 1; a miss only sets Hitted = 0 and leaves the other result fields as they were.  (gk_raycast above is the CPU
 * variant RayCastInCPU: tmax 2000, face normal, no material.)  `io` is a host array updated in place. */
GkStatus gk_raycast_task(GkContext* ctx, GkRayCastIO* io, uint32_t count);
/* Closest-hit queries on arbitrary rays: 8 floats per ray {O.xyz, tmin, D.xyz, tmax};
 * out_tuv 3 floats, out_ids {triangle, node-proxy index} per ray.  Host pointers. */
GkStatus gk_intersect(GkContext* ctx, const float* rays, uint32_t count, float* out_tuv, uint32_t* out_ids);
/* Same on device-resident buffers (no copies); rays must stay valid until gk_synchronize. */
GkStatus gk_intersect_device(GkContext* ctx, const void* d_rays, uint32_t count, void* d_out_tuv, void* d_out_ids, int anyHit);

/* Plane access. gk_readback/gk_upload_plane copy the whole plane (bytes must match
 * gk_plane_bytes).  gk_plane_device returns the device pointer so that frame-end collectives
 * (NCCL) can run on the planes in place. */
size_t gk_plane_bytes(const GkContext* ctx, GkPlane plane);
GkStatus gk_readback(GkContext* ctx, GkPlane plane, void* dst, size_t bytes);
GkStatus gk_upload_plane(GkContext* ctx, GkPlane plane, const void* src, size_t bytes);
/* Read-back that overlaps the next frame: the copy is queued on a second stream behind the work
 * submitted so far and the call returns at once; `dst` should be page-locked (gk_host_alloc).
 * Kernels of later frames that overwrite the plane wait for the copy on the device.  One copy can be
 * in flight per context; gk_readback_wait blocks until `dst` is complete. */
GkStatus gk_readback_async(GkContext* ctx, GkPlane plane, void* dst, size_t bytes);
GkStatus gk_readback_wait(GkContext* ctx);
void* gk_plane_device(GkContext* ctx, GkPlane plane);

/* Frame-end exchange of the tile-partitioned frame (multi-GPU compositor).  gk_exchange_pack
 * gathers the rows this context owns of the six integrator planes the filters read (diffuse,
 * specular, albedo, normal, object id, motion: 44 B/pixel) into a device staging buffer of
 * gk_exchange_bytes() bytes; after an all-gather of the staging buffers over NCCL,
 * gk_exchange_unpack scatters tileCount x gk_exchange_bytes() bytes (rank-major) back into the
 * full planes.  Device pointers; both calls enqueue on the context stream. */
size_t gk_exchange_bytes(const GkContext* ctx);
GkStatus gk_exchange_pack(GkContext* ctx, void* d_staging);
GkStatus gk_exchange_unpack(GkContext* ctx, const void* d_all);

/* Peer-to-peer form of the same exchange (one process per GPU on one NVLink/NVSwitch node):
 *   gk_exchange_ipc_handles  writes GK_EXCHANGE_IPC_BYTES bytes of CUDA IPC handles of this context's exchange planes;
 *   gk_exchange_open_peers   takes the handles of all `world` ranks (rank-major, world == GkConfig.tileCount) and maps them;
 *   gk_exchange_push         one kernel stores the rows this rank owns into the planes of every peer.
 * The caller orders it between two cross-rank barriers on gk_stream(): peers must have finished
 * filtering the previous frame before the push, and all pushes must have landed before gk_filter_frame. */
#define GK_EXCHANGE_IPC_BYTES (8 * 64)
GkStatus gk_exchange_ipc_handles(GkContext* ctx, void* out, size_t bytes);
GkStatus gk_exchange_open_peers(GkContext* ctx, const void* handles_all, uint32_t world);
GkStatus gk_exchange_push(GkContext* ctx);
/* Unmaps the peers' planes and gather buffers (CUDA IPC).  Memory that another process has mapped must not be freed under
 * it: before gk_resize / gk_destroy of a context whose planes were exported, EVERY rank calls this and the caller
 * synchronises the ranks (compositor.release does both); afterwards the peer exchange has to be set up again. */
GkStatus gk_exchange_close_peers(GkContext* ctx);
/* Progressive rendering without the denoiser (the state gkNextBenchmark runs in, gkNextBenchmark.cpp:16-31)
 * filters every pixel on its own, so a tile-partitioned frame needs no exchange before the filters:
 *   gk_filter_frame_owned     runs the accumulate + compose passes on the rows this context owns
 *                             (GK_ERR_UNSUPPORTED unless ProgressiveRender != 0 and BFSize == 0);
 *   gk_exchange_push_final    stores the owned rows of rtDenoised into rank `dst_rank` (-1: every peer). */
GkStatus gk_filter_frame_owned(GkContext* ctx);
GkStatus gk_exchange_push_final(GkContext* ctx, int dst_rank);

/* Frame-sharded progressive rendering (ProgressiveRender != 0, BFSize == 0; contexts created with
 * GK_CFG_TRACE_ALL_ROWS): in a super-step rank r traces the WHOLE frame number f0 + r (its own TotalFrames, i.e. its
 * own random sequence); then every rank receives, for the rows it owns, the three source planes of all `world` frames
 * and applies the progressive accumulation lerp(history, src_s, 1/TemporalFrames) for s = 0..world-1 in frame order
 * (rounded to RGBA16F after every step, exactly what `world` consecutive single-GPU frames do), composes and tonemaps
 * the last one.  gk_exchange_push_final then delivers the finished rows to the presenting rank.
 *   gk_frame_shard_handle      allocates the gather buffer and writes its 64-byte CUDA IPC handle
 *   gk_frame_shard_open        maps the gather buffers of all ranks (rank-major handles)
 *   gk_frame_shard_push        one kernel: rows of this rank's frame go to the rank that owns them (NVLink stores)
 *   gk_frame_shard_accumulate  the `world` lerps + compose on the owned rows
 * The caller puts stream barriers between push and accumulate, as for gk_exchange_push. */
GkStatus gk_frame_shard_handle(GkContext* ctx, void* out, size_t bytes);
GkStatus gk_frame_shard_open(GkContext* ctx, const void* handles_all, uint32_t world);
GkStatus gk_frame_shard_push(GkContext* ctx);
GkStatus gk_frame_shard_accumulate(GkContext* ctx);

/* Page-locked host memory for arrays that are uploaded every frame (the node proxies: the reference
 * writes them straight into a mapped device buffer, src/Assets/Scene.cpp:464-511).  Returns NULL when
 * no CUDA device is usable; the caller then keeps using ordinary memory. */
void* gk_host_alloc(size_t bytes);
void gk_host_free(void* p);

GkStatus gk_synchronize(GkContext* ctx);
GkStatus gk_get_stats(GkContext* ctx, GkFrameStats* out);
GkStatus gk_get_bvh_info(GkContext* ctx, GkBvhInfo* out);
/* Tuning hooks (the defaults are the measured optima; DESIGN.md lists the sweeps).  Unknown names return
 * GK_ERR_INVALID_ARGUMENT.  Names: "trace_variant" (0 while-while lane kernel + cooperative kernel, 1 persistent
 * vote-scheduled kernel), "sched_refill_min", "sched_bias_node", "sched_keep_node", "sched_keep_tri", "sched_min_rays", "micro_tiles", "coop_divisor", "tail_divisor", "coop_threshold", "primary_lane_kernel", "tail_threshold",
 * "tail_fraction", "concurrent_shadow", "wave_lookahead". */
GkStatus gk_set_option(GkContext* ctx, const char* name, double value);
/* Measurement aid for the roofline of the traversal kernels ("peak measured once with an L2-resident
 * stream microbench, not assumed"): streams 128-bit loads that bypass L1 over a scratch buffer of `bytes`, `reps` times
 * per launch, and returns the best GB/s of five launches.  bytes well below the 126 MB L2 measures L2 read bandwidth,
 * bytes far above it measures HBM read bandwidth. */
GkStatus gk_measure_read_bandwidth(GkContext* ctx, size_t bytes, int reps, float* out_gbps);
/* Enables node-visit / triangle-test counters in the traversal kernels (slower). */
GkStatus gk_set_traversal_stats(GkContext* ctx, int enabled);
/* CUDA stream used by the context (cudaStream_t), for callers that order their own work. */
void* gk_stream(GkContext* ctx);
/* Dumps the extension-ray queue of wave `wave` of the last frame (8 floats per ray) into a
 * host buffer of `capacity` rays; returns the number of rays written through *count.  Used to
 * hand the CPU baseline the very rays the GPU traced. Requires gk_set_ray_capture(ctx, wave). */
GkStatus gk_set_ray_capture(GkContext* ctx, int wave);
GkStatus gk_get_captured_rays(GkContext* ctx, float* rays, uint32_t capacity, uint32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* GKNEXT_CUDA_H_ */
