// orc_pt.h — CPU restatement of the reference's path-tracing megakernel.
// TEST INFRASTRUCTURE ONLY (see orc_math.h).
//
// Follows, statement by statement:
//   assets/shaders/Core.PathTracing.comp.slang:31-102            (per-pixel driver)
//   assets/shaders/common/Shading.slang:285-434                  (primary caster; the raster
//        visibility buffer is replaced by a closest-hit primary ray)
//   assets/shaders/common/GeneralFunc.slang:33-83                (get_material_data)
//   assets/shaders/common/Shading.slang:50-58, 68-98             (motion vector, G-buffer)
//   assets/shaders/common/Shading.slang:659-758                  (FHardwareRayTracer)
//   assets/shaders/common/Shading.slang:826-850                  (sun direct illumination)
//   assets/shaders/common/Shading.slang:930-1082                 (FPathTracingRenderer)
//   assets/shaders/common/Const_Func.slang:8-49, 227-354         (Schlick, ONB, RNG, sampling)
//   assets/shaders/common/AmbientCube.slang:71-78, 178-223, 254-364 (probe read side)
//
// PARITY UNPINNED for radiance: the reference has no golden images and its shaders cannot
// be compiled here (no slangc / Vulkan), so this restatement is pinned only by review
// against the cited lines.  Hit ids *are* pinned (orc_bvh vs. the real tinybvh).
#pragma once
#include "orc_scene.h"

namespace orc {

struct PtOutputs {
    float* diffuse;  // 4/px: OutImage       (rgb = demodulated diffuse radiance, a = |pixelOffset|)
    float* spec;     // 4/px: OutImageSpec
    float* albedo;   // 4/px: OutAlbedoBuffer
    float* normal;   // 4/px: OutNormalBuffer (xyz normal, w roughness)
    float* motion;   // 2/px: OutMotionVector.rg (pixels)
    float* depth;    // 1/px: OutDepthBuffer (NDC z)
    uint32_t* objectId; // ObjectId0 (65535 = miss)
    uint32_t* primIds;  // 2/px: primary {prim, node index}
    uint32_t* rayCount; // rays traced for this pixel (primary + extension + shadow)
};

void renderFrame(const Scene& S, const GkUniformBufferObject& ubo, uint32_t W, uint32_t H, const GkAmbientCube* cubes,
                 const GkVoxelData* voxels, PtOutputs& out, int threads);

// Bake.HwAmbientCube.comp.slang:29-46 + FGpuProbeGenerator::Render (AmbientCube.slang:574-629) on probes [first, first + count)
void bakeProbes(const Scene& S, const GkUniformBufferObject& U, GkAmbientCube* cubes, GkVoxelData* voxels, uint32_t first, uint32_t count, int threads);

// the GPU ray-cast task (Task.RayCast.comp.slang:31-55) on in-place records
void rayCastTask(const Scene& S, GkRayCastIO* io, uint32_t n);

} // namespace orc
