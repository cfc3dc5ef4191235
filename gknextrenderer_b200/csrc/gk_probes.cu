// gk_probes.cu — the ambient-cube probe baker on the CUDA traversal.
//
// Replaces Bake.HwAmbientCube.comp.slang:29-46 (one thread per probe, dispatched over a slice of the 192 x 48 x 192 grid
// every frame by RayTraceBaseRenderer.cpp:244-291) and FGpuProbeGenerator::Render (common/AmbientCube.slang:574-629):
//   * six axis rays classify the probe (InsideGeometry :547-571): distance to the nearest surface per axis, "inside
//     geometry" when a back face or an emitter is closer than one cell; eight diagonal rays refine the distance of probes
//     that saw nothing (DetectDistance :534-545);
//   * the voxel record gets the distance in cells, the product of the six clamped distances ("inside" byte) and the six
//     per-axis distances as bytes (:612-615) - the read side (interpolateAmbientCubes, gk_shading.cuh) uses them as the
//     visibility test between a shading point and a probe;
//   * probes within reach of a surface run FaceTask (:459-532) for their six faces: 16 rays over a 4x4 grid jittered by the
//     probe's age (grid3x3[age % 9]), a hit gathers albedo x the DIRECT light stored in the probes around the hit point
//     (interpolateAmbientCubes<DIAmbientCubeSampler> x 1.25), a miss gathers the sky; plus the first area light
//     (TraceSegment) and the sun (TraceOcclusion); results are blended into the RGB10A2 faces with weight 1/8.
// The path tracer's terminator (Shading.slang:1054, gk_integrator.cu PH_POST_NEE) then reads non-zero probes.
//
// Mapping to the hardware: eight lanes per probe.  All eight run the probe's (cheap, scalar) control flow redundantly,
// so that they always hold the same ray; the traversal is the cooperative one (traverseCoop: lane j tests child j /
// triangle j) because a probe's ~14 ... 122 rays depend on one another and the launch is latency-bound; lane 0 stores.
//
// Stated deviations (the same as in the path tracer): no textures (albedo = Material.Diffuse), the sky is the constant
// BackGroundColor instead of the SH-projected HDR map (SampleIBLRough, Shading.slang:142-146).  InsideGeometry's `out`
// distance is left at its initial 255 when the ray misses (the shader leaves it unassigned).
#include "gk_context.h"
#include "gk_shading.cuh"

namespace gk {

namespace {

constexpr float kCubeUnit = GK_CUBE_UNIT;
constexpr float kFastMaxTrace = 20.f; // FAST_MAX_TRACE_DISTANCE, Shading.slang:16
constexpr float kMaxIlluminance = 512.f;

struct c4 {
    float x, y, z, w;
};
__device__ __forceinline__ c4 mk4c(float x, float y, float z, float w) { return c4{x, y, z, w}; }
__device__ __forceinline__ c4 operator+(c4 a, c4 b) { return mk4c(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ c4 operator*(c4 a, c4 b) { return mk4c(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ c4 operator*(c4 a, float s) { return mk4c(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ c4 operator/(c4 a, float s) { return mk4c(a.x / s, a.y / s, a.z / s, a.w / s); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }

__device__ __forceinline__ c4 unpackColor(uint32_t p) // unpackRGB10A2, AmbientCube.slang:71-78
{
    return mk4c(float(p & 0x3FF) / 1023.0f, float((p >> 10) & 0x3FF) / 1023.0f, float((p >> 20) & 0x3FF) / 1023.0f, 0.0f) * kMaxIlluminance;
}
__device__ __forceinline__ uint32_t packColor(c4 c) // packRGB10A2, AmbientCube.slang:59-69
{
    const float r = clampx(c.x / kMaxIlluminance, 0.f, 1.f), g = clampx(c.y / kMaxIlluminance, 0.f, 1.f), b = clampx(c.z / kMaxIlluminance, 0.f, 1.f),
                a = clampx(c.w / kMaxIlluminance, 0.f, 1.f);
    return (uint32_t)(r * 1023.0f) | ((uint32_t)(g * 1023.0f) << 10) | ((uint32_t)(b * 1023.0f) << 20) | ((uint32_t)(a * 3.0f) << 30);
}
__device__ __forceinline__ uint32_t lerpPackedColorAlt(uint32_t c0, c4 c1, float t) // AmbientCube.slang:119-126
{
    const c4 a = unpackColor(c0);
    return packColor(mk4c(mixf(a.x, c1.x, t), mixf(a.y, c1.y, t), mixf(a.z, c1.z, t), mixf(a.w, c1.w, t)));
}

// sampleAmbientCubeHL2_DI, AmbientCube.slang:128-151: the direct-light faces only
__device__ __forceinline__ c4 sampleCubeDI(const GkAmbientCube& cb, f3 n)
{
    const float wx = fmaxx(n.x, 0.f), wnx = fmaxx(-n.x, 0.f), wy = fmaxx(n.y, 0.f), wny = fmaxx(-n.y, 0.f), wz = fmaxx(n.z, 0.f), wnz = fmaxx(-n.z, 0.f);
    const float sum = wx + wnx + wy + wny + wz + wnz;
    c4 col = mk4c(0, 0, 0, 0);
    col = col + unpackColor(cb.PosX_D) * wx;
    col = col + unpackColor(cb.NegX_D) * wnx;
    col = col + unpackColor(cb.PosY_D) * wy;
    col = col + unpackColor(cb.NegY_D) * wny;
    col = col + unpackColor(cb.PosZ_D) * wz;
    col = col + unpackColor(cb.NegZ_D) * wnz;
    return col * ((sum > 0.0f) ? (1.0f / sum) : 1.0f);
}

// interpolateAmbientCubes<DIAmbientCubeSampler>, AmbientCube.slang:275-364
__device__ c4 interpolateDI(const GkAmbientCube* cubes, const GkVoxelData* voxels, f3 pos, f3 normal)
{
    const f3 off = mk3(-float(GK_CUBE_SIZE_XY / 2), -1.375f, -float(GK_CUBE_SIZE_XY / 2)) * kCubeUnit;
    const f3 np = (pos - off) / kCubeUnit;
    if (np.x < 0 || np.y < 0 || np.z < 0 || np.x > GK_CUBE_SIZE_XY - 1 || np.y > GK_CUBE_SIZE_Z - 1 || np.z > GK_CUBE_SIZE_XY - 1) return mk4c(0, 0, 0, 1);
    const int bx = (int)floorf(np.x), by = (int)floorf(np.y), bz = (int)floorf(np.z);
    const f3 fr = mk3(np.x - floorf(np.x), np.y - floorf(np.y), np.z - floorf(np.z));
    float total = 0;
    c4 result = mk4c(0, 0, 0, 0);
    for (int i = 0; i < 8; ++i) {
        const int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
        const int idx = (by + oy) * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + (bz + oz) * GK_CUBE_SIZE_XY + (bx + ox);
        const GkVoxelData vx = voxels[idx];
        const uint32_t p0 = vx.distanceToSolid_gg_z01, p1 = vx.distanceToSolid_x01_y01;
        if (float((p0 >> 8) & 0xFF) / 255.0f < 0.01f) continue;
        const float dPZ = float((p0 >> 16) & 0xFF) / 255.0f, dNZ = float((p0 >> 24) & 0xFF) / 255.0f;
        const float dPX = float(p1 & 0xFF) / 255.0f, dNX = float((p1 >> 8) & 0xFF) / 255.0f, dPY = float((p1 >> 16) & 0xFF) / 255.0f, dNY = float((p1 >> 24) & 0xFF) / 255.0f;
        const f3 ptl = fr - mk3((float)ox, (float)oy, (float)oz);
        const float dist = length3(ptl);
        const f3 dir = normalize3(ptl);
        const float hitLen = sqrtf(fmaxx(dir.x, 0.f)) * dPX + sqrtf(fmaxx(-dir.x, 0.f)) * dNX + sqrtf(fmaxx(dir.y, 0.f)) * dPY + sqrtf(fmaxx(-dir.y, 0.f)) * dNY +
                             sqrtf(fmaxx(dir.z, 0.f)) * dPZ + sqrtf(fmaxx(-dir.z, 0.f)) * dNZ;
        if (dist > hitLen + 0.05f) continue;
        const float wx = ox == 0 ? (1.0f - fr.x) : fr.x, wy = oy == 0 ? (1.0f - fr.y) : fr.y, wz = oz == 0 ? (1.0f - fr.z) : fr.z;
        const float w = wx * wy * wz;
        result = result + sampleCubeDI(cubes[idx], normal) * w;
        total += w;
    }
    return total > 0.0f ? result / total : mk4c(0, 0, 0, 0);
}

struct BakeArgs {
    SceneView V;
    ShadeScene SS;
    GkAmbientCube* cubes;
    GkVoxelData* voxels;
    const GkAmbientCube* cubesPrev; // the probe state before this call: what the gathers read (see bakeProbes)
    const GkVoxelData* voxelsPrev;
    const GkLightObject* lights;
    uint32_t first, count;
};

// The tracer of a probe group (FHardwareRayTracer, Shading.slang:659-758).  RayQuery distances are in units of the given
// direction, the traversal's in world units (it normalises, as tinybvh does): tmin / tmax are scaled by |direction| and the
// hit distance is scaled back, so un-normalised directions (DetectDistance's diagonals) behave as in the shader.
struct ProbeTracer {
    const SceneView& V;
    const ShadeScene& SS;
    uint2* stackRow;
    __device__ __forceinline__ bool traceRay(f3 ro, f3 rd, float maxDistance, Vtx& out) const // :708-750
    {
        const float len = length3(rd);
        Hit h{maxDistance * len, 0.f, 0.f, kInvalid, kInvalid};
        if (!traverseCoop<false, false>(V, ro, normalizeRayDir(rd), kEps * len, h, stackRow, nullptr)) return false;
        resolveHit(SS, ro, rd, h.t / len, h.u, h.v, h.prim, h.inst, out);
        return true;
    }
    __device__ __forceinline__ bool traceOcclusion(f3 ro, f3 rd) const // :661-681
    {
        const float len = length3(rd);
        Hit h{kMaxTrace * len, 0.f, 0.f, kInvalid, kInvalid};
        return traverseCoop<true, false>(V, ro, normalizeRayDir(rd), kEps * len, h, stackRow, nullptr);
    }
    __device__ __forceinline__ bool traceSegment(f3 ro, f3 target, float epsilon) const // :683-706
    {
        const f3 dir = target - ro;
        const float len = length3(dir);
        const f3 d = dir / len;
        const float dl = length3(d);
        Hit h{(len - epsilon) * dl, 0.f, 0.f, kInvalid, kInvalid};
        return traverseCoop<true, false>(V, ro, normalizeRayDir(d), epsilon * dl, h, stackRow, nullptr);
    }
};

__constant__ float2 kGrid3x3[9] = {{-0.667f, -0.667f}, {0.0f, -0.667f}, {0.667f, -0.667f}, {-0.667f, 0.0f}, {0.0f, 0.0f}, {0.667f, 0.0f}, {-0.667f, 0.667f}, {0.0f, 0.667f}, {0.667f, 0.667f}};
__constant__ float2 kGrid4x4[16] = {{-0.75f, -0.75f}, {-0.25f, -0.75f}, {0.25f, -0.75f}, {0.75f, -0.75f}, {-0.75f, -0.25f}, {-0.25f, -0.25f}, {0.25f, -0.25f}, {0.75f, -0.25f},
                                    {-0.75f, 0.25f},  {-0.25f, 0.25f},  {0.25f, 0.25f},  {0.75f, 0.25f},  {-0.75f, 0.75f},  {-0.25f, 0.75f},  {0.25f, 0.75f},  {0.75f, 0.75f}};

// FaceTask, AmbientCube.slang:459-532
__device__ __noinline__ void faceTask(const GkUniformBufferObject& U, const BakeArgs& A, const ProbeTracer& T, f3 origin, f3 basis, uint32_t iterate, uint32_t& directLight, uint32_t& indirectLight,
                         uint32_t& skyVisOut, uint32_t& sunVisOut)
{
    origin = origin + basis * kCubeUnit * 0.25f;
    c4 directColor = mk4c(0, 0, 0, 0), bounceColor = mk4c(0, 0, 0, 0);
    float skyVisibility = 0.0f;
    const float2 jit = kGrid3x3[iterate % 9];
    const float offx = jit.x * 0.25f, offy = jit.y * 0.25f;
    for (uint32_t i = 0; i < 16; ++i) {
        const f3 hemiVec = normalize3(mk3(kGrid4x4[i].x + offx, kGrid4x4[i].y + offy, 1.0f));
        const f3 rayDir = alignWithNormal(hemiVec, basis);
        Vtx hv;
        if (T.traceRay(origin, rayDir, kFastMaxTrace, hv)) {
            const GkMaterial& hm = A.SS.materials[hv.MaterialIndex];
            const c4 albedo = mk4c(hm.Diffuse[0], hm.Diffuse[1], hm.Diffuse[2], hm.Diffuse[3]);
            bounceColor = bounceColor + albedo * interpolateDI(A.cubesPrev, A.voxelsPrev, hv.Position, hv.Normal) * 1.25f; // "magic bounce twice"
        } else {
            const float k = U.HasSky ? U.SkyIntensity : 0.0f;
            directColor = directColor + mk4c(U.BackGroundColor[0], U.BackGroundColor[1], U.BackGroundColor[2], 1.0f) * k; // constant sky (see header)
            skyVisibility += 1.0f;
        }
    }
    directColor = directColor / 16.0f;
    bounceColor = bounceColor / 16.0f;
    if (U.LightCount > 0) { // the first parametric light only (:503-514)
        const GkLightObject& L = A.lights[0];
        const GkMaterial& lm = A.SS.materials[L.lightMatIdx];
        const c4 lightPower = mk4c(lm.Diffuse[0], lm.Diffuse[1], lm.Diffuse[2], lm.Diffuse[3]);
        const f3 p1 = mk3(L.p1[0], L.p1[1], L.p1[2]), p3 = mk3(L.p3[0], L.p3[1], L.p3[2]);
        const f3 lightPos = mk3(mixf(p1.x, p3.x, 0.5f), mixf(p1.y, p3.y, 0.5f), mixf(p1.z, p3.z, 0.5f));
        const float lightAtten = T.traceSegment(origin, lightPos, kCubeUnit * 0.5f) ? 0.0f : 1.0f;
        const f3 lightDir = normalize3(lightPos - origin);
        const float ndotl = clampx(dot3(basis, lightDir), 0.0f, 1.0f);
        const float distance = length3(lightPos - origin);
        const float attenuation = ndotl * L.normal_area[3] / (distance * distance * 3.14159f);
        directColor = directColor + lightPower * attenuation * lightAtten;
    }
    if (U.HasSun) { // :517-524
        const f3 sunDir = mk3(U.SunDirection[0], U.SunDirection[1], U.SunDirection[2]);
        const float sunAtten = T.traceOcclusion(origin, sunDir) ? 0.0f : 1.0f;
        const float ndotl = clampx(dot3(basis, sunDir), 0.0f, 1.0f);
        sunVisOut = sunAtten > 0.0f ? 1u : 0u;
        directColor = directColor + mk4c(U.SunColor[0], U.SunColor[1], U.SunColor[2], U.SunColor[3]) * sunAtten * ndotl * (U.HasSun ? 1.0f : 0.0f) * 0.25f;
    }
    const float currWeight = 0.125f; // the GPU baker keeps ~8 frames
    skyVisOut = (uint32_t)mixf((float)skyVisOut, 255.0f * skyVisibility / 16.0f, currWeight);
    directLight = lerpPackedColorAlt(directLight, directColor, currWeight);
    indirectLight = lerpPackedColorAlt(indirectLight, bounceColor, currWeight);
}

// InsideGeometry, AmbientCube.slang:547-571
__device__ __noinline__ bool insideGeometry(const BakeArgs& A, const ProbeTracer& T, f3 origin, f3 rayDir, uint32_t& outMaterialId, float& outDistanceToSolid)
{
    Vtx hv;
    if (T.traceRay(origin, rayDir, kCubeUnit * 64, hv)) {
        const float hitDist = length3(hv.Position - origin);
        outDistanceToSolid = hitDist;
        if (outDistanceToSolid <= kCubeUnit) {
            const GkMaterial& hm = A.SS.materials[hv.MaterialIndex];
            outMaterialId = hv.MaterialIndex;
            if (dot3(hv.Normal, rayDir) > 0.0f || ((hm.MaterialModel == GK_MAT_DIFFUSE_LIGHT) && hitDist < 0.02f)) { // voxel inclusive
                outDistanceToSolid = 0;
                return true;
            }
        }
    }
    return false;
}

__device__ __noinline__ float detectDistance(const ProbeTracer& T, f3 origin, f3 rayDir) // :534-545
{
    Vtx hv;
    if (T.traceRay(origin, rayDir, kCubeUnit * 64, hv)) return length3(hv.Position - origin);
    return 255.f;
}

__global__ void __launch_bounds__(256, 2) k_bake_probes(const GkUniformBufferObject* __restrict__ ubo, BakeArgs A)
{
    __shared__ uint2 stack[kRaysPerBlock * kStackStride];
    const GkUniformBufferObject& U = *ubo;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; // one group of eight lanes per probe
    if (g >= A.count) return;                                        // whole groups leave together
    const bool leader = (threadIdx.x & 7u) == 0;
    const uint32_t gIdx = A.first + g;
    const uint32_t y = gIdx / (GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY);
    const uint32_t z = (gIdx - y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY) / GK_CUBE_SIZE_XY;
    const uint32_t x = gIdx - y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY - z * GK_CUBE_SIZE_XY;
    const f3 cubeOffset = mk3(-float(GK_CUBE_SIZE_XY / 2), -1.375f, -float(GK_CUBE_SIZE_XY / 2)) * kCubeUnit;
    const f3 origin = mk3((float)x, (float)y, (float)z) * kCubeUnit + cubeOffset;
    const ProbeTracer T{A.V, A.SS, stack + (threadIdx.x >> 3) * kStackStride};

    // ---- FGpuProbeGenerator::Render, AmbientCube.slang:574-629 (every lane of the group holds the same values)
    GkVoxelData vox = A.voxels[gIdx];
    GkAmbientCube cube = A.cubes[gIdx];
    vox.matId = 0;
    float distPY = 255.0f, distNY = 255.0f, distPX = 255.0f, distNX = 255.0f, distPZ = 255.0f, distNZ = 255.0f;
    insideGeometry(A, T, origin, mk3(0, 1, 0), vox.matId, distPY);
    insideGeometry(A, T, origin, mk3(0, -1, 0), vox.matId, distNY);
    insideGeometry(A, T, origin, mk3(1, 0, 0), vox.matId, distPX);
    insideGeometry(A, T, origin, mk3(-1, 0, 0), vox.matId, distNX);
    insideGeometry(A, T, origin, mk3(0, 0, 1), vox.matId, distPZ);
    insideGeometry(A, T, origin, mk3(0, 0, -1), vox.matId, distNZ);
    float minDist = fminx(fminx(fminx(distPY, distNY), fminx(distPX, distNX)), fminx(distPZ, distNZ));
    if (minDist > 254.0f) { // the eight calls of the shader (two directions appear twice)
        minDist = fminx(minDist, detectDistance(T, origin, mk3(1, 1, 1)));
        minDist = fminx(minDist, detectDistance(T, origin, mk3(-1, 1, 1)));
        minDist = fminx(minDist, detectDistance(T, origin, mk3(-1, -1, 1)));
        minDist = fminx(minDist, detectDistance(T, origin, mk3(-1, 1, 1)));
        minDist = fminx(minDist, detectDistance(T, origin, mk3(1, 1, -1)));
        minDist = fminx(minDist, detectDistance(T, origin, mk3(-1, 1, -1)));
        minDist = fminx(minDist, detectDistance(T, origin, mk3(-1, -1, -1)));
        minDist = fminx(minDist, detectDistance(T, origin, mk3(-1, 1, -1)));
    }
    distPY = saturatef(distPY * 4.0f), distNY = saturatef(distNY * 4.0f), distPX = saturatef(distPX * 4.0f);
    distNX = saturatef(distNX * 4.0f), distPZ = saturatef(distPZ * 4.0f), distNZ = saturatef(distNZ * 4.0f);
    const float inside = distPY * distNY * distPX * distNX * distPZ * distNZ;
    auto pack4 = [](uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return (a & 0xFF) | ((b & 0xFF) << 8) | ((c & 0xFF) << 16) | ((d & 0xFF) << 24); };
    vox.distanceToSolid_gg_z01 = pack4((uint32_t)(minDist / kCubeUnit), (uint32_t)(inside * 255.0f), (uint32_t)(distPZ * 255.0f), (uint32_t)(distNZ * 255.0f));
    vox.distanceToSolid_x01_y01 = pack4((uint32_t)(distPX * 255.0f), (uint32_t)(distNX * 255.0f), (uint32_t)(distPY * 255.0f), (uint32_t)(distNY * 255.0f));
    if (minDist < 4) { // surface probes only (the shader compares metres here)
        const uint32_t iterate = vox.age;
        vox.age = vox.age + 1;
        uint32_t sv0[4] = {cube.skyVisibility_pznzpyny & 0xFF, (cube.skyVisibility_pznzpyny >> 8) & 0xFF, (cube.skyVisibility_pznzpyny >> 16) & 0xFF, (cube.skyVisibility_pznzpyny >> 24) & 0xFF};
        uint32_t sv1[4] = {cube.skyVisibility_pxnxs0s1 & 0xFF, (cube.skyVisibility_pxnxs0s1 >> 8) & 0xFF, (cube.skyVisibility_pxnxs0s1 >> 16) & 0xFF, (cube.skyVisibility_pxnxs0s1 >> 24) & 0xFF};
        uint32_t sunvis = 0;
        faceTask(U, A, T, origin, mk3(0, 1, 0), iterate, cube.PosY_D, cube.PosY, sv0[2], sv1[2]);
        faceTask(U, A, T, origin, mk3(0, -1, 0), iterate, cube.NegY_D, cube.NegY, sv0[3], sunvis);
        faceTask(U, A, T, origin, mk3(1, 0, 0), iterate, cube.PosX_D, cube.PosX, sv1[0], sunvis);
        faceTask(U, A, T, origin, mk3(-1, 0, 0), iterate, cube.NegX_D, cube.NegX, sv1[1], sunvis);
        faceTask(U, A, T, origin, mk3(0, 0, 1), iterate, cube.PosZ_D, cube.PosZ, sv0[0], sunvis);
        faceTask(U, A, T, origin, mk3(0, 0, -1), iterate, cube.NegZ_D, cube.NegZ, sv0[1], sunvis);
        cube.skyVisibility_pznzpyny = pack4(sv0[0], sv0[1], sv0[2], sv0[3]);
        cube.skyVisibility_pxnxs0s1 = pack4(sv1[0], sv1[1], sv1[2], sv1[3]);
        if (leader) A.cubes[gIdx] = cube;
    }
    if (leader) A.voxels[gIdx] = vox;
}

} // namespace

GkStatus bakeProbes(Context& c, uint32_t first, uint32_t count)
{
    const uint32_t total = GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z;
    if (!c.haveScene || !c.haveInstances || !c.haveUbo) {
        setLastError("gk_bake_probes: scene, instances and UBO must be set first");
        return GK_ERR_NOT_READY;
    }
    if (first >= total || count == 0 || count > total - first) {
        setLastError("gk_bake_probes: probe range outside the 192 x 48 x 192 grid");
        return GK_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = c.stream;
    if (!c.haveProbes) { // un-baked probes are all zero (the state the reference starts from)
        GK_CUDA(c.dCubes.reserve(total));
        GK_CUDA(c.dVoxels.reserve(total));
        GK_CUDA(cudaMemsetAsync(c.dCubes.p, 0, sizeof(GkAmbientCube) * (size_t)total, st));
        GK_CUDA(cudaMemsetAsync(c.dVoxels.p, 0, sizeof(GkVoxelData) * (size_t)total, st));
        c.haveProbes = true;
    }
    // The shader gathers light from the probes around every hit point while other threads of the same dispatch rewrite them,
    // so its result depends on the execution order.  This backend defines the update of a call on the state BEFORE the call
    // (a device copy, 127 MB, ~0.05 ms): deterministic, order-free, same fixed point.
    GK_CUDA(c.dCubesPrev.reserve(total));
    GK_CUDA(c.dVoxelsPrev.reserve(total));
    GK_CUDA(cudaMemcpyAsync(c.dCubesPrev.p, c.dCubes.p, sizeof(GkAmbientCube) * (size_t)total, cudaMemcpyDeviceToDevice, st));
    GK_CUDA(cudaMemcpyAsync(c.dVoxelsPrev.p, c.dVoxels.p, sizeof(GkVoxelData) * (size_t)total, cudaMemcpyDeviceToDevice, st));
    GK_CUDA(cudaMemcpyAsync(c.dUbo, &c.ubo, sizeof(GkUniformBufferObject), cudaMemcpyHostToDevice, st));
    BakeArgs A;
    A.V = c.view();
    A.SS.verts = c.dGpuVerts.p, A.SS.indices = c.dIndices.p, A.SS.models = c.dModels.p, A.SS.materials = c.dMaterials.p, A.SS.nodes = c.dNodes.p, A.SS.inst = c.dInst.p;
    A.SS.cubes = c.dCubes.p, A.SS.voxels = c.dVoxels.p, A.SS.materialCount = c.materialCount;
    A.cubes = c.dCubes.p, A.voxels = c.dVoxels.p, A.lights = c.dLights.p;
    A.cubesPrev = c.dCubesPrev.p, A.voxelsPrev = c.dVoxelsPrev.p;
    A.first = first, A.count = count;
    const unsigned grid = (unsigned)(((size_t)count * 8 + 255) / 256);
    k_bake_probes<<<grid, 256, 0, st>>>(c.dUbo, A);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return GK_OK;
}

GkStatus getProbes(Context& c, GkAmbientCube* cubes, GkVoxelData* voxels, size_t count)
{
    const size_t total = (size_t)GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z;
    if (count != total || (!cubes && !voxels)) {
        setLastError("gk_get_probes: the probe grid is 192 x 48 x 192");
        return GK_ERR_INVALID_ARGUMENT;
    }
    if (!c.haveProbes) { // nothing baked or uploaded: all zero
        if (cubes) memset(cubes, 0, sizeof(GkAmbientCube) * total);
        if (voxels) memset(voxels, 0, sizeof(GkVoxelData) * total);
        return GK_OK;
    }
    if (cubes) GK_CUDA(cudaMemcpyAsync(cubes, c.dCubes.p, sizeof(GkAmbientCube) * total, cudaMemcpyDeviceToHost, c.stream));
    if (voxels) GK_CUDA(cudaMemcpyAsync(voxels, c.dVoxels.p, sizeof(GkVoxelData) * total, cudaMemcpyDeviceToHost, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return GK_OK;
}

} // namespace gk
