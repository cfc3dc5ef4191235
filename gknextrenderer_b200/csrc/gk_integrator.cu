// gk_integrator.cu — the wavefront path tracer: generate / extend / shadow / shade / accumulate.
//
// Replaces Core.PathTracing.comp.slang (assets/shaders/Core.PathTracing.comp.slang:31-102) and the
// megakernel it calls (FPathTracingRenderer, assets/shaders/common/Shading.slang:930-1082), which
// the reference dispatches as 8x8 groups over the render extent
// (src/Rendering/PathTracing/PathTracingRenderer.cpp:104-114).
//
// The megakernel's per-pixel program is cut at every ray cast into a small state machine; one
// path per pixel walks it and carries the pcg4d state, so random numbers are drawn in exactly
// the reference's per-pixel order (including the data-dependent draws):
//
//   generate : camera ray per owned pixel                      -> extend queue
//   extend   : closest hit over TLAS/BLAS (gk_bvh.cuh)          -> hit records
//   shadow   : any hit (sun next-event estimation, direct sun)  -> occlusion flags
//   shade    : consumes hit / occlusion records, evaluates the material, draws the next
//              direction or the NEE sample and appends to the extend / shadow queue of the next
//              wave with one warp-aggregated atomic per warp (ballot + popc ranks)
//   accumulate: writes the finished pixel to the RGBA16F planes (+ fp32 parity planes)
//
// Queues are SoA (origin|tmin, direction|tmax, path id; t|u|v|prim, instance), 32 + 4 + 20 bytes
// per ray.  All kernels are one thread per queue entry, 256 threads per block.
#include "gk_context.h"
#include "gk_shading.cuh"
#include "gk_trace_sched.cuh"
#include <cstdio>
#include <cstdlib>

namespace gk {

enum PathStateId : uint32_t { ST_PRIMARY = 0, ST_PRIMARY_DOF = 1, ST_BOUNCE = 2, ST_NEE = 3, ST_DIRECT = 4, ST_DONE = 5 };

constexpr uint32_t F_CHANCE_REFLECT = 1u << 3, F_TERMINATE_ON_HIT = 1u << 4, F_HIT_REFLECT = 1u << 5, F_HIT_METAL = 1u << 6, F_PRIM_DIELECTRIC = 1u << 7,
                   F_SRC_DIELECTRIC = 1u << 8;
GK_HD uint32_t flagsState(uint32_t f) { return f & 7u; }
GK_HD uint32_t flagsBounce(uint32_t f) { return (f >> 9) & 0x7fu; }
GK_HD uint32_t flagsSample(uint32_t f) { return f >> 16; }
GK_HD uint32_t packFlags(uint32_t state, uint32_t bits, uint32_t bounce, uint32_t sample) { return state | bits | (bounce << 9) | (sample << 16); }

struct FrameParams {
    uint32_t width, height;
    uint32_t tileIndex, tileCount, tileRows;
    uint32_t pathCount;
    uint32_t microTiles; // paths of a warp cover an 8 x 4 pixel block instead of a 32 x 1 strip (see pathToPixel)
};

struct PlaneView {
    __half* outDiffuse;
    __half* outSpec;
    __half* albedo;
    __half* normal;
    uint32_t* objectId0;
    float2* motion;
    float* depth;
    float4* radDiffuse;
    float4* radSpec;
    uint2* primaryIds;
    float* primaryT;
    uint32_t* rayCount;
};

__device__ __forceinline__ void storeHalf4(__half* plane, uint32_t pixel, float x, float y, float z, float w)
{
    const __half2 a = __floats2half2_rn(x, y), b = __floats2half2_rn(z, w);
    uint2 v;
    v.x = *reinterpret_cast<const uint32_t*>(&a), v.y = *reinterpret_cast<const uint32_t*>(&b);
    reinterpret_cast<uint2*>(plane)[pixel] = v;
}

__device__ __forceinline__ uint32_t pathToPixel(const FrameParams& P, uint32_t path)
{
    uint32_t lr = path / P.width, x = path - lr * P.width;
    if (P.microTiles) {
        // Within every band of four rows the paths run through 8 x 4 pixel blocks, one block per warp: camera rays of a warp
        // then span a compact window of the image (more coherent traversal, more shared cache lines of the scene) and the G-buffer
        // stores of a warp are four 64-byte row segments.  Needs width % 8 == 0 and tileRows % 4 == 0 (else the linear order).
        if (P.microTiles == 2) {
            // 16-row bands: a thread block (256 paths) covers a 16 x 16 pixel square made of 2 x 4 warp blocks
            const uint32_t band = lr >> 4, q = (lr & 15u) * P.width + x;
            const uint32_t sq = q >> 8, t = q & 255u, w = t >> 5, l = t & 31u;
            x = sq * 16u + (w & 1u) * 8u + (l & 7u);
            lr = band * 16u + (w >> 1) * 4u + (l >> 3);
        } else {
            const uint32_t band = lr >> 2, q = (lr & 3u) * P.width + x; // position inside the band of 4 rows
            const uint32_t blk8 = q >> 5, l = q & 31u;
            x = blk8 * 8u + (l & 7u);
            lr = band * 4u + (l >> 3);
        }
    }
    const uint32_t blk = lr / P.tileRows, within = lr - blk * P.tileRows;
    const uint32_t row = (blk * P.tileCount + P.tileIndex) * P.tileRows + within;
    return row < P.height ? row * P.width + x : kInvalid;
}

__device__ __forceinline__ f3 cameraDir(const GkUniformBufferObject& U, int px, int py, uint32_t W, uint32_t H)
{
    const float ux = (float(px) / float(W)) * 2.0f - 1.0f, uy = (float(py) / float(H)) * 2.0f - 1.0f;
    const f4 target = mulM(U.ProjectionInverse, mk4(ux, uy, 1, 1));
    const f3 tn = normalize3(xyz(target));
    const f4 dir = mulM(U.ModelViewInverse, mk4(tn.x, tn.y, tn.z, 0));
    return normalize3(xyz(dir));
}

// -------------------------------------------------------------------------------- generate
__global__ void __launch_bounds__(256) k_generate(const GkUniformBufferObject* __restrict__ ubo, FrameParams P, PathState S, RayQueue Q)
{
    const uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
    if (path >= P.pathCount) return;
    const GkUniformBufferObject& U = *ubo;
    const uint32_t pixel = pathToPixel(P, path);
    S.pixel[path] = pixel;
    S.rays[path] = 0;
    uint32_t x = 0, y = 0;
    f3 dir = mk3(0, 0, 1);
    const f3 origin = xyz(mulM(U.ModelViewInverse, mk4(0, 0, 0, 1)));
    if (pixel != kInvalid) {
        y = pixel / P.width, x = pixel - y * P.width;
        dir = cameraDir(U, (int)x, (int)y, P.width, P.height);
    }
    S.rng[path] = make_uint4(x, y, U.TotalFrames, 0); // InitRandomSeed, Const_Func.slang:237-240
    S.nrmFlags[path] = make_float4(0, 0, 0, __uint_as_float(packFlags(pixel == kInvalid ? ST_DONE : ST_PRIMARY, 0, 0, 0)));
    S.accDiffuse[path] = make_float4(0, 0, 0, 0);
    S.accSpec[path] = make_float4(0, 0, 0, 0);
    // every path owns slot `path` of the first extend queue; rows past the image get a null ray
    Q.o_tmin[path] = make_float4(origin.x, origin.y, origin.z, 0.0f);
    Q.d_tmax[path] = make_float4(dir.x, dir.y, dir.z, pixel == kInvalid ? 0.0f : kPrimaryTMax);
    Q.path[path] = path;
    if (path == 0) *Q.count = P.pathCount;
}

// -------------------------------------------------------------------------------- extend / shadow
// Two mappings of rays to lanes (gk_bvh.cuh):
//   kCoop = false : one ray per lane, 256 rays per block (large waves)
//   kCoop = true  : eight lanes per ray, 32 rays per block (small waves: ~5x lower latency per ray,
//                   which is what bounds the long tail of nearly empty waves)
__device__ __forceinline__ uint2* stackRowOf(uint2* stack) { return stack + (threadIdx.x >> 3) * kStackStride; }

struct QueueIO { // RayIO over a RayQueue (closest hit)
    RayQueue Q;
    __device__ __forceinline__ bool load(uint32_t i, f3& O, f3& D, float& tmin, float& tmax) const
    {
        const float4 o = __ldg(Q.o_tmin + i), d = __ldg(Q.d_tmax + i);
        O = mk3(o.x, o.y, o.z), D = mk3(d.x, d.y, d.z), tmin = o.w, tmax = d.w;
        return d.w > 0.0f;
    }
    __device__ __forceinline__ void store(uint32_t i, const Hit& h, bool) const
    {
        Q.hit_tuvp[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
        Q.hit_inst[i] = h.inst;
    }
};
struct ShadowIO { // RayIO over a RayQueue (any hit: only the occlusion flag is stored)
    RayQueue Q;
    __device__ __forceinline__ bool load(uint32_t i, f3& O, f3& D, float& tmin, float& tmax) const
    {
        const float4 o = __ldg(Q.o_tmin + i), d = __ldg(Q.d_tmax + i);
        O = mk3(o.x, o.y, o.z), D = mk3(d.x, d.y, d.z), tmin = o.w, tmax = d.w;
        return d.w > 0.0f;
    }
    __device__ __forceinline__ void store(uint32_t i, const Hit&, bool occluded) const { Q.hit_inst[i] = occluded ? 1u : 0u; }
};
template <bool kAny> struct ArrayIO { // RayIO over interleaved {O|tmin, D|tmax} records (gk_intersect)
    const float4* rays;
    float* tuv;
    uint32_t* ids;
    __device__ __forceinline__ bool load(uint32_t i, f3& O, f3& D, float& tmin, float& tmax) const
    {
        const float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
        O = mk3(o.x, o.y, o.z), D = mk3(d.x, d.y, d.z), tmin = o.w, tmax = d.w;
        return true;
    }
    __device__ __forceinline__ void store(uint32_t i, const Hit& h, bool hit) const
    {
        if (tuv) tuv[3 * i] = h.t, tuv[3 * i + 1] = h.u, tuv[3 * i + 2] = h.v;
        if (ids) {
            if (kAny) ids[2 * i] = hit ? 1u : 0u, ids[2 * i + 1] = hit ? 1u : 0u;
            else ids[2 * i] = h.inst == kInvalid ? kInvalid : h.prim, ids[2 * i + 1] = h.inst;
        }
    }
};

// One kernel body for extend / shadow / intersect.
template <bool kAnyHit, bool kCoop, bool kStats, class RayIO>
__global__ void __launch_bounds__(256, kCoop ? 2 : 4) k_trace(SceneView V, RayIO io, uint32_t count, TraversalStats* stats, const uint32_t* __restrict__ countPtr)
{
    __shared__ uint2 stack[kCoop ? kRaysPerBlock * kStackStride : 1];
    if (countPtr) count = *countPtr; // device-driven wave loop: `count` only sized the grid
    TraversalStats local{0, 0};
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = kCoop ? (gt >> 3) : gt;
    if (i >= count) return; // cooperative mapping: whole 8-lane groups leave together
    f3 O = mk3(0, 0, 0), D = mk3(0, 0, 1);
    float tmin = 0.f, tmax = 0.f;
    const bool live = io.load(i, O, D, tmin, tmax);
    Hit h{tmax, 0.f, 0.f, kInvalid, kInvalid};
    bool hit = false;
    if (live) {
        const f3 dn = normalizeRayDir(D);
        if (kCoop) hit = traverseCoop<kAnyHit, kStats>(V, O, dn, tmin, h, stackRowOf(stack), &local);
        else hit = traverseLane<kAnyHit, kStats>(V, O, dn, tmin, h, &local);
    }
    if (!kCoop || (threadIdx.x & 7u) == 0) {
        io.store(i, h, hit);
        if (kStats) {
            atomicAdd(&stats->nodeVisits, local.nodeVisits);
            atomicAdd(&stats->triTests, local.triTests);
            atomicAdd(&stats->tlasVisits, local.tlasVisits);
            atomicAdd(&stats->instanceEntries, local.instanceEntries);
            atomicMax(&stats->maxStack, local.maxStack);
        }
    }
}

// The scheduled (persistent, vote-driven) traversal kernel: gk_trace_sched.cuh.  `countPtr` (device) overrides
// `countImm` so that a wave can be launched before the host knows its size.
template <bool kAnyHit, bool kStats, class RayIO>
__global__ void __launch_bounds__(kSchedBlock, 6) k_trace_sched(SceneView V, RayIO io, const uint32_t* __restrict__ countPtr, uint32_t countImm, uint32_t* __restrict__ cursor,
                                                              SchedParams prm, TraversalStats* stats, SchedStats* sched)
{
    __shared__ uint32_t sMem[kSchedSmemWords];
    const uint32_t count = countPtr ? *countPtr : countImm;
    traverseScheduled<kAnyHit, kStats, RayIO>(V, io, count, cursor, prm, sMem, stats, sched);
}

// -------------------------------------------------------------------------------- shade
struct Emit {
    int kind; // 0 none, 1 extend, 2 shadow
    f3 o, d;
    float tmin, tmax;
};

// First half of GetRayColor (Shading.slang:934-965): draw the lobe, the direction, emit the ray.
__device__ __forceinline__ void setupBounce(const ShadeScene& SS, u4& rng, const f3 pos, const f3 nrm, uint32_t matIdx, f3& dir, uint32_t& bits, bool firstBounce,
                                            Emit& e)
{
    const GkMaterial& mat = SS.materials[matIdx];
    const bool dielectric = mat.MaterialModel == GK_MAT_DIELECTRIC;
    const float startPosOffset = dielectric ? 0.0f : 1.0f;
    const float roughness = mat.Fuzziness;
    const float dotValue = dot3(dir, nrm);
    const bool backFace = dotValue > 0;
    const f3 outwardNormal = backFace ? -nrm : nrm;
    const float niOverNt = backFace ? mat.RefractionIndex2 : (1 / mat.RefractionIndex2);
    const float cosine = dotValue > 0 ? mat.RefractionIndex * dotValue : -dotValue;
    const float reflectProb = schlick(cosine, mat.RefractionIndex);
    const float metalProb = mat.Metalness;
    const bool chanceReflect = randomFloat(rng) < reflectProb;
    const bool chanceMetal = randomFloat(rng) < metalProb;
    const bool chanceGGX = chanceReflect || chanceMetal;
    const f3 traceNext = chanceGGX ? reflect3(dir, outwardNormal) : outwardNormal;
    f3 traceDir = chanceGGX ? ggxSampling(rng, sqrtf(roughness), traceNext) : alignWithNormal(randomInHemiSphere1(rng), traceNext);
    if (dielectric && !chanceReflect) traceDir = refract3(dir, outwardNormal, niOverNt);
    bits &= ~(F_CHANCE_REFLECT | F_TERMINATE_ON_HIT | F_SRC_DIELECTRIC);
    if (chanceReflect) bits |= F_CHANCE_REFLECT;
    if (backFace && !dielectric) bits |= F_TERMINATE_ON_HIT;
    if (firstBounce) { // only the first GetRayColor of a sample reports its lobe (Shading.slang:1025 vs :1032)
        bits &= ~(F_HIT_REFLECT | F_HIT_METAL);
        if (chanceGGX) bits |= F_HIT_REFLECT;
        if (chanceMetal) bits |= F_HIT_METAL;
    }
    dir = traceDir;
    e.kind = 1;
    e.o = pos + nrm * kTraceOffset * startPosOffset;
    e.d = traceDir;
    e.tmin = kEps, e.tmax = kMaxTrace;
}

// What a path needs to know about the ray it was waiting for.
struct QueueSrc { // results sitting in the wave's queues
    RayQueue inE, inS;
    uint32_t slot;
    __device__ __forceinline__ float4 hitTuvp() const { return inE.hit_tuvp[slot]; }
    __device__ __forceinline__ uint32_t hitInst() const { return inE.hit_inst[slot]; }
    __device__ __forceinline__ float4 rayO() const { return inE.o_tmin[slot]; }
    __device__ __forceinline__ float4 rayD() const { return inE.d_tmax[slot]; }
    __device__ __forceinline__ bool occluded() const { return inS.hit_inst[slot] != 0; }
};
struct RegSrc { // results held in registers (tail kernel)
    float4 o, d, tuvp;
    uint32_t inst;
    bool occ;
    __device__ __forceinline__ float4 hitTuvp() const { return tuvp; }
    __device__ __forceinline__ uint32_t hitInst() const { return inst; }
    __device__ __forceinline__ float4 rayO() const { return o; }
    __device__ __forceinline__ float4 rayD() const { return d; }
    __device__ __forceinline__ bool occluded() const { return occ; }
};

// One step of a path's state machine: consume the result of its outstanding ray, run the control
// flow of FPathTracingRenderer::Render up to the next ray (returned in `e`) or to the end of the path.
template <class Src>
__device__ __forceinline__ void shadePath(const GkUniformBufferObject& U, const FrameParams& P, const ShadeScene& SS, const PathState& S, const PlaneView& PL,
                                          uint32_t path, const Src& src, Emit& e)
{
    {
        const float4 nf = S.nrmFlags[path];
        uint32_t flags = __float_as_uint(nf.w);
        uint32_t state = flagsState(flags);
        uint32_t bits = flags & (F_CHANCE_REFLECT | F_TERMINATE_ON_HIT | F_HIT_REFLECT | F_HIT_METAL | F_PRIM_DIELECTRIC | F_SRC_DIELECTRIC);
        uint32_t bounce = flagsBounce(flags), sample = flagsSample(flags);
        if (state != ST_DONE) {
            const uint4 r4 = S.rng[path];
            u4 rng{r4.x, r4.y, r4.z, r4.w};
            const uint32_t pixel = S.pixel[path];
            uint32_t rays = S.rays[path] + 1; // the ray whose result we are consuming
            const f3 eye = xyz(mulM(U.ModelViewInverse, mk4(0, 0, 0, 1)));
            const uint32_t samples = U.FastGather ? 1u : U.NumberOfSamples;

            // working registers (loaded lazily per state)
            f3 vpos = mk3(0, 0, 0), vnrm = mk3(nf.x, nf.y, nf.z), dir = mk3(0, 0, 0), color = mk3(1, 1, 1);
            uint32_t vmat = 0;
            f3 ppos = mk3(0, 0, 0), pnrm = mk3(0, 0, 0);
            uint32_t pmat = 0;
            float offLen = 0.f;
            f3 accD = mk3(0, 0, 0), accS = mk3(0, 0, 0);
            float shadowTerm = 0.f;
            bool finished = false; // emissive primary: accD already holds the final value
            bool accDirty = false;

            enum Phase { PH_START_SAMPLE, PH_SETUP_BOUNCE, PH_POST_NEE, PH_END_SAMPLE, PH_AFTER_SAMPLES, PH_FINAL, PH_EXIT };
            Phase phase = PH_EXIT;

            // primary vertex and accumulators are fetched only by the steps that use them (sample
            // boundaries); a bounce in the middle of a sample touches neither
            bool havePrimary = false, haveAcc = false;
            float accW = 0.f; // .w of both accumulators: length of the depth-of-field pixel offset
            auto needPrimary = [&]() {
                if (havePrimary) return;
                const float4 a = S.primPosMat[path], b = S.primNrm[path];
                ppos = mk3(a.x, a.y, a.z), pmat = __float_as_uint(a.w);
                pnrm = mk3(b.x, b.y, b.z), offLen = b.w;
                havePrimary = true;
            };
            auto needAcc = [&]() {
                accDirty = true;
                if (haveAcc) return;
                const float4 a = S.accDiffuse[path], b = S.accSpec[path];
                accD = mk3(a.x, a.y, a.z), accS = mk3(b.x, b.y, b.z), accW = a.w;
                haveAcc = true;
            };
            auto loadCurrent = [&]() {
                const float4 a = S.posMat[path], c = S.dirT[path], t = S.throughput[path];
                vpos = mk3(a.x, a.y, a.z), vmat = __float_as_uint(a.w);
                dir = mk3(c.x, c.y, c.z), color = mk3(t.x, t.y, t.z);
            };

            if (state == ST_PRIMARY || state == ST_PRIMARY_DOF) {
                // ---- Core.PathTracing main :46-79 with FVisibilityBufferRayCaster (Shading.slang:287-434);
                //      the visibility id comes from the traced primary ray instead of the raster pass.
                const float4 hr = src.hitTuvp();
                const uint32_t hinst = src.hitInst(), hprim = __float_as_uint(hr.w);
                const uint32_t py = pixel / P.width, px = pixel - py * P.width;
                const f3 rayDir0 = cameraDir(U, (int)px, (int)py, P.width, P.height);
                bool resolved = false;
                Vtx hitV;
                uint32_t hitNode = 0;
                hitV.MaterialIndex = 0;
                hitV.Position = hitV.Normal = mk3(0, 0, 0);
                hitV.TexCoord = f2{0, 0};
                if (state == ST_PRIMARY) {
                    PL.primaryIds[pixel] = make_uint2(hinst == kInvalid ? kInvalid : hprim, hinst);
                    PL.primaryT[pixel] = hr.x;
                    if (hinst == kInvalid) {
                        // miss: Core.PathTracing :53-64
                        const f3 sky = skyColor(U);
                        const float skyA = U.HasSky ? fminx(1.f, U.BackGroundColor[3]) * U.SkyIntensity : 0.f;
                        PL.motion[pixel] = make_float2(0, 0);
                        storeHalf4(PL.albedo, pixel, 1, 1, 1, 1);
                        storeHalf4(PL.normal, pixel, 0, 1, 0, 1);
                        PL.objectId0[pixel] = 65535u;
                        PL.depth[pixel] = 0.f;
                        S.accDiffuse[path] = make_float4(sky.x, sky.y, sky.z, skyA);
                        S.accSpec[path] = make_float4(0, 0, 0, 0);
                        state = ST_DONE;
                    } else {
                        uint32_t rawMat;
                        const Vtx initial = getMaterialData(SS, hinst, hprim, eye, rayDir0, rawMat);
                        const float vertexDistance = length3(initial.Position - eye);
                        float coc = 0.0f;
                        if (fabsf(vertexDistance - U.FocusDistance) > 0.001f) coc = (U.Aperture * fabsf(vertexDistance - U.FocusDistance)) / vertexDistance;
                        float pox = 0.f, poy = 0.f;
                        if (coc > 0.001f) {
                            const f2 disk = concentricDisk(randomFloat2(rng));
                            const f3 right = normalize3(cross3(rayDir0, mk3(0, 1, 0)));
                            const f3 edge = initial.Position + right * coc;
                            const f4 cp = mulM(U.ViewProjection, mk4(initial.Position.x, initial.Position.y, initial.Position.z, 1));
                            const f4 ep = mulM(U.ViewProjection, mk4(edge.x, edge.y, edge.z, 1));
                            const float dx = ep.x / ep.w - cp.x / cp.w, dy = ep.y / ep.w - cp.y / cp.w;
                            const float ssr = sqrtf(dx * dx + dy * dy);
                            pox = disk.x * ssr * float(P.width) * 0.5f, poy = disk.y * ssr * float(P.height) * 0.5f;
                        }
                        offLen = sqrtf(pox * pox + poy * poy);
                        int ox = (int)px + (int)pox, oy = (int)py + (int)poy;
                        ox = ox < 0 ? 0 : (ox > (int)P.width - 1 ? (int)P.width - 1 : ox);
                        oy = oy < 0 ? 0 : (oy > (int)P.height - 1 ? (int)P.height - 1 : oy);
                        if (ox == (int)px && oy == (int)py) {
                            // the offset pixel is this pixel: its visibility id is the one we hold
                            hitNode = hinst;
                            hitV.Position = initial.Position, hitV.Normal = normalize3(initial.Normal), hitV.TexCoord = initial.TexCoord;
                            hitV.MaterialIndex = SS.nodes[hinst].matId[rawMat & 15];
                            resolved = true;
                        } else {
                            // depth of field: fetch the visibility id of the offset pixel by tracing its camera ray
                            S.dofVertex[path] = make_float4(initial.Position.x, initial.Position.y, initial.Position.z, __uint_as_float(rawMat));
                            S.dofNormal[path] = make_float4(initial.Normal.x, initial.Normal.y, initial.Normal.z, __uint_as_float(hinst));
                            S.primNrm[path] = make_float4(__int_as_float(ox), __int_as_float(oy), vertexDistance, offLen);
                            const f3 fd = cameraDir(U, ox, oy, P.width, P.height);
                            e.kind = 1, e.o = eye, e.d = fd, e.tmin = 0.f, e.tmax = kPrimaryTMax;
                            state = ST_PRIMARY_DOF;
                        }
                    }
                } else {
                    // ST_PRIMARY_DOF: Shading.slang:356-431
                    const float4 dv = S.dofVertex[path], dn = S.dofNormal[path], aux = S.primNrm[path];
                    const uint32_t rawMat0 = __float_as_uint(dv.w), node0 = __float_as_uint(dn.w);
                    const int ox = __float_as_int(aux.x), oy = __float_as_int(aux.y);
                    const float vertexDistance = aux.z;
                    offLen = aux.w;
                    bool useInitial = (hinst == kInvalid);
                    Vtx finalV;
                    uint32_t finalRaw = 0;
                    if (!useInitial) {
                        const f3 fd = cameraDir(U, ox, oy, P.width, P.height);
                        finalV = getMaterialData(SS, hinst, hprim, eye, fd, finalRaw);
                        if (fabsf(vertexDistance - U.FocusDistance) < U.FocusDistance * 0.1f) {
                            const float fdist = length3(finalV.Position - eye);
                            if (vertexDistance - fdist > U.FocusDistance * 0.05f) useInitial = true;
                        }
                    }
                    if (useInitial) {
                        hitNode = node0;
                        hitV.Position = mk3(dv.x, dv.y, dv.z), hitV.Normal = normalize3(mk3(dn.x, dn.y, dn.z));
                        hitV.MaterialIndex = SS.nodes[node0].matId[rawMat0 & 15];
                    } else {
                        hitNode = hinst;
                        hitV.Position = finalV.Position, hitV.Normal = normalize3(finalV.Normal);
                        hitV.MaterialIndex = SS.nodes[hinst].matId[finalRaw & 15];
                    }
                    resolved = true;
                }
                if (resolved) {
                    const GkNodeProxy& hn = SS.nodes[hitNode];
                    { // CalculateMotionVector, Shading.slang:50-58
                        const f4 cur = mulM(U.ViewProjectionUnJit, mk4(hitV.Position.x, hitV.Position.y, hitV.Position.z, 1));
                        const float cx = cur.x / cur.w * 0.5f, cy = cur.y / cur.w * 0.5f;
                        float PM[16];
                        for (int c = 0; c < 4; ++c)
                            for (int r = 0; r < 4; ++r) {
                                float s = 0;
                                for (int k = 0; k < 4; ++k) s += U.PrevViewProjectionUnJit[k * 4 + r] * hn.combinedPrevTS[c * 4 + k];
                                PM[c * 4 + r] = s;
                            }
                        const f4 prev = mulM(PM, mk4(hitV.Position.x, hitV.Position.y, hitV.Position.z, 1));
                        const float qx = prev.x / prev.w * 0.5f, qy = prev.y / prev.w * 0.5f;
                        PL.motion[pixel] = make_float2((qx - cx) * float(P.width), (qy - cy) * float(P.height));
                    }
                    const GkMaterial& mat = SS.materials[hitV.MaterialIndex];
                    storeHalf4(PL.albedo, pixel, mat.Diffuse[0], mat.Diffuse[1], mat.Diffuse[2], mat.Diffuse[3]);
                    storeHalf4(PL.normal, pixel, hitV.Normal.x, hitV.Normal.y, hitV.Normal.z, mat.Fuzziness);
                    PL.objectId0[pixel] = hn.instanceId;
                    {
                        const f4 clip = mulM(U.ViewProjection, mk4(hitV.Position.x, hitV.Position.y, hitV.Position.z, 1));
                        PL.depth[pixel] = clip.z / clip.w;
                    }
                    ppos = hitV.Position, pnrm = hitV.Normal, pmat = hitV.MaterialIndex;
                    S.primPosMat[path] = make_float4(ppos.x, ppos.y, ppos.z, __uint_as_float(pmat));
                    S.primNrm[path] = make_float4(pnrm.x, pnrm.y, pnrm.z, offLen);
                    bits = (mat.MaterialModel == GK_MAT_DIELECTRIC) ? F_PRIM_DIELECTRIC : 0u;
                    sample = 0;
                    accDirty = true, haveAcc = true, havePrimary = true, accW = offLen;
                    if (mat.MaterialModel == GK_MAT_DIFFUSE_LIGHT) { // Shading.slang:1003-1008
                        accD = mk3(mat.Diffuse[0], mat.Diffuse[1], mat.Diffuse[2]), accS = mk3(0, 0, 0);
                        finished = true;
                        phase = PH_FINAL;
                    } else {
                        phase = samples > 0 ? PH_START_SAMPLE : PH_AFTER_SAMPLES;
                    }
                }
            } else if (state == ST_BOUNCE) {
                // ---- second half of GetRayColor (Shading.slang:965-995)
                loadCurrent();
                const float4 hr = src.hitTuvp();
                const uint32_t hinst = src.hitInst();
                const uint32_t maxBounces = (bits & F_PRIM_DIELECTRIC) ? U.MaxNumberOfBounces : U.NumberOfBounces;
                bool terminated;
                if (hinst != kInvalid) {
                    const float4 ro = src.rayO(), rd = src.rayD();
                    Vtx hv;
                    resolveHit(SS, mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), hr.x, hr.y, hr.z, __float_as_uint(hr.w), hinst, hv);
                    vpos = hv.Position, vnrm = hv.Normal, vmat = hv.MaterialIndex;
                    const GkMaterial& hm = SS.materials[vmat];
                    const bool light = hm.MaterialModel == GK_MAT_DIFFUSE_LIGHT;
                    if (light || !(bits & F_CHANCE_REFLECT)) color = color * mk3(hm.Diffuse[0], hm.Diffuse[1], hm.Diffuse[2]);
                    if (bits & F_TERMINATE_ON_HIT) {
                        color = mk3(0, 0, 0);
                        terminated = true;
                    } else terminated = light;
                } else {
                    color = color * skyColor(U);
                    terminated = true;
                }
                if (terminated) phase = PH_END_SAMPLE;
                else if (bounce == 0) {
                    bounce = 1;
                    phase = (1u < maxBounces) ? PH_SETUP_BOUNCE : PH_END_SAMPLE;
                } else {
                    // sun next-event estimation, Shading.slang:1037-1046 (guard variable is never written: always taken)
                    if (U.HasSun && (randomFloat(rng) < 0.5f)) {
                        const f3 lv = mk3(U.SunDirection[0], U.SunDirection[1], U.SunDirection[2]);
                        const f3 cone = alignWithNormal(randomInCone(rng, cosf(0.25f / 180.f * kPi)), lv);
                        e.kind = 2, e.o = vpos + vnrm * kTraceOffset, e.d = cone, e.tmin = kEps, e.tmax = kMaxTrace;
                        state = ST_NEE;
                    } else phase = PH_POST_NEE;
                }
            } else if (state == ST_NEE) {
                loadCurrent();
                const bool occluded = src.occluded();
                if (!occluded) {
                    color = color * mk3(U.SunColor[0], U.SunColor[1], U.SunColor[2]);
                    phase = PH_END_SAMPLE;
                } else phase = PH_POST_NEE;
            } else if (state == ST_DIRECT) {
                shadowTerm = src.occluded() ? 0.f : 1.f;
                phase = PH_FINAL;
            }

            // ---- run the control flow of FPathTracingRenderer::Render until the next ray (Shading.slang:1010-1081)
            while (phase != PH_EXIT) {
                const uint32_t maxBounces = (bits & F_PRIM_DIELECTRIC) ? U.MaxNumberOfBounces : U.NumberOfBounces;
                switch (phase) {
                case PH_START_SAMPLE:
                    needPrimary();
                    color = mk3(1, 1, 1);
                    dir = normalize3(ppos - eye);
                    vpos = ppos, vnrm = pnrm, vmat = pmat;
                    bounce = 0;
                    phase = PH_SETUP_BOUNCE;
                    break;
                case PH_SETUP_BOUNCE:
                    setupBounce(SS, rng, vpos, vnrm, vmat, dir, bits, bounce == 0, e);
                    state = ST_BOUNCE;
                    phase = PH_EXIT;
                    break;
                case PH_POST_NEE: {
                    // early exit, Shading.slang:1049-1056 (the random number is drawn only for non-dielectric primaries)
                    const bool earlyExit = !(bits & F_PRIM_DIELECTRIC) && (randomFloat(rng) < 0.5f);
                    if (bounce == maxBounces - 1 || earlyExit) {
                        color = color * interpolateAmbientCubes(SS, vpos, vnrm);
                        phase = PH_END_SAMPLE;
                    } else {
                        ++bounce;
                        phase = PH_SETUP_BOUNCE;
                    }
                    break;
                }
                case PH_END_SAMPLE: {
                    needAcc();
                    if (bits & F_HIT_METAL) {
                        needPrimary();
                        const GkMaterial& pm = SS.materials[pmat];
                        color = color * mk3(pm.Diffuse[0], pm.Diffuse[1], pm.Diffuse[2]);
                    }
                    if (bits & F_HIT_REFLECT) accS = accS + color;
                    else accD = accD + color;
                    ++sample;
                    phase = sample < samples ? PH_START_SAMPLE : PH_AFTER_SAMPLES;
                    break;
                }
                case PH_AFTER_SAMPLES: {
                    needAcc();
                    if (samples != 1u) { // x / 1.0f == x bit for bit: the 1-spp real-time frame skips two IEEE vector divisions per path
                        accD = accD / float(samples);
                        accS = accS / float(samples);
                    }
                    if (U.HasSun) { // DirectIlluminate, Shading.slang:826-845
                        needPrimary();
                        const f3 lv = mk3(U.SunDirection[0], U.SunDirection[1], U.SunDirection[2]);
                        const f3 cone = alignWithNormal(randomInCone(rng, cosf(0.25f / 180.f * kPi)), lv);
                        e.kind = 2, e.o = ppos, e.d = cone, e.tmin = kEps, e.tmax = kMaxTrace;
                        state = ST_DIRECT;
                        phase = PH_EXIT;
                    } else {
                        shadowTerm = 0.f;
                        phase = PH_FINAL;
                    }
                    break;
                }
                case PH_FINAL: {
                    if (!finished) {
                        needPrimary();
                        needAcc();
                        const f3 lv = mk3(U.SunDirection[0], U.SunDirection[1], U.SunDirection[2]);
                        const float d = fmaxx(dot3(lv, normalize3(pnrm)), 0.0f) * kInvPi;
                        accD = accD + mk3(U.SunColor[0], U.SunColor[1], U.SunColor[2]) * d * shadowTerm;
                    }
                    state = ST_DONE;
                    phase = PH_EXIT;
                    break;
                }
                default: phase = PH_EXIT; break;
                }
            }

            // ---- write back
            S.rng[path] = make_uint4(rng.x, rng.y, rng.z, rng.w);
            S.rays[path] = rays;
            if (state == ST_BOUNCE || state == ST_NEE) {
                S.posMat[path] = make_float4(vpos.x, vpos.y, vpos.z, __uint_as_float(vmat));
                S.dirT[path] = make_float4(dir.x, dir.y, dir.z, 0);
                S.throughput[path] = make_float4(color.x, color.y, color.z, 0);
            }
            if (accDirty) {
                S.accDiffuse[path] = make_float4(accD.x, accD.y, accD.z, accW);
                S.accSpec[path] = make_float4(accS.x, accS.y, accS.z, accW);
            }
            S.nrmFlags[path] = make_float4(vnrm.x, vnrm.y, vnrm.z, __uint_as_float(packFlags(state, bits, bounce, sample)));
        }
    }

}

// Appends the rays of a warp to the next wave's queues: one atomic per warp and queue.
__device__ __forceinline__ void appendRay(const Emit& e, uint32_t path, const RayQueue& outE, const RayQueue& outS)
{
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int kind = 1; kind <= 2; ++kind) {
        const RayQueue& Q = kind == 1 ? outE : outS;
        const unsigned m = __ballot_sync(full, e.kind == kind);
        if (m) {
            uint32_t base = 0;
            if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(Q.count, (uint32_t)__popc(m));
            base = __shfl_sync(full, base, __ffs(m) - 1);
            if (e.kind == kind) {
                const uint32_t dst = base + __popc(m & ((1u << lane) - 1u));
                Q.o_tmin[dst] = make_float4(e.o.x, e.o.y, e.o.z, e.tmin);
                Q.d_tmax[dst] = make_float4(e.d.x, e.d.y, e.d.z, e.tmax);
                Q.path[dst] = path;
            }
        }
    }
}

template <int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) k_shade(const GkUniformBufferObject* __restrict__ ubo, FrameParams P, ShadeScene SS, PathState S, PlaneView PL, RayQueue inE,
                                               uint32_t countE, RayQueue inS, uint32_t countS, RayQueue outE, RayQueue outS)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    Emit e;
    e.kind = 0;
    uint32_t path = 0;
    if (i < countE + countS) {
        const bool fromExtend = i < countE;
        const uint32_t slot = fromExtend ? i : i - countE;
        path = fromExtend ? inE.path[slot] : inS.path[slot];
        shadePath(*ubo, P, SS, S, PL, path, QueueSrc{inE, inS, slot}, e);
    }
    appendRay(e, path, outE, outS);
}

// Device-driven form of the same kernel: the queue sizes are read from device memory and the blocks stride over the
// queue, so a wave can be enqueued before the host knows how many rays the previous wave produced.
// The 784-byte UBO travels as a kernel parameter (constant bank): ncu attributed 5 % of this kernel's stall samples to the first
// global load of a UBO matrix; parameter space makes every U.* operand a constant-cache / uniform-register read.
template <int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) k_shade_stream(const __grid_constant__ GkUniformBufferObject U, FrameParams P, ShadeScene SS, PathState S, PlaneView PL, RayQueue inE,
                                                      RayQueue inS, RayQueue outE, RayQueue outS)
{
    const uint32_t countE = *inE.count, countS = *inS.count, total = countE + countS;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) { // block-uniform trip count
        const uint32_t i = base + threadIdx.x;
        Emit e;
        e.kind = 0;
        uint32_t path = 0;
        if (i < total) {
            const bool fromExtend = i < countE;
            const uint32_t slot = fromExtend ? i : i - countE;
            path = fromExtend ? inE.path[slot] : inS.path[slot];
            shadePath(U, P, SS, S, PL, path, QueueSrc{inE, inS, slot}, e);
        }
        appendRay(e, path, outE, outS);
    }
}

// End of a wave: publishes the sizes of the next wave's queues to the host (mapped pinned memory; the tag goes last).
// The host polls these slots a few waves behind the device instead of synchronising the stream after every wave.
__global__ void k_wave_end(const uint32_t* __restrict__ nextE, const uint32_t* __restrict__ nextS, volatile uint32_t* hostSlot, uint32_t tag)
{
    hostSlot[0] = *nextE;
    hostSlot[1] = *nextS;
    __threadfence_system();
    hostSlot[2] = tag;
}

// Tail of the frame: when few paths are still alive, one launch walks every one of them to its end
// (trace -> shade -> trace ...), one path per lane, instead of paying three launches and a queue
// count read-back per wave.  counters[0] / [1] receive the extension / shadow rays traced here.
__global__ void __launch_bounds__(128, 4) k_tail(const GkUniformBufferObject* __restrict__ ubo, FrameParams P, SceneView V, ShadeScene SS, PathState S, PlaneView PL,
                                                 RayQueue inE, uint32_t countE, RayQueue inS, uint32_t countS, unsigned long long* __restrict__ counters, int countsOnDevice)
{
    if (countsOnDevice) countE = *inE.count, countS = *inS.count; // streamed wave loop: the host only knows an upper bound (it sized the grid)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nE = 0, nS = 0;
    if (i < countE + countS) {
        const bool fromExtend = i < countE;
        const uint32_t slot = fromExtend ? i : i - countE;
        const RayQueue& Q = fromExtend ? inE : inS;
        const uint32_t path = Q.path[slot];
        RegSrc r;
        r.o = Q.o_tmin[slot], r.d = Q.d_tmax[slot];
        int kind = fromExtend ? 1 : 2;
        for (int guard = 0; guard < 65536 && kind != 0; ++guard) {
            Hit h{r.d.w, 0.f, 0.f, kInvalid, kInvalid};
            bool hit = false;
            if (r.d.w > 0.0f) {
                const f3 O = mk3(r.o.x, r.o.y, r.o.z), dn = normalizeRayDir(mk3(r.d.x, r.d.y, r.d.z));
                if (kind == 1) hit = traverseLane<false, false>(V, O, dn, r.o.w, h, nullptr), ++nE;
                else hit = traverseLane<true, false>(V, O, dn, r.o.w, h, nullptr), ++nS;
            }
            r.tuvp = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
            r.inst = h.inst;
            r.occ = hit;
            Emit e;
            e.kind = 0;
            shadePath(*ubo, P, SS, S, PL, path, r, e);
            kind = e.kind;
            r.o = make_float4(e.o.x, e.o.y, e.o.z, e.tmin), r.d = make_float4(e.d.x, e.d.y, e.d.z, e.tmax);
        }
        if (kind != 0) counters[2] = 1ull; // the step guard expired with the path still alive: the host reports it
    }
    const unsigned full = 0xffffffffu;
    for (int o = 16; o; o >>= 1) nE += __shfl_xor_sync(full, nE, o), nS += __shfl_xor_sync(full, nS, o);
    if ((threadIdx.x & 31u) == 0) {
        if (nE) atomicAdd(&counters[0], (unsigned long long)nE);
        if (nS) atomicAdd(&counters[1], (unsigned long long)nS);
    }
}

// The same tail with eight lanes per path: what is left at the end of a frame is a few hundred long paths (dielectric
// primaries run to MaxNumberOfBounces) whose rays depend on one another, so the launch is bound by the latency of a single
// ray.  The cooperative traversal (lane j tests child j / triangle j) cuts that latency several times; lane 0 of a group runs
// the path's state machine and hands the next ray to its group through shuffles.
__global__ void __launch_bounds__(256, 2) k_tail_coop(const __grid_constant__ GkUniformBufferObject U, FrameParams P, SceneView V, ShadeScene SS, PathState S, PlaneView PL,
                                                      RayQueue inE, RayQueue inS, unsigned long long* __restrict__ counters)
{
    __shared__ uint2 stack[kRaysPerBlock * kStackStride];
    const uint32_t countE = *inE.count, countS = *inS.count;
    const unsigned lane = threadIdx.x & 31u, sub = lane & 7u, shift = lane & 24u;
    const unsigned gmask = 0xffu << shift;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; // one group of eight lanes per queue entry
    uint32_t nE = 0, nS = 0;
    if (i < countE + countS) { // whole groups take the branch together
        const bool fromExtend = i < countE;
        const uint32_t slot = fromExtend ? i : i - countE;
        const RayQueue& Q = fromExtend ? inE : inS;
        const uint32_t path = Q.path[slot];
        RegSrc r;
        r.o = Q.o_tmin[slot], r.d = Q.d_tmax[slot];
        int kind = fromExtend ? 1 : 2;
        int guard = 0;
        for (; guard < 65536 && kind != 0; ++guard) {
            Hit h{r.d.w, 0.f, 0.f, kInvalid, kInvalid};
            bool hit = false;
            if (r.d.w > 0.0f) { // uniform within the group: every lane holds the same ray
                const f3 O = mk3(r.o.x, r.o.y, r.o.z), dn = normalizeRayDir(mk3(r.d.x, r.d.y, r.d.z));
                if (kind == 1) hit = traverseCoop<false, false>(V, O, dn, r.o.w, h, stackRowOf(stack), nullptr);
                else hit = traverseCoop<true, false>(V, O, dn, r.o.w, h, stackRowOf(stack), nullptr);
                if (sub == 0) (kind == 1 ? nE : nS)++;
            }
            Emit e;
            e.kind = 0, e.o = e.d = mk3(0, 0, 0), e.tmin = e.tmax = 0.f;
            if (sub == 0) {
                r.tuvp = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
                r.inst = h.inst;
                r.occ = hit;
                shadePath(U, P, SS, S, PL, path, r, e);
            }
            const int src = (int)shift; // lane 0 of the group
            kind = __shfl_sync(gmask, e.kind, src);
            r.o = make_float4(__shfl_sync(gmask, e.o.x, src), __shfl_sync(gmask, e.o.y, src), __shfl_sync(gmask, e.o.z, src), __shfl_sync(gmask, e.tmin, src));
            r.d = make_float4(__shfl_sync(gmask, e.d.x, src), __shfl_sync(gmask, e.d.y, src), __shfl_sync(gmask, e.d.z, src), __shfl_sync(gmask, e.tmax, src));
        }
        if (kind != 0 && sub == 0) counters[2] = 1ull; // the step guard expired with the path still alive: the host reports it
    }
    if (nE) atomicAdd(&counters[0], (unsigned long long)nE);
    if (nS) atomicAdd(&counters[1], (unsigned long long)nS);
}

// -------------------------------------------------------------------------------- accumulate
// Core.PathTracing :88-100 — final stores of the pixel (RGBA16F render targets + fp32 parity copies)
__global__ void __launch_bounds__(256) k_accumulate(FrameParams P, PathState S, PlaneView PL)
{
    const uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
    if (path >= P.pathCount) return;
    const uint32_t pixel = S.pixel[path];
    if (pixel == kInvalid) return;
    const float4 d = S.accDiffuse[path], s = S.accSpec[path];
    PL.radDiffuse[pixel] = d;
    PL.radSpec[pixel] = s;
    storeHalf4(PL.outDiffuse, pixel, d.x, d.y, d.z, d.w);
    storeHalf4(PL.outSpec, pixel, s.x, s.y, s.z, s.w);
    PL.rayCount[pixel] = S.rays[path];
}

// -------------------------------------------------------------------------------- utilities
template <bool kAnyHit, bool kCoop, class RayIO>
static void launchMapped(cudaStream_t st, unsigned grid, unsigned block, const SceneView& V, const RayIO& io, uint32_t count, TraversalStats* ts, const uint32_t* countPtr = nullptr)
{
    if (ts) k_trace<kAnyHit, kCoop, true, RayIO><<<grid, block, 0, st>>>(V, io, count, ts, countPtr);
    else k_trace<kAnyHit, kCoop, false, RayIO><<<grid, block, 0, st>>>(V, io, count, ts, countPtr);
}

// Small waves of the streamed loop: the eight-lanes-per-ray kernel (lowest latency per ray), sized by an upper bound.
template <bool kAnyHit, class RayIO>
static void launchCoopBounded(Context& c, const SceneView& V, const RayIO& io, uint32_t bound, cudaStream_t stream, const uint32_t* countPtr)
{
    const unsigned grid = std::max(1u, (bound + kRaysPerBlock - 1) / kRaysPerBlock);
    launchMapped<kAnyHit, true>(stream, grid, 256, V, io, bound, c.travStats ? c.dTravStats : nullptr, countPtr);
}

// Mid-size waves of the streamed loop (sched_min_rays): the one-ray-per-lane kernel, sized by an upper bound.
template <bool kAnyHit, class RayIO>
static void launchLaneBounded(Context& c, const SceneView& V, const RayIO& io, uint32_t bound, cudaStream_t stream, const uint32_t* countPtr)
{
    launchMapped<kAnyHit, false>(stream, std::max(1u, (bound + c.laneBlock - 1) / c.laneBlock), c.laneBlock, V, io, bound, c.travStats ? c.dTravStats : nullptr, countPtr);
}

// Next zeroed fetch cursor of the frame (the block of cursors is cleared once per frame / per intersect call).
static uint32_t* nextCursor(Context& c, cudaStream_t)
{
    if (c.cursorNext >= Context::kCursorCount) { // more launches than cursors since the last clear (frames with hundreds of waves): drain and clear
        cudaStreamSynchronize(c.stream);
        if (c.stream2) cudaStreamSynchronize(c.stream2);
        cudaMemsetAsync(c.dCursors, 0, sizeof(uint32_t) * Context::kCursorCount, c.stream);
        cudaStreamSynchronize(c.stream);
        c.cursorNext = 0;
    }
    return c.dCursors + c.cursorNext++;
}

// `count` sizes the grid (an upper bound is enough when `countPtr` supplies the exact number on the device).
template <bool kAnyHit, class RayIO>
static void launchSched(Context& c, const SceneView& V, const RayIO& io, uint32_t count, cudaStream_t stream, const uint32_t* countPtr = nullptr)
{
    if (c.schedBlocksPerSm == 0) {
        int nb = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_sched<false, false, QueueIO>, kSchedBlock, 0);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
        c.schedBlocksPerSm = nb > 0 ? nb : 1, c.smCount = sms > 0 ? sms : 1;
    }
    const unsigned resident = (unsigned)(c.schedBlocksPerSm * c.smCount);
    const unsigned grid = std::max(1u, std::min(resident, (count + kSchedBlock - 1) / kSchedBlock));
    const SchedParams prm{std::min(32u, std::max(1u, c.schedRefillMin)), c.schedBiasN, std::min(33u, std::max(1u, c.schedKeepN)), std::min(33u, std::max(1u, c.schedKeepT))};
    uint32_t* cursor = nextCursor(c, stream);
    if (c.travStats) k_trace_sched<kAnyHit, true, RayIO><<<grid, kSchedBlock, 0, stream>>>(V, io, countPtr, count, cursor, prm, c.dTravStats, c.dSchedStats);
    else k_trace_sched<kAnyHit, false, RayIO><<<grid, kSchedBlock, 0, stream>>>(V, io, countPtr, count, cursor, prm, nullptr, nullptr);
}

template <bool kAnyHit, class RayIO>
static void launchTraceIO(Context& c, const SceneView& V, const RayIO& io, uint32_t count, cudaStream_t stream = nullptr)
{
    if (!stream) stream = c.stream;
    if (c.traceVariant == 1 && count >= c.schedMinRays) {
        launchSched<kAnyHit>(c, V, io, count, stream);
        return;
    }
    const bool coop = count < std::min(c.coopThreshold, 65536u); // one-shot batches: the cooperative mapping pays below 65 k rays (r01 sweep)
    const unsigned per = coop ? kRaysPerBlock : c.laneBlock; // rays per block
    const unsigned grid = (count + per - 1) / per;
    TraversalStats* ts = c.travStats ? c.dTravStats : nullptr;
    if (coop) launchMapped<kAnyHit, true>(stream, grid, 256, V, io, count, ts);
    else launchMapped<kAnyHit, false>(stream, grid, c.laneBlock, V, io, count, ts);
}

template <bool kShadow>
static void launchTrace(Context& c, const SceneView& V, const RayQueue& Q, uint32_t count, cudaStream_t stream = nullptr)
{
    if (kShadow) launchTraceIO<true>(c, V, ShadowIO{Q}, count, stream);
    else launchTraceIO<false>(c, V, QueueIO{Q}, count, stream);
}

// -------------------------------------------------------------------------------- host side
static inline unsigned gridFor(size_t n, unsigned block = 256) { return (unsigned)((n + block - 1) / block); }

template <class T> static cudaError_t allocInto(std::vector<void*>& pool, T*& p, size_t n)
{
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e == cudaSuccess) {
        pool.push_back(q);
        p = reinterpret_cast<T*>(q);
    }
    return e;
}

void freeFrameResources(Context& c)
{
    if (c.asyncCopySrc) cudaEventSynchronize(c.evCopyDone), c.asyncCopySrc = nullptr;
    exchangeClosePeers(c); // the mapped peer planes belong to the extent being torn down
    frameShardRelease(c);
    c.peers.myId0 = c.peers.myId1 = nullptr;
    for (void* p : c.pathAllocs) cudaFree(p);
    for (void* p : c.queueAllocs) cudaFree(p);
    c.pathAllocs.clear(), c.queueAllocs.clear();
    for (int i = 0; i < GK_PLANE_COUNT; ++i) {
        if (c.planes.p[i]) cudaFree(c.planes.p[i]);
        c.planes.p[i] = nullptr, c.planes.bytes[i] = 0;
    }
    if (c.dUbo) cudaFree(c.dUbo);
    c.dUbo = nullptr;
    if (c.hCounts) cudaFreeHost(c.hCounts);
    c.hCounts = nullptr;
    if (c.dTravStats) cudaFree(c.dTravStats);
    c.dTravStats = nullptr;
    if (c.dTailCounters) cudaFree(c.dTailCounters);
    c.dTailCounters = nullptr;
    if (c.dCursors) cudaFree(c.dCursors);
    c.dCursors = nullptr;
    if (c.dSchedStats) cudaFree(c.dSchedStats);
    c.dSchedStats = nullptr;
    if (c.dOverflow) cudaFree(c.dOverflow);
    c.dOverflow = nullptr;
    if (c.hWave) cudaFreeHost((void*)c.hWave);
    c.hWave = nullptr, c.dWave = nullptr;
}

static size_t planePixelBytes(int plane)
{
    switch (plane) {
    case GK_PLANE_OBJECT_ID0:
    case GK_PLANE_OBJECT_ID1:
    case GK_PLANE_DEPTH:
    case GK_PLANE_PRIMARY_T:
    case GK_PLANE_RAY_COUNT: return 4;
    case GK_PLANE_RADIANCE_DIFFUSE_F32:
    case GK_PLANE_RADIANCE_SPECULAR_F32: return 16;
    default: return 8; // RGBA16F, RG32F, 2 x u32
    }
}

GkStatus allocFrameResources(Context& c)
{
    freeFrameResources(c);
    const size_t px = (size_t)c.width * c.height;
    // owned rows of this rank's tile set
    uint32_t owned = 0;
    for (uint32_t r = 0; r < c.height; ++r)
        if ((r / c.tileRows) % c.traceTileCount == c.traceTileIndex) ++owned;
    // paths are laid out in whole tile-row blocks so that pathToPixel stays arithmetic
    const uint32_t blocks = (c.height + c.tileRows * c.traceTileCount - 1) / (c.tileRows * c.traceTileCount);
    c.ownedRows = owned;
    c.pathCount = blocks * c.tileRows * c.width;
    const size_t n = c.pathCount;
    for (int i = 0; i < GK_PLANE_COUNT; ++i) {
        c.planes.bytes[i] = px * planePixelBytes(i);
        GK_CUDA(cudaMalloc(&c.planes.p[i], c.planes.bytes[i]));
        GK_CUDA(cudaMemsetAsync(c.planes.p[i], 0, c.planes.bytes[i], c.stream));
    }
    PathState& S = c.paths;
    GK_CUDA(allocInto(c.pathAllocs, S.rng, n));
    GK_CUDA(allocInto(c.pathAllocs, S.posMat, n));
    GK_CUDA(allocInto(c.pathAllocs, S.nrmFlags, n));
    GK_CUDA(allocInto(c.pathAllocs, S.dirT, n));
    GK_CUDA(allocInto(c.pathAllocs, S.throughput, n));
    GK_CUDA(allocInto(c.pathAllocs, S.primPosMat, n));
    GK_CUDA(allocInto(c.pathAllocs, S.primNrm, n));
    GK_CUDA(allocInto(c.pathAllocs, S.accDiffuse, n));
    GK_CUDA(allocInto(c.pathAllocs, S.accSpec, n));
    GK_CUDA(allocInto(c.pathAllocs, S.pixel, n));
    GK_CUDA(allocInto(c.pathAllocs, S.rays, n));
    GK_CUDA(allocInto(c.pathAllocs, S.dofVertex, n));
    GK_CUDA(allocInto(c.pathAllocs, S.dofNormal, n));
    for (int k = 0; k < 4; ++k) {
        RayQueue& Q = k < 2 ? c.extendQ[k] : c.shadowQ[k - 2];
        GK_CUDA(allocInto(c.queueAllocs, Q.o_tmin, n));
        GK_CUDA(allocInto(c.queueAllocs, Q.d_tmax, n));
        GK_CUDA(allocInto(c.queueAllocs, Q.path, n));
        GK_CUDA(allocInto(c.queueAllocs, Q.hit_tuvp, n));
        GK_CUDA(allocInto(c.queueAllocs, Q.hit_inst, n));
        GK_CUDA(allocInto(c.queueAllocs, Q.count, 4));
        GK_CUDA(cudaMemsetAsync(Q.count, 0, 16, c.stream));
    }
    GK_CUDA(cudaMalloc(&c.dUbo, sizeof(GkUniformBufferObject)));
    GK_CUDA(cudaMallocHost(&c.hCounts, 64));
    GK_CUDA(cudaMalloc(&c.dTailCounters, 3 * sizeof(unsigned long long)));
    GK_CUDA(cudaMalloc(&c.dTravStats, sizeof(TraversalStats)));
    GK_CUDA(cudaMemsetAsync(c.dTravStats, 0, sizeof(TraversalStats), c.stream));
    GK_CUDA(cudaMalloc(&c.dCursors, sizeof(uint32_t) * Context::kCursorCount));
    GK_CUDA(cudaMemsetAsync(c.dCursors, 0, sizeof(uint32_t) * Context::kCursorCount, c.stream));
    c.cursorNext = 0;
    GK_CUDA(cudaMalloc(&c.dSchedStats, sizeof(SchedStats)));
    GK_CUDA(cudaMemsetAsync(c.dSchedStats, 0, sizeof(SchedStats), c.stream));
    GK_CUDA(cudaMalloc(&c.dOverflow, sizeof(uint32_t)));
    GK_CUDA(cudaMemsetAsync(c.dOverflow, 0, sizeof(uint32_t), c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return GK_OK;
}

static ShadeScene shadeSceneOf(const Context& c)
{
    ShadeScene s;
    s.verts = c.dGpuVerts.p, s.indices = c.dIndices.p, s.models = c.dModels.p, s.materials = c.dMaterials.p, s.nodes = c.dNodes.p, s.inst = c.dInst.p;
    s.cubes = c.haveProbes ? c.dCubes.p : nullptr, s.voxels = c.haveProbes ? c.dVoxels.p : nullptr;
    s.materialCount = c.materialCount;
    return s;
}

static PlaneView planeViewOf(const Context& c)
{
    PlaneView v;
    v.outDiffuse = (__half*)c.planes.p[GK_PLANE_OUTPUT_DIFFUSE], v.outSpec = (__half*)c.planes.p[GK_PLANE_OUTPUT_SPECULAR];
    v.albedo = (__half*)c.planes.p[GK_PLANE_ALBEDO], v.normal = (__half*)c.planes.p[GK_PLANE_NORMAL];
    v.objectId0 = (uint32_t*)c.planes.p[GK_PLANE_OBJECT_ID0], v.motion = (float2*)c.planes.p[GK_PLANE_MOTION], v.depth = (float*)c.planes.p[GK_PLANE_DEPTH];
    v.radDiffuse = (float4*)c.planes.p[GK_PLANE_RADIANCE_DIFFUSE_F32], v.radSpec = (float4*)c.planes.p[GK_PLANE_RADIANCE_SPECULAR_F32];
    v.primaryIds = (uint2*)c.planes.p[GK_PLANE_PRIMARY_IDS], v.primaryT = (float*)c.planes.p[GK_PLANE_PRIMARY_T], v.rayCount = (uint32_t*)c.planes.p[GK_PLANE_RAY_COUNT];
    return v;
}

static cudaEvent_t poolEvent(Context& c, size_t i)
{
    while (c.evPool.size() <= i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c.evPool.push_back(e);
    }
    return c.evPool[i];
}

// Waits until the device has published wave slot `w` of this frame (k_wave_end); polls pinned memory, no stream sync.
static GkStatus waitWaveSlot(Context& c, uint32_t w, uint32_t tag, uint32_t& countE, uint32_t& countS)
{
    volatile uint32_t* slot = c.hWave + 4 * (size_t)w;
    for (uint64_t spin = 0; slot[2] != tag; ++spin) {
        __builtin_ia32_pause();
        if ((spin & 0xfffffu) == 0xfffffu) { // every ~million polls: has the stream died?
            const cudaError_t e = cudaStreamQuery(c.stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                setLastError(std::string("gk_trace_frame: ") + cudaGetErrorString(e));
                return GK_ERR_CUDA;
            }
        }
    }
    countE = slot[0], countS = slot[1];
    return GK_OK;
}

// The wave loop of the scheduled kernels: every kernel takes its queue sizes from device memory, so the host enqueues
// wave w without knowing how many rays wave w-1 produced.  It learns the sizes `waveLookahead` waves late from slots the
// device writes into mapped pinned memory (k_wave_end) and stops enqueuing once a wave was empty; at most `waveLookahead`
// empty waves (a few microseconds of launches each) are enqueued.  No stream synchronisation inside the frame.
static GkStatus traceFrameStreamed(Context& c)
{
    cudaStream_t st = c.stream;
    const uint32_t n = c.pathCount;
    FrameParams P{c.width, c.height, c.traceTileIndex, c.traceTileCount, c.tileRows, n, (c.microTiles == 2 && c.width % 16 == 0 && c.tileRows % 16 == 0) ? 2u : (c.microTiles && c.width % 8 == 0 && c.tileRows % 4 == 0) ? 1u : 0u};
    const SceneView V = c.view();
    const ShadeScene SS = shadeSceneOf(c);
    const PlaneView PL = planeViewOf(c);
    GkFrameStats& fs = c.stats;
    fs.primaryRays = fs.extensionRays = fs.shadowRays = 0;
    fs.waves = fs.launches = 0;
    fs.msGenerate = fs.msExtend = fs.msShade = fs.msShadow = fs.msAccumulate = fs.msTail = fs.msTrace = 0;
    fs.tailPaths = 0, fs.tailExtensionRays = fs.tailShadowRays = 0;
    size_t ev = 0;
    struct Span { size_t a, b; int kind; };
    std::vector<Span> spans;
    auto mark = [&]() { cudaEvent_t e = poolEvent(c, ev); cudaEventRecord(e, st); return ev++; };
    auto markOn = [&](cudaStream_t s2) { cudaEvent_t e = poolEvent(c, ev); cudaEventRecord(e, s2); return ev++; };
    if (c.concurrentShadow && !c.stream2) {
        GK_CUDA(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
        GK_CUDA(cudaEventCreateWithFlags(&c.evFork, cudaEventDisableTiming));
        GK_CUDA(cudaEventCreateWithFlags(&c.evJoin, cudaEventDisableTiming));
    }
    if (!c.hWave) {
        GK_CUDA(cudaHostAlloc((void**)&c.hWave, sizeof(uint32_t) * 4 * Context::kWaveLimit, cudaHostAllocMapped));
        memset((void*)c.hWave, 0, sizeof(uint32_t) * 4 * Context::kWaveLimit);
        GK_CUDA(cudaHostGetDevicePointer((void**)&c.dWave, (void*)c.hWave, 0));
    }
    const uint32_t tag = ++c.frameTag ? c.frameTag : ++c.frameTag; // never 0
    GK_CUDA(cudaMemcpyAsync(c.dUbo, &c.ubo, sizeof(GkUniformBufferObject), cudaMemcpyHostToDevice, st));
    if (c.travStats) {
        GK_CUDA(cudaMemsetAsync(c.dTravStats, 0, sizeof(TraversalStats), st));
        GK_CUDA(cudaMemsetAsync(c.dSchedStats, 0, sizeof(SchedStats), st));
    }
    if (c.cursorNext) { // queue cursors of the previous frame's launches
        GK_CUDA(cudaMemsetAsync(c.dCursors, 0, sizeof(uint32_t) * c.cursorNext, st));
        c.cursorNext = 0;
    }
    const size_t evStart = mark();
    k_generate<<<gridFor(n), 256, 0, st>>>(c.dUbo, P, c.paths, c.extendQ[0]);
    fs.launches++;
    GK_CUDA(cudaMemsetAsync(c.shadowQ[0].count, 0, sizeof(uint32_t), st));
    const size_t evGen = mark();
    spans.push_back({evStart, evGen, 0});
    fs.primaryRays = (uint64_t)c.ownedRows * c.width;
    c.capturedCount = 0;
    // capture / statistics frames keep the host in lock step (it needs the exact size of the wave it is about to enqueue)
    const uint32_t look = (c.captureWave >= 0) ? 0u : c.waveLookahead;
    int shadeBlocksPerSm = c.shadeMinBlocks >= 4 ? 4 : c.shadeMinBlocks >= 3 ? 3 : 2;
    if (c.smCount == 0) cudaDeviceGetAttribute(&c.smCount, cudaDevAttrMultiProcessorCount, c.device);
    uint32_t bound = n;      // upper bound of the rays of the wave being enqueued (paths alive never increase)
    uint32_t enqueued = 0;   // waves enqueued so far
    uint32_t known = 0;      // slots read so far: slot w holds the queue sizes of wave w + 1
    bool drained = false, tailRan = false;
    size_t tailSpan[2] = {0, 0};
    auto consume = [&](uint32_t upTo) -> GkStatus { // reads slots [known, upTo)
        for (; known < upTo; ++known) {
            uint32_t e = 0, s2 = 0;
            const GkStatus r = waitWaveSlot(c, known, tag, e, s2);
            if (r != GK_OK) return r;
            // the rays the tail launch starts from are counted by the tail itself (it counts every ray it traces)
            if (!(tailRan && known + 1 == enqueued)) fs.extensionRays += e, fs.shadowRays += s2;
            if (e + s2) fs.waves = known + 2;
            else drained = true;
            bound = e + s2;
        }
        return GK_OK;
    };
    fs.waves = 1;
    for (uint32_t wave = 0; wave < Context::kWaveLimit; ++wave) {
        if (wave > look) {
            const GkStatus r = consume(wave - look);
            if (r != GK_OK) return r;
            if (drained) break;
        }
        const int cur = (int)(wave & 1u), nxt = cur ^ 1;
        const uint32_t sizeE = wave == 0 ? n : bound, sizeS = wave == 0 ? 0u : bound;
        if (wave > 0 && bound <= std::min(c.streamTailPaths, n / std::max(1u, c.tailDivisor)) && c.captureWave < 0 && !c.travStats) {
            // a handful of long paths is left (dielectric primaries run to MaxNumberOfBounces): one launch walks each of them to
            // its end (trace -> shade -> trace ..., one path per lane) instead of a dozen waves of a few rays at ~65 us each
            const size_t ta = mark();
            GK_CUDA(cudaMemsetAsync(c.dTailCounters, 0, 3 * sizeof(unsigned long long), st));
            if (c.tailCoop) k_tail_coop<<<gridFor((size_t)2 * bound * 8, 256), 256, 0, st>>>(c.ubo, P, V, SS, c.paths, PL, c.extendQ[cur], c.shadowQ[cur], c.dTailCounters);
            else k_tail<<<gridFor((size_t)2 * bound, 128), 128, 0, st>>>(c.dUbo, P, V, SS, c.paths, PL, c.extendQ[cur], 0, c.shadowQ[cur], 0, c.dTailCounters, 1);
            fs.launches++;
            const size_t tb = mark();
            tailSpan[0] = ta, tailSpan[1] = tb;
            GK_CUDA(cudaMemcpyAsync((void*)(c.hCounts + 2), c.dTailCounters, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            fs.tailPaths = bound;
            tailRan = true;
            break;
        }
        const size_t a = mark();
        if ((int)wave == c.captureWave && look == 0) {
            const uint32_t countE = wave == 0 ? n : c.hWave[4 * (size_t)(wave - 1)];
            if (countE) {
                GK_CUDA(c.dCapture.reserve(2 * (size_t)countE));
                GK_CUDA(cudaMemcpy2DAsync(c.dCapture.p, 32, c.extendQ[cur].o_tmin, 16, 16, countE, cudaMemcpyDeviceToDevice, st));
                GK_CUDA(cudaMemcpy2DAsync(c.dCapture.p + 1, 32, c.extendQ[cur].d_tmax, 16, 16, countE, cudaMemcpyDeviceToDevice, st));
            }
            c.capturedCount = countE;
        }
        // extend and shadow rays of a wave are independent: two streams, so that one kernel's blocks fill the SMs the other's tail leaves idle
        const bool fork = sizeS && c.concurrentShadow;
        // waves below the threshold: the eight-lanes-per-ray kernel (a lone ray finishes ~4x sooner than on one lane); `bound` is
        // the size of an earlier wave, so a wave is only classed small when it certainly is
        const bool small = wave > 0 && bound < std::min(c.coopThreshold, n / std::max(1u, c.coopDivisor)); // relative too: a rank of an 8-GPU frame has 0.26 M paths in all
        // waves too small to keep the persistent warps of the scheduled kernel fed (a rank's share of a multi-GPU frame, Cornell):
        // its refill/vote machinery then costs more than the divergence it removes, and the plain lane kernel is faster
        const bool mid = !small && wave > 0 && bound < c.schedMinRays && !c.travStats;
        auto traceE = [&](cudaStream_t on) {
            if (small) launchCoopBounded<false>(c, V, QueueIO{c.extendQ[cur]}, sizeE, on, c.extendQ[cur].count);
            else if (mid) launchLaneBounded<false>(c, V, QueueIO{c.extendQ[cur]}, sizeE, on, c.extendQ[cur].count);
            else launchSched<false>(c, V, QueueIO{c.extendQ[cur]}, sizeE, on, c.extendQ[cur].count);
        };
        auto traceS = [&](cudaStream_t on) {
            if (small) launchCoopBounded<true>(c, V, ShadowIO{c.shadowQ[cur]}, sizeS, on, c.shadowQ[cur].count);
            else if (mid) launchLaneBounded<true>(c, V, ShadowIO{c.shadowQ[cur]}, sizeS, on, c.shadowQ[cur].count);
            else launchSched<true>(c, V, ShadowIO{c.shadowQ[cur]}, sizeS, on, c.shadowQ[cur].count);
        };
        size_t b, d, s0 = 0, s1 = 0;
        if (fork) {
            GK_CUDA(cudaEventRecord(c.evFork, st));
            GK_CUDA(cudaStreamWaitEvent(c.stream2, c.evFork, 0));
            s0 = markOn(c.stream2);
            traceS(c.stream2);
            s1 = markOn(c.stream2);
            GK_CUDA(cudaEventRecord(c.evJoin, c.stream2));
            traceE(st);
            b = mark();
            GK_CUDA(cudaStreamWaitEvent(st, c.evJoin, 0));
            d = mark();
            fs.launches += 2;
        } else {
            // camera rays are coherent: the while-while lane kernel needs ~30 % fewer instructions on them (3.1 vs 2.4 Grays/s on C2)
            if (!small && wave == 0 && c.primaryLaneKernel && !c.travStats) launchMapped<false, false>(st, gridFor(sizeE, c.laneBlock), c.laneBlock, V, QueueIO{c.extendQ[cur]}, sizeE, nullptr);
            else traceE(st);
            fs.launches++;
            b = mark();
            if (sizeS) {
                traceS(st);
                fs.launches++;
            }
            d = mark();
        }
        GK_CUDA(cudaMemsetAsync(c.extendQ[nxt].count, 0, sizeof(uint32_t), st));
        GK_CUDA(cudaMemsetAsync(c.shadowQ[nxt].count, 0, sizeof(uint32_t), st));
        const unsigned shadeGrid = std::max(1u, std::min((unsigned)(c.smCount * shadeBlocksPerSm * 2), (unsigned)gridFor((size_t)sizeE + sizeS)));
        if (shadeBlocksPerSm == 4) k_shade_stream<4><<<shadeGrid, 256, 0, st>>>(c.ubo, P, SS, c.paths, PL, c.extendQ[cur], c.shadowQ[cur], c.extendQ[nxt], c.shadowQ[nxt]);
        else if (shadeBlocksPerSm == 3) k_shade_stream<3><<<shadeGrid, 256, 0, st>>>(c.ubo, P, SS, c.paths, PL, c.extendQ[cur], c.shadowQ[cur], c.extendQ[nxt], c.shadowQ[nxt]);
        else k_shade_stream<2><<<shadeGrid, 256, 0, st>>>(c.ubo, P, SS, c.paths, PL, c.extendQ[cur], c.shadowQ[cur], c.extendQ[nxt], c.shadowQ[nxt]);
        const size_t f = mark();
        k_wave_end<<<1, 1, 0, st>>>(c.extendQ[nxt].count, c.shadowQ[nxt].count, c.dWave + 4 * (size_t)wave, tag);
        fs.launches += 2;
        spans.push_back({a, b, 1}), spans.push_back({d, f, 3}), spans.push_back({a, d, 6});
        if (fork) spans.push_back({s0, s1, 2});
        else spans.push_back({b, d, 2});
        enqueued = wave + 1;
    }
    {
        const GkStatus r = consume(enqueued);
        if (r != GK_OK) return r;
    }
    if (!drained && !tailRan) {
        GK_CUDA(cudaStreamSynchronize(st));
        setLastError("gk_trace_frame: paths were still alive after the wave limit (" + std::to_string(Context::kWaveLimit) + "); lower NumberOfSamples / bounces");
        return GK_ERR_UNSUPPORTED;
    }
    const size_t g = mark();
    k_accumulate<<<gridFor(n), 256, 0, st>>>(P, c.paths, PL);
    fs.launches++;
    const size_t hEnd = mark();
    spans.push_back({g, hEnd, 4});
    GK_CUDA(cudaGetLastError());
    GK_CUDA(cudaMemcpyAsync(c.hCounts + 12, c.dOverflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, st)); // traversal stack overflow flag
    GK_CUDA(cudaStreamSynchronize(st));
    for (const Span& sp : spans) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c.evPool[sp.a], c.evPool[sp.b]);
        if (sp.kind == 0) fs.msGenerate += ms;
        else if (sp.kind == 1) fs.msExtend += ms;
        else if (sp.kind == 2) fs.msShadow += ms;
        else if (sp.kind == 3) fs.msShade += ms;
        else if (sp.kind == 6) fs.msTrace += ms;
        else fs.msAccumulate += ms;
    }
    cudaEventElapsedTime(&fs.msTotal, c.evPool[evStart], c.evPool[hEnd]);
    if (tailRan) {
        cudaEventElapsedTime(&fs.msTail, c.evPool[tailSpan[0]], c.evPool[tailSpan[1]]);
        unsigned long long t[3];
        memcpy(t, (const void*)(c.hCounts + 2), sizeof(t));
        if (t[2]) {
            setLastError("gk_trace_frame: paths were still alive when the tail kernel's step guard (65536) expired; lower NumberOfSamples / bounces");
            return GK_ERR_UNSUPPORTED;
        }
        fs.extensionRays += t[0], fs.shadowRays += t[1];
        fs.tailExtensionRays = t[0], fs.tailShadowRays = t[1];
    }
    if (getenv("GK_WAVE_LOG")) {
        if (tailRan) fprintf(stderr, "[gk tail] %u paths (bound)  %.3f ms  extension %llu shadow %llu rays\n", fs.tailPaths, fs.msTail, (unsigned long long)fs.tailExtensionRays, (unsigned long long)fs.tailShadowRays); // per-wave device times (diagnostics): trace span, shade span, queue sizes
        size_t k = 1;
        for (uint32_t w = 0; w < enqueued && k + 2 < spans.size(); ++w, k += 4) {
            float tr = 0, sh = 0;
            cudaEventElapsedTime(&tr, c.evPool[spans[k + 2].a], c.evPool[spans[k + 2].b]);
            cudaEventElapsedTime(&sh, c.evPool[spans[k + 1].a], c.evPool[spans[k + 1].b]);
            const uint32_t e = w == 0 ? n : c.hWave[4 * (size_t)(w - 1)], s2 = w == 0 ? 0u : c.hWave[4 * (size_t)(w - 1) + 1];
            fprintf(stderr, "[gk wave %2u] extend %8u shadow %8u rays  trace %.3f ms  shade %.3f ms\n", w, e, s2, tr, sh);
        }
    }
    if (c.travStats) {
        TraversalStats h;
        GK_CUDA(cudaMemcpy(&h, c.dTravStats, sizeof(h), cudaMemcpyDeviceToHost));
        fs.nodeVisits = h.nodeVisits, fs.triTests = h.triTests, fs.tlasVisits = h.tlasVisits, fs.instanceEntries = h.instanceEntries, fs.maxStack = (uint32_t)h.maxStack;
        SchedStats ss;
        GK_CUDA(cudaMemcpy(&ss, c.dSchedStats, sizeof(ss), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; ++k) fs.schedIters[k] = ss.iters[k], fs.schedLanes[k] = ss.lanes[k];
        fs.schedRefills = ss.refills, fs.schedRefillLanes = ss.refillLanes, fs.schedPopIters = ss.popIters, fs.schedPopLanes = ss.popLanes;
    }
    if (c.hCounts[12]) {
        c.hCounts[12] = 0;
        return checkTraversalOverflow(c);
    }
    return GK_OK;
}

GkStatus traceFrame(Context& c)
{
    if (!c.haveScene || !c.haveInstances || !c.haveUbo) {
        setLastError("gk_trace_frame: scene, instances and UBO must be set first");
        return GK_ERR_NOT_READY;
    }
    cudaStream_t st = c.stream;
    applyPendingHistorySwap(c);
    {
        static const GkPlane written[] = {GK_PLANE_OUTPUT_DIFFUSE, GK_PLANE_OUTPUT_SPECULAR, GK_PLANE_ALBEDO, GK_PLANE_NORMAL, GK_PLANE_OBJECT_ID0, GK_PLANE_MOTION, GK_PLANE_DEPTH,
                                          GK_PLANE_RADIANCE_DIFFUSE_F32, GK_PLANE_RADIANCE_SPECULAR_F32, GK_PLANE_PRIMARY_IDS, GK_PLANE_PRIMARY_T, GK_PLANE_RAY_COUNT};
        const void* bufs[sizeof(written) / sizeof(written[0])];
        for (size_t i = 0; i < sizeof(written) / sizeof(written[0]); ++i) bufs[i] = c.planes.p[written[i]];
        waitAsyncCopyBeforeWriting(c, bufs, (int)(sizeof(written) / sizeof(written[0])));
    }
    c.tracedSinceFilter = true;
    if (c.traceVariant == 1) return traceFrameStreamed(c);
    const uint32_t n = c.pathCount;
    FrameParams P{c.width, c.height, c.traceTileIndex, c.traceTileCount, c.tileRows, n, (c.microTiles == 2 && c.width % 16 == 0 && c.tileRows % 16 == 0) ? 2u : (c.microTiles && c.width % 8 == 0 && c.tileRows % 4 == 0) ? 1u : 0u};
    const SceneView V = c.view();
    const ShadeScene SS = shadeSceneOf(c);
    const PlaneView PL = planeViewOf(c);
    GkFrameStats& fs = c.stats;
    fs.primaryRays = fs.extensionRays = fs.shadowRays = 0;
    fs.waves = fs.launches = 0;
    fs.msGenerate = fs.msExtend = fs.msShade = fs.msShadow = fs.msAccumulate = fs.msTail = fs.msTrace = 0;
    size_t ev = 0;
    struct Span { size_t a, b; int kind; };
    std::vector<Span> spans;
    auto mark = [&]() { cudaEvent_t e = poolEvent(c, ev); cudaEventRecord(e, st); return ev++; };
    auto markOn = [&](cudaStream_t s) { cudaEvent_t e = poolEvent(c, ev); cudaEventRecord(e, s); return ev++; };
    if (c.concurrentShadow && !c.stream2) {
        GK_CUDA(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
        GK_CUDA(cudaEventCreateWithFlags(&c.evFork, cudaEventDisableTiming));
        GK_CUDA(cudaEventCreateWithFlags(&c.evJoin, cudaEventDisableTiming));
    }

    GK_CUDA(cudaMemcpyAsync(c.dUbo, &c.ubo, sizeof(GkUniformBufferObject), cudaMemcpyHostToDevice, st));
    if (c.travStats) {
        GK_CUDA(cudaMemsetAsync(c.dTravStats, 0, sizeof(TraversalStats), st));
        GK_CUDA(cudaMemsetAsync(c.dSchedStats, 0, sizeof(SchedStats), st));
    }
    if (c.cursorNext) { // queue cursors of the previous frame's launches
        GK_CUDA(cudaMemsetAsync(c.dCursors, 0, sizeof(uint32_t) * c.cursorNext, st));
        c.cursorNext = 0;
    }
    const size_t evStart = mark();
    int cur = 0;
    k_generate<<<gridFor(n), 256, 0, st>>>(c.dUbo, P, c.paths, c.extendQ[cur]);
    fs.launches++;
    GK_CUDA(cudaMemsetAsync(c.shadowQ[cur].count, 0, sizeof(uint32_t), st));
    const size_t evGen = mark();
    spans.push_back({evStart, evGen, 0});
    uint32_t countE = n, countS = 0;
    bool tailRan = false;
    // the tail launch pays off once the waves are short: a quarter of the paths (swept on 1, 2 and 4 GPUs), capped per SM
    // (the scheduled kernel balances short waves by itself and is the only kernel that reports stack overflow: no tail launch with it)
    const uint32_t tailLimit = c.traceVariant == 1 ? 0u : std::min(c.tailThreshold, (uint32_t)(c.tailFraction * (float)n));
    fs.tailPaths = 0;
    fs.tailExtensionRays = fs.tailShadowRays = 0;
    fs.primaryRays = (uint64_t)c.ownedRows * c.width;
    c.capturedCount = 0;
    for (uint32_t wave = 0; wave < 4096; ++wave) {
        if (countE == 0 && countS == 0) break;
        if (wave > 0 && countE + countS <= tailLimit && !c.travStats && c.captureWave < 0) {
            // few paths left: finish them in one launch
            const size_t a = mark();
            GK_CUDA(cudaMemsetAsync(c.dTailCounters, 0, 3 * sizeof(unsigned long long), st));
            k_tail<<<gridFor((size_t)countE + countS, 128), 128, 0, st>>>(c.dUbo, P, V, SS, c.paths, PL, c.extendQ[cur], countE, c.shadowQ[cur], countS, c.dTailCounters, 0);
            fs.launches++;
            const size_t b = mark();
            spans.push_back({a, b, 5});
            GK_CUDA(cudaMemcpyAsync(c.hCounts + 2, c.dTailCounters, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            fs.tailPaths = countE + countS;
            tailRan = true;
            fs.waves++;
            break;
        }
        const size_t a = mark();
        if (countE) {
            if ((int)wave == c.captureWave) {
                GK_CUDA(c.dCapture.reserve(2 * (size_t)countE));
                // interleave origin/direction records for the CPU baseline
                GK_CUDA(cudaMemcpy2DAsync(c.dCapture.p, 32, c.extendQ[cur].o_tmin, 16, 16, countE, cudaMemcpyDeviceToDevice, st));
                GK_CUDA(cudaMemcpy2DAsync(c.dCapture.p + 1, 32, c.extendQ[cur].d_tmax, 16, 16, countE, cudaMemcpyDeviceToDevice, st));
                c.capturedCount = countE;
            }
        }
        // extend and shadow rays of a wave are independent: the shadow kernel runs on a second stream so
        // that its blocks fill the SMs the long tail of the extend kernel leaves idle (and vice versa)
        const bool fork = countE && countS && c.concurrentShadow;
        size_t b, d, s0 = 0, s1 = 0;
        if (fork) {
            GK_CUDA(cudaEventRecord(c.evFork, st));
            GK_CUDA(cudaStreamWaitEvent(c.stream2, c.evFork, 0));
            s0 = markOn(c.stream2);
            launchTrace<true>(c, V, c.shadowQ[cur], countS, c.stream2);
            s1 = markOn(c.stream2);
            GK_CUDA(cudaEventRecord(c.evJoin, c.stream2));
            launchTrace<false>(c, V, c.extendQ[cur], countE);
            b = mark();
            GK_CUDA(cudaStreamWaitEvent(st, c.evJoin, 0));
            d = mark();
            fs.launches += 2;
        } else {
            if (countE) {
                launchTrace<false>(c, V, c.extendQ[cur], countE);
                fs.launches++;
            }
            b = mark();
            if (countS) {
                launchTrace<true>(c, V, c.shadowQ[cur], countS);
                fs.launches++;
            }
            d = mark();
        }
        const int nxt = cur ^ 1;
        GK_CUDA(cudaMemsetAsync(c.extendQ[nxt].count, 0, sizeof(uint32_t), st));
        GK_CUDA(cudaMemsetAsync(c.shadowQ[nxt].count, 0, sizeof(uint32_t), st));
        if (c.shadeMinBlocks >= 4)
            k_shade<4><<<gridFor((size_t)countE + countS), 256, 0, st>>>(c.dUbo, P, SS, c.paths, PL, c.extendQ[cur], countE, c.shadowQ[cur], countS, c.extendQ[nxt],
                                                                     c.shadowQ[nxt]);
        else if (c.shadeMinBlocks >= 3)
            k_shade<3><<<gridFor((size_t)countE + countS), 256, 0, st>>>(c.dUbo, P, SS, c.paths, PL, c.extendQ[cur], countE, c.shadowQ[cur], countS, c.extendQ[nxt],
                                                                     c.shadowQ[nxt]);
        else
            k_shade<2><<<gridFor((size_t)countE + countS), 256, 0, st>>>(c.dUbo, P, SS, c.paths, PL, c.extendQ[cur], countE, c.shadowQ[cur], countS, c.extendQ[nxt],
                                                                     c.shadowQ[nxt]);
        fs.launches++;
        const size_t f = mark();
        spans.push_back({a, b, 1}), spans.push_back({d, f, 3}), spans.push_back({a, d, 6});
        if (fork) spans.push_back({s0, s1, 2}); // overlaps the extend span
        else spans.push_back({b, d, 2});
        GK_CUDA(cudaMemcpyAsync(c.hCounts, c.extendQ[nxt].count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaMemcpyAsync(c.hCounts + 1, c.shadowQ[nxt].count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
        if (wave > 0) fs.extensionRays += countE;
        fs.shadowRays += countS;
        countE = c.hCounts[0], countS = c.hCounts[1];
        fs.waves++;
        cur = nxt;
    }
    if (!tailRan && (countE || countS)) {
        // the wave limit was hit with paths still alive (NumberOfSamples x bounces far beyond what a frame is): never
        // accumulate unfinished paths
        GK_CUDA(cudaStreamSynchronize(st));
        setLastError("gk_trace_frame: " + std::to_string(countE + countS) + " paths were still alive after the wave limit (4096); lower NumberOfSamples / bounces");
        return GK_ERR_UNSUPPORTED;
    }
    const size_t g = mark();
    k_accumulate<<<gridFor(n), 256, 0, st>>>(P, c.paths, PL);
    fs.launches++;
    const size_t hEnd = mark();
    spans.push_back({g, hEnd, 4});
    GK_CUDA(cudaGetLastError());
    GK_CUDA(cudaMemcpyAsync(c.hCounts + 12, c.dOverflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, st)); // traversal stack overflow flag
    GK_CUDA(cudaStreamSynchronize(st));
    for (const Span& s : spans) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c.evPool[s.a], c.evPool[s.b]);
        if (s.kind == 0) fs.msGenerate += ms;
        else if (s.kind == 1) fs.msExtend += ms;
        else if (s.kind == 2) fs.msShadow += ms;
        else if (s.kind == 3) fs.msShade += ms;
        else if (s.kind == 5) fs.msTail += ms;
        else if (s.kind == 6) fs.msTrace += ms;
        else fs.msAccumulate += ms;
    }
    cudaEventElapsedTime(&fs.msTotal, c.evPool[evStart], c.evPool[hEnd]);
    if (tailRan) {
        unsigned long long t[3];
        memcpy(t, c.hCounts + 2, sizeof(t));
        if (t[2]) {
            setLastError("gk_trace_frame: paths were still alive when the tail kernel's step guard (65536) expired; lower NumberOfSamples / bounces");
            return GK_ERR_UNSUPPORTED;
        }
        fs.extensionRays += t[0], fs.shadowRays += t[1];
        fs.tailExtensionRays = t[0], fs.tailShadowRays = t[1];
    }
    if (c.travStats) {
        TraversalStats h;
        GK_CUDA(cudaMemcpy(&h, c.dTravStats, sizeof(h), cudaMemcpyDeviceToHost));
        fs.nodeVisits = h.nodeVisits, fs.triTests = h.triTests, fs.tlasVisits = h.tlasVisits, fs.instanceEntries = h.instanceEntries, fs.maxStack = (uint32_t)h.maxStack;
        SchedStats ss;
        GK_CUDA(cudaMemcpy(&ss, c.dSchedStats, sizeof(ss), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; ++k) fs.schedIters[k] = ss.iters[k], fs.schedLanes[k] = ss.lanes[k];
        fs.schedRefills = ss.refills, fs.schedRefillLanes = ss.refillLanes, fs.schedPopIters = ss.popIters, fs.schedPopLanes = ss.popLanes;
    }
    if (c.hCounts[12]) {
        c.hCounts[12] = 0;
        return checkTraversalOverflow(c);
    }
    return GK_OK;
}

// A traversal that ran out of stack (GK_TRAVERSAL_STACK entries) dropped an entry and may have missed a hit: that is an
// error of the call, not a statistic.  The flag costs one 4-byte read-back per frame / intersect call.
GkStatus checkTraversalOverflow(Context& c)
{
    uint32_t flag = 0;
    GK_CUDA(cudaMemcpyAsync(&flag, c.dOverflow, sizeof(flag), cudaMemcpyDeviceToHost, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    if (flag) {
        GK_CUDA(cudaMemsetAsync(c.dOverflow, 0, sizeof(uint32_t), c.stream));
        setLastError("traversal stack overflow: the scene needs more than GK_TRAVERSAL_STACK (" + std::to_string(kStackSize) + ") entries for some ray; hits may be missing");
        return GK_ERR_UNSUPPORTED;
    }
    return GK_OK;
}

GkStatus intersectDevice(Context& c, const float4* rays, uint32_t n, float* tuv, uint32_t* ids, bool anyHit)
{
    if (!c.haveScene || !c.haveInstances) {
        setLastError("gk_intersect: scene and instances must be set first");
        return GK_ERR_NOT_READY;
    }
    if (n == 0) return GK_OK;
    const SceneView V = c.view();
    if (anyHit) launchTraceIO<true>(c, V, ArrayIO<true>{rays, tuv, ids}, n);
    else launchTraceIO<false>(c, V, ArrayIO<false>{rays, tuv, ids}, n);
    GK_CUDA(cudaGetLastError());
    return GK_OK;
}

} // namespace gk
