// orc_api.cpp — C entry points of the CPU oracle (ctypes-friendly).
// TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py.  Never by the product.
#include "orc_filters.h"
#include "orc_pt.h"
#include "orc_scene.h"
#include <chrono>
#include <thread>
#include <vector>

using namespace orc;

extern "C" {

void* orc_scene_create(const GkSceneDesc* d, const GkNodeProxy* nodes, uint32_t nodeCount)
{
    Scene* S = new Scene();
    S->load(*d);
    S->setNodes(nodes, nodeCount);
    return S;
}
void orc_scene_set_nodes(void* h, const GkNodeProxy* nodes, uint32_t nodeCount) { ((Scene*)h)->setNodes(nodes, nodeCount); }
void orc_scene_destroy(void* h) { delete (Scene*)h; }

// Same contract as ref_intersect (oracle/ref_glue.cpp) except that tmin (ray[3]) is honoured.
// stats (optional, 2 words): total node visits and triangle tests over the batch.
double orc_intersect(void* h, const float* rays, uint32_t n, float* out_tuv, uint32_t* out_ids, int threads, uint64_t* stats)
{
    const Scene* S = (const Scene*)h;
    if (threads < 1) threads = 1;
    std::vector<uint64_t> nv(threads, 0), nt(threads, 0);
    auto work = [&](int tid, uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) {
            const float* r = rays + 8 * (size_t)i;
            Hit hit;
            bool ok = S->trace(f3(r[0], r[1], r[2]), f3(r[4], r[5], r[6]), r[3], r[7], hit, stats ? &nv[tid] : nullptr, stats ? &nt[tid] : nullptr);
            if (ok) {
                if (out_tuv) out_tuv[3 * (size_t)i] = hit.t, out_tuv[3 * (size_t)i + 1] = hit.u, out_tuv[3 * (size_t)i + 2] = hit.v;
                if (out_ids) out_ids[2 * (size_t)i] = hit.prim, out_ids[2 * (size_t)i + 1] = hit.inst;
            } else {
                if (out_tuv) out_tuv[3 * (size_t)i] = r[7], out_tuv[3 * (size_t)i + 1] = 0, out_tuv[3 * (size_t)i + 2] = 0;
                if (out_ids) out_ids[2 * (size_t)i] = 0xffffffffu, out_ids[2 * (size_t)i + 1] = 0xffffffffu;
            }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) work(0, 0, n);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t)
            pool.emplace_back(work, t, (uint32_t)((uint64_t)n * t / threads), (uint32_t)((uint64_t)n * (t + 1) / threads));
        for (auto& th : pool) th.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats[0] = stats[1] = 0;
        for (int t = 0; t < threads; ++t) stats[0] += nv[t], stats[1] += nt[t];
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// Brute-force closest hit over every triangle of every ray-visible instance with the same
// per-triangle arithmetic: independent of any BVH, used to classify id mismatches
// (exact-t tie vs. a box test that dropped the true nearest triangle).
static void bruteforceRange(const Scene* S, const float* rays, uint32_t lo, uint32_t hi, float* out_tuv, uint32_t* out_ids)
{
    for (uint32_t i = lo; i < hi; ++i) {
        const float* r = rays + 8 * (size_t)i;
        RayQ q = makeRay(f3(r[0], r[1], r[2]), f3(r[4], r[5], r[6]), r[3], r[7]);
        Hit best = q.hit;
        uint32_t bestNode = 0xffffffffu;
        for (size_t k = 0; k < S->tlas.inst.size(); ++k) {
            const Instance& in = S->tlas.inst[k];
            const Blas& B = S->blas[in.blas];
            const f3 O = xformPoint(q.O, in.invT), D = xformVector(q.D, in.invT);
            for (uint32_t p = 0; p < B.tri.size() / 3; ++p) {
                const f3 v0 = B.tri[p * 3].xyz(), e1 = B.tri[p * 3 + 1].xyz() - v0, e2 = B.tri[p * 3 + 2].xyz() - v0;
                const f3 hh = cross(D, e2);
                const float a = dot(e1, hh);
                if (fabsf(a) < 0.0000001f) continue;
                const float f = 1 / a;
                const f3 s = O - v0;
                const float u = f * dot(s, hh);
                if (u < 0 || u > 1) continue;
                const f3 qq = cross(s, e1);
                const float v = f * dot(D, qq);
                if (v < 0 || u + v > 1) continue;
                const float t = f * dot(e2, qq);
                if (t > q.tmin && t < best.t) best.t = t, best.u = u, best.v = v, best.prim = p, bestNode = S->instToNode[k];
            }
        }
        out_tuv[3 * (size_t)i] = best.t, out_tuv[3 * (size_t)i + 1] = best.u, out_tuv[3 * (size_t)i + 2] = best.v;
        out_ids[2 * (size_t)i] = bestNode == 0xffffffffu ? 0xffffffffu : best.prim, out_ids[2 * (size_t)i + 1] = bestNode;
    }
}

void orc_intersect_bruteforce_mt(void* h, const float* rays, uint32_t n, float* out_tuv, uint32_t* out_ids, int threads)
{
    const Scene* S = (const Scene*)h;
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > n) threads = (int)(n ? n : 1);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        const uint32_t lo = (uint32_t)((uint64_t)n * t / threads), hi = (uint32_t)((uint64_t)n * (t + 1) / threads);
        pool.emplace_back(bruteforceRange, S, rays, lo, hi, out_tuv, out_ids);
    }
    for (auto& th : pool) th.join();
}

void orc_intersect_bruteforce(void* h, const float* rays, uint32_t n, float* out_tuv, uint32_t* out_ids)
{
    bruteforceRange((const Scene*)h, rays, 0, n, out_tuv, out_ids);
}

void orc_raycast(void* h, const float* origin_dir /* 6 floats per ray */, uint32_t n, GkRayCastResult* out)
{
    const Scene* S = (const Scene*)h;
    for (uint32_t i = 0; i < n; ++i) {
        const float* r = origin_dir + 6 * (size_t)i;
        out[i] = S->rayCastInCPU(f3(r[0], r[1], r[2]), f3(r[3], r[4], r[5]));
    }
}

void orc_bake_probes(void* h, const GkUniformBufferObject* ubo, GkAmbientCube* cubes, GkVoxelData* voxels, uint32_t first, uint32_t count, int threads)
{
    bakeProbes(*(const Scene*)h, *ubo, cubes, voxels, first, count, threads);
}

void orc_raycast_task(void* h, GkRayCastIO* io, uint32_t n) { rayCastTask(*(const Scene*)h, io, n); }

uint32_t orc_blas_node_count(void* h, uint32_t m) { return (uint32_t)((Scene*)h)->blas[m].bvh.nodes.size(); }
void orc_blas_nodes(void* h, uint32_t m, float* out8) { memcpy(out8, ((Scene*)h)->blas[m].bvh.nodes.data(), ((Scene*)h)->blas[m].bvh.nodes.size() * 32); }
uint32_t orc_tlas_node_count(void* h) { return (uint32_t)((Scene*)h)->tlas.bvh.nodes.size(); }
void orc_tlas_nodes(void* h, float* out8) { memcpy(out8, ((Scene*)h)->tlas.bvh.nodes.data(), ((Scene*)h)->tlas.bvh.nodes.size() * 32); }

// ---- path tracer (orc_pt.cpp) ----
// Renders one frame of Core.PathTracing over the full image.  Planes are fp32, un-quantised:
//   diffuse/spec/albedo/normal: 4 floats per pixel, motion: 2, depth: 1, objectId: u32,
//   primIds: {prim, node index} per pixel (0xffffffff on miss), rayCount: rays traced per pixel.
void orc_render(void* h, const GkUniformBufferObject* ubo, uint32_t width, uint32_t height, const GkAmbientCube* cubes,
                const GkVoxelData* voxels, float* diffuse, float* spec, float* albedo, float* normal, float* motion, float* depth,
                uint32_t* objectId, uint32_t* primIds, uint32_t* rayCount, int threads)
{
    PtOutputs o{diffuse, spec, albedo, normal, motion, depth, objectId, primIds, rayCount};
    renderFrame(*(const Scene*)h, *ubo, width, height, cubes, voxels, o, threads);
}

// ---- filters (orc_filters.cpp) ----
void orc_reproject(const GkUniformBufferObject* ubo, uint32_t width, uint32_t height, int needClamp, int needSpatio, const uint16_t* src,
                   const uint16_t* history, const float* motion, const uint32_t* objId0, const uint32_t* objId1, const uint16_t* normal,
                   uint16_t* out)
{
    reproject(*ubo, width, height, needClamp != 0, needSpatio != 0, src, history, motion, objId0, objId1, normal, out);
}

void orc_denoise_jbf(const GkUniformBufferObject* ubo, uint32_t width, uint32_t height, const uint16_t* diffuse, const uint16_t* spec,
                     const uint16_t* normal, const uint32_t* objId0, const uint32_t* objId1, const uint16_t* albedo, uint16_t* out)
{
    denoiseJBF(*ubo, width, height, diffuse, spec, normal, objId0, objId1, albedo, out);
}

uint16_t orc_float_to_half(float f) { return float_to_half_rne(f); }
float orc_half_to_float(uint16_t h) { return half_to_float(h); }
uint16_t orc_glm_to_half(float f) { return glm_to_half(f); }

} // extern "C"
