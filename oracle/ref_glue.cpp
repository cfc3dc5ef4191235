// ref_glue.cpp — thin C shim over the REAL tinybvh that the reference vendors.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; the only reference code it pulls
// in is src/ThirdParty/tinybvh/tiny_bvh.h, included from where it lies under
// /root/reference (see oracle/Makefile: -I$(REF)/src/ThirdParty/tinybvh).  The
// build product goes to oracle/_ref/ (git-ignored) and travels to the GPU box as
// a prebuilt .so.  It drives tinybvh exactly the way
// FCPUAccelerationStructure does (src/Assets/CPUAccelerationStructure.cpp:171-307):
// per-model BLAS over de-indexed fp32 triangles via BVH::Build(verts, n), a TLAS via
// BVH::Build(BLASInstance*, n, BVHBase**, m) with transposed world matrices, and
// queries through tinybvh::Ray + BVH::Intersect.
//
// Used for (1) pinning oracle/orc_bvh.cpp (bit-exact t,u,v,prim,inst) and
// (2) the "reference" CPU baseline in bench.py.
#define TINYBVH_IMPLEMENTATION
#include "tiny_bvh.h"

#include "../include/gknext_types.h"
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct RefScene {
    std::vector<std::vector<tinybvh::bvhvec4>> tris;
    std::vector<tinybvh::BVH*> blas;
    std::vector<tinybvh::BVHBase*> blasBase;
    std::vector<tinybvh::BLASInstance> inst;
    std::vector<uint32_t> instToNode;
    tinybvh::BVH tlas;
    bool hasTlas = false;
    ~RefScene()
    {
        for (auto* b : blas) delete b;
    }
};

} // namespace

extern "C" {

void* ref_scene_create(const GkSceneDesc* d, const GkNodeProxy* nodes, uint32_t nodeCount)
{
    RefScene* S = new RefScene();
    S->tris.resize(d->modelCount);
    for (uint32_t m = 0; m < d->modelCount; ++m) {
        const GkModelDesc& md = d->models[m];
        auto& T = S->tris[m];
        for (uint32_t i = 0; i + 2 < md.indexCount; i += 3)
            for (int k = 0; k < 3; ++k) {
                const GkVertex& v = md.vertices[md.indices[i + k]];
                T.push_back(tinybvh::bvhvec4(v.Position[0], v.Position[1], v.Position[2], 0));
            }
        tinybvh::BVH* b = new tinybvh::BVH();
        b->Build(T.data(), (uint32_t)(T.size() / 3));
        S->blas.push_back(b);
        S->blasBase.push_back(b);
    }
    for (uint32_t i = 0; i < nodeCount; ++i) {
        const GkNodeProxy& p = nodes[i];
        if (!p.visible || p.nort) continue;
        tinybvh::BLASInstance in;
        in.blasIdx = p.modelId / 10;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) in.transform[r * 4 + c] = p.worldTS[c * 4 + r];
        S->inst.push_back(in);
        S->instToNode.push_back(i);
    }
    if (!S->inst.empty()) {
        S->tlas.Build(S->inst.data(), (uint32_t)S->inst.size(), S->blasBase.data(), (uint32_t)S->blasBase.size());
        S->hasTlas = true;
    }
    return S;
}

void ref_scene_destroy(void* h) { delete (RefScene*)h; }

// rays: 8 floats each {Ox,Oy,Oz,tmin(ignored: tinybvh accepts t>0),Dx,Dy,Dz,tmax}
// out_tuv: 3 floats per ray, out_ids: {prim, node index} per ray (0xffffffff on miss).
// Returns seconds spent inside the traversal loop (steady_clock, all threads joined).
double ref_intersect(void* h, const float* rays, uint32_t n, float* out_tuv, uint32_t* out_ids, int threads)
{
    RefScene* S = (RefScene*)h;
    if (threads < 1) threads = 1;
    auto work = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) {
            const float* r = rays + 8 * (size_t)i;
            tinybvh::Ray ray(tinybvh::bvhvec3(r[0], r[1], r[2]), tinybvh::bvhvec3(r[4], r[5], r[6]), r[7]);
            if (S->hasTlas) S->tlas.Intersect(ray);
            if (ray.hit.t < r[7]) {
                if (out_tuv) out_tuv[3 * (size_t)i] = ray.hit.t, out_tuv[3 * (size_t)i + 1] = ray.hit.u, out_tuv[3 * (size_t)i + 2] = ray.hit.v;
                if (out_ids) out_ids[2 * (size_t)i] = ray.hit.prim, out_ids[2 * (size_t)i + 1] = S->instToNode[ray.hit.inst];
            } else {
                if (out_tuv) out_tuv[3 * (size_t)i] = r[7], out_tuv[3 * (size_t)i + 1] = 0, out_tuv[3 * (size_t)i + 2] = 0;
                if (out_ids) out_ids[2 * (size_t)i] = 0xffffffffu, out_ids[2 * (size_t)i + 1] = 0xffffffffu;
            }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) work(0, n);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) {
            uint32_t lo = (uint32_t)((uint64_t)n * t / threads), hi = (uint32_t)((uint64_t)n * (t + 1) / threads);
            pool.emplace_back(work, lo, hi);
        }
        for (auto& th : pool) th.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Dump of the built structures, so the restatement can be compared node by node.
uint32_t ref_blas_node_count(void* h, uint32_t m) { return ((RefScene*)h)->blas[m]->usedNodes; }
void ref_blas_nodes(void* h, uint32_t m, float* out8 /* usedNodes * 8 words */)
{
    tinybvh::BVH* b = ((RefScene*)h)->blas[m];
    memcpy(out8, b->bvhNode, (size_t)b->usedNodes * 32);
}
uint32_t ref_tlas_node_count(void* h) { return ((RefScene*)h)->hasTlas ? ((RefScene*)h)->tlas.usedNodes : 0; }
void ref_tlas_nodes(void* h, float* out8)
{
    RefScene* S = (RefScene*)h;
    if (S->hasTlas) memcpy(out8, S->tlas.bvhNode, (size_t)S->tlas.usedNodes * 32);
}

const char* ref_version() { return "tinybvh " "1.3.8" " (vendored by gkNextRenderer, src/ThirdParty/tinybvh/tiny_bvh.h)"; }

} // extern "C"
