// orc_scene.h — scene ingestion for the CPU oracle.  TEST INFRASTRUCTURE ONLY.
//
// Restates the glue of FCPUAccelerationStructure::InitBVH / UpdateBVH
// (src/Assets/CPUAccelerationStructure.cpp:171-281) and the GPU-buffer side of
// Scene::RebuildMeshBuffer (src/Assets/Scene.cpp:101-196) on top of orc_bvh.
#pragma once
#include "../include/gknext_types.h"
#include "orc_bvh.h"
#include <vector>

namespace orc {

struct TriExt { // FCPUBLASVertInfo: face normal + material slot of vertex 0
    f3 normal;
    uint32_t matIdx;
};

struct Model {
    std::vector<GkGPUVertex> gpuVerts; // MakeVertex, Vertex.hpp:80-99
    std::vector<uint32_t> indices;     // original order (== PrimAddress triangle order)
    std::vector<TriExt> ext;
};

struct Scene {
    std::vector<Model> models;
    std::vector<Blas> blas;
    std::vector<GkMaterial> materials;
    std::vector<GkLightObject> lights;
    std::vector<GkNodeProxy> nodes;   // as uploaded
    std::vector<uint32_t> instToNode; // TLAS instance -> index into nodes
    Tlas tlas;

    void load(const GkSceneDesc& d);
    void setNodes(const GkNodeProxy* n, uint32_t count);

    // closest hit in world space; hit.inst is an index into `nodes`
    bool trace(f3 O, f3 D, float tmin, float tmax, Hit& out, uint64_t* nv = nullptr, uint64_t* nt = nullptr) const;
    bool anyHit(f3 O, f3 D, float tmin, float tmax) const;
    GkRayCastResult rayCastInCPU(f3 O, f3 D) const; // CPUAccelerationStructure.cpp:283-307
};

GkGPUVertex makeGpuVertex(const GkVertex& v);

} // namespace orc
