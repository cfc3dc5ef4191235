// orc_scene.cpp — see orc_scene.h.  TEST INFRASTRUCTURE ONLY.
#include "orc_scene.h"

namespace orc {

GkGPUVertex makeGpuVertex(const GkVertex& v) // src/Assets/Vertex.hpp:80-99
{
    GkGPUVertex g;
    g.posx = glm_to_half(v.Position[0]);
    g.posy = glm_to_half(v.Position[1]);
    g.posz = glm_to_half(v.Position[2]);
    g.texcoordx = glm_to_half(v.TexCoord[0]);
    g.normalx = glm_to_half(v.Normal[0]);
    g.normaly = glm_to_half(v.Normal[1]);
    g.normalz = glm_to_half(v.Normal[2]);
    g.texcoordy = glm_to_half(v.TexCoord[1]);
    g.tangentx = glm_to_half(v.Tangent[0]);
    g.tangenty = glm_to_half(v.Tangent[1]);
    g.tangentz = glm_to_half(v.Tangent[2]);
    g.tangentw = (uint16_t)(((v.Tangent[3] > 0 ? 2 : 0) << 8) | (uint16_t)v.MaterialIndex);
    return g;
}

void Scene::load(const GkSceneDesc& d)
{
    models.clear();
    blas.clear();
    models.resize(d.modelCount);
    blas.resize(d.modelCount);
    for (uint32_t m = 0; m < d.modelCount; ++m) {
        const GkModelDesc& md = d.models[m];
        Model& M = models[m];
        M.gpuVerts.resize(md.vertexCount);
        for (uint32_t i = 0; i < md.vertexCount; ++i) M.gpuVerts[i] = makeGpuVertex(md.vertices[i]);
        M.indices.assign(md.indices, md.indices + md.indexCount);
        Blas& B = blas[m];
        for (uint32_t i = 0; i + 2 < md.indexCount; i += 3) { // CPUAccelerationStructure.cpp:185-204
            const GkVertex& a = md.vertices[md.indices[i]];
            const GkVertex& b = md.vertices[md.indices[i + 1]];
            const GkVertex& c = md.vertices[md.indices[i + 2]];
            const f3 pa(a.Position[0], a.Position[1], a.Position[2]);
            const f3 pb(b.Position[0], b.Position[1], b.Position[2]);
            const f3 pc(c.Position[0], c.Position[1], c.Position[2]);
            const f3 e1 = pb - pa, e2 = pc - pb;
            const f3 cr = cross(e1, e2);
            const float inv = 1.0f / sqrtf(dot(cr, cr)); // glm::normalize = v * inversesqrt(dot(v,v))
            M.ext.push_back({cr * inv, a.MaterialIndex});
            B.tri.push_back(f4(pa, 0));
            B.tri.push_back(f4(pb, 0));
            B.tri.push_back(f4(pc, 0));
        }
        B.build();
    }
    materials.assign(d.materials, d.materials + d.materialCount);
    lights.assign(d.lights, d.lights + d.lightCount);
}

void Scene::setNodes(const GkNodeProxy* n, uint32_t count) // CPUAccelerationStructure.cpp:236-281
{
    nodes.assign(n, n + count);
    tlas.inst.clear();
    instToNode.clear();
    for (uint32_t i = 0; i < count; ++i) {
        const GkNodeProxy& p = nodes[i];
        // ray-visible instances: RayTraceBaseRenderer.cpp:188-189 (mask = visible && !nort)
        if (!p.visible || p.nort) continue;
        Instance in;
        in.blas = p.modelId / 10;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) in.T[r * 4 + c] = p.worldTS[c * 4 + r];
        tlas.inst.push_back(in);
        instToNode.push_back(i);
    }
    if (!tlas.inst.empty()) tlas.build(blas);
}

bool Scene::trace(f3 O, f3 D, float tmin, float tmax, Hit& out, uint64_t* nv, uint64_t* nt) const
{
    if (tlas.inst.empty()) return false;
    RayQ r = makeRay(O, D, tmin, tmax);
    tlas.intersect(blas, r, nv, nt);
    if (!(r.hit.t < tmax)) return false;
    out = r.hit;
    out.inst = instToNode[r.hit.inst];
    return true;
}

bool Scene::anyHit(f3 O, f3 D, float tmin, float tmax) const
{
    if (tlas.inst.empty()) return false;
    RayQ r = makeRay(O, D, tmin, tmax);
    return tlas.occluded(blas, r);
}

GkRayCastResult Scene::rayCastInCPU(f3 O, f3 D) const
{
    GkRayCastResult R;
    memset(&R, 0, sizeof(R));
    Hit h;
    if (!trace(O, D, 0.0f, 2000.0f, h)) return R;
    const GkNodeProxy& p = nodes[h.inst];
    const f3 n = models[p.modelId / 10].ext[h.prim].normal;
    const f3 hp = O + D * h.t;
    R.HitPoint[0] = hp.x, R.HitPoint[1] = hp.y, R.HitPoint[2] = hp.z, R.HitPoint[3] = 0;
    // vec4(n,0) * transposedWorld == world * n, glm row-vector product = per-column dot,
    // each dot summed as (x+y)+(z+w)  (CPUAccelerationStructure.cpp:297-298)
    const float* W = p.worldTS; // column-major
    R.Normal[0] = (W[0] * n.x + W[4] * n.y) + (W[8] * n.z + W[12] * 0.0f);
    R.Normal[1] = (W[1] * n.x + W[5] * n.y) + (W[9] * n.z + W[13] * 0.0f);
    R.Normal[2] = (W[2] * n.x + W[6] * n.y) + (W[10] * n.z + W[14] * 0.0f);
    R.Normal[3] = (W[3] * n.x + W[7] * n.y) + (W[11] * n.z + W[15] * 0.0f);
    R.T = h.t;
    R.InstanceId = p.instanceId;
    R.MaterialId = 0;
    R.Hitted = 1;
    return R;
}

} // namespace orc
