// gk_scene.cu — scene upload to device buffers.
//
// Replaces the host loops of Scene::RebuildMeshBuffer (src/Assets/Scene.cpp:118-196) and the
// triangle gathering of FCPUAccelerationStructure::InitBVH
// (src/Assets/CPUAccelerationStructure.cpp:182-204): the raw fp32 Vertex arrays are copied
// once, then two kernels produce
//   - the fp16 GPUVertex shading buffer (Assets::MakeVertex, src/Assets/Vertex.hpp:80-99),
//   - the de-indexed fp32 triangle list + per-triangle boxes the BLAS builder consumes.
// Both kernels are pure streaming (coalesced 52-byte reads would straddle, so a warp reads
// its 32 vertices as 13 coalesced 128-byte words through shared memory).
#include "gk_context.h"

namespace gk {

__global__ void k_convert_vertices(const GkVertex* __restrict__ in, GkGPUVertex* __restrict__ out, uint32_t n)
{
    // stage 32 vertices (1664 B = 416 words) per warp through shared memory
    __shared__ uint32_t sm[8][416];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (blockIdx.x * 8 + warp) * 32;
    if (base >= n) return;
    const uint32_t cnt = min(32u, n - base);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + base);
    for (uint32_t w = lane; w < cnt * 13; w += 32) sm[warp][w] = src[w];
    __syncwarp();
    if ((uint32_t)lane < cnt) {
        const float* v = reinterpret_cast<const float*>(&sm[warp][lane * 13]);
        GkGPUVertex g;
        g.posx = glmToHalf(v[0]), g.posy = glmToHalf(v[1]), g.posz = glmToHalf(v[2]);
        g.texcoordx = glmToHalf(v[10]);
        g.normalx = glmToHalf(v[3]), g.normaly = glmToHalf(v[4]), g.normalz = glmToHalf(v[5]);
        g.texcoordy = glmToHalf(v[11]);
        g.tangentx = glmToHalf(v[6]), g.tangenty = glmToHalf(v[7]), g.tangentz = glmToHalf(v[8]);
        const uint32_t mat = sm[warp][lane * 13 + 12];
        g.tangentw = (uint16_t)(((v[9] > 0 ? 2 : 0) << 8) | (mat & 0xffffu));
        // 24-byte record: three 8-byte stores
        uint2* o = reinterpret_cast<uint2*>(out + base + lane);
        const uint2* s = reinterpret_cast<const uint2*>(&g);
        o[0] = s[0], o[1] = s[1], o[2] = s[2];
    }
}

// One thread per triangle: gather the three fp32 positions (index order preserved, as the CPU
// BVH does), emit the primitive box and the owning model.
__global__ void k_gather_triangles(const GkVertex* __restrict__ verts, const uint32_t* __restrict__ indices, const ModelInfo* __restrict__ models,
                                   uint32_t modelCount, uint32_t triCount, float4* __restrict__ triP, float4* __restrict__ plo,
                                   float4* __restrict__ phi, uint32_t* __restrict__ group, float4* __restrict__ faceNormal)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= triCount) return;
    // binary search the model owning triangle t
    uint32_t lo = 0, hi = modelCount - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (models[mid].triOffset <= t) lo = mid;
        else hi = mid - 1;
    }
    const ModelInfo M = models[lo];
    const uint32_t local = t - M.triOffset;
    float3 p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t vi = M.vertexOffset + indices[M.indexOffset + local * 3 + k];
        const float* v = reinterpret_cast<const float*>(verts + vi);
        p[k] = make_float3(v[0], v[1], v[2]);
        triP[(size_t)t * 3 + k] = make_float4(p[k].x, p[k].y, p[k].z, 0.f);
    }
    plo[t] = make_float4(fminf(p[0].x, fminf(p[1].x, p[2].x)), fminf(p[0].y, fminf(p[1].y, p[2].y)), fminf(p[0].z, fminf(p[1].z, p[2].z)), 0.f);
    phi[t] = make_float4(fmaxf(p[0].x, fmaxf(p[1].x, p[2].x)), fmaxf(p[0].y, fmaxf(p[1].y, p[2].y)), fmaxf(p[0].z, fmaxf(p[1].z, p[2].z)), 0.f);
    group[t] = lo;
    // face normal as InitBVH computes it (CPUAccelerationStructure.cpp:193-195): normalize(cross(v1-v0, v2-v1)),
    // glm::normalize = v * inversesqrt(dot(v,v))
    const float e1x = p[1].x - p[0].x, e1y = p[1].y - p[0].y, e1z = p[1].z - p[0].z;
    const float e2x = p[2].x - p[1].x, e2y = p[2].y - p[1].y, e2z = p[2].z - p[1].z;
    const float cx = e1y * e2z - e1z * e2y, cy = e1z * e2x - e1x * e2z, cz = e1x * e2y - e1y * e2x;
    const float inv = 1.0f / sqrtf(cx * cx + cy * cy + cz * cz);
    faceNormal[t] = make_float4(cx * inv, cy * inv, cz * inv, 0.f);
}

// scratch shared with the BLAS builder
DevBuf<float4>& sceneTriPositions()
{
    static thread_local DevBuf<float4> buf;
    return buf;
}

GkStatus uploadScene(Context& c, const GkSceneDesc& d)
{
    if (d.modelCount == 0 || !d.models) {
        setLastError("gk_upload_scene: scene has no models");
        return GK_ERR_INVALID_ARGUMENT;
    }
    c.models.assign(d.modelCount, ModelInfo{});
    uint64_t vtx = 0, idx = 0, tri = 0;
    for (uint32_t m = 0; m < d.modelCount; ++m) {
        const GkModelDesc& md = d.models[m];
        if (md.indexCount % 3 != 0 || (md.vertexCount && !md.vertices) || (md.indexCount && !md.indices)) {
            setLastError("gk_upload_scene: model " + std::to_string(m) + " is malformed");
            return GK_ERR_INVALID_ARGUMENT;
        }
        ModelInfo& M = c.models[m];
        M.vertexOffset = (uint32_t)vtx, M.vertexCount = md.vertexCount;
        M.indexOffset = (uint32_t)idx, M.indexCount = md.indexCount;
        M.triOffset = (uint32_t)tri, M.triCount = md.indexCount / 3;
        M.blasRoot = kInvalid;
        vtx += md.vertexCount, idx += md.indexCount, tri += md.indexCount / 3;
    }
    if (tri == 0 || tri >= (1ull << 28) || vtx >= (1ull << 31)) {
        setLastError("gk_upload_scene: triangle/vertex count out of range");
        return GK_ERR_INVALID_ARGUMENT;
    }
    c.totalTris = tri;

    DevBuf<GkVertex> raw;
    GK_CUDA(raw.reserve(vtx));
    GK_CUDA(c.dGpuVerts.reserve(vtx));
    GK_CUDA(c.dIndices.reserve(idx));
    GK_CUDA(c.dModels.reserve(d.modelCount));
    for (uint32_t m = 0; m < d.modelCount; ++m) {
        const GkModelDesc& md = d.models[m];
        const ModelInfo& M = c.models[m];
        if (md.vertexCount) GK_CUDA(cudaMemcpyAsync(raw.p + M.vertexOffset, md.vertices, sizeof(GkVertex) * md.vertexCount, cudaMemcpyHostToDevice, c.stream));
        if (md.indexCount) GK_CUDA(cudaMemcpyAsync(c.dIndices.p + M.indexOffset, md.indices, sizeof(uint32_t) * md.indexCount, cudaMemcpyHostToDevice, c.stream));
    }
    GK_CUDA(cudaMemcpyAsync(c.dModels.p, c.models.data(), sizeof(ModelInfo) * d.modelCount, cudaMemcpyHostToDevice, c.stream));

    c.materialCount = d.materialCount;
    GK_CUDA(c.dMaterials.reserve(d.materialCount ? d.materialCount : 1));
    if (d.materialCount) GK_CUDA(cudaMemcpyAsync(c.dMaterials.p, d.materials, sizeof(GkMaterial) * d.materialCount, cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(c.dLights.reserve(d.lightCount ? d.lightCount : 1));
    if (d.lightCount) GK_CUDA(cudaMemcpyAsync(c.dLights.p, d.lights, sizeof(GkLightObject) * d.lightCount, cudaMemcpyHostToDevice, c.stream));

    k_convert_vertices<<<(unsigned)((vtx + 255) / 256), 256, 0, c.stream>>>(raw.p, c.dGpuVerts.p, (uint32_t)vtx);

    Lbvh& T = c.blasTree;
    T.n = (uint32_t)tri;
    GK_CUDA(T.plo.reserve(tri));
    GK_CUDA(T.phi.reserve(tri));
    GK_CUDA(T.group.reserve(tri));
    GK_CUDA(sceneTriPositions().reserve(tri * 3));
    GK_CUDA(c.dFaceNormals.reserve(tri));
    k_gather_triangles<<<(unsigned)((tri + 255) / 256), 256, 0, c.stream>>>(raw.p, c.dIndices.p, c.dModels.p, d.modelCount, (uint32_t)tri,
                                                                           sceneTriPositions().p, T.plo.p, T.phi.p, T.group.p, c.dFaceNormals.p);
    GK_CUDA(cudaGetLastError());
    GkStatus s = buildBlasForest(c);
    GK_CUDA(cudaStreamSynchronize(c.stream));
    raw.release();
    sceneTriPositions().release();
    if (s != GK_OK) return s;
    c.haveScene = true;
    c.haveInstances = false;
    return GK_OK;
}

} // namespace gk
