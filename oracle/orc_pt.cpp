// orc_pt.cpp — see orc_pt.h.  TEST INFRASTRUCTURE ONLY.
#include "orc_pt.h"
#include <thread>

namespace orc {

namespace {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.283185307179586476925f;
constexpr float kInvPi = 0.318309886183790671538f;
constexpr float kPiOver4 = 0.785398163397448309616f;
constexpr float kEps = 1e-3f;                // PreProcessor.slang:13
constexpr float kMaxTrace = 1000.f;          // Shading.slang:15
constexpr float kTraceOffset = 0.001f;       // Shading.slang:18
constexpr float kPrimaryTMax = 2000.f;       // RayCastInCPU, CPUAccelerationStructure.cpp:287
constexpr float kNearZero = 1e-35f;

struct u4 {
    uint32_t x, y, z, w;
};

// Const_Func.slang:227-233
inline void pcg4d(u4& v)
{
    v.x = v.x * 1664525u + 1013904223u;
    v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u;
    v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16u; v.y ^= v.y >> 16u; v.z ^= v.z >> 16u; v.w ^= v.w >> 16u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
}
inline float u2f(uint32_t x) // Const_Func.slang:235
{
    uint32_t b = 0x3f800000u | (x >> 9);
    float f;
    memcpy(&f, &b, 4);
    return f - 1.0f;
}
inline float randomFloat(u4& s) { pcg4d(s); return u2f(s.x); }
inline f2 randomFloat2(u4& s) { pcg4d(s); return f2(u2f(s.x), u2f(s.y)); }

inline float pow5(float x) { return x * x * x * x * x; }
inline float schlick(float cosine, float ri) // Const_Func.slang:8-14
{
    float r0 = (1 - ri) / (1 + ri);
    r0 *= r0;
    return r0 + (1 - r0) * pow5(1 - cosine);
}
inline void onb(f3 n, f3& b1, f3& b2) // Const_Func.slang:17-22
{
    float signZ = n.z < 0.f ? -1.f : 1.f;
    float a = -1.0f / (signZ + n.z);
    b2 = f3(n.x * n.y * a, signZ + n.y * n.y * a, -n.y);
    b1 = f3(1.0f + signZ * n.x * n.x * a, signZ * b2.x, -signZ * n.x);
}
// to_world(v,T,B,N) = mul(v, float3x3(T,B,N)) = v.x*T + v.y*B + v.z*N   (PreProcessor.slang:2)
inline f3 toWorld(f3 v, f3 T, f3 B, f3 N)
{
    return f3(v.x * T.x + v.y * B.x + v.z * N.x, v.x * T.y + v.y * B.y + v.z * N.y, v.x * T.z + v.y * B.z + v.z * N.z);
}
// to_local(v,t,b,n) = mul(float3x3(t,b,n), v) = (t.v, b.v, n.v)          (PreProcessor.slang:1)
inline f3 toLocal(f3 v, f3 t, f3 b, f3 n) { return f3(dot(t, v), dot(b, v), dot(n, v)); }
inline f3 alignWithNormal(f3 ray, f3 normal) // Const_Func.slang:44-49
{
    f3 T, B;
    onb(normal, T, B);
    return toWorld(ray, T, B, normal);
}
inline f2 concentricDisk(f2 o) // Const_Func.slang:260-275
{
    o = f2(o.x + (o.x - 1.0f), o.y + (o.y - 1.0f));
    auto isZero = [](float x) { return x > -kNearZero && x < kNearZero; };
    if (isZero(o.x) && isZero(o.y)) return f2(0, 0);
    float theta;
    if (fabsf(o.x) > fabsf(o.y)) {
        theta = kPiOver4 * o.y / o.x;
        return f2(o.x * cosf(theta), o.x * sinf(theta));
    }
    float ct = sinf(kPiOver4 * o.x / o.y);
    return f2(o.y * ct, o.y * sqrtf(1.f - ct * ct));
}
inline f3 randomInCone(u4& s, float cosTheta) // Const_Func.slang:282-289
{
    const f2 u = randomFloat2(s);
    float phi = kTwoPi * u.x;
    cosTheta = 1.0f + u.y * (cosTheta - 1.f);
    float r = sqrtf(1.0f - cosTheta * cosTheta);
    return f3(r * cosf(phi), r * sinf(phi), cosTheta);
}
inline f3 randomInHemiSphere1(u4& s) // Const_Func.slang:299-305
{
    const f2 u = randomFloat2(s);
    float phi = kTwoPi * u.x;
    float r = sqrtf(u.y);
    return f3(r * cosf(phi), r * sinf(phi), sqrtf(1.0f - u.y));
}
inline float saturate(float x) { return clampf(x, 0.0f, 1.0f); }
inline f3 ggxSampleVndf(f2 alpha, f3 wi_, f2 uv) // Const_Func.slang:310-328
{
    f3 wi = normalize(f3(wi_.x * alpha.x, wi_.y * alpha.y, wi_.z));
    float b = wi.z;
    if (wi_.z > 0.f) {
        float a = saturate(fminf_(alpha.x, alpha.y));
        float awiz_s = a * wi_.z / (1.0f + sqrtf(wi_.x * wi_.x + wi_.y * wi_.y));
        b *= ((1.0f - a * a) / (1.0f + awiz_s * awiz_s));
    }
    float z = fmaf(1.0f - uv.y, 1.0f + b, -b);
    float phi = kTwoPi * uv.x;
    float r = sqrtf(saturate(1.0f - z * z));
    f3 o_std(r * cosf(phi), r * sinf(phi), z);
    f3 m_std = wi + o_std;
    return normalize(f3(m_std.x * alpha.x, m_std.y * alpha.y, m_std.z));
}
inline f3 ggxSampling(u4& s, float roughness, f3 normal) // Const_Func.slang:330-340
{
    f3 t, b;
    onb(normal, t, b);
    f3 wm = ggxSampleVndf(f2(roughness * roughness, roughness * roughness), toLocal(normal, t, b, normal), randomFloat2(s));
    return toWorld(wm, t, b, normal);
}
inline f3 reflect3(f3 i, f3 n) { return i - n * (2.0f * dot(n, i)); }
inline f3 refract3(f3 i, f3 n, float eta)
{
    float d = dot(n, i);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return f3(0, 0, 0);
    return i * eta - n * (eta * d + sqrtf(k));
}

struct Vtx { // the fields of the shader's Vertex that the path tracer reads
    f3 Position, Normal;
    f2 TexCoord;
    uint32_t MaterialIndex;
};

struct UnpackedV {
    f3 P, N;
    f2 uv;
    uint32_t mat;
};
inline UnpackedV unpackVertex(const GkGPUVertex& g) // Const_Func.slang:342-354
{
    UnpackedV v;
    v.P = f3(half_to_float(g.posx), half_to_float(g.posy), half_to_float(g.posz));
    v.N = f3(half_to_float(g.normalx), half_to_float(g.normaly), half_to_float(g.normalz));
    v.uv = f2(half_to_float(g.texcoordx), half_to_float(g.texcoordy));
    v.mat = g.tangentw & 0xFF;
    return v;
}

inline m4 asM4(const float* p) { m4 M; memcpy(M.m, p, 64); return M; }
inline f3 fma3(float s, f3 a, f3 c) { return f3(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z)); }

struct Ctx {
    const Scene& S;
    const GkUniformBufferObject& U;
    const GkAmbientCube* cubes;
    const GkVoxelData* voxels;
    uint32_t rays;
};

// GeneralFunc.slang:33-83 — shading vertex of a primary hit from its visibility id.
Vtx getMaterialData(const Ctx& c, uint32_t node, uint32_t prim, f3 ro, f3 rd, uint32_t& rawMat)
{
    const GkNodeProxy& px = c.S.nodes[node];
    const Model& M = c.S.models[px.modelId / 10];
    const m4 W = asM4(px.worldTS);
    f3 P[3], N[3];
    f2 T[3];
    uint32_t matid = 0;
    for (int i = 0; i < 3; ++i) {
        UnpackedV v = unpackVertex(M.gpuVerts[M.indices[prim * 3 + i]]);
        P[i] = mul(W, f4(v.P, 1)).xyz();
        N[i] = mul(W, f4(v.N, 0)).xyz();
        T[i] = v.uv;
        if (i == 0) matid = v.mat;
    }
    f3 e0 = P[1] - P[0], e1 = P[2] - P[0];
    f3 rce1 = cross(rd, e1);
    float rcpDet = 1.0f / dot(e0, rce1);
    f3 r0 = ro - P[0];
    float by = rcpDet * dot(r0, rce1);
    f3 e0c0 = cross(e0, r0);
    float bz = -rcpDet * dot(rd, e0c0);
    float bx = 1.0f - (by + bz);
    Vtx r;
    r.Position = fma3(bx, P[0], fma3(by, P[1], P[2] * bz));
    r.Normal = normalize(fma3(bx, N[0], fma3(by, N[1], N[2] * bz)));
    r.TexCoord = f2(fmaf(bx, T[0].x, fmaf(by, T[1].x, bz * T[2].x)), fmaf(bx, T[0].y, fmaf(by, T[1].y, bz * T[2].y)));
    r.MaterialIndex = matid;
    rawMat = matid;
    return r;
}

// Shading.slang:708-750 — closest hit resolved to a shading vertex.
bool traceRay(Ctx& c, f3 ro, f3 rd, float tmax, Vtx& out, uint32_t& outNode)
{
    ++c.rays;
    Hit h;
    if (!c.S.trace(ro, rd, kEps, tmax, h)) return false;
    const GkNodeProxy& px = c.S.nodes[h.inst];
    const Model& M = c.S.models[px.modelId / 10];
    UnpackedV v0 = unpackVertex(M.gpuVerts[M.indices[h.prim * 3]]);
    UnpackedV v1 = unpackVertex(M.gpuVerts[M.indices[h.prim * 3 + 1]]);
    UnpackedV v2 = unpackVertex(M.gpuVerts[M.indices[h.prim * 3 + 2]]);
    // Mix(a,b,c,bary) = a + (b-a)*bary.y + (c-a)*bary.z with bary = (1-u-v, u, v)
    f3 n = v0.N + (v1.N - v0.N) * h.u + (v2.N - v0.N) * h.v;
    // mul(WorldToObject(4x3), n).xyz = transpose(inverse(world3x3)) * n
    float inv[16], T[16];
    for (int r = 0; r < 4; ++r)
        for (int cc = 0; cc < 4; ++cc) T[r * 4 + cc] = px.worldTS[cc * 4 + r];
    for (int k = 0; k < 16; ++k) inv[k] = (k % 5 == 0) ? 1.f : 0.f;
    invert4x4RowMajor(T, inv);
    f3 nw(inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z, inv[2] * n.x + inv[6] * n.y + inv[10] * n.z);
    out.Normal = normalize(nw);
    out.TexCoord = f2(v0.uv.x + (v1.uv.x - v0.uv.x) * h.u + (v2.uv.x - v0.uv.x) * h.v, v0.uv.y + (v1.uv.y - v0.uv.y) * h.u + (v2.uv.y - v0.uv.y) * h.v);
    out.Position = ro + rd * h.t;
    out.MaterialIndex = px.matId[v0.mat & 15];
    outNode = h.inst;
    return true;
}

bool traceOcclusion(Ctx& c, f3 ro, f3 rd) // Shading.slang:661-681
{
    ++c.rays;
    return c.S.anyHit(ro, rd, kEps, kMaxTrace);
}

// Shading.slang:148-153 with the lat-long texture replaced by a constant texel
// (ubo.BackGroundColor); see DESIGN.md "sky".
inline f3 skyColor(const GkUniformBufferObject& U)
{
    if (!U.HasSky) return f3(0, 0, 0);
    return f3(fminf_(10.f, U.BackGroundColor[0]), fminf_(10.f, U.BackGroundColor[1]), fminf_(10.f, U.BackGroundColor[2])) * U.SkyIntensity;
}

inline f3 unpackRGB10(uint32_t p) // AmbientCube.slang:71-78
{
    return f3(float(p & 0x3FF) / 1023.0f, float((p >> 10) & 0x3FF) / 1023.0f, float((p >> 20) & 0x3FF) / 1023.0f) * 512.f;
}

f3 sampleCubeFull(const GkAmbientCube& cb, f3 n) // AmbientCube.slang:178-223 (rgb part)
{
    float wx = fmaxf_(n.x, 0.f), wnx = fmaxf_(-n.x, 0.f), wy = fmaxf_(n.y, 0.f), wny = fmaxf_(-n.y, 0.f), wz = fmaxf_(n.z, 0.f), wnz = fmaxf_(-n.z, 0.f);
    float sum = wx + wnx + wy + wny + wz + wnz;
    f3 col(0, 0, 0);
    col = col + unpackRGB10(cb.PosX_D) * wx;
    col = col + unpackRGB10(cb.NegX_D) * wnx;
    col = col + unpackRGB10(cb.PosY_D) * wy;
    col = col + unpackRGB10(cb.NegY_D) * wny;
    col = col + unpackRGB10(cb.PosZ_D) * wz;
    col = col + unpackRGB10(cb.NegZ_D) * wnz;
    col = col + unpackRGB10(cb.PosX) * wx;
    col = col + unpackRGB10(cb.NegX) * wnx;
    col = col + unpackRGB10(cb.PosY) * wy;
    col = col + unpackRGB10(cb.NegY) * wny;
    col = col + unpackRGB10(cb.PosZ) * wz;
    col = col + unpackRGB10(cb.NegZ) * wnz;
    col = col * ((sum > 0.0f) ? (1.0f / sum) : 1.0f);
    return col;
}

// AmbientCube.slang:275-364.  With no probe data bound (cubes == nullptr) every probe reads
// as zero, the state the reference is in while progressive / benchmarking
// (RayTraceBaseRenderer.cpp:239): the function then returns rgb = 0.
f3 interpolateAmbientCubes(const Ctx& c, f3 pos, f3 normal)
{
    const f3 off = f3(-float(GK_CUBE_SIZE_XY / 2), -1.375f, -float(GK_CUBE_SIZE_XY / 2)) * GK_CUBE_UNIT;
    f3 np = (pos - off) / GK_CUBE_UNIT;
    if (np.x < 0 || np.y < 0 || np.z < 0 || np.x > GK_CUBE_SIZE_XY - 1 || np.y > GK_CUBE_SIZE_Z - 1 || np.z > GK_CUBE_SIZE_XY - 1) return f3(0, 0, 0);
    if (!c.cubes || !c.voxels) return f3(0, 0, 0);
    int bx = (int)floorf(np.x), by = (int)floorf(np.y), bz = (int)floorf(np.z);
    f3 fr(np.x - floorf(np.x), np.y - floorf(np.y), np.z - floorf(np.z));
    float total = 0;
    f3 result(0, 0, 0);
    for (int i = 0; i < 8; ++i) {
        int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
        int idx = (by + oy) * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + (bz + oz) * GK_CUBE_SIZE_XY + (bx + ox);
        const GkAmbientCube& cb = c.cubes[idx];
        const GkVoxelData& vx = c.voxels[idx];
        uint32_t p0 = vx.distanceToSolid_gg_z01, p1 = vx.distanceToSolid_x01_y01;
        float d0y = float((p0 >> 8) & 0xFF) / 255.0f;
        if (d0y < 0.01f) continue;
        float dPZ = float((p0 >> 16) & 0xFF) / 255.0f, dNZ = float((p0 >> 24) & 0xFF) / 255.0f;
        float dPX = float(p1 & 0xFF) / 255.0f, dNX = float((p1 >> 8) & 0xFF) / 255.0f, dPY = float((p1 >> 16) & 0xFF) / 255.0f, dNY = float((p1 >> 24) & 0xFF) / 255.0f;
        f3 ptl = fr - f3((float)ox, (float)oy, (float)oz);
        float dist = length(ptl);
        f3 dir = normalize(ptl);
        float hitLen = sqrtf(fmaxf_(dir.x, 0.f)) * dPX + sqrtf(fmaxf_(-dir.x, 0.f)) * dNX + sqrtf(fmaxf_(dir.y, 0.f)) * dPY + sqrtf(fmaxf_(-dir.y, 0.f)) * dNY +
                       sqrtf(fmaxf_(dir.z, 0.f)) * dPZ + sqrtf(fmaxf_(-dir.z, 0.f)) * dNZ;
        if (dist > hitLen + 0.05f) continue;
        float wx = ox == 0 ? (1.0f - fr.x) : fr.x, wy = oy == 0 ? (1.0f - fr.y) : fr.y, wz = oz == 0 ? (1.0f - fr.z) : fr.z;
        float w = wx * wy * wz;
        result = result + sampleCubeFull(cb, normal) * w;
        total += w;
    }
    return total > 0.0f ? result / total : f3(0, 0, 0);
}

// Shading.slang:932-996.  Returns true when the sample's path ends here.
bool getRayColor(Ctx& c, Vtx& v, f3& rayDir, f3& rayColor, u4& seed, bool& hitReflect, bool& hitMetal)
{
    const GkMaterial& mat = c.S.materials[v.MaterialIndex];
    const float startPosOffset = mat.MaterialModel == GK_MAT_DIELECTRIC ? 0.0f : 1.0f;
    const float roughness = mat.Fuzziness;
    const float dotValue = dot(rayDir, v.Normal);
    const bool backFace = dotValue > 0;
    const f3 outwardNormal = backFace ? -v.Normal : v.Normal;
    const float niOverNt = backFace ? mat.RefractionIndex2 : (1 / mat.RefractionIndex2);
    const float cosine = dotValue > 0 ? mat.RefractionIndex * dotValue : -dotValue;
    const float reflectProb = schlick(cosine, mat.RefractionIndex);
    const float metalProb = mat.Metalness;

    const bool chanceReflect = randomFloat(seed) < reflectProb;
    const bool chanceMetal = randomFloat(seed) < metalProb;
    const bool chanceGGX = chanceReflect || chanceMetal;
    const f3 traceNext = chanceGGX ? reflect3(rayDir, outwardNormal) : outwardNormal;
    f3 traceDir = chanceGGX ? ggxSampling(seed, sqrtf(roughness), traceNext) : alignWithNormal(randomInHemiSphere1(seed), traceNext);

    hitReflect = chanceGGX;
    hitMetal = chanceMetal; // NB: callers may alias both flags (Shading.slang:1032)

    if (mat.MaterialModel == GK_MAT_DIELECTRIC && !chanceReflect) traceDir = refract3(rayDir, outwardNormal, niOverNt);

    rayDir = traceDir;

    uint32_t hitNode;
    const f3 origin = v.Position + v.Normal * kTraceOffset * startPosOffset;
    if (traceRay(c, origin, traceDir, kMaxTrace, v, hitNode)) {
        const GkMaterial& hm = c.S.materials[v.MaterialIndex];
        const f3 albedo(hm.Diffuse[0], hm.Diffuse[1], hm.Diffuse[2]);
        if (hm.MaterialModel == GK_MAT_DIFFUSE_LIGHT || !chanceReflect) rayColor = rayColor * albedo;
        if (backFace && mat.MaterialModel != GK_MAT_DIELECTRIC) {
            rayColor = f3(0, 0, 0);
            return true;
        }
        if (hm.MaterialModel == GK_MAT_DIFFUSE_LIGHT) return true;
        return false;
    }
    rayColor = rayColor * skyColor(c.U);
    return true;
}

void renderPixel(Ctx& c, uint32_t x, uint32_t y, uint32_t W, uint32_t H, const uint32_t* vis, PtOutputs& o)
{
    const GkUniformBufferObject& U = c.U;
    const size_t pi = (size_t)y * W + x;
    c.rays = 1; // the primary ray, traced in the visibility pre-pass
    u4 seed{x, y, U.TotalFrames, 0};

    const m4 MVI = asM4(U.ModelViewInverse), PI = asM4(U.ProjectionInverse);
    auto cameraDir = [&](int px, int py) {
        f2 uv((float(px) / float(W)) * 2.0f - 1.0f, (float(py) / float(H)) * 2.0f - 1.0f);
        f4 target = mul(PI, f4(uv.x, uv.y, 1, 1));
        f4 dir = mul(MVI, f4(normalize(target.xyz()), 0));
        return normalize(dir.xyz());
    };
    const f3 origin = mul(MVI, f4(0, 0, 0, 1)).xyz();
    const f3 rayDir0 = cameraDir((int)x, (int)y);

    auto writeMiss = [&]() {
        f3 sky = skyColor(U);
        o.diffuse[4 * pi] = sky.x, o.diffuse[4 * pi + 1] = sky.y, o.diffuse[4 * pi + 2] = sky.z, o.diffuse[4 * pi + 3] = U.HasSky ? fminf_(1.f, U.BackGroundColor[3]) * U.SkyIntensity : 0.f;
        for (int k = 0; k < 4; ++k) o.spec[4 * pi + k] = 0.f; // not written by the shader; defined as 0 here
        o.motion[2 * pi] = o.motion[2 * pi + 1] = 0.f;
        for (int k = 0; k < 4; ++k) o.albedo[4 * pi + k] = 1.f;
        o.normal[4 * pi] = 0, o.normal[4 * pi + 1] = 1, o.normal[4 * pi + 2] = 0, o.normal[4 * pi + 3] = 1;
        o.objectId[pi] = 65535;
        o.depth[pi] = 0.f;
        o.rayCount[pi] = c.rays;
    };

    uint32_t prim = vis[2 * pi], node = vis[2 * pi + 1];
    if (node == 0xffffffffu) { writeMiss(); return; }

    // Shading.slang:287-434
    uint32_t rawMat;
    Vtx initial = getMaterialData(c, node, prim, origin, rayDir0, rawMat);
    const float vertexDistance = length(initial.Position - origin);
    float cocRadius = 0.0f;
    if (fabsf(vertexDistance - U.FocusDistance) > 0.001f) cocRadius = (U.Aperture * fabsf(vertexDistance - U.FocusDistance)) / vertexDistance;
    f2 pixelOffset(0, 0);
    if (cocRadius > 0.001f) {
        f2 disk = concentricDisk(randomFloat2(seed));
        f3 right = normalize(cross(rayDir0, f3(0, 1, 0)));
        f3 edge = initial.Position + right * cocRadius;
        const m4 VP = asM4(U.ViewProjection);
        f4 cp = mul(VP, f4(initial.Position, 1)), ep = mul(VP, f4(edge, 1));
        f2 cxy(cp.x / cp.w, cp.y / cp.w), exy(ep.x / ep.w, ep.y / ep.w);
        float ssr = length(exy - cxy);
        pixelOffset = f2(disk.x * ssr * float(W) * 0.5f, disk.y * ssr * float(H) * 0.5f);
    }
    int ox = (int)x + (int)pixelOffset.x, oy = (int)y + (int)pixelOffset.y;
    ox = ox < 0 ? 0 : (ox > (int)W - 1 ? (int)W - 1 : ox);
    oy = oy < 0 ? 0 : (oy > (int)H - 1 ? (int)H - 1 : oy);
    const size_t opi = (size_t)oy * W + ox;
    const uint32_t fprim = vis[2 * opi], fnode = vis[2 * opi + 1];
    Vtx hitV;
    uint32_t hitNode;
    f3 rayDir;
    bool useInitial = (fnode == 0xffffffffu);
    Vtx finalV;
    uint32_t finalRaw = 0;
    f3 finalDir = rayDir0;
    if (!useInitial) {
        finalDir = cameraDir(ox, oy);
        finalV = getMaterialData(c, fnode, fprim, origin, finalDir, finalRaw);
        const float distanceToFocus = fabsf(vertexDistance - U.FocusDistance);
        if (distanceToFocus < U.FocusDistance * 0.1f) {
            float fd = length(finalV.Position - origin);
            if (vertexDistance - fd > U.FocusDistance * 0.05f) useInitial = true;
        }
    }
    if (useInitial) {
        hitNode = node;
        hitV.Position = initial.Position;
        hitV.Normal = normalize(initial.Normal);
        hitV.TexCoord = initial.TexCoord;
        hitV.MaterialIndex = c.S.nodes[node].matId[rawMat & 15];
        rayDir = rayDir0;
    } else {
        hitNode = fnode;
        hitV.Position = finalV.Position;
        hitV.Normal = normalize(finalV.Normal);
        hitV.TexCoord = finalV.TexCoord;
        hitV.MaterialIndex = c.S.nodes[fnode].matId[finalRaw & 15];
        rayDir = finalDir;
    }
    (void)rayDir;
    const GkNodeProxy& hn = c.S.nodes[hitNode];

    // Shading.slang:50-58
    {
        const m4 VPu = asM4(U.ViewProjectionUnJit), PVPu = asM4(U.PrevViewProjectionUnJit), prevTS = asM4(hn.combinedPrevTS);
        f4 cur = mul(VPu, f4(hitV.Position, 1));
        f2 curf(cur.x / cur.w * 0.5f, cur.y / cur.w * 0.5f);
        f4 prev = mul(matmul(PVPu, prevTS), f4(hitV.Position, 1));
        f2 prevf(prev.x / prev.w * 0.5f, prev.y / prev.w * 0.5f);
        o.motion[2 * pi] = (prevf.x - curf.x) * float(W);
        o.motion[2 * pi + 1] = (prevf.y - curf.y) * float(H);
    }

    // Shading.slang:68-98 (no textures bound: SURVEY.md §8f N4)
    const GkMaterial& mat = c.S.materials[hitV.MaterialIndex];
    const f4 albedo(mat.Diffuse[0], mat.Diffuse[1], mat.Diffuse[2], mat.Diffuse[3]);
    o.albedo[4 * pi] = albedo.x, o.albedo[4 * pi + 1] = albedo.y, o.albedo[4 * pi + 2] = albedo.z, o.albedo[4 * pi + 3] = albedo.w;
    o.normal[4 * pi] = hitV.Normal.x, o.normal[4 * pi + 1] = hitV.Normal.y, o.normal[4 * pi + 2] = hitV.Normal.z, o.normal[4 * pi + 3] = mat.Fuzziness;
    o.objectId[pi] = hn.instanceId;

    // Shading.slang:998-1082
    f3 finalColor(0, 0, 0), finalRefl(0, 0, 0);
    if (mat.MaterialModel == GK_MAT_DIFFUSE_LIGHT) {
        finalColor = albedo.xyz();
    } else {
        const uint32_t samples = U.FastGather ? 1 : U.NumberOfSamples;
        for (uint32_t i = 0; i < samples; ++i) {
            f3 rayColor(1, 1, 1);
            f3 direction = normalize(hitV.Position - origin);
            Vtx vs = hitV;
            const uint32_t maxBounces = mat.MaterialModel == GK_MAT_DIELECTRIC ? U.MaxNumberOfBounces : U.NumberOfBounces;
            bool hitReflect = false, hitMetal = false;
            const bool exitFirst = getRayColor(c, vs, direction, rayColor, seed, hitReflect, hitMetal);
            if (!exitFirst) {
                for (uint32_t b = 1; b < maxBounces; ++b) {
                    bool dontCare = false; // both inout flags alias this one variable (Shading.slang:1030-1032)
                    if (getRayColor(c, vs, direction, rayColor, seed, dontCare, dontCare)) break;
                    // `hitReflectDontCare` is never written, so the guard is always true (Shading.slang:1037)
                    if (U.HasSun && (randomFloat(seed) < 0.5f)) {
                        const f3 lv(U.SunDirection[0], U.SunDirection[1], U.SunDirection[2]);
                        const f3 cone = alignWithNormal(randomInCone(seed, cosf(0.25f / 180.f * kPi)), lv);
                        if (!traceOcclusion(c, vs.Position + vs.Normal * kTraceOffset, cone)) {
                            rayColor = rayColor * f3(U.SunColor[0], U.SunColor[1], U.SunColor[2]);
                            break;
                        }
                    }
                    const bool earlyExit = (mat.MaterialModel != GK_MAT_DIELECTRIC) && (randomFloat(seed) < 0.5f);
                    if (b == maxBounces - 1 || earlyExit) {
                        rayColor = rayColor * interpolateAmbientCubes(c, vs.Position, vs.Normal);
                        break;
                    }
                }
            }
            if (hitMetal) rayColor = rayColor * albedo.xyz();
            if (hitReflect) finalRefl = finalRefl + rayColor;
            else finalColor = finalColor + rayColor;
        }
        finalColor = finalColor / float(samples);
        finalRefl = finalRefl / float(samples);

        // Shading.slang:826-850
        const f3 lv(U.SunDirection[0], U.SunDirection[1], U.SunDirection[2]);
        const float d = fmaxf_(dot(lv, normalize(hitV.Normal)), 0.0f) * kInvPi;
        float shadow = 0.0f;
        if (U.HasSun) {
            const f3 cone = alignWithNormal(randomInCone(seed, cosf(0.25f / 180.f * kPi)), lv);
            shadow = 1;
            if (traceOcclusion(c, hitV.Position, cone)) shadow = 0;
        }
        finalColor = finalColor + f3(U.SunColor[0], U.SunColor[1], U.SunColor[2]) * d * shadow;
    }

    const float offLen = length(pixelOffset);
    o.diffuse[4 * pi] = finalColor.x, o.diffuse[4 * pi + 1] = finalColor.y, o.diffuse[4 * pi + 2] = finalColor.z, o.diffuse[4 * pi + 3] = offLen;
    o.spec[4 * pi] = finalRefl.x, o.spec[4 * pi + 1] = finalRefl.y, o.spec[4 * pi + 2] = finalRefl.z, o.spec[4 * pi + 3] = offLen;
    {
        const m4 VP = asM4(U.ViewProjection);
        f4 clip = mul(VP, f4(hitV.Position, 1));
        o.depth[pi] = clip.z / clip.w;
    }
    o.rayCount[pi] = c.rays;
}

} // namespace

void renderFrame(const Scene& S, const GkUniformBufferObject& U, uint32_t W, uint32_t H, const GkAmbientCube* cubes, const GkVoxelData* voxels,
                 PtOutputs& o, int threads)
{
    if (threads < 1) threads = 1;
    // Visibility pre-pass (stands in for Rast.VisibilityPass + the first `extend` wave).
    const m4 MVI = asM4(U.ModelViewInverse), PI = asM4(U.ProjectionInverse);
    const f3 origin = mul(MVI, f4(0, 0, 0, 1)).xyz();
    auto rows = [&](auto fn) {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([&, t]() {
                for (uint32_t y = (uint32_t)t; y < H; y += (uint32_t)threads) fn(y);
            });
        for (auto& th : pool) th.join();
    };
    rows([&](uint32_t y) {
        for (uint32_t x = 0; x < W; ++x) {
            f2 uv((float(x) / float(W)) * 2.0f - 1.0f, (float(y) / float(H)) * 2.0f - 1.0f);
            f4 target = mul(PI, f4(uv.x, uv.y, 1, 1));
            f4 dir = mul(MVI, f4(normalize(target.xyz()), 0));
            f3 rd = normalize(dir.xyz());
            Hit h;
            size_t pi = (size_t)y * W + x;
            if (S.trace(origin, rd, 0.0f, kPrimaryTMax, h)) o.primIds[2 * pi] = h.prim, o.primIds[2 * pi + 1] = h.inst;
            else o.primIds[2 * pi] = o.primIds[2 * pi + 1] = 0xffffffffu;
        }
    });
    rows([&](uint32_t y) {
        Ctx c{S, U, cubes, voxels, 0};
        for (uint32_t x = 0; x < W; ++x) renderPixel(c, x, y, W, H, o.primIds, o);
    });
}

// ---------------------------------------------------------------------------------------------------------------
// Probe baker: Bake.HwAmbientCube.comp.slang:29-46 + FGpuProbeGenerator (common/AmbientCube.slang:453-675), restated.
// Same stated deviations as the path tracer (no textures, constant sky instead of the SH sky of SampleIBLRough), and the
// gathers read the probe state as it was before the call (the shader's read-while-write order is undefined).
namespace {

struct c4 {
    float x, y, z, w;
    c4() : x(0), y(0), z(0), w(0) {}
    c4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
inline c4 operator+(c4 a, c4 b) { return c4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline c4 operator*(c4 a, c4 b) { return c4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline c4 operator*(c4 a, float t) { return c4(a.x * t, a.y * t, a.z * t, a.w * t); }
inline c4 operator/(c4 a, float t) { return c4(a.x / t, a.y / t, a.z / t, a.w / t); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; } // lerp as FMix: x(1-a) + ya

inline c4 unpackRGB10A2(uint32_t p) // AmbientCube.slang:71-78 (MAX_ILLUMINANCE 512)
{
    return c4(float(p & 0x3FF) / 1023.0f, float((p >> 10) & 0x3FF) / 1023.0f, float((p >> 20) & 0x3FF) / 1023.0f, 0.0f) * 512.f;
}
inline uint32_t packRGB10A2(c4 c) // AmbientCube.slang:59-69
{
    float r = clampf(c.x / 512.f, 0.f, 1.f), g = clampf(c.y / 512.f, 0.f, 1.f), b = clampf(c.z / 512.f, 0.f, 1.f), a = clampf(c.w / 512.f, 0.f, 1.f);
    return (uint32_t)(r * 1023.0f) | ((uint32_t)(g * 1023.0f) << 10) | ((uint32_t)(b * 1023.0f) << 20) | ((uint32_t)(a * 3.0f) << 30);
}
inline uint32_t lerpPackedColorAlt(uint32_t c0, c4 c1, float t) // AmbientCube.slang:119-126
{
    c4 a = unpackRGB10A2(c0);
    return packRGB10A2(c4(mixf(a.x, c1.x, t), mixf(a.y, c1.y, t), mixf(a.z, c1.z, t), mixf(a.w, c1.w, t)));
}
c4 sampleCubeDI(const GkAmbientCube& cb, f3 n) // sampleAmbientCubeHL2_DI, AmbientCube.slang:128-151
{
    float wx = fmaxf_(n.x, 0.f), wnx = fmaxf_(-n.x, 0.f), wy = fmaxf_(n.y, 0.f), wny = fmaxf_(-n.y, 0.f), wz = fmaxf_(n.z, 0.f), wnz = fmaxf_(-n.z, 0.f);
    float sum = wx + wnx + wy + wny + wz + wnz;
    c4 col;
    col = col + unpackRGB10A2(cb.PosX_D) * wx;
    col = col + unpackRGB10A2(cb.NegX_D) * wnx;
    col = col + unpackRGB10A2(cb.PosY_D) * wy;
    col = col + unpackRGB10A2(cb.NegY_D) * wny;
    col = col + unpackRGB10A2(cb.PosZ_D) * wz;
    col = col + unpackRGB10A2(cb.NegZ_D) * wnz;
    return col * ((sum > 0.0f) ? (1.0f / sum) : 1.0f);
}
c4 interpolateDI(const GkAmbientCube* cubes, const GkVoxelData* voxels, f3 pos, f3 normal) // AmbientCube.slang:275-364, T = DIAmbientCubeSampler
{
    const f3 off = f3(-float(GK_CUBE_SIZE_XY / 2), -1.375f, -float(GK_CUBE_SIZE_XY / 2)) * GK_CUBE_UNIT;
    f3 np = (pos - off) / GK_CUBE_UNIT;
    if (np.x < 0 || np.y < 0 || np.z < 0 || np.x > GK_CUBE_SIZE_XY - 1 || np.y > GK_CUBE_SIZE_Z - 1 || np.z > GK_CUBE_SIZE_XY - 1) return c4(0, 0, 0, 1);
    int bx = (int)floorf(np.x), by = (int)floorf(np.y), bz = (int)floorf(np.z);
    f3 fr(np.x - floorf(np.x), np.y - floorf(np.y), np.z - floorf(np.z));
    float total = 0;
    c4 result;
    for (int i = 0; i < 8; ++i) {
        int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
        int idx = (by + oy) * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + (bz + oz) * GK_CUBE_SIZE_XY + (bx + ox);
        const GkVoxelData& vx = voxels[idx];
        uint32_t p0 = vx.distanceToSolid_gg_z01, p1 = vx.distanceToSolid_x01_y01;
        if (float((p0 >> 8) & 0xFF) / 255.0f < 0.01f) continue;
        float dPZ = float((p0 >> 16) & 0xFF) / 255.0f, dNZ = float((p0 >> 24) & 0xFF) / 255.0f;
        float dPX = float(p1 & 0xFF) / 255.0f, dNX = float((p1 >> 8) & 0xFF) / 255.0f, dPY = float((p1 >> 16) & 0xFF) / 255.0f, dNY = float((p1 >> 24) & 0xFF) / 255.0f;
        f3 ptl = fr - f3((float)ox, (float)oy, (float)oz);
        float dist = length(ptl);
        f3 dir = normalize(ptl);
        float hitLen = sqrtf(fmaxf_(dir.x, 0.f)) * dPX + sqrtf(fmaxf_(-dir.x, 0.f)) * dNX + sqrtf(fmaxf_(dir.y, 0.f)) * dPY + sqrtf(fmaxf_(-dir.y, 0.f)) * dNY +
                       sqrtf(fmaxf_(dir.z, 0.f)) * dPZ + sqrtf(fmaxf_(-dir.z, 0.f)) * dNZ;
        if (dist > hitLen + 0.05f) continue;
        float wx = ox == 0 ? (1.0f - fr.x) : fr.x, wy = oy == 0 ? (1.0f - fr.y) : fr.y, wz = oz == 0 ? (1.0f - fr.z) : fr.z;
        float w = wx * wy * wz;
        result = result + sampleCubeDI(cubes[idx], normal) * w;
        total += w;
    }
    return total > 0.0f ? result / total : c4(0, 0, 0, 0);
}

// FHardwareRayTracer for the baker.  RayQuery measures t in units of the given direction; Scene::trace normalises (tinybvh),
// so tmin / tmax are scaled by |direction| and the hit distance is scaled back.
struct BakeTracer {
    Ctx& c;
    bool traceRay(f3 ro, f3 rd, float maxDistance, Vtx& out) // Shading.slang:708-750
    {
        const float len = length(rd);
        Hit h;
        if (!c.S.trace(ro, rd, kEps * len, maxDistance * len, h)) return false;
        const GkNodeProxy& px = c.S.nodes[h.inst];
        const Model& M = c.S.models[px.modelId / 10];
        UnpackedV v0 = unpackVertex(M.gpuVerts[M.indices[h.prim * 3]]), v1 = unpackVertex(M.gpuVerts[M.indices[h.prim * 3 + 1]]), v2 = unpackVertex(M.gpuVerts[M.indices[h.prim * 3 + 2]]);
        f3 n = v0.N + (v1.N - v0.N) * h.u + (v2.N - v0.N) * h.v;
        float inv[16], T[16];
        for (int r = 0; r < 4; ++r)
            for (int cc = 0; cc < 4; ++cc) T[r * 4 + cc] = px.worldTS[cc * 4 + r];
        for (int k = 0; k < 16; ++k) inv[k] = (k % 5 == 0) ? 1.f : 0.f;
        invert4x4RowMajor(T, inv);
        f3 nw(inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z, inv[2] * n.x + inv[6] * n.y + inv[10] * n.z);
        out.Normal = normalize(nw);
        out.Position = ro + rd * (h.t / len);
        out.MaterialIndex = px.matId[v0.mat & 15];
        return true;
    }
    bool traceOcclusion(f3 ro, f3 rd) // :661-681
    {
        const float len = length(rd);
        return c.S.anyHit(ro, rd, kEps * len, kMaxTrace * len);
    }
    bool traceSegment(f3 ro, f3 target, float epsilon) // :683-706
    {
        f3 dir = target - ro;
        float len = length(dir);
        f3 d = dir / len;
        float dl = length(d);
        return c.S.anyHit(ro, d, epsilon * dl, (len - epsilon) * dl);
    }
};

const f2 kGrid3x3[9] = {f2(-0.667f, -0.667f), f2(0.0f, -0.667f), f2(0.667f, -0.667f), f2(-0.667f, 0.0f), f2(0.0f, 0.0f), f2(0.667f, 0.0f), f2(-0.667f, 0.667f), f2(0.0f, 0.667f), f2(0.667f, 0.667f)};
const f2 kGrid4x4[16] = {f2(-0.75f, -0.75f), f2(-0.25f, -0.75f), f2(0.25f, -0.75f), f2(0.75f, -0.75f), f2(-0.75f, -0.25f), f2(-0.25f, -0.25f), f2(0.25f, -0.25f), f2(0.75f, -0.25f),
                         f2(-0.75f, 0.25f),  f2(-0.25f, 0.25f),  f2(0.25f, 0.25f),  f2(0.75f, 0.25f),  f2(-0.75f, 0.75f),  f2(-0.25f, 0.75f),  f2(0.25f, 0.75f),  f2(0.75f, 0.75f)};

// FaceTask, AmbientCube.slang:459-532
void faceTask(BakeTracer& T, const GkAmbientCube* cubesPrev, const GkVoxelData* voxelsPrev, f3 origin, f3 basis, uint32_t iterate, uint32_t& directLight, uint32_t& indirectLight,
              uint32_t& skyVisOut, uint32_t& sunVisOut)
{
    const GkUniformBufferObject& U = T.c.U;
    const Scene& S = T.c.S;
    origin = origin + basis * GK_CUBE_UNIT * 0.25f;
    c4 directColor, bounceColor;
    float skyVisibility = 0.0f;
    const f2 jit = kGrid3x3[iterate % 9];
    const float offx = jit.x * 0.25f, offy = jit.y * 0.25f;
    for (uint32_t i = 0; i < 16; ++i) {
        f3 hemiVec = normalize(f3(kGrid4x4[i].x + offx, kGrid4x4[i].y + offy, 1.0f));
        f3 rayDir = alignWithNormal(hemiVec, basis);
        Vtx hv;
        if (T.traceRay(origin, rayDir, 20.f /* FAST_MAX_TRACE_DISTANCE */, hv)) {
            const GkMaterial& hm = S.materials[hv.MaterialIndex];
            c4 albedo(hm.Diffuse[0], hm.Diffuse[1], hm.Diffuse[2], hm.Diffuse[3]);
            bounceColor = bounceColor + albedo * interpolateDI(cubesPrev, voxelsPrev, hv.Position, hv.Normal) * 1.25f;
        } else {
            float k = U.HasSky ? U.SkyIntensity : 0.0f;
            directColor = directColor + c4(U.BackGroundColor[0], U.BackGroundColor[1], U.BackGroundColor[2], 1.0f) * k;
            skyVisibility += 1.0f;
        }
    }
    directColor = directColor / 16.0f;
    bounceColor = bounceColor / 16.0f;
    if (U.LightCount > 0) {
        const GkLightObject& L = S.lights[0];
        const GkMaterial& lm = S.materials[L.lightMatIdx];
        c4 lightPower(lm.Diffuse[0], lm.Diffuse[1], lm.Diffuse[2], lm.Diffuse[3]);
        f3 p1(L.p1[0], L.p1[1], L.p1[2]), p3(L.p3[0], L.p3[1], L.p3[2]);
        f3 lightPos(mixf(p1.x, p3.x, 0.5f), mixf(p1.y, p3.y, 0.5f), mixf(p1.z, p3.z, 0.5f));
        float lightAtten = T.traceSegment(origin, lightPos, GK_CUBE_UNIT * 0.5f) ? 0.0f : 1.0f;
        f3 lightDir = normalize(lightPos - origin);
        float ndotl = clampf(dot(basis, lightDir), 0.0f, 1.0f);
        float distance = length(lightPos - origin);
        float attenuation = ndotl * L.normal_area[3] / (distance * distance * 3.14159f);
        directColor = directColor + lightPower * attenuation * lightAtten;
    }
    if (U.HasSun) {
        f3 sunDir(U.SunDirection[0], U.SunDirection[1], U.SunDirection[2]);
        float sunAtten = T.traceOcclusion(origin, sunDir) ? 0.0f : 1.0f;
        float ndotl = clampf(dot(basis, sunDir), 0.0f, 1.0f);
        sunVisOut = sunAtten > 0.0f ? 1u : 0u;
        directColor = directColor + c4(U.SunColor[0], U.SunColor[1], U.SunColor[2], U.SunColor[3]) * sunAtten * ndotl * (U.HasSun ? 1.0f : 0.0f) * 0.25f;
    }
    const float currWeight = 0.125f;
    skyVisOut = (uint32_t)mixf((float)skyVisOut, 255.0f * skyVisibility / 16.0f, currWeight);
    directLight = lerpPackedColorAlt(directLight, directColor, currWeight);
    indirectLight = lerpPackedColorAlt(indirectLight, bounceColor, currWeight);
}

bool insideGeometry(BakeTracer& T, f3 origin, f3 rayDir, uint32_t& outMaterialId, float& outDistanceToSolid) // AmbientCube.slang:547-571
{
    Vtx hv;
    if (T.traceRay(origin, rayDir, GK_CUBE_UNIT * 64, hv)) {
        float hitDist = length(hv.Position - origin);
        outDistanceToSolid = hitDist;
        if (outDistanceToSolid <= GK_CUBE_UNIT) {
            const GkMaterial& hm = T.c.S.materials[hv.MaterialIndex];
            outMaterialId = hv.MaterialIndex;
            if (dot(hv.Normal, rayDir) > 0.0f || ((hm.MaterialModel == GK_MAT_DIFFUSE_LIGHT) && hitDist < 0.02f)) {
                outDistanceToSolid = 0;
                return true;
            }
        }
    }
    return false;
}
float detectDistance(BakeTracer& T, f3 origin, f3 rayDir) // AmbientCube.slang:534-545
{
    Vtx hv;
    if (T.traceRay(origin, rayDir, GK_CUBE_UNIT * 64, hv)) return length(hv.Position - origin);
    return 255.f;
}
inline uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return (a & 0xFF) | ((b & 0xFF) << 8) | ((c & 0xFF) << 16) | ((d & 0xFF) << 24); }

} // namespace

// Bake.HwAmbientCube main + FGpuProbeGenerator::Render (AmbientCube.slang:574-629) for probes [first, first + count)
void bakeProbes(const Scene& S, const GkUniformBufferObject& U, GkAmbientCube* cubes, GkVoxelData* voxels, uint32_t first, uint32_t count, int threads)
{
    const size_t total = (size_t)GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z;
    std::vector<GkAmbientCube> cubesPrev(cubes, cubes + total);
    std::vector<GkVoxelData> voxelsPrev(voxels, voxels + total);
    auto work = [&](uint32_t lo, uint32_t hi) {
        Ctx ctx{S, U, nullptr, nullptr, 0};
        BakeTracer T{ctx};
        for (uint32_t gIdx = lo; gIdx < hi; ++gIdx) {
            const uint32_t y = gIdx / (GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY);
            const uint32_t z = (gIdx - y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY) / GK_CUBE_SIZE_XY;
            const uint32_t x = gIdx - y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY - z * GK_CUBE_SIZE_XY;
            const f3 cubeOffset = f3(-float(GK_CUBE_SIZE_XY / 2), -1.375f, -float(GK_CUBE_SIZE_XY / 2)) * GK_CUBE_UNIT;
            const f3 origin = f3((float)x, (float)y, (float)z) * GK_CUBE_UNIT + cubeOffset;
            GkVoxelData& vox = voxels[gIdx];
            GkAmbientCube& cube = cubes[gIdx];
            vox.matId = 0;
            float distPY = 255.0f, distNY = 255.0f, distPX = 255.0f, distNX = 255.0f, distPZ = 255.0f, distNZ = 255.0f;
            insideGeometry(T, origin, f3(0, 1, 0), vox.matId, distPY);
            insideGeometry(T, origin, f3(0, -1, 0), vox.matId, distNY);
            insideGeometry(T, origin, f3(1, 0, 0), vox.matId, distPX);
            insideGeometry(T, origin, f3(-1, 0, 0), vox.matId, distNX);
            insideGeometry(T, origin, f3(0, 0, 1), vox.matId, distPZ);
            insideGeometry(T, origin, f3(0, 0, -1), vox.matId, distNZ);
            float minDist = fminf_(fminf_(fminf_(distPY, distNY), fminf_(distPX, distNX)), fminf_(distPZ, distNZ));
            if (minDist > 254.0f) {
                const f3 diag[8] = {f3(1, 1, 1), f3(-1, 1, 1), f3(-1, -1, 1), f3(-1, 1, 1), f3(1, 1, -1), f3(-1, 1, -1), f3(-1, -1, -1), f3(-1, 1, -1)};
                for (int k = 0; k < 8; ++k) minDist = fminf_(minDist, detectDistance(T, origin, diag[k]));
            }
            distPY = clampf(distPY * 4.0f, 0.f, 1.f), distNY = clampf(distNY * 4.0f, 0.f, 1.f), distPX = clampf(distPX * 4.0f, 0.f, 1.f);
            distNX = clampf(distNX * 4.0f, 0.f, 1.f), distPZ = clampf(distPZ * 4.0f, 0.f, 1.f), distNZ = clampf(distNZ * 4.0f, 0.f, 1.f);
            const float inside = distPY * distNY * distPX * distNX * distPZ * distNZ;
            vox.distanceToSolid_gg_z01 = pack4((uint32_t)(minDist / GK_CUBE_UNIT), (uint32_t)(inside * 255.0f), (uint32_t)(distPZ * 255.0f), (uint32_t)(distNZ * 255.0f));
            vox.distanceToSolid_x01_y01 = pack4((uint32_t)(distPX * 255.0f), (uint32_t)(distNX * 255.0f), (uint32_t)(distPY * 255.0f), (uint32_t)(distNY * 255.0f));
            if (minDist < 4) {
                const uint32_t iterate = vox.age;
                vox.age = vox.age + 1;
                uint32_t sv0[4] = {cube.skyVisibility_pznzpyny & 0xFF, (cube.skyVisibility_pznzpyny >> 8) & 0xFF, (cube.skyVisibility_pznzpyny >> 16) & 0xFF, (cube.skyVisibility_pznzpyny >> 24) & 0xFF};
                uint32_t sv1[4] = {cube.skyVisibility_pxnxs0s1 & 0xFF, (cube.skyVisibility_pxnxs0s1 >> 8) & 0xFF, (cube.skyVisibility_pxnxs0s1 >> 16) & 0xFF, (cube.skyVisibility_pxnxs0s1 >> 24) & 0xFF};
                uint32_t sunvis = 0;
                faceTask(T, cubesPrev.data(), voxelsPrev.data(), origin, f3(0, 1, 0), iterate, cube.PosY_D, cube.PosY, sv0[2], sv1[2]);
                faceTask(T, cubesPrev.data(), voxelsPrev.data(), origin, f3(0, -1, 0), iterate, cube.NegY_D, cube.NegY, sv0[3], sunvis);
                faceTask(T, cubesPrev.data(), voxelsPrev.data(), origin, f3(1, 0, 0), iterate, cube.PosX_D, cube.PosX, sv1[0], sunvis);
                faceTask(T, cubesPrev.data(), voxelsPrev.data(), origin, f3(-1, 0, 0), iterate, cube.NegX_D, cube.NegX, sv1[1], sunvis);
                faceTask(T, cubesPrev.data(), voxelsPrev.data(), origin, f3(0, 0, 1), iterate, cube.PosZ_D, cube.PosZ, sv0[0], sunvis);
                faceTask(T, cubesPrev.data(), voxelsPrev.data(), origin, f3(0, 0, -1), iterate, cube.NegZ_D, cube.NegZ, sv0[1], sunvis);
                cube.skyVisibility_pznzpyny = pack4(sv0[0], sv0[1], sv0[2], sv0[3]);
                cube.skyVisibility_pxnxs0s1 = pack4(sv1[0], sv1[1], sv1[2], sv1[3]);
            }
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        const uint32_t lo = first + (uint32_t)((uint64_t)count * t / threads), hi = first + (uint32_t)((uint64_t)count * (t + 1) / threads);
        pool.emplace_back(work, lo, hi);
    }
    for (auto& th : pool) th.join();
}

// Task.RayCast.comp.slang:31-55 — the GPU ray-cast task, one in-place RayCastIO record per ray:
// FHardwareRayTracer::TraceRay(Origin, Direction, 10000) (Shading.slang:708-750, tmin = EPS), HitPoint = Origin + Direction * t,
// interpolated world-space normal, T = |HitPoint - Origin|, InstanceId = node instance id, MaterialId = the node's material of
// the triangle's slot; a miss only clears Hitted.
void rayCastTask(const Scene& S, GkRayCastIO* io, uint32_t n)
{
    GkUniformBufferObject U{};
    Ctx c{S, U, nullptr, nullptr, 0};
    for (uint32_t i = 0; i < n; ++i) {
        GkRayCastIO& R = io[i];
        const f3 ro(R.Context.Origin[0], R.Context.Origin[1], R.Context.Origin[2]), rd(R.Context.Direction[0], R.Context.Direction[1], R.Context.Direction[2]);
        Vtx v;
        uint32_t node = 0;
        if (traceRay(c, ro, rd, 10000.0f, v, node)) {
            R.Result.HitPoint[0] = v.Position.x, R.Result.HitPoint[1] = v.Position.y, R.Result.HitPoint[2] = v.Position.z, R.Result.HitPoint[3] = 1.0f;
            R.Result.Normal[0] = v.Normal.x, R.Result.Normal[1] = v.Normal.y, R.Result.Normal[2] = v.Normal.z, R.Result.Normal[3] = 0.0f;
            R.Result.Hitted = 1;
            R.Result.T = length(v.Position - ro);
            R.Result.InstanceId = S.nodes[node].instanceId;
            R.Result.MaterialId = v.MaterialIndex;
        } else R.Result.Hitted = 0;
    }
}

} // namespace orc
