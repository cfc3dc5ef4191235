// gk_shading.cuh — material / light / sampling math of the path tracer, as device functions.
//
// Follows the reference shaders (citations per function):
//   assets/shaders/common/Const_Func.slang   RNG :227-258, sampling :260-340, Schlick :8-14, ONB :17-22
//   assets/shaders/common/GeneralFunc.slang  get_material_data :33-83
//   assets/shaders/common/Shading.slang      FHardwareRayTracer hit resolve :725-747, sky :148-153
//   assets/shaders/common/AmbientCube.slang  probe read side :71-78, :178-223, :275-364
// This translation unit is compiled with -fmad=false: a*b+c stays two roundings unless fmaf
// is written out (where the shader writes mad()/fma()), so the arithmetic follows the same
// order as a plain fp32 CPU evaluation of the shader source.
#pragma once
#include "gk_bvh.cuh"

namespace gk {

// ---- small vector algebra (plain operators; contraction is disabled for this file) ----
GK_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
GK_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
GK_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
GK_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
GK_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
GK_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
GK_HD float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GK_HD f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GK_HD float length3(f3 a) { return sqrtf(dot3(a, a)); }
GK_HD f3 normalize3(f3 a)
{
    const float rl = 1.0f / length3(a);
    return mk3(a.x * rl, a.y * rl, a.z * rl);
}
GK_HD f3 fma3(float s, f3 a, f3 c) { return mk3(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z)); }
GK_HD f4 mulM(const float* M, f4 v) // column-major mat4 * vec4, (c0*x + c1*y) + (c2*z + c3*w)
{
    f4 r;
    r.x = (M[0] * v.x + M[4] * v.y) + (M[8] * v.z + M[12] * v.w);
    r.y = (M[1] * v.x + M[5] * v.y) + (M[9] * v.z + M[13] * v.w);
    r.z = (M[2] * v.x + M[6] * v.y) + (M[10] * v.z + M[14] * v.w);
    r.w = (M[3] * v.x + M[7] * v.y) + (M[11] * v.z + M[15] * v.w);
    return r;
}
GK_HD f3 xyz(f4 v) { return mk3(v.x, v.y, v.z); }

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.283185307179586476925f;
constexpr float kInvPi = 0.318309886183790671538f;
constexpr float kPiOver4 = 0.785398163397448309616f;
constexpr float kNearZero = 1e-35f;

// ---- RNG (Const_Func.slang:227-258) ----
struct u4 {
    uint32_t x, y, z, w;
};
GK_HD void pcg4d(u4& v)
{
    v.x = v.x * 1664525u + 1013904223u;
    v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u;
    v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w, v.y += v.z * v.x, v.z += v.x * v.y, v.w += v.y * v.z;
    v.x ^= v.x >> 16u, v.y ^= v.y >> 16u, v.z ^= v.z >> 16u, v.w ^= v.w >> 16u;
    v.x += v.y * v.w, v.y += v.z * v.x, v.z += v.x * v.y, v.w += v.y * v.z;
}
GK_HD float u2f(uint32_t x) { return asFloat(0x3f800000u | (x >> 9)) - 1.0f; }
GK_HD float randomFloat(u4& s)
{
    pcg4d(s);
    return u2f(s.x);
}
struct f2 {
    float x, y;
};
GK_HD f2 randomFloat2(u4& s)
{
    pcg4d(s);
    return f2{u2f(s.x), u2f(s.y)};
}

// ---- sampling ----
GK_HD float pow5(float x) { return x * x * x * x * x; }
GK_HD float schlick(float cosine, float ri)
{
    float r0 = (1 - ri) / (1 + ri);
    r0 *= r0;
    return r0 + (1 - r0) * pow5(1 - cosine);
}
GK_HD void onb(f3 n, f3& b1, f3& b2)
{
    const float signZ = n.z < 0.f ? -1.f : 1.f;
    const float a = -1.0f / (signZ + n.z);
    b2 = mk3(n.x * n.y * a, signZ + n.y * n.y * a, -n.y);
    b1 = mk3(1.0f + signZ * n.x * n.x * a, signZ * b2.x, -signZ * n.x);
}
GK_HD f3 toWorld(f3 v, f3 T, f3 B, f3 N)
{
    return mk3(v.x * T.x + v.y * B.x + v.z * N.x, v.x * T.y + v.y * B.y + v.z * N.y, v.x * T.z + v.y * B.z + v.z * N.z);
}
GK_HD f3 toLocal(f3 v, f3 t, f3 b, f3 n) { return mk3(dot3(t, v), dot3(b, v), dot3(n, v)); }
GK_HD f3 alignWithNormal(f3 ray, f3 normal)
{
    f3 T, B;
    onb(normal, T, B);
    return toWorld(ray, T, B, normal);
}
GK_HD f2 concentricDisk(f2 o)
{
    o = f2{o.x + (o.x - 1.0f), o.y + (o.y - 1.0f)};
    const bool zx = o.x > -kNearZero && o.x < kNearZero, zy = o.y > -kNearZero && o.y < kNearZero;
    if (zx && zy) return f2{0, 0};
    if (fabsf(o.x) > fabsf(o.y)) {
        const float theta = kPiOver4 * o.y / o.x;
        return f2{o.x * cosf(theta), o.x * sinf(theta)};
    }
    const float ct = sinf(kPiOver4 * o.x / o.y);
    return f2{o.y * ct, o.y * sqrtf(1.f - ct * ct)};
}
GK_HD f3 randomInCone(u4& s, float cosTheta)
{
    const f2 u = randomFloat2(s);
    const float phi = kTwoPi * u.x;
    cosTheta = 1.0f + u.y * (cosTheta - 1.f);
    const float r = sqrtf(1.0f - cosTheta * cosTheta);
    return mk3(r * cosf(phi), r * sinf(phi), cosTheta);
}
GK_HD f3 randomInHemiSphere1(u4& s)
{
    const f2 u = randomFloat2(s);
    const float phi = kTwoPi * u.x;
    const float r = sqrtf(u.y);
    return mk3(r * cosf(phi), r * sinf(phi), sqrtf(1.0f - u.y));
}
GK_HD float saturatef(float x) { return clampx(x, 0.0f, 1.0f); }
GK_HD f3 ggxSampleVndf(f2 alpha, f3 wi_, f2 uv) // Eto & Tokuyoshi 2023, Const_Func.slang:310-328
{
    const f3 wi = normalize3(mk3(wi_.x * alpha.x, wi_.y * alpha.y, wi_.z));
    float b = wi.z;
    if (wi_.z > 0.f) {
        const float a = saturatef(fminx(alpha.x, alpha.y));
        const float awiz_s = a * wi_.z / (1.0f + sqrtf(wi_.x * wi_.x + wi_.y * wi_.y));
        b *= ((1.0f - a * a) / (1.0f + awiz_s * awiz_s));
    }
    const float z = fmaf(1.0f - uv.y, 1.0f + b, -b);
    const float phi = kTwoPi * uv.x;
    const float r = sqrtf(saturatef(1.0f - z * z));
    const f3 o_std = mk3(r * cosf(phi), r * sinf(phi), z);
    const f3 m_std = wi + o_std;
    return normalize3(mk3(m_std.x * alpha.x, m_std.y * alpha.y, m_std.z));
}
GK_HD f3 ggxSampling(u4& s, float roughness, f3 normal)
{
    f3 t, b;
    onb(normal, t, b);
    const f3 wm = ggxSampleVndf(f2{roughness * roughness, roughness * roughness}, toLocal(normal, t, b, normal), randomFloat2(s));
    return toWorld(wm, t, b, normal);
}
GK_HD f3 reflect3(f3 i, f3 n) { return i - n * (2.0f * dot3(n, i)); }
GK_HD f3 refract3(f3 i, f3 n, float eta)
{
    const float d = dot3(n, i);
    const float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return mk3(0, 0, 0);
    return i * eta - n * (eta * d + sqrtf(k));
}

// ---- scene access ----
struct ShadeScene {
    const GkGPUVertex* verts;
    const uint32_t* indices;
    const ModelInfo* models;
    const GkMaterial* materials;
    const GkNodeProxy* nodes;
    const InstRecord* inst;
    const GkAmbientCube* cubes; // may be null
    const GkVoxelData* voxels;  // may be null
    uint32_t materialCount;
};

struct UnpackedV {
    f3 P, N;
    f2 uv;
    uint32_t mat;
};
GK_HD UnpackedV unpackVertex(const GkGPUVertex* verts, uint32_t index) // Const_Func.slang:342-354
{
    // 24-byte record read as three 8-byte words
    const uint2* p = reinterpret_cast<const uint2*>(verts + index);
#ifdef __CUDA_ARCH__
    const uint2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
#else
    const uint2 a = p[0], b = p[1], c = p[2];
#endif
    UnpackedV v;
    v.P = mk3(halfBitsToFloat((uint16_t)(a.x & 0xffff)), halfBitsToFloat((uint16_t)(a.x >> 16)), halfBitsToFloat((uint16_t)(a.y & 0xffff)));
    v.N = mk3(halfBitsToFloat((uint16_t)(b.x & 0xffff)), halfBitsToFloat((uint16_t)(b.x >> 16)), halfBitsToFloat((uint16_t)(b.y & 0xffff)));
    v.uv = f2{halfBitsToFloat((uint16_t)(a.y >> 16)), halfBitsToFloat((uint16_t)(b.y >> 16))};
    v.mat = (c.y >> 16) & 0xFFu;
    return v;
}

struct Vtx {
    f3 Position, Normal;
    f2 TexCoord;
    uint32_t MaterialIndex;
};

// GeneralFunc.slang:33-83 — shading vertex of a primary hit from its visibility id.
GK_HD Vtx getMaterialData(const ShadeScene& S, uint32_t node, uint32_t prim, f3 ro, f3 rd, uint32_t& rawMat)
{
    const GkNodeProxy& px = S.nodes[node];
    const InstRecord& M = S.inst[node];
    const float* W = px.worldTS;
    f3 P[3], N[3];
    f2 T[3];
    uint32_t matid = 0;
    for (int i = 0; i < 3; ++i) {
        const UnpackedV v = unpackVertex(S.verts, M.vertexOffset + S.indices[M.indexOffset + prim * 3 + i]);
        P[i] = xyz(mulM(W, mk4(v.P.x, v.P.y, v.P.z, 1)));
        N[i] = xyz(mulM(W, mk4(v.N.x, v.N.y, v.N.z, 0)));
        T[i] = v.uv;
        if (i == 0) matid = v.mat;
    }
    const f3 e0 = P[1] - P[0], e1 = P[2] - P[0];
    const f3 rce1 = cross3(rd, e1);
    const float rcpDet = 1.0f / dot3(e0, rce1);
    const f3 r0 = ro - P[0];
    const float by = rcpDet * dot3(r0, rce1);
    const f3 e0c0 = cross3(e0, r0);
    const float bz = -rcpDet * dot3(rd, e0c0);
    const float bx = 1.0f - (by + bz);
    Vtx r;
    r.Position = fma3(bx, P[0], fma3(by, P[1], P[2] * bz));
    r.Normal = normalize3(fma3(bx, N[0], fma3(by, N[1], N[2] * bz)));
    r.TexCoord = f2{fmaf(bx, T[0].x, fmaf(by, T[1].x, bz * T[2].x)), fmaf(bx, T[0].y, fmaf(by, T[1].y, bz * T[2].y))};
    r.MaterialIndex = matid;
    rawMat = matid;
    return r;
}

// Shading.slang:725-747 — shading vertex of a traced hit.
GK_HD void resolveHit(const ShadeScene& S, f3 ro, f3 rd, float t, float u, float v, uint32_t prim, uint32_t node, Vtx& out)
{
    const InstRecord& I = S.inst[node];
    const uint32_t* idx = S.indices + I.indexOffset + prim * 3;
    const UnpackedV v0 = unpackVertex(S.verts, I.vertexOffset + idx[0]);
    const UnpackedV v1 = unpackVertex(S.verts, I.vertexOffset + idx[1]);
    const UnpackedV v2 = unpackVertex(S.verts, I.vertexOffset + idx[2]);
    const f3 n = v0.N + (v1.N - v0.N) * u + (v2.N - v0.N) * v;
    const float* inv = I.invT; // row-major inverse world transform
    const f3 nw = mk3(inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z, inv[2] * n.x + inv[6] * n.y + inv[10] * n.z);
    out.Normal = normalize3(nw);
    out.TexCoord = f2{v0.uv.x + (v1.uv.x - v0.uv.x) * u + (v2.uv.x - v0.uv.x) * v, v0.uv.y + (v1.uv.y - v0.uv.y) * u + (v2.uv.y - v0.uv.y) * v};
    out.Position = ro + rd * t;
    out.MaterialIndex = S.nodes[node].matId[v0.mat & 15];
}

GK_HD f3 skyColor(const GkUniformBufferObject& U) // Shading.slang:148-153 with a constant texel
{
    if (!U.HasSky) return mk3(0, 0, 0);
    return mk3(fminx(10.f, U.BackGroundColor[0]), fminx(10.f, U.BackGroundColor[1]), fminx(10.f, U.BackGroundColor[2])) * U.SkyIntensity;
}

GK_HD f3 unpackRGB10(uint32_t p)
{
    return mk3(float(p & 0x3FF) / 1023.0f, float((p >> 10) & 0x3FF) / 1023.0f, float((p >> 20) & 0x3FF) / 1023.0f) * 512.f;
}

GK_HD f3 sampleCubeFull(const GkAmbientCube& cb, f3 n)
{
    const float wx = fmaxx(n.x, 0.f), wnx = fmaxx(-n.x, 0.f), wy = fmaxx(n.y, 0.f), wny = fmaxx(-n.y, 0.f), wz = fmaxx(n.z, 0.f), wnz = fmaxx(-n.z, 0.f);
    const float sum = wx + wnx + wy + wny + wz + wnz;
    f3 col = mk3(0, 0, 0);
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
    return col * ((sum > 0.0f) ? (1.0f / sum) : 1.0f);
}

GK_HD f3 interpolateAmbientCubes(const ShadeScene& S, f3 pos, f3 normal)
{
    const f3 off = mk3(-float(GK_CUBE_SIZE_XY / 2), -1.375f, -float(GK_CUBE_SIZE_XY / 2)) * GK_CUBE_UNIT;
    const f3 np = (pos - off) / GK_CUBE_UNIT;
    if (np.x < 0 || np.y < 0 || np.z < 0 || np.x > GK_CUBE_SIZE_XY - 1 || np.y > GK_CUBE_SIZE_Z - 1 || np.z > GK_CUBE_SIZE_XY - 1) return mk3(0, 0, 0);
    if (!S.cubes || !S.voxels) return mk3(0, 0, 0);
    const int bx = (int)floorf(np.x), by = (int)floorf(np.y), bz = (int)floorf(np.z);
    const f3 fr = mk3(np.x - floorf(np.x), np.y - floorf(np.y), np.z - floorf(np.z));
    float total = 0;
    f3 result = mk3(0, 0, 0);
    for (int i = 0; i < 8; ++i) {
        const int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
        const int idx = (by + oy) * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + (bz + oz) * GK_CUBE_SIZE_XY + (bx + ox);
        const GkVoxelData vx = S.voxels[idx];
        const uint32_t p0 = vx.distanceToSolid_gg_z01, p1 = vx.distanceToSolid_x01_y01;
        const float d0y = float((p0 >> 8) & 0xFF) / 255.0f;
        if (d0y < 0.01f) continue;
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
        result = result + sampleCubeFull(S.cubes[idx], normal) * w;
        total += w;
    }
    return total > 0.0f ? result / total : mk3(0, 0, 0);
}

} // namespace gk
