// orc_math.h — tiny fp32 vector helpers for the CPU oracle.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it.
//
// Every operation is plain IEEE fp32 in source order.  The oracle is compiled
// with -ffp-contract=off so that no multiply-add is fused: the hit ids the
// oracle returns then depend only on the arithmetic written here.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct f3 {
    float x, y, z;
    f3() : x(0), y(0), z(0) {}
    f3(float a) : x(a), y(a), z(a) {}
    f3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct f4 {
    float x, y, z, w;
    f4() : x(0), y(0), z(0), w(0) {}
    f4(float a) : x(a), y(a), z(a), w(a) {}
    f4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    f4(f3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    f3 xyz() const { return f3(x, y, z); }
};
struct f2 {
    float x, y;
    f2() : x(0), y(0) {}
    f2(float a, float b) : x(a), y(b) {}
};

inline f3 operator+(f3 a, f3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline f3 operator-(f3 a, f3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline f3 operator*(f3 a, f3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline f3 operator*(f3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
inline f3 operator*(float s, f3 a) { return f3(a.x * s, a.y * s, a.z * s); }
inline f3 operator/(f3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
inline f3 operator-(f3 a) { return f3(-a.x, -a.y, -a.z); }
inline f4 operator+(f4 a, f4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline f4 operator-(f4 a, f4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline f4 operator*(f4 a, f4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline f4 operator*(f4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline f4 operator/(f4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline f2 operator+(f2 a, f2 b) { return f2(a.x + b.x, a.y + b.y); }
inline f2 operator-(f2 a, f2 b) { return f2(a.x - b.x, a.y - b.y); }
inline f2 operator*(f2 a, float s) { return f2(a.x * s, a.y * s); }

inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(f2 a, f2 b) { return a.x * b.x + a.y * b.y; }
inline f3 cross(f3 a, f3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(f3 a) { return sqrtf(dot(a, a)); }
inline float length(f2 a) { return sqrtf(dot(a, a)); }
// normalize: the shader compiler is free to emit v * rsqrt(dot(v,v)); we *define* it as
// v * (1 / sqrt(dot(v,v))) with IEEE sqrt and divide, the same source order the CUDA
// kernels use (and the order tinybvh_normalize uses, tiny_bvh.h:391-395).
inline f3 normalize(f3 a) { float l = length(a); float rl = 1.0f / l; return f3(a.x * rl, a.y * rl, a.z * rl); }
inline f3 vmin(f3 a, f3 b) { return f3(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z); }
inline f3 vmax(f3 a, f3 b) { return f3(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z); }
inline float fminf_(float a, float b) { return a < b ? a : b; }
inline float fmaxf_(float a, float b) { return a > b ? a : b; }
inline float clampf(float v, float lo, float hi) { return fminf_(fmaxf_(v, lo), hi); }
inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }
inline f3 lerp3(f3 a, f3 b, float t) { return a + (b - a) * t; }

// Column-major 4x4 (glm layout): element(row r, col c) = m[c*4+r].
struct m4 {
    float m[16];
};
inline f4 mul(const m4& M, f4 v)
{
    // glm mat4 * vec4: sum of columns scaled, evaluated as (c0*x + c1*y) + (c2*z + c3*w)
    // (glm/detail/type_mat4x4.inl operator*); Slang mul(M, v) on the GPU has no
    // defined association, so this order is *our* definition, mirrored in CUDA.
    f4 r;
    r.x = (M.m[0] * v.x + M.m[4] * v.y) + (M.m[8] * v.z + M.m[12] * v.w);
    r.y = (M.m[1] * v.x + M.m[5] * v.y) + (M.m[9] * v.z + M.m[13] * v.w);
    r.z = (M.m[2] * v.x + M.m[6] * v.y) + (M.m[10] * v.z + M.m[14] * v.w);
    r.w = (M.m[3] * v.x + M.m[7] * v.y) + (M.m[11] * v.z + M.m[15] * v.w);
    return r;
}
inline m4 matmul(const m4& A, const m4& B)
{
    m4 R;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s += A.m[k * 4 + r] * B.m[c * 4 + k];
            R.m[c * 4 + r] = s;
        }
    return R;
}

// IEEE binary16 helpers.  f32 -> f16 follows glm::detail::toFloat16
// (glm/detail/type_half.inl; un-vendored dependency of the reference, used at
// src/Assets/Vertex.hpp:83-95): round-half-up on the magnitude, flush below 2^-25.
inline uint16_t glm_to_half(float f)
{
    uint32_t i;
    memcpy(&i, &f, 4);
    int s = (int)((i >> 16) & 0x8000u);
    int e = (int)((i >> 23) & 0xffu) - (127 - 15);
    int m = (int)(i & 0x007fffffu);
    if (e <= 0) {
        if (e < -10) return (uint16_t)s;
        m = (m | 0x00800000) >> (1 - e);
        if (m & 0x00001000) m += 0x00002000;
        return (uint16_t)(s | (m >> 13));
    } else if (e == 0xff - (127 - 15)) {
        if (m == 0) return (uint16_t)(s | 0x7c00);
        m >>= 13;
        return (uint16_t)(s | 0x7c00 | m | (m == 0));
    } else {
        if (m & 0x00001000) {
            m += 0x00002000;
            if (m & 0x00800000) { m = 0; e += 1; }
        }
        if (e > 30) return (uint16_t)(s | 0x7c00);
        return (uint16_t)(s | (e << 10) | (m >> 13));
    }
}
inline float half_to_float(uint16_t h)
{
    uint32_t s = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t o;
    if (e == 0) {
        if (m == 0) o = s;
        else {
            // subnormal: normalise
            int sh = 0;
            while (!(m & 0x400u)) { m <<= 1; ++sh; }
            m &= 0x3ffu;
            o = s | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13);
        }
    } else if (e == 31) o = s | 0x7f800000u | (m << 13);
    else o = s | ((e + (127 - 15)) << 23) | (m << 13);
    float f;
    memcpy(&f, &o, 4);
    return f;
}
// Storage rounding for RGBA16F images (round-to-nearest-even, what image stores do).
inline uint16_t float_to_half_rne(float f)
{
    _Float16 h = (_Float16)f;
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
inline float quant_half(float f) { return half_to_float(float_to_half_rne(f)); }

} // namespace orc
