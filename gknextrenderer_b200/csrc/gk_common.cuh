// gk_common.cuh — shared device/host helpers of the sm_100a backend.
//
// Arithmetic policy.  Everything that decides WHICH triangle a ray hits — direction
// normalisation, the instance-space ray transform, the 4x4 inverse, the Möller–Trumbore
// test — is written with explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn/...)
// in the operation order of the reference's CPU query (tinybvh, see gk_traverse.cuh), so
// nvcc can never contract it into FMAs and the hit distances are bit-identical to the
// CPU's.  Box tests are free to use FMA: they are conservative and only prune.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gknext_types.h"

#define GK_HD __host__ __device__ __forceinline__
#define GK_D __device__ __forceinline__

namespace gk {

constexpr float kFar = 1e30f;      // BVH_FAR (tiny_bvh.h:129)
constexpr float kEps = 1e-3f;      // EPS, PreProcessor.slang:13
constexpr float kMaxTrace = 1000.f;   // PT_MAX_TRACE_DISTANCE, Shading.slang:15
constexpr float kTraceOffset = 0.001f; // TRACE_CORRECTION_OFFSET, Shading.slang:18
constexpr float kPrimaryTMax = 2000.f; // RayCastInCPU, CPUAccelerationStructure.cpp:287
constexpr uint32_t kInvalid = 0xffffffffu;

struct f3 {
    float x, y, z;
};
GK_HD f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
GK_HD f3 mk3(float a) { return f3{a, a, a}; }

// exact (never contracted) scalar ops
#ifdef __CUDA_ARCH__
GK_D float xmul(float a, float b) { return __fmul_rn(a, b); }
GK_D float xadd(float a, float b) { return __fadd_rn(a, b); }
GK_D float xsub(float a, float b) { return __fsub_rn(a, b); }
GK_D float xdiv(float a, float b) { return __fdiv_rn(a, b); }
GK_D float xsqrt(float a) { return __fsqrt_rn(a); }
GK_D float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
inline float xmul(float a, float b) { return a * b; }
inline float xadd(float a, float b) { return a + b; }
inline float xsub(float a, float b) { return a - b; }
inline float xdiv(float a, float b) { return a / b; }
inline float xsqrt(float a) { return sqrtf(a); }
inline float xfma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

// exact vector helpers (plain left-to-right fp32, one rounding per operation)
GK_HD f3 xadd3(f3 a, f3 b) { return mk3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)); }
GK_HD f3 xsub3(f3 a, f3 b) { return mk3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
GK_HD f3 xmul3(f3 a, f3 b) { return mk3(xmul(a.x, b.x), xmul(a.y, b.y), xmul(a.z, b.z)); }
GK_HD f3 xscale(f3 a, float s) { return mk3(xmul(a.x, s), xmul(a.y, s), xmul(a.z, s)); }
GK_HD f3 xneg(f3 a) { return mk3(-a.x, -a.y, -a.z); }
GK_HD float xdot(f3 a, f3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
GK_HD f3 xcross(f3 a, f3 b)
{
    return mk3(xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)), xsub(xmul(a.x, b.y), xmul(a.y, b.x)));
}
GK_HD float xlength(f3 a) { return xsqrt(xdot(a, a)); }
GK_HD f3 xnormalize(f3 a)
{
    const float rl = xdiv(1.0f, xlength(a));
    return xscale(a, rl);
}
GK_HD float fminx(float a, float b) { return a < b ? a : b; } // tinybvh_min semantics (NaN -> b)
GK_HD float fmaxx(float a, float b) { return a > b ? a : b; }
GK_HD float clampx(float v, float lo, float hi) { return fminx(fmaxx(v, lo), hi); }

GK_HD float safeRcp(float x) // tiny_bvh.h:329
{
    if (x > 1e-12f) return xdiv(1.0f, x);
    if (x < -1e-12f) return xdiv(1.0f, x);
    return kFar;
}

// column-major 4x4 times vec4, summed as (c0*x + c1*y) + (c2*z + c3*w)
struct f4 {
    float x, y, z, w;
};
GK_HD f4 mk4(float x, float y, float z, float w) { return f4{x, y, z, w}; }
GK_HD f4 xmulM(const float* M, f4 v)
{
    f4 r;
    r.x = xadd(xadd(xmul(M[0], v.x), xmul(M[4], v.y)), xadd(xmul(M[8], v.z), xmul(M[12], v.w)));
    r.y = xadd(xadd(xmul(M[1], v.x), xmul(M[5], v.y)), xadd(xmul(M[9], v.z), xmul(M[13], v.w)));
    r.z = xadd(xadd(xmul(M[2], v.x), xmul(M[6], v.y)), xadd(xmul(M[10], v.z), xmul(M[14], v.w)));
    r.w = xadd(xadd(xmul(M[3], v.x), xmul(M[7], v.y)), xadd(xmul(M[11], v.z), xmul(M[15], v.w)));
    return r;
}

// row-major 4x4 point / vector transforms, tiny_bvh.h:396-409
GK_HD f3 xformPoint(f3 v, const float* T)
{
    f3 res = mk3(xadd(xadd(xadd(xmul(T[0], v.x), xmul(T[1], v.y)), xmul(T[2], v.z)), T[3]),
                 xadd(xadd(xadd(xmul(T[4], v.x), xmul(T[5], v.y)), xmul(T[6], v.z)), T[7]),
                 xadd(xadd(xadd(xmul(T[8], v.x), xmul(T[9], v.y)), xmul(T[10], v.z)), T[11]));
    const float w = xadd(xadd(xadd(xmul(T[12], v.x), xmul(T[13], v.y)), xmul(T[14], v.z)), T[15]);
    if (w == 1) return res;
    return xscale(res, xdiv(1.f, w));
}
GK_HD f3 xformVector(f3 v, const float* T)
{
    return mk3(xadd(xadd(xmul(T[0], v.x), xmul(T[1], v.y)), xmul(T[2], v.z)), xadd(xadd(xmul(T[4], v.x), xmul(T[5], v.y)), xmul(T[6], v.z)),
               xadd(xadd(xmul(T[8], v.x), xmul(T[9], v.y)), xmul(T[10], v.z)));
}

// IEEE binary16 -> binary32 is exact on every implementation
GK_HD float halfBitsToFloat(uint16_t h)
{
#ifdef __CUDA_ARCH__
    return __half2float(__ushort_as_half(h));
#else
    uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu, o;
    if (e == 0) {
        if (m == 0) o = s;
        else {
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
#endif
}

// f32 -> f16 with glm::detail::toFloat16's rounding (round half up on the magnitude), the
// conversion Assets::MakeVertex applies (src/Assets/Vertex.hpp:83-95).
GK_HD uint16_t glmToHalf(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t i = __float_as_uint(f);
#else
    uint32_t i;
    memcpy(&i, &f, 4);
#endif
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

} // namespace gk
