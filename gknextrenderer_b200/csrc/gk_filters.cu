// gk_filters.cu — temporal reprojection, joint-bilateral denoise + compose, history hand-over.
//
// Replaces
//   Process.ReProject.comp.slang :60-181, dispatched three times per frame for diffuse, specular
//     and albedo with push constants {needClamp, needSpatio} = {0,1},{0,1},{1,0}
//     (src/Rendering/PathTracing/PathTracingRenderer.cpp:117-144)           -> k_reproject (ONE launch)
//   Process.DenoiseJBF.comp.slang :96-195 (PathTracingRenderer.cpp:146-169)  -> k_denoise_jbf
//   the three vkCmdCopyImage "copy pass" calls (:171-195) and the ObjectId0 -> ObjectId1 copy
//     (src/Rendering/VulkanBaseRenderer.cpp:1287-1303)                       -> pointer swaps, 0 bytes
//
// Both kernels are HBM-bound streaming kernels.  A block owns a 32x8 pixel tile and stages the
// halo it needs in shared memory with coalesced 8-byte (RGBA16F) row reads:
//   reproject: source colour of the three channel sets, object ids and normals with a 2-pixel halo
//              (the 5x5 spatial fallback and the 5x5 YCoCg clamp read the same 25 texels);
//              history is gathered through L2 at the motion-vector target (4 taps).
//   denoise  : diffuse with a 5-pixel halo (taps at -5,-3,..,5); spec/albedo/ids are centre-only.
// Algorithmic bytes per pixel: reproject 96 B, denoise 48 B.
// Images outside [0,W)x[0,H) read as 0 and are never written (Vulkan robust image access).
// lerp(a,b,t) = a*(1-t) + b*t.
#include "gk_context.h"
#include <cuda_fp16.h>

namespace gk {

namespace {

constexpr int TX = 32, TY = 8;

// Row ownership of a tile-partitioned frame (GkConfig): count == 1 means "every row".
struct RowTiles {
    uint32_t index, count, rows;
    __device__ __forceinline__ bool owns(int y) const { return count <= 1u || ((uint32_t)y / rows) % count == index; }
    __device__ __forceinline__ bool ownsAny(int y0, int n) const
    {
        if (count <= 1u) return true;
        for (int i = 0; i < n; ++i)
            if (owns(y0 + i)) return true;
        return false;
    }
};

struct c3 {
    float x, y, z;
};
__device__ __forceinline__ c3 mk(float x, float y, float z) { return c3{x, y, z}; }
__device__ __forceinline__ c3 operator+(c3 a, c3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ c3 operator*(c3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ c3 operator*(c3 a, c3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ c3 operator/(c3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ c3 mix3(c3 a, c3 b, float t) { return mk(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
__device__ __forceinline__ c3 min3(c3 a, c3 b) { return mk(fminx(a.x, b.x), fminx(a.y, b.y), fminx(a.z, b.z)); }
__device__ __forceinline__ c3 max3(c3 a, c3 b) { return mk(fmaxx(a.x, b.x), fmaxx(a.y, b.y), fmaxx(a.z, b.z)); }
__device__ __forceinline__ c3 clamp3(c3 v, c3 lo, c3 hi) { return min3(max3(v, lo), hi); }
__device__ __forceinline__ c3 rgb2ycocg(c3 c) { return mk(0.25f * c.x + 0.5f * c.y + 0.25f * c.z, 0.5f * c.x - 0.5f * c.z, -0.25f * c.x + 0.5f * c.y - 0.25f * c.z); }
__device__ __forceinline__ c3 ycocg2rgb(c3 c) { return mk(c.x + c.y - c.z, c.x + c.z, c.x - c.y - c.z); }

__device__ __forceinline__ uint2 loadPx(const uint2* img, int x, int y, int W, int H)
{
    if (x < 0 || y < 0 || x >= W || y >= H) return make_uint2(0u, 0u);
    return __ldg(img + (size_t)y * W + x);
}
__device__ __forceinline__ c3 unpackRgb(uint2 v)
{
    const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return mk(fa.x, fa.y, fb.x);
}
__device__ __forceinline__ uint2 packRgba(c3 c, float a)
{
    const __half2 lo = __floats2half2_rn(c.x, c.y), hi = __floats2half2_rn(c.z, a);
    uint2 v;
    v.x = *reinterpret_cast<const uint32_t*>(&lo), v.y = *reinterpret_cast<const uint32_t*>(&hi);
    return v;
}
__device__ __forceinline__ uint32_t loadId(const uint32_t* img, int x, int y, int W, int H)
{
    if (x < 0 || y < 0 || x >= W || y >= H) return 0u;
    return __ldg(img + (size_t)y * W + x);
}

// ReProject:40-55 with the two divisions by tap constants turned into multiplications by their reciprocals and fused adds
// (<= 2 ulp of fp32 from the reference's weight: four orders below the RGBA16F step of the accumulated images)
__device__ __forceinline__ float calculateWeightFast(float rcpDenom, bool sameObject, bool isCenter, float nd)
{
    if (isCenter) return 0.4f;
    if (!sameObject) return 0.0f;
    nd = clampx(nd, 0.0f, 1.0f);
    const float th = 0.98f;
    if (nd < th) return 0.0f;
    const float nw = (nd - th) * (1.0f / (1.0f - th));
    return nw * 2.0f * rcpDenom;
}
struct ReprojectArgs {
    const uint2* src[3];   // rtOutputDiffuse, rtOutputSpecular, rtAlbedo_
    const uint2* hist[3];  // rtPingPong0/1/3
    uint2* out[3];         // rtAccumlatedDiffuse/Specular/Albedo_
    const float2* motion;
    const uint32_t* id0;
    const uint32_t* id1;
    const uint2* normal;
    int W, H;
    RowTiles tiles; // rows to produce (the whole frame unless gk_filter_frame_owned)
};

constexpr int RH = 2; // halo of the 5x5 windows
constexpr int RW = TX + 2 * RH, RHT = TY + 2 * RH;

// One block = 32 x 8 pixels.  Staged in shared memory with a 2-pixel halo: the three source planes, normals, object ids.
// The 5x5 YCoCg clamp box of the albedo (ReProject:158-174) is separable (min / max): a horizontal pass over the staged
// albedo leaves per-row minima / maxima in shared memory and every pixel folds five of them, 10 + 7.5 taps instead of 25.
// The four history taps of the three planes (twelve gathers at the motion-vector target) are issued before the weight loops so
// that their latency overlaps the arithmetic; the 5x5 spatial estimate (needed only where a history tap is rejected) sums the
// three planes while it forms each weight, so no weight array lives in registers.
__global__ void __launch_bounds__(TX* TY, 4) k_reproject(const __grid_constant__ GkUniformBufferObject U, ReprojectArgs A) // UBO in parameter space (constant bank), see k_shade_stream
{
    __shared__ uint2 sSrc[3][RHT][RW];
    __shared__ float4 sNrm[RHT][RW]; // normal, unpacked once per texel (25 pixels read it); w = object id bits
    __shared__ float4 sYc[RHT][RW];  // YCoCg of the albedo texel
    __shared__ float4 sMn[RHT][TX], sMx[RHT][TX]; // horizontal 5-tap min / max of the albedo's YCoCg
    const int W = A.W, H = A.H;
    const int vx = (int)U.ViewportRect[0], vy = (int)U.ViewportRect[1];
    const int bx = blockIdx.x * TX + vx, by = blockIdx.y * TY + vy;
    if (!A.tiles.ownsAny(by, TY)) return;
    const bool progressive = U.ProgressiveRender != 0;
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    const bool mine = x < W && y < H && A.tiles.owns(y);
    const size_t pi = (size_t)y * W + x;
    if (progressive) { // ReProject:76-82
        if (!mine) return;
        const float t = clampx(1.0f / float(U.TemporalFrames), 0.0f, 1.0f);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const c3 s = unpackRgb(__ldg(A.src[ch] + pi)), h = unpackRgb(__ldg(A.hist[ch] + pi));
            A.out[ch][pi] = packRgba(mix3(h, s, t), 1.0f);
        }
        return;
    }
    // ---- early, independent of the tile: motion vector and the history gathers
    float2 motion = make_float2(0.f, 0.f);
    if (mine) motion = __ldg(A.motion + pi);
    const float fxp = float(x) + motion.x, fyp = float(y) + motion.y;
    const int px = (int)floorf(fxp), py = (int)floorf(fyp);
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    uint2 hraw[3][4];
    if (mine) {
        p0 = loadId(A.id1, px, py, W, H), p1 = loadId(A.id1, px + 1, py, W, H), p2 = loadId(A.id1, px, py + 1, W, H), p3 = loadId(A.id1, px + 1, py + 1, W, H);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            hraw[ch][0] = loadPx(A.hist[ch], px, py, W, H), hraw[ch][1] = loadPx(A.hist[ch], px + 1, py, W, H);
            hraw[ch][2] = loadPx(A.hist[ch], px, py + 1, W, H), hraw[ch][3] = loadPx(A.hist[ch], px + 1, py + 1, W, H);
        }
    }
    // ---- stage the tile
    for (int ly = threadIdx.y; ly < RHT; ly += TY)
      for (int lx = threadIdx.x; lx < RW; lx += TX) {
        const int gx = bx + lx - RH, gy = by + ly - RH;
        sSrc[0][ly][lx] = loadPx(A.src[0], gx, gy, W, H);
        sSrc[1][ly][lx] = loadPx(A.src[1], gx, gy, W, H);
        const uint2 alb = loadPx(A.src[2], gx, gy, W, H);
        sSrc[2][ly][lx] = alb;
        const c3 yc = rgb2ycocg(unpackRgb(alb));
        sYc[ly][lx] = make_float4(yc.x, yc.y, yc.z, 0.f);
        const c3 nn = unpackRgb(loadPx(A.normal, gx, gy, W, H));
        sNrm[ly][lx] = make_float4(nn.x, nn.y, nn.z, __uint_as_float(loadId(A.id0, gx, gy, W, H)));
    }
    __syncthreads();
    for (int ly = threadIdx.y; ly < RHT; ly += TY) { // horizontal pass of the clamp box
        const int cx = threadIdx.x;
        float4 t = sYc[ly][cx];
        c3 mn = mk(t.x, t.y, t.z), mx = mn;
#pragma unroll
        for (int dx = 1; dx < 5; ++dx) {
            t = sYc[ly][cx + dx];
            const c3 yc = mk(t.x, t.y, t.z);
            mn = min3(mn, yc), mx = max3(mx, yc);
        }
        sMn[ly][cx] = make_float4(mn.x, mn.y, mn.z, 0.f), sMx[ly][cx] = make_float4(mx.x, mx.y, mx.z, 0.f);
    }
    __syncthreads();
    if (!mine) return;
    const int lx = threadIdx.x + RH, ly = threadIdx.y + RH;
    const int vEndX = (int)(U.ViewportRect[0] + U.ViewportRect[2]), vEndY = (int)(U.ViewportRect[1] + U.ViewportRect[3]);
    const bool inside = (px < vEndX && py < vEndY) && (px >= vx - 1 && py >= vy - 1);
    const float4 cnr = sNrm[ly][lx];
    const uint32_t cur0 = __float_as_uint(cnr.w);
    const bool useHistory = !(cur0 == 65535u || U.TotalFrames == 0 || !inside);
    if (!useHistory) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) A.out[ch][pi] = packRgba(unpackRgb(sSrc[ch][ly][lx]), 1.0f);
        return;
    }
    if (sqrtf(motion.x * motion.x + motion.y * motion.y) < 0.02f) p0 = p1 = p2 = p3 = cur0;
    // The 5x5 spatial estimate only replaces history taps whose object id differs (ReProject:139-147):
    // where all four taps are accepted it is never read, so it is not computed.
    const bool needSpatial = !(cur0 == p0 && cur0 == p1 && cur0 == p2 && cur0 == p3);
    const int R = U.DisableSpatialReuse ? 0 : 2;
    c3 spatial[3] = {mk(0, 0, 0), mk(0, 0, 0), mk(0, 0, 0)};
    if (needSpatial) { // weights are shared by the three channel sets (ReProject:93-121); summed in the reference's tap order
        float total = 0.f;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                if (dx >= -R && dx <= R && dy >= -R && dy <= R) {
                    const float cd = sqrtf(float(dx) * float(dx) + float(dy) * float(dy)); // folded at compile time
                    const float4 nt = sNrm[ly + dy][lx + dx];
                    const float wt = calculateWeightFast(1.0f / (cd * 1.5f + 4.0f), __float_as_uint(nt.w) == cur0, dx == 0 && dy == 0, nt.x * cnr.x + nt.y * cnr.y + nt.z * cnr.z);
                    total += wt;
                    if (wt != 0.f) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const c3 t = unpackRgb(sSrc[ch][ly + dy][lx + dx]);
                            spatial[ch] = mk(fmaf(t.x, wt, spatial[ch].x), fmaf(t.y, wt, spatial[ch].y), fmaf(t.z, wt, spatial[ch].z));
                        }
                    }
                }
            }
        const float rt = 1.0f / total;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) spatial[ch] = spatial[ch] * rt;
    }
    const float sx = fxp - floorf(fxp), sy = fyp - floorf(fyp);
    const uint32_t tf = U.TemporalFrames > 1 ? U.TemporalFrames : 1;
    const float keep = clampx(1.0f / float(tf), 0.0f, 1.0f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const bool needClamp = (ch == 2); // PathTracingRenderer.cpp:122,131,139
        const c3 src = unpackRgb(sSrc[ch][ly][lx]);
        const c3 h0 = cur0 == p0 ? unpackRgb(hraw[ch][0]) : spatial[ch];
        const c3 h1 = cur0 == p1 ? unpackRgb(hraw[ch][1]) : spatial[ch];
        const c3 h2 = cur0 == p2 ? unpackRgb(hraw[ch][2]) : spatial[ch];
        const c3 h3 = cur0 == p3 ? unpackRgb(hraw[ch][3]) : spatial[ch];
        c3 history = mix3(mix3(h0, h1, sx), mix3(h2, h3, sx), sy);
        history = clamp3(history, mk(0.f, 0.f, 0.f), mk(1600.f, 1600.f, 1600.f));
        if (needClamp) {
            c3 mn = rgb2ycocg(src), mx = mn;
#pragma unroll
            for (int dy = -2; dy <= 2; ++dy) { // vertical pass: rows ly-2..ly+2 of the horizontal minima / maxima starting at column lx-2
                const float4 a = sMn[ly + dy][threadIdx.x], b = sMx[ly + dy][threadIdx.x];
                mn = min3(mn, mk(a.x, a.y, a.z)), mx = max3(mx, mk(b.x, b.y, b.z));
            }
            history = ycocg2rgb(clamp3(rgb2ycocg(history), mn, mx));
        }
        A.out[ch][pi] = packRgba(mix3(history, src, keep), 1.0f);
    }
}

// ---- tonemappers (Const_Func.slang:51-68, 84-126) ----
__device__ __forceinline__ float W_f(float x, float e0, float e1)
{
    if (x <= e0) return 0;
    if (x >= e1) return 1;
    const float a = (x - e0) / (e1 - e0);
    return a * a * (3 - 2 * a);
}
__device__ __forceinline__ float H_f(float x, float e0, float e1)
{
    if (x <= e0) return 0;
    if (x >= e1) return 1;
    return (x - e0) / (e1 - e0);
}
__device__ __forceinline__ float granTurismo(float x)
{
    const float e = 2.71828f;
    const float P = 1, a = 0.7f, m = 0.22f, l = 0.4f, c = 1.33f, b = 0;
    const float l0 = (P - m) * l / a;
    const float L_x = m + a * (x - m);
    const float T_x = m * powf(x / m, c) + b;
    const float S0 = m + l0;
    const float S1 = m + a * l0;
    const float C2 = a * P / (P - S1);
    const float S_x = P - (P - S1) * powf(e, -(C2 * (x - S0) / P));
    const float w0 = 1 - W_f(x, 0, m);
    const float w2 = H_f(x, m + l0, m + l0);
    const float w1 = 1 - w0 - w2;
    return T_x * w0 + L_x * w1 + S_x * w2;
}
__device__ __forceinline__ c3 gt3(c3 v) { return mk(granTurismo(v.x), granTurismo(v.y), granTurismo(v.z)); }
__device__ __forceinline__ float st2084one(float v)
{
    const float m1 = 0.1593017578125f, m2 = 78.84375f, c1 = 0.8359375f, c2 = 18.8515625f, c3_ = 18.6875f, C = 10000.f;
    const float L = v / C;
    const float Lm = powf(L, m1);
    const float N1 = c1 + c2 * Lm, N2 = 1.0f + c3_ * Lm;
    const float N = N1 * (1.0f / N2);
    return powf(N, m2);
}

__device__ __forceinline__ bool edgeDetect(uint32_t center, const uint32_t* img, int x, int y, int W, int H) // DenoiseJBF:55-68
{
    const uint32_t a = loadId(img, x + 1, y + 1, W, H), b = loadId(img, x - 1, y - 1, W, H), c = loadId(img, x - 1, y + 1, W, H), d = loadId(img, x + 1, y - 1, W, H);
    const bool e0 = a != center || b != center || c != center || d != center;
    const bool e1 = a == center || b == center || c == center || d == center;
    return e0 && e1;
}

struct DenoiseArgs {
    const uint2* diffuse; // rtAccumlatedDiffuse ("FinalImage")
    const uint2* spec;    // rtAccumlatedSpecular
    const uint2* albedo;  // rtAccumlatedAlbedo_
    const uint32_t* id0;
    const uint32_t* id1;
    uint2* out; // rtDenoised
    int W, H;
    RowTiles tiles;
};

constexpr int DH = 5;
constexpr int JX = 32, JY = 16; // pixels per block of the compose kernel
constexpr int DW = JX + 2 * DH, DHT = JY + 2 * DH;

// 2^x by the SFU (ex2.approx without flush-to-zero: gradual underflow kept).  2 ulp; used only where the result is weighed
// against terms >= 1e-25 (see k_denoise_jbf).
__device__ __forceinline__ float fastExp2(float x)
{
    float y;
    asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Flush-to-zero form (one MUFU.EX2): results below 2^-126 become 0.  Only for terms that are weighed against >= 1e-25.
__device__ __forceinline__ float fastExp2Ftz(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fastLog2(float x)
{
    float y;
    asm("lg2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// x^c and e^-k through lg2/ex2 (relative error ~1e-6: three orders below the half-precision output step)
__device__ __forceinline__ float granTurismoFast(float x)
{
    const float P = 1, a = 0.7f, m = 0.22f, l = 0.4f, c = 1.33f;
    const float l0 = (P - m) * l / a;
    const float S0 = m + l0, S1 = m + a * l0;
    const float C2 = a * P / (P - S1);
    if (!(x > 0.f) || !(x < 1e25f)) return granTurismo(x); // zero, negative, NaN, huge (the unused toe term overflows to inf * 0): the exact routine defines the result
    if (x < m) { // toe blended into the linear part: w2 = 0
        const float r = x * (1.0f / m);                // x / m, also the smoothstep argument (x - 0) / (m - 0)
        const float T_x = m * fastExp2(c * fastLog2(r));
        const float L_x = fmaf(a, x - m, m);
        const float w0 = 1 - r * r * (3 - 2 * r);
        return fmaf(T_x, w0, L_x * (1 - w0));
    }
    if (x > m + l0) // shoulder: w0 = 0, w2 = 1 (H_f is 0 AT the edge), w1 = 0.  pow(2.71828f, k) = 2^(k log2(2.71828f)), log2(2.71828f) = 1.44269411...
        return P - (P - S1) * fastExp2((x - S0) * (-C2 * 1.4426941f));
    return fmaf(a, x - m, m); // linear section: w0 = 0, w1 = 1, w2 = 0
}
__device__ __forceinline__ c3 gt3fast(c3 v) { return mk(granTurismoFast(v.x), granTurismoFast(v.y), granTurismoFast(v.z)); }

// Compose (DenoiseJBF:96-195).  A block owns 32 x 16 pixels and stages diffuse + 0.001 (DenoiseJBF:37) and its luminance with a
// 5-pixel halo in shared memory (overfetch 2.1x, served by L2).  36 taps per pixel.  The luminance weight
// exp(-dl^4 / (2 sigmaL^2)) is taken from the SFU (ex2.approx) and the products are fused; a pixel whose summed weight is
// below 1e-25 (every tap in the denormal range of the weight: a bright outlier among dark neighbours), zero or NaN is
// recomputed with IEEE expf and the reference's exact grouping, because there single denormal quanta decide the result
// (and where the reference's 0/0 NaNs appear).  Above that threshold the SFU error (4e-6 relative per weight) is three
// orders below the half-precision step of the output.
__global__ void __launch_bounds__(JX* JY) k_denoise_jbf(const __grid_constant__ GkUniformBufferObject U, DenoiseArgs A)
{
    __shared__ float4 sDif[DHT][DW]; // diffuse + bias, w = luminance
    __shared__ float sFi[36];        // spatial weights of the 36 taps: they depend on the tap only
    __shared__ float sLf[36];        // ... and their base-2 logarithms (folded into the exponent of the luminance weight)
    const int W = A.W, H = A.H;
    const int vx = (int)U.ViewportRect[0], vy = (int)U.ViewportRect[1];
    const int bx = blockIdx.x * JX + vx, by = blockIdx.y * JY + vy;
    if (!A.tiles.ownsAny(by, JY)) return;
    const bool filter = U.BFSize > 0;
    if (filter) {
        const int t = threadIdx.y * JX + threadIdx.x;
        for (int ly = threadIdx.y; ly < DHT; ly += JY)
            for (int lx = threadIdx.x; lx < DW; lx += JX) {
                const c3 c = unpackRgb(loadPx(A.diffuse, bx + lx - DH, by + ly - DH, W, H)) + mk(0.001f, 0.001f, 0.001f);
                sDif[ly][lx] = make_float4(c.x, c.y, c.z, c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f);
            }
        if (t < 36) {
            const int i = 2 * (t / 6) - 5, j = 2 * (t % 6) - 5;
            const float dist = clampx(float(i * i + j * j) / float(5 * 5), 0.0f, 1.0f);
            sFi[t] = expf(-dist * dist / (2.0f * U.BFSigma * U.BFSigma));
            sLf[t] = (-dist * dist / (2.0f * U.BFSigma * U.BFSigma)) * 1.4426950408889634f;
        }
        __syncthreads();
    }
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    if (x >= W || y >= H || !A.tiles.owns(y)) return;
    const size_t pi = (size_t)y * W + x;
    const c3 bias = mk(0.001f, 0.001f, 0.001f);
    c3 Total = mk(0, 0, 0);
    const c3 specC = unpackRgb(__ldg(A.spec + pi)), albC = unpackRgb(__ldg(A.albedo + pi));
    if (filter) {
        const int lx = threadIdx.x + DH, ly = threadIdx.y + DH;
        const float sigmaL = U.BFSigmaLum * 100.0f;
        const float invL = 1.0f / (2.0f * sigmaL * sigmaL);
        const float kx = -invL * 1.4426950408889634f; // exp(-q invL) = 2^(q kx)
        const c3 cs = specC + bias;
        const float clum = sDif[ly][lx].w;
        float Weight = 0;
        float tx = 0, ty = 0, tz = 0;
#pragma unroll
        for (int i = -5; i <= 5; i += 2)
#pragma unroll
            for (int j = -5; j <= 5; j += 2) {
                const float4 tap = sDif[ly + i][lx + j];
                const float dq = clum - tap.w, dl = dq * dq;
                const float w = fastExp2Ftz(fmaf(dl * dl, kx, sLf[((i + 5) / 2) * 6 + (j + 5) / 2])); // Fi * Li = 2^(log2 Fi + dl^2 kx)
                tx = fmaf(tap.x, w, tx), ty = fmaf(tap.y, w, ty), tz = fmaf(tap.z, w, tz);
                Weight += w;
            }
        if (Weight >= 1e-25f && Weight < 1e30f) {
            const float rw = 1.0f / Weight;
            Total = mk(tx * rw, ty * rw, tz * rw);
        } else {
            // rare: the reference's exact arithmetic decides (denormal weights, 0/0)
            Weight = 0;
            c3 T = mk(0, 0, 0);
            for (int i = -5; i <= 5; i += 2)
                for (int j = -5; j <= 5; j += 2) {
                    const float4 tap = sDif[ly + i][lx + j];
                    const c3 Ci = mk(tap.x, tap.y, tap.z);
                    const float dl = (clum - tap.w) * (clum - tap.w);
                    const float Fi = sFi[((i + 5) / 2) * 6 + (j + 5) / 2];
                    const float Li = expf(-dl * dl * invL);
                    T = T + Ci * Fi * Li; // the reference's grouping
                    Weight += Fi * Li;
                }
            Total = T / Weight;
        }
        if (!U.DebugDraw_Lighting) Total = Total * albC + cs;
    } else {
        const c3 d = unpackRgb(__ldg(A.diffuse + pi));
        if (U.DebugDraw_Lighting) Total = d * mk(0.5f, 0.5f, 0.5f) + specC;
        else Total = d * albC + specC;
    }
    if (U.SelectedId != 0xFFFFFFFFu) { // object ids never equal 0xFFFFFFFF ("nothing selected"): no outline, no id fetches
        const float eThis = edgeDetect(U.SelectedId, A.id0, x, y, W, H) ? 0.5f : 0.0f;
        const float ePrev = edgeDetect(U.SelectedId, A.id1, x, y, W, H) ? 0.5f : 0.0f;
        if (eThis + eThis > 0) Total = mix3(Total, mk(150, 100, 0), eThis + ePrev);
    }
    c3 o;
    if (U.HDR) {
        Total = Total / 2000.f;
        Total = gt3(Total);
        Total = Total * 2000.f;
        const c3 v = Total * U.PaperWhiteNit / 230.0f;
        o = mk(st2084one(v.x), st2084one(v.y), st2084one(v.z));
    } else {
        o = gt3fast(Total * (U.PaperWhiteNit * (1.0f / 40000.0f)));
    }
    A.out[pi] = packRgba(o, 1.0f);
}

} // namespace

static GkStatus runFilters(Context& c, bool ownedRowsOnly);

GkStatus filterFrame(Context& c) { return runFilters(c, false); }

// Tile-partitioned progressive frames: every pass is per pixel (ReProject:76-82, DenoiseJBF with BFSize == 0),
// so each rank accumulates and composes only the rows it traced; history stays on the rank that owns the row.
GkStatus filterFrameOwnedRows(Context& c)
{
    if (c.haveUbo && c.tileCount > 1 && !(c.ubo.ProgressiveRender != 0 && c.ubo.BFSize == 0)) {
        setLastError("gk_filter_frame_owned: needs ProgressiveRender != 0 and BFSize == 0 (the spatial passes read rows of other ranks)");
        return GK_ERR_UNSUPPORTED;
    }
    // the selection outline (DenoiseJBF:55-68, 173-179) reads object ids at y +- 1, i.e. rows of other ranks on tile borders, which
    // are neither traced nor exchanged in this mode: only "nothing selected" composes correctly on owned rows alone
    if (c.haveUbo && c.tileCount > 1 && c.ubo.SelectedId != 0xFFFFFFFFu) {
        setLastError("gk_filter_frame_owned: SelectedId must be 0xFFFFFFFF (the selection outline reads object ids of neighbouring rows owned by other ranks)");
        return GK_ERR_UNSUPPORTED;
    }
    return runFilters(c, true);
}

static GkStatus runFilters(Context& c, bool ownedRowsOnly)
{
    if (!c.haveUbo) {
        setLastError("gk_filter_frame: UBO must be set first");
        return GK_ERR_NOT_READY;
    }
    cudaStream_t st = c.stream;
    if (!c.tracedSinceFilter) applyPendingHistorySwap(c); // filter-only use: each call is a new frame
    {
        const void* bufs[] = {c.planes.p[GK_PLANE_ACCUM_DIFFUSE], c.planes.p[GK_PLANE_ACCUM_SPECULAR], c.planes.p[GK_PLANE_ACCUM_ALBEDO], c.planes.p[GK_PLANE_DENOISED]};
        waitAsyncCopyBeforeWriting(c, bufs, 4);
    }
    c.tracedSinceFilter = false;
    GK_CUDA(cudaMemcpyAsync(c.dUbo, &c.ubo, sizeof(GkUniformBufferObject), cudaMemcpyHostToDevice, st));
    ScopedEvents evs(3);
    cudaEvent_t e0 = evs.e[0], e1 = evs.e[1], e2 = evs.e[2];
    void** P = c.planes.p;
    ReprojectArgs R;
    R.src[0] = (const uint2*)P[GK_PLANE_OUTPUT_DIFFUSE], R.src[1] = (const uint2*)P[GK_PLANE_OUTPUT_SPECULAR], R.src[2] = (const uint2*)P[GK_PLANE_ALBEDO];
    R.hist[0] = (const uint2*)P[GK_PLANE_HISTORY_DIFFUSE], R.hist[1] = (const uint2*)P[GK_PLANE_HISTORY_SPECULAR], R.hist[2] = (const uint2*)P[GK_PLANE_HISTORY_ALBEDO];
    R.out[0] = (uint2*)P[GK_PLANE_ACCUM_DIFFUSE], R.out[1] = (uint2*)P[GK_PLANE_ACCUM_SPECULAR], R.out[2] = (uint2*)P[GK_PLANE_ACCUM_ALBEDO];
    R.motion = (const float2*)P[GK_PLANE_MOTION], R.id0 = (const uint32_t*)P[GK_PLANE_OBJECT_ID0], R.id1 = (const uint32_t*)P[GK_PLANE_OBJECT_ID1];
    R.normal = (const uint2*)P[GK_PLANE_NORMAL];
    R.W = (int)c.width, R.H = (int)c.height;
    R.tiles = ownedRowsOnly ? RowTiles{c.tileIndex, c.tileCount, c.tileRows} : RowTiles{0, 1, 1};
    const dim3 block(TX, TY), grid((c.width + TX - 1) / TX, (c.height + TY - 1) / TY);
    cudaEventRecord(e0, st);
    k_reproject<<<grid, block, 0, st>>>(c.ubo, R);
    cudaEventRecord(e1, st);
    DenoiseArgs D;
    D.diffuse = R.out[0], D.spec = R.out[1], D.albedo = R.out[2], D.id0 = R.id0, D.id1 = R.id1, D.out = (uint2*)P[GK_PLANE_DENOISED];
    D.W = R.W, D.H = R.H;
    D.tiles = R.tiles;
    const dim3 jblock(JX, JY), jgrid((c.width + JX - 1) / JX, (c.height + JY - 1) / JY);
    k_denoise_jbf<<<jgrid, jblock, 0, st>>>(c.ubo, D);
    cudaEventRecord(e2, st);
    GK_CUDA(cudaGetLastError());
    GK_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c.stats.msReproject, e0, e1);
    cudaEventElapsedTime(&c.stats.msDenoise, e1, e2);
    c.stats.launches += 2;
    // "copy pass": the accumulated images become next frame's history and ObjectId0 becomes
    // ObjectId1.  No bytes move: the plane pointers are exchanged when the next frame starts
    // (applyPendingHistorySwap); until then reads of the history planes resolve to the
    // accumulated ones, which is what the reference's images hold after its copies.
    c.pendingHistorySwap = true;
    return GK_OK;
}

// Compose + tonemap (k_denoise_jbf) of the rows this context owns from the accumulated planes, then the
// history hand-over.  Used by the frame-sharded accumulation, which writes the accumulated planes itself.
GkStatus composeOwnedRows(Context& c)
{
    cudaStream_t st = c.stream;
    void** P = c.planes.p;
    DenoiseArgs D;
    D.diffuse = (const uint2*)P[GK_PLANE_ACCUM_DIFFUSE], D.spec = (const uint2*)P[GK_PLANE_ACCUM_SPECULAR], D.albedo = (const uint2*)P[GK_PLANE_ACCUM_ALBEDO];
    D.id0 = (const uint32_t*)P[GK_PLANE_OBJECT_ID0], D.id1 = (const uint32_t*)P[GK_PLANE_OBJECT_ID1], D.out = (uint2*)P[GK_PLANE_DENOISED];
    D.W = (int)c.width, D.H = (int)c.height;
    D.tiles = RowTiles{c.tileIndex, c.tileCount, c.tileRows};
    {
        const void* bufs[] = {P[GK_PLANE_DENOISED]};
        waitAsyncCopyBeforeWriting(c, bufs, 1);
    }
    const dim3 jblock(JX, JY), jgrid((c.width + JX - 1) / JX, (c.height + JY - 1) / JY);
    k_denoise_jbf<<<jgrid, jblock, 0, st>>>(c.ubo, D);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    c.tracedSinceFilter = false;
    c.pendingHistorySwap = true;
    return GK_OK;
}

void applyPendingHistorySwap(Context& c)
{
    if (!c.pendingHistorySwap) return;
    void** P = c.planes.p;
    std::swap(P[GK_PLANE_ACCUM_DIFFUSE], P[GK_PLANE_HISTORY_DIFFUSE]);
    std::swap(P[GK_PLANE_ACCUM_SPECULAR], P[GK_PLANE_HISTORY_SPECULAR]);
    std::swap(P[GK_PLANE_ACCUM_ALBEDO], P[GK_PLANE_HISTORY_ALBEDO]);
    std::swap(P[GK_PLANE_OBJECT_ID0], P[GK_PLANE_OBJECT_ID1]);
    c.pendingHistorySwap = false;
}

} // namespace gk
