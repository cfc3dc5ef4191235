// gk_bvh.cuh — acceleration-structure layouts and the two-level traversal used by the
// extend / shadow kernels.
//
// What it replaces: the Vulkan RayQuery against the driver's TLAS/BLAS
// (assets/shaders/common/Shading.slang:659-758) on the GPU side, and tinybvh's
// BVH::Intersect / IntersectTLAS (src/ThirdParty/tinybvh/tiny_bvh.h:2245-2353) whose
// per-triangle arithmetic it reproduces bit for bit:
//   ray setup ............ tiny_bvh.h:562-567 (normalise, safe reciprocal :329)
//   instance transform ... tiny_bvh.h:2311-2315 (direction NOT re-normalised: t stays world-space)
//   triangle test ........ tiny_bvh.h:6815-6843 (|a|<1e-7 reject, u,v in [0,1], u+v<=1, tmin<t<hit.t)
// The tree itself is ours: an 8-wide node with child boxes quantised to 8 bits against the
// node's own box (96 bytes used of a 128-byte line), built on the GPU (gk_bvh_build.cu).
//
// HBM/L2 layout
//   WideNode  128 B stride, one cache line per node, read as 6 x 128-bit loads
//   TriRecord  48 B: v0 | e1 = v1-v0 | e2 = v2-v0 (the exact fp32 differences tinybvh
//              forms per test), w lanes carry the original triangle index
//   InstRecord 80 B: row-major inverse transform (64 B) + BLAS root reference + node index
#pragma once
#include "gk_common.cuh"

namespace gk {

// child / root reference: bit31 = leaf.
//  BLAS leaf: bits[30:3] first triangle (sorted order), bits[2:0] = count-1
//  TLAS leaf: bits[30:0] instance index
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kSentinel = 0xfffffffeu; // "return to the TLAS" stack marker

struct __align__(16) WideNode {
    float ox, oy, oz; // box origin (one quantisation step below the true minimum)
    uint8_t ex, ey, ez, count; // biased exponents of the per-axis step, valid children
    uint32_t child[8];
    uint8_t qlo[3][8];
    uint8_t qhi[3][8];
    uint32_t src[8]; // binary-tree node each slot was made from (refit re-quantises from these)
};
static_assert(sizeof(WideNode) == 128, "WideNode must fill one cache line");

struct __align__(16) TriRecord {
    float v0x, v0y, v0z;
    uint32_t prim;
    float e1x, e1y, e1z;
    uint32_t pad0;
    float e2x, e2y, e2z;
    uint32_t pad1;
};
static_assert(sizeof(TriRecord) == 48, "TriRecord");

struct __align__(16) InstRecord {
    float invT[16];    // row-major inverse world transform (tinybvh BLASInstance::invTransform)
    uint32_t blasRoot; // reference into the BLAS node array (may be a leaf reference)
    uint32_t node;     // index into the NodeProxy array
    uint32_t model;
    uint32_t pad;
};
static_assert(sizeof(InstRecord) == 80, "InstRecord");

struct SceneView { // everything a traversal needs, passed by value to kernels
    const WideNode* tlasNodes;
    const WideNode* blasNodes;
    const TriRecord* tris;
    const InstRecord* inst;
    uint32_t tlasRoot; // reference (leaf bit possible when there is a single instance)
    uint32_t instanceCount;
};

struct Hit {
    float t, u, v;
    uint32_t prim, inst;
};

struct TraversalStats {
    unsigned long long nodeVisits, triTests;
};

// ---- exact triangle test -------------------------------------------------------------
// Returns true and shortens `hit` when tmin < t < hit.t.  O, D are the (instance-space) ray.
GK_HD bool triangleTest(const TriRecord& T, f3 O, f3 D, float tmin, Hit& hit, uint32_t instIdx)
{
    const f3 e1 = mk3(T.e1x, T.e1y, T.e1z), e2 = mk3(T.e2x, T.e2y, T.e2z);
    const f3 h = xcross(D, e2);
    const float a = xdot(e1, h);
    if (fabsf(a) < 0.0000001f) return false;
    const float f = xdiv(1.0f, a);
    const f3 s = xsub3(O, mk3(T.v0x, T.v0y, T.v0z));
    const float u = xmul(f, xdot(s, h));
    if (u < 0 || u > 1) return false;
    const f3 q = xcross(s, e1);
    const float v = xmul(f, xdot(D, q));
    if (v < 0 || xadd(u, v) > 1) return false;
    const float t = xmul(f, xdot(e2, q));
    if (t > tmin && t < hit.t) {
        hit.t = t, hit.u = u, hit.v = v, hit.prim = T.prim, hit.inst = instIdx;
        return true;
    }
    return false;
}

// ---- wide-node box tests ---------------------------------------------------------------
struct NodeTest {
    float t[8]; // entry distance per slot, kFar when missed
};

GK_HD float byteToFloat(uint32_t word, int k)
{
#ifdef __CUDA_ARCH__
    // 0x4B0000qq is 2^23 + q exactly; one PRMT + one FADD, both full-rate
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540 + k)) - 8388608.0f;
#else
    return (float)((word >> (8 * k)) & 0xffu);
#endif
}

GK_HD float expToFloat(uint32_t e8)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(e8 << 23);
#else
    uint32_t b = e8 << 23;
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

struct NodeWords { // one WideNode pulled into registers with 6 x 128-bit loads
    uint4 q0, q1, q2, q3, q4, q5;
};
GK_HD NodeWords loadNode(const WideNode* n)
{
    NodeWords w;
    const uint4* p = reinterpret_cast<const uint4*>(n);
#ifdef __CUDA_ARCH__
    w.q0 = __ldg(p + 0), w.q1 = __ldg(p + 1), w.q2 = __ldg(p + 2), w.q3 = __ldg(p + 3), w.q4 = __ldg(p + 4), w.q5 = __ldg(p + 5);
#else
    w.q0 = p[0], w.q1 = p[1], w.q2 = p[2], w.q3 = p[3], w.q4 = p[4], w.q5 = p[5];
#endif
    return w;
}
GK_HD uint32_t nodeChild(const NodeWords& w, int i)
{
    switch (i) {
    case 0: return w.q1.x;
    case 1: return w.q1.y;
    case 2: return w.q1.z;
    case 3: return w.q1.w;
    case 4: return w.q2.x;
    case 5: return w.q2.y;
    case 6: return w.q2.z;
    default: return w.q2.w;
    }
}
GK_HD float asFloat(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// Slab test of all 8 quantised child boxes.  Conservative: boxes carry >= 1/64 step of
// slack from the builder and the comparison allows 4 ulp on the exit distance.
GK_HD uint32_t testWideNode(const NodeWords& N, f3 O, f3 rD, float tmin, float tmax, NodeTest& out)
{
    const uint32_t ec = N.q0.w;
    const uint32_t count = ec >> 24;
    const float sx = expToFloat(ec & 0xffu) * rD.x, sy = expToFloat((ec >> 8) & 0xffu) * rD.y, sz = expToFloat((ec >> 16) & 0xffu) * rD.z;
    const float bx = (asFloat(N.q0.x) - O.x) * rD.x, by = (asFloat(N.q0.y) - O.y) * rD.y, bz = (asFloat(N.q0.z) - O.z) * rD.z;
    // per axis pick which byte plane is the entry side
    const bool nx = rD.x < 0, ny = rD.y < 0, nz = rD.z < 0;
    const uint32_t loX[2] = {N.q3.x, N.q3.y}, loY[2] = {N.q3.z, N.q3.w}, loZ[2] = {N.q4.x, N.q4.y};
    const uint32_t hiX[2] = {N.q4.z, N.q4.w}, hiY[2] = {N.q5.x, N.q5.y}, hiZ[2] = {N.q5.z, N.q5.w};
    uint32_t mask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t nearX = nx ? hiX[half] : loX[half], farX = nx ? loX[half] : hiX[half];
        const uint32_t nearY = ny ? hiY[half] : loY[half], farY = ny ? loY[half] : hiY[half];
        const uint32_t nearZ = nz ? hiZ[half] : loZ[half], farZ = nz ? loZ[half] : hiZ[half];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float t0x = fmaf(byteToFloat(nearX, k), sx, bx), t1x = fmaf(byteToFloat(farX, k), sx, bx);
            const float t0y = fmaf(byteToFloat(nearY, k), sy, by), t1y = fmaf(byteToFloat(farY, k), sy, by);
            const float t0z = fmaf(byteToFloat(nearZ, k), sz, bz), t1z = fmaf(byteToFloat(farZ, k), sz, bz);
            const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
            const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
            const int slot = half * 4 + k;
            const bool hitBox = (tn <= tf * 1.0000005f) && (slot < (int)count);
            out.t[slot] = hitBox ? tn : kFar;
            mask |= hitBox ? (1u << slot) : 0u;
        }
    }
    return mask;
}

// ---- two-level traversal ---------------------------------------------------------------
constexpr int kStackSize = 48;

struct StackEntry {
    uint32_t ref;
    float t;
};

// Closest hit (anyHit = false) or first hit (anyHit = true).  `hit.t` must hold tmax on entry.
// Returns true if something was hit.  Dn is the NORMALISED world direction (tinybvh normalises
// in the Ray constructor), O the world origin.
template <bool kAnyHit, bool kStats>
GK_HD bool traverseScene(const SceneView& S, f3 O, f3 Dn, float tmin, Hit& hit, TraversalStats* stats)
{
    StackEntry stack[kStackSize];
    int sp = 0;
    const float tmax0 = hit.t;
    // current (TLAS = world, or instance) ray
    f3 o = O, d = Dn;
    f3 rd = mk3(safeRcp(Dn.x), safeRcp(Dn.y), safeRcp(Dn.z));
    bool inBlas = false;
    uint32_t curInst = 0;
    uint32_t cur = S.tlasRoot;
    if (S.instanceCount == 0) return false;

    for (;;) {
        if (!(cur & kLeafBit)) {
            const NodeWords N = loadNode(inBlas ? S.blasNodes + cur : S.tlasNodes + cur);
            if (kStats) stats->nodeVisits++;
            NodeTest nt;
            uint32_t mask = testWideNode(N, o, rd, tmin, hit.t, nt);
            if (mask) {
                // continue with the nearest child, push the rest with their entry distances
                int best = -1;
                float bt = kFar;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if ((mask >> i) & 1u) {
                        if (nt.t[i] < bt || best < 0) bt = nt.t[i], best = i;
                    }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (((mask >> i) & 1u) && i != best && sp < kStackSize) stack[sp].ref = nodeChild(N, i), stack[sp].t = nt.t[i], ++sp;
                cur = nodeChild(N, best);
                continue;
            }
        } else if (!inBlas) {
            // TLAS leaf: enter the instance (tiny_bvh.h:2305-2315)
            const uint32_t ii = cur & 0x7fffffffu;
            const InstRecord& I = S.inst[ii];
            if (sp < kStackSize) stack[sp].ref = kSentinel, stack[sp].t = 0.f, ++sp;
            o = xformPoint(O, I.invT);
            d = xformVector(Dn, I.invT);
            rd = mk3(safeRcp(d.x), safeRcp(d.y), safeRcp(d.z));
            inBlas = true;
            curInst = I.node;
            cur = I.blasRoot;
            continue;
        } else {
            // BLAS leaf: 1..8 consecutive triangle records
            const uint32_t first = (cur & 0x7fffffffu) >> 3, cnt = (cur & 7u) + 1u;
            for (uint32_t k = 0; k < cnt; ++k) {
                if (kStats) stats->triTests++;
                const bool h = triangleTest(S.tris[first + k], o, d, tmin, hit, curInst);
                if (kAnyHit && h) return true;
            }
        }
        // pop
        for (;;) {
            if (sp == 0) return hit.t < tmax0;
            --sp;
            const uint32_t r = stack[sp].ref;
            if (r == kSentinel) {
                o = O, d = Dn;
                rd = mk3(safeRcp(Dn.x), safeRcp(Dn.y), safeRcp(Dn.z));
                inBlas = false;
                continue;
            }
            if (stack[sp].t < hit.t) {
                cur = r;
                break;
            }
        }
    }
}

// tinybvh's Ray constructor (tiny_bvh.h:562-567, :391-395)
GK_HD f3 normalizeRayDir(f3 D)
{
    const float l = xsqrt(xadd(xadd(xmul(D.x, D.x), xmul(D.y, D.y)), xmul(D.z, D.z)));
    const float rl = (l == 0) ? 0.0f : xdiv(1.0f, l);
    return mk3(xmul(D.x, rl), xmul(D.y, rl), xmul(D.z, rl));
}

} // namespace gk
