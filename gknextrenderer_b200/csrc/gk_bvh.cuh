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
// node's own box, one 128-byte cache line per node, built on the GPU (gk_bvh_build.cu).
//
// B200 has no RT cores; traversal is issue-bound shader code (profiles/), so the design goal is
// fewest instructions per box test and highest lane utilisation:
//   * box test = 6 PRMT + 6 FFMA + 2 three-input min/max + compare per child.  A child plane byte
//     q is turned into the float 2^23+q by ONE byte-permute (the constant bytes 0x00,0x4B live in
//     the child record itself), the -2^23 is folded into the FMA addend, and the permute selectors
//     pick the entry / exit plane by the sign of the ray direction, so no per-axis min/max remains.
//     The folded constant costs at most half a quantisation step of precision; the builder pads
//     every box by one full step, so the test stays conservative.
//   * two ray-to-lane mappings:
//       one ray per lane (large waves): "while-while" traversal, every lane descends to a leaf,
//         then the warp processes leaves together; stack in local memory;
//       cooperative groups (small waves): eight lanes own one ray, lane j tests child j / triangle j,
//         ballots and warp reductions pick the nearest; ~5x lower latency per ray, which bounds the
//         long tail of nearly empty waves.
//
// HBM/L2 layout
//   WideNode  128 B: 16-B header (origin, per-axis step exponents, child count) + 8 child
//             records of 12 B {reference | qlo.xyz qhi.x | qhi.yz 0x00 0x4B}
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
constexpr uint32_t kBlasLeafMax = 8;        // triangles per BLAS leaf (one per lane in the cooperative mapping)

struct WideChild {
    uint32_t ref;
    uint8_t qlo[3];
    uint8_t qhi[3];
    uint8_t k00, k4b; // constant bytes 0x00 and 0x4B: the upper bytes of the float 2^23 + q
};
static_assert(sizeof(WideChild) == 12, "WideChild");

struct __align__(16) WideNode {
    float ox, oy, oz;          // box origin (two quantisation steps below the true minimum)
    uint8_t ex, ey, ez, count; // biased exponents of the per-axis step, valid children
    WideChild c[8];
    uint32_t spare[4];
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
    uint32_t indexOffset;  // first index / first vertex of the instance's model: the shading fetch goes
    uint32_t vertexOffset; // instance -> indices -> vertices without passing through NodeProxy and ModelInfo
};
static_assert(sizeof(InstRecord) == 80, "InstRecord");

struct SceneView { // everything a traversal needs, passed by value to kernels
    const WideNode* tlasNodes;
    const WideNode* blasNodes;
    const TriRecord* tris;
    const InstRecord* inst;
    uint32_t tlasRoot; // reference (leaf bit possible when there is a single instance)
    uint32_t instanceCount;
    uint32_t* overflowFlag; // set to 1 by a traversal that had to drop a stack entry (the host turns it into an error)
};

struct Hit {
    float t, u, v;
    uint32_t prim, inst;
};

struct TraversalStats {
    unsigned long long nodeVisits, triTests, tlasVisits, instanceEntries;
    unsigned long long maxStack; // deepest traversal stack seen (entries); must stay below kStackSize
};

// ---- exact triangle test -------------------------------------------------------------
// Returns true when tmin < t < tmax; O, D are the (instance-space) ray.
GK_HD bool triangleTest(const TriRecord& T, f3 O, f3 D, float tmin, float tmax, float& t, float& u, float& v)
{
    const f3 e1 = mk3(T.e1x, T.e1y, T.e1z), e2 = mk3(T.e2x, T.e2y, T.e2z);
    const f3 h = xcross(D, e2);
    const float a = xdot(e1, h);
    if (fabsf(a) < 0.0000001f) return false;
    const float f = xdiv(1.0f, a);
    const f3 s = xsub3(O, mk3(T.v0x, T.v0y, T.v0z));
    u = xmul(f, xdot(s, h));
    if (u < 0 || u > 1) return false;
    const f3 q = xcross(s, e1);
    v = xmul(f, xdot(D, q));
    if (v < 0 || xadd(u, v) > 1) return false;
    t = xmul(f, xdot(e2, q));
    return t > tmin && t < tmax;
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

// tinybvh's Ray constructor (tiny_bvh.h:562-567, :391-395)
GK_HD f3 normalizeRayDir(f3 D)
{
    const float l = xsqrt(xadd(xadd(xmul(D.x, D.x), xmul(D.y, D.y)), xmul(D.z, D.z)));
    const float rl = (l == 0) ? 0.0f : xdiv(1.0f, l);
    return mk3(xmul(D.x, rl), xmul(D.y, rl), xmul(D.z, rl));
}

#ifdef __CUDACC__
constexpr int kStackSize = 48;
constexpr int kStackStride = kStackSize + 1; // odd row stride: rows of different rays start in different banks
constexpr int kRaysPerBlock = 32;            // cooperative mapping: 256 threads, 8 lanes per ray
constexpr float kBoxTolerance = 1.000001f;   // slab comparison slack on the exit distance

// Reciprocal direction for the BOX tests only (the triangle test never uses it, tiny_bvh.h:6815-6843):
// one MUFU.RCP instead of the IEEE division sequence.  Its 1-ulp error is covered by the padding of
// the quantised boxes and the tolerance of the slab comparison.
__device__ __forceinline__ float boxRcp(float x)
{
    if (fabsf(x) > 1e-12f) return __fdividef(1.0f, x);
    return kFar;
}
__device__ __forceinline__ f3 boxRcp3(f3 d) { return mk3(boxRcp(d.x), boxRcp(d.y), boxRcp(d.z)); }

// Byte-permute selectors for the six planes of a child record {w1 = qlo.x qlo.y qlo.z qhi.x,
// w2 = qhi.y qhi.z 0x00 0x4B}: result bytes = [plane byte, 0x00, 0x00, 0x4B] = float(2^23 + q).
struct PlaneSel {
    uint32_t nx, fx, ny, fy, nz, fz; // entry ("near") and exit ("far") plane per axis
};
__device__ __forceinline__ PlaneSel makePlaneSel(f3 rd)
{
    PlaneSel s;
    const bool x = rd.x < 0, y = rd.y < 0, z = rd.z < 0;
    s.nx = 0x7660u | (x ? 3u : 0u), s.fx = 0x7660u | (x ? 0u : 3u);
    s.ny = 0x7660u | (y ? 4u : 1u), s.fy = 0x7660u | (y ? 1u : 4u);
    s.nz = 0x7660u | (z ? 5u : 2u), s.fz = 0x7660u | (z ? 2u : 5u);
    return s;
}

struct NodeFrame { // per node visit: t(q) = fma(2^23 + q, s, b)
    float sx, sy, sz, bx, by, bz;
};
__device__ __forceinline__ NodeFrame makeNodeFrame(const uint4 hdr, f3 o, f3 rd)
{
    NodeFrame F;
    const uint32_t ec = hdr.w;
    F.sx = __uint_as_float((ec & 0xffu) << 23) * rd.x, F.sy = __uint_as_float(((ec >> 8) & 0xffu) << 23) * rd.y, F.sz = __uint_as_float(((ec >> 16) & 0xffu) << 23) * rd.z;
    F.bx = fmaf(-8388608.0f, F.sx, (__uint_as_float(hdr.x) - o.x) * rd.x);
    F.by = fmaf(-8388608.0f, F.sy, (__uint_as_float(hdr.y) - o.y) * rd.y);
    F.bz = fmaf(-8388608.0f, F.sz, (__uint_as_float(hdr.z) - o.z) * rd.z);
    return F;
}
// true when the ray enters the child box before leaving it; tn = entry distance
__device__ __forceinline__ bool childTest(const NodeFrame& F, const PlaneSel& S, uint32_t w1, uint32_t w2, float tmin, float tmax, float& tn)
{
    const float t0x = fmaf(__uint_as_float(__byte_perm(w1, w2, S.nx)), F.sx, F.bx), t1x = fmaf(__uint_as_float(__byte_perm(w1, w2, S.fx)), F.sx, F.bx);
    const float t0y = fmaf(__uint_as_float(__byte_perm(w1, w2, S.ny)), F.sy, F.by), t1y = fmaf(__uint_as_float(__byte_perm(w1, w2, S.fy)), F.sy, F.by);
    const float t0z = fmaf(__uint_as_float(__byte_perm(w1, w2, S.nz)), F.sz, F.bz), t1z = fmaf(__uint_as_float(__byte_perm(w1, w2, S.fz)), F.sz, F.bz);
    tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
    return tn <= tf * kBoxTolerance;
}

// NOTE (round 2): traverseCoop / traverseLane are the round-1 kernels (camera rays, small and mid-size waves, the tail, ray
// casts, the probe baker; the scheduled kernel in gk_trace_sched.cuh carries the large waves and reports stack overflow through
// SceneView::overflowFlag).  The only change since round 1: child pushes stop one entry short of the end, so that the
// return-to-TLAS sentinel always finds a slot - a full stack can drop a far child (a lost hit, flagged by the statistics)
// but can no longer lose the marker and index the BLAS with a TLAS entry.  Adding an overflow store to traverseLane made ptxas address the local-memory stack
// through a uniform register that the code after the loop reuses; lanes leaving the any-hit loop early then corrupted
// the stack base of the lanes still inside (compute-sanitizer: invalid __local__ read at the tuv pointer's low word +
// 8*sp).  These two functions only report a dropped entry through the traversal statistics (maxStack).
// ---- cooperative mapping: eight lanes per ray ---------------------------------------------------
// Every lane of the group passes the same ray and receives the same result.  `stackRow` points at
// the group's row of kStackSize uint2 entries in shared memory.  Dn is the NORMALISED world
// direction, hit.t must hold tmax on entry.
template <bool kAnyHit, bool kStats>
__device__ __forceinline__ bool traverseCoop(const SceneView& S, f3 O, f3 Dn, float tmin, Hit& hit, uint2* stackRow, TraversalStats* stats)
{
    const unsigned lane = threadIdx.x & 31u, sub = lane & 7u, shift = lane & 24u;
    const unsigned gmask = 0xffu << shift;
    const float tmax0 = hit.t;
    if (S.instanceCount == 0) return false;
    f3 o = O, d = Dn;
    const f3 rdWorld = boxRcp3(Dn);
    const PlaneSel selWorld = makePlaneSel(rdWorld);
    f3 rd = rdWorld;
    PlaneSel sel = selWorld;
    bool inBlas = false;
    uint32_t curInst = 0;
    uint32_t cur = S.tlasRoot;
    int sp = 0;

    for (;;) {
        if (!(cur & kLeafBit)) {
            // ---- node visit: lane `sub` tests child `sub`
            const WideNode* N = (inBlas ? S.blasNodes : S.tlasNodes) + cur;
            const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(N));
            const uint32_t* cw = reinterpret_cast<const uint32_t*>(N) + 4 + 3 * sub;
            const uint32_t ref = __ldg(cw), w1 = __ldg(cw + 1), w2 = __ldg(cw + 2);
            if (kStats && sub == 0) { stats->nodeVisits++; if (!inBlas) stats->tlasVisits++; }
            const NodeFrame F = makeNodeFrame(hdr, o, rd);
            float tn;
            const bool hitBox = childTest(F, sel, w1, w2, tmin, hit.t, tn) && (ref != kInvalid);
            const unsigned m = (__ballot_sync(gmask, hitBox) >> shift) & 0xffu;
            if (m) {
                // nearest child: min over (entry distance | lane) keys; tn >= tmin >= 0 so the bits are monotone
                const uint32_t key = __reduce_min_sync(gmask, hitBox ? ((__float_as_uint(tn) & ~7u) | sub) : 0xffffffffu);
                const unsigned near = key & 7u;
                const unsigned others = m & ~(1u << near);
                const int room = kStackSize - 1 - sp; // the last slot is reserved for the return-to-TLAS sentinel
                if (hitBox && sub != near) {
                    const int rank = __popc(others & ((1u << sub) - 1u));
                    if (rank < room) stackRow[sp + rank] = make_uint2(ref, __float_as_uint(tn));
                }
                if (kStats && sub == 0) {
                    if (__popc(others) > room) stats->maxStack = kStackSize + 1;
                    else if ((unsigned long long)(sp + __popc(others)) > stats->maxStack) stats->maxStack = sp + __popc(others);
                }
                sp += min(__popc(others), room);
                cur = __shfl_sync(gmask, ref, near + shift);
                continue;
            }
        } else if (!inBlas) {
            // ---- TLAS leaf: enter the instance (tiny_bvh.h:2305-2315); every lane transforms the ray
            const uint32_t ii = cur & 0x7fffffffu;
            const float4* ip = reinterpret_cast<const float4*>(S.inst + ii);
            const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
            const uint4 tail = __ldg(reinterpret_cast<const uint4*>(ip + 4));
            const float T[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
            if (kStats && sub == 0) stats->instanceEntries++;
            if (sub == 0 && sp < kStackSize) stackRow[sp] = make_uint2(kSentinel, 0u);
            sp = min(sp + 1, kStackSize);
            o = xformPoint(O, T);
            d = xformVector(Dn, T);
            rd = boxRcp3(d);
            sel = makePlaneSel(rd);
            inBlas = true;
            curInst = tail.y;
            cur = tail.x;
            __syncwarp(gmask);
            continue;
        } else {
            // ---- BLAS leaf: lane `sub` tests triangle `sub` of up to 8 consecutive records
            const uint32_t first = (cur & 0x7fffffffu) >> 3, cnt = (cur & 7u) + 1u;
            float t = 0.f, u = 0.f, v = 0.f;
            uint32_t prim = 0;
            bool ok = false;
            if (sub < cnt) {
                const float4* tp = reinterpret_cast<const float4*>(S.tris + first + sub);
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                TriRecord T;
                T.v0x = a.x, T.v0y = a.y, T.v0z = a.z, T.prim = __float_as_uint(a.w);
                T.e1x = b.x, T.e1y = b.y, T.e1z = b.z, T.e2x = c.x, T.e2y = c.y, T.e2z = c.z;
                ok = triangleTest(T, o, d, tmin, hit.t, t, u, v);
                prim = T.prim;
            }
            if (kStats && sub == 0) stats->triTests += cnt;
            const unsigned okm = (__ballot_sync(gmask, ok) >> shift) & 0xffu;
            if (okm) {
                if (kAnyHit) return true;
                // closest accepted triangle; equal distances resolve to the first triangle of the leaf,
                // which is what a sequential strict-less scan would keep
                const uint32_t key = ok ? __float_as_uint(t) : 0x7f800000u;
                const uint32_t kmin = __reduce_min_sync(gmask, key);
                const unsigned win = (__ballot_sync(gmask, ok && key == kmin) >> shift) & 0xffu;
                const int src = (__ffs(win) - 1) + shift;
                hit.t = __shfl_sync(gmask, t, src);
                hit.u = __shfl_sync(gmask, u, src);
                hit.v = __shfl_sync(gmask, v, src);
                hit.prim = __shfl_sync(gmask, prim, src);
                hit.inst = curInst;
            }
        }
        // ---- pop (the group barrier orders the pushes of the other lanes before these reads)
        __syncwarp(gmask);
        for (;;) {
            if (sp == 0) return hit.t < tmax0;
            --sp;
            const uint2 e = stackRow[sp];
            if (e.x == kSentinel) {
                o = O, d = Dn, rd = rdWorld, sel = selWorld;
                inBlas = false;
                continue;
            }
            if (__uint_as_float(e.y) < hit.t) {
                cur = e.x;
                break;
            }
        }
    }
}

// ---- one ray per lane ------------------------------------------------------------------------------
// Used for the large waves, where 32 rays per warp amortise the node decode.  "while-while" form:
// every lane first descends through inner nodes until it holds a leaf, then the warp processes
// leaves together.  The stack lives in local memory.  (A persistent variant that refilled finished
// lanes from a per-warp queue slice was measured and dropped: it de-phases the lanes of coherent
// waves and lost 2x on camera rays for a 5 % gain on diffuse bounces; see DESIGN.md.)
struct LaneStack {
    uint2 e[kStackSize];
};

template <bool kAnyHit, bool kStats>
__device__ __forceinline__ bool traverseLane(const SceneView& S, f3 O, f3 Dn, float tmin, Hit& hit, TraversalStats* stats)
{
    LaneStack stk;
    int sp = 0;
    const float tmax0 = hit.t;
    if (S.instanceCount == 0) return false;
    f3 o = O, d = Dn;
    f3 rd = boxRcp3(Dn);
    PlaneSel sel = makePlaneSel(rd);
    bool inBlas = false;
    uint32_t curInst = 0;
    uint32_t cur = S.tlasRoot;

    // pops the next live entry into `cur` (kInvalid when the stack is empty)
#define GK_POP()                                                                                          \
    for (;;) {                                                                                            \
        if (sp == 0) { cur = kInvalid; break; }                                                           \
        --sp;                                                                                             \
        const uint2 e_ = stk.e[sp];                                                                       \
        if (e_.x == kSentinel) { o = O, d = Dn, rd = boxRcp3(Dn), sel = makePlaneSel(rd), inBlas = false; continue; } \
        if (__uint_as_float(e_.y) < hit.t) { cur = e_.x; break; }                                         \
    }

    for (;;) {
        // ---- (1) descend through inner nodes until this lane holds a leaf
        while (!(cur & kLeafBit)) {
            const uint4* np = reinterpret_cast<const uint4*>((inBlas ? S.blasNodes : S.tlasNodes) + cur);
            const uint4 hdr = __ldg(np);
            if (kStats) { stats->nodeVisits++; if (!inBlas) stats->tlasVisits++; }
            const uint32_t count = hdr.w >> 24;
            const NodeFrame F = makeNodeFrame(hdr, o, rd);
            float bestT = kFar;
            uint32_t bestRef = kInvalid;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half == 1 && count <= 4) break;
                // 4 child records = 12 words = 3 x 128-bit loads
                const uint4 q0 = __ldg(np + 1 + 3 * half), q1 = __ldg(np + 2 + 3 * half), q2 = __ldg(np + 3 + 3 * half);
                const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t ref = w[3 * k];
                    float tn;
                    if (childTest(F, sel, w[3 * k + 1], w[3 * k + 2], tmin, hit.t, tn) && (ref != kInvalid)) {
                        // keep the nearest child in registers, push the other one
                        uint32_t pr = ref;
                        float pt = tn;
                        if (tn < bestT) { pr = bestRef, pt = bestT, bestRef = ref, bestT = tn; }
                        if (pr != kInvalid && sp < kStackSize - 1) stk.e[sp++] = make_uint2(pr, __float_as_uint(pt)); // last slot: sentinel only
                        else if (kStats && pr != kInvalid) stats->maxStack = kStackSize + 1; // an entry was dropped
                    }
                }
            }
            if (kStats && (unsigned long long)sp > stats->maxStack) stats->maxStack = sp;
            if (bestRef != kInvalid) cur = bestRef;
            else { GK_POP() }
        }
        if (cur == kInvalid) break;
        // ---- (2) one leaf step
        if (!inBlas) {
            const uint32_t ii = cur & 0x7fffffffu;
            const float4* ip = reinterpret_cast<const float4*>(S.inst + ii);
            const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
            const uint4 tail = __ldg(reinterpret_cast<const uint4*>(ip + 4));
            const float T[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
            if (kStats) stats->instanceEntries++;
            if (sp < kStackSize) stk.e[sp++] = make_uint2(kSentinel, 0u);
            o = xformPoint(O, T);
            d = xformVector(Dn, T);
            rd = boxRcp3(d);
            sel = makePlaneSel(rd);
            inBlas = true;
            curInst = tail.y;
            cur = tail.x;
            continue;
        }
        {
            const uint32_t first = (cur & 0x7fffffffu) >> 3, cnt = (cur & 7u) + 1u;
            for (uint32_t k = 0; k < cnt; ++k) {
                const float4* tp = reinterpret_cast<const float4*>(S.tris + first + k);
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                TriRecord T;
                T.v0x = a.x, T.v0y = a.y, T.v0z = a.z, T.prim = __float_as_uint(a.w);
                T.e1x = b.x, T.e1y = b.y, T.e1z = b.z, T.e2x = c.x, T.e2y = c.y, T.e2z = c.z;
                float t, u, v;
                if (kStats) stats->triTests++;
                if (triangleTest(T, o, d, tmin, hit.t, t, u, v)) {
                    if (kAnyHit) return true;
                    hit.t = t, hit.u = u, hit.v = v, hit.prim = T.prim, hit.inst = curInst;
                }
            }
        }
        GK_POP()
        if (cur == kInvalid) break;
    }
#undef GK_POP
    return hit.t < tmax0;
}

#endif // __CUDACC__

} // namespace gk
