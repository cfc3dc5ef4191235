// gk_trace_sched.cuh — the scheduled traversal kernel: persistent warps, dynamic ray fetch and a
// per-warp vote that decides which of the three traversal steps the warp runs next.
//
// Same job and same arithmetic as traverseLane (gk_bvh.cuh): closest / any hit over the two-level
// 8-wide BVH with tinybvh's exact triangle test (tiny_bvh.h:2245-2353, 6815-6843), replacing the
// RayQuery of Shading.slang:659-758.  What changes is how a warp spends its issue slots.
//
// Evidence (profiles/r01_k_trace_full.md + the per-instruction ncu page of the same capture): on a
// bounce wave the while-while kernel issued 52 node-loop iterations per warp for rays that need 17
// node visits each: 17 of 32 lanes were still alive on average (rays of a warp finish at different
// times) and 10.5 of those were in the node loop (the others waited with a leaf in hand); the
// per-child push blocks ran at 4 lanes, triangle tests at 5.  Useful lane throughput: 25 %.
//
// This kernel keeps one ray per lane but
//   * is persistent: a warp takes rays from the wave's queue through a cursor (one atomic per 128
//     rays), and refills its idle lanes whenever `refillMin` of them have finished: lanes do not
//     wait for the slowest ray of "their" warp;
//   * gives every lane an explicit state — N: holds an inner node, T: holds a triangle of a BLAS
//     leaf, I: holds a TLAS leaf (instance to enter) — and every iteration the warp votes
//     (three ballots) and runs ONE step for the largest class.  Lanes of the other classes keep
//     their item; classes fill up until they win.  A step therefore runs at the occupancy of
//     the biggest class instead of at whatever is left inside nested divergent loops;
//   * tests ONE triangle per T step (a leaf of k triangles is k steps), so leaves of different
//     sizes do not idle each other;
//   * has no divergent block per child: the eight slab tests produce sortable keys
//     (entry-distance bits | slot), a min picks the nearest child, the others are pushed with
//     predicated shared-memory stores;
//   * keeps the stack in shared memory, [entry][lane] layout: conflict-free whatever the depth
//     of the individual lanes (the local-memory stack of the lane kernel turns divergent depths
//     into up to 32 L1 wavefronts per access).  Entries beyond kSmemStack spill to local memory;
//   * marks "return to the TLAS" by the stack depth at instance entry instead of a sentinel entry.
//
// Each lane still walks its own ray's items in an order that depends on that ray alone (the vote
// only delays), so results do not depend on which rays share a warp: frames stay deterministic and
// tile partitions compose bit-exactly.
#pragma once
#include "gk_bvh.cuh"

namespace gk {

#ifdef __CUDACC__

constexpr int kSchedBlock = 128;  // threads per block
constexpr int kSmemStack = 20;    // stack entries per lane held in shared memory (kStackSize in total)
constexpr int kWorldWords = 9;    // world-space ray kept in shared memory per lane: origin, normalised direction, box reciprocal
constexpr uint32_t kFetchChunk = 128; // most rays a warp takes from the queue per atomic (small waves: fewer, see fetchChunk)
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kSchedSmemWords = (2 * kSmemStack + kWorldWords) * kSchedBlock;

struct SchedParams {
    uint32_t refillMin;  // refill the warp's idle lanes once this many have finished (1..32)
    uint32_t biasN;      // vote: the node class wins against another class unless that one is larger by more than this many lanes
    uint32_t keepN;      // a node phase repeats node steps while at least this many lanes hold a node
    uint32_t keepT;      // a triangle phase repeats while at least this many lanes hold a triangle
};

struct SchedStats { // per launch, summed over warps (kStats only)
    unsigned long long iters[3];  // steps run per class N, T, I
    unsigned long long lanes[3];  // lanes that took part
    unsigned long long refills, refillLanes;
    unsigned long long popIters, popLanes; // votes taken, lanes alive at the vote
    unsigned long long overflow;  // stack entries dropped (must stay 0)
};

// Shared-memory access by 32-bit shared-space address (one address register + immediate offsets): plain C++ indexing made
// the compiler rebuild the base address and branch around every push.
__device__ __forceinline__ void stackStore(uint32_t addr, uint32_t ref, uint32_t key)
{
    asm volatile("st.shared.u32 [%0], %1;\n\tst.shared.u32 [%0+%3], %2;" ::"r"(addr), "r"(ref), "r"(key), "n"(kSmemStack * kSchedBlock * 4));
}
__device__ __forceinline__ void stackStoreRef(uint32_t addr, uint32_t ref) // occlusion queries: entries carry no key
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(ref));
}
__device__ __forceinline__ uint32_t stackLoadRef(uint32_t addr)
{
    uint32_t ref;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ref) : "r"(addr));
    return ref;
}
__device__ __forceinline__ void stackLoad(uint32_t addr, uint32_t& ref, uint32_t& key)
{
    asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+%3];" : "=r"(ref), "=r"(key) : "r"(addr), "n"(kSmemStack * kSchedBlock * 4));
}
template <int kWord> __device__ __forceinline__ void worldStore(uint32_t addr, float v)
{
    asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(addr), "f"(v), "n"(kWord * kSchedBlock * 4));
}
template <int kWord> __device__ __forceinline__ float worldLoad(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(kWord * kSchedBlock * 4));
    return v;
}
// (a & b) | c in one LOP3 (the compiler splits it in two when b and c are both immediates)
__device__ __forceinline__ uint32_t andOr(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Slab test of the scheduled kernel: as childTest (gk_bvh.cuh) without the tolerance multiply.  The builder pads every
// child box by one whole quantisation step on both sides (k_quantise); the decode error of the fused t = (2^23+q)*s + b
// with a MUFU reciprocal is below 0.6 step for any origin within 2^18 steps of the node (chooseExponent guarantees it).
__device__ __forceinline__ bool childTestPadded(const NodeFrame& F, const PlaneSel& S, uint32_t w1, uint32_t w2, float tmin, float tmax, float& tn)
{
    const float t0x = fmaf(__uint_as_float(__byte_perm(w1, w2, S.nx)), F.sx, F.bx), t1x = fmaf(__uint_as_float(__byte_perm(w1, w2, S.fx)), F.sx, F.bx);
    const float t0y = fmaf(__uint_as_float(__byte_perm(w1, w2, S.ny)), F.sy, F.by), t1y = fmaf(__uint_as_float(__byte_perm(w1, w2, S.fy)), F.sy, F.by);
    const float t0z = fmaf(__uint_as_float(__byte_perm(w1, w2, S.nz)), F.sz, F.bz), t1z = fmaf(__uint_as_float(__byte_perm(w1, w2, S.fz)), F.sz, F.bz);
    tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
    return tn <= tf;
}

// One warp-scheduled traversal over `count` rays of `io`.  cursor: device counter, zero at launch.
// Shared memory per lane (word w of lane `tid` lives at sMem[w * kSchedBlock + tid], conflict-free for any mix of depths):
//   words [0, kSmemStack)              stack references      } the lane keeps `top`, the shared-space byte address of its next
//   words [kSmemStack, 2 kSmemStack)   stack keys            } free entry: a push is two stores and one add, "empty" is top == bottom
//   words [2 kSmemStack, +9)           world-space ray: O, Dn, 1/Dn (read when an instance is entered or left)
#ifndef GK_ANYHIT_UNORDERED
#define GK_ANYHIT_UNORDERED 1
#endif
constexpr bool kAnyHitUnordered = GK_ANYHIT_UNORDERED != 0;
template <bool kAnyHit, bool kStats, class RayIO>
__device__ __forceinline__ void traverseScheduled(const SceneView& V, const RayIO& io, uint32_t count, uint32_t* __restrict__ cursor, const SchedParams prm,
                                                  uint32_t* sMem, TraversalStats* stats, SchedStats* sched)
{
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    constexpr uint32_t kEntryStride = kSchedBlock * 4; // bytes between consecutive words of a lane
    const uint32_t bottom = (uint32_t)__cvta_generic_to_shared(sMem) + tid * 4u;
    const uint32_t world = bottom + 2 * kSmemStack * kEntryStride;
    const uint32_t smemLimit = bottom + kSmemStack * kEntryStride;       // first entry outside shared memory
    const uint32_t deepLimit = bottom + (kSmemStack - 8) * kEntryStride; // a node step stores at most 8 entries (7 pushes + one dead store)
    const uint32_t hardLimit = bottom + kStackSize * kEntryStride;
    // ---- lane state
    bool alive = false, inBlas = false;
    uint32_t cur = kNone, curInst = 0, rayIdx = 0;
    uint32_t top = bottom, topBase = bottom;
    f3 o = mk3(0, 0, 0), d = mk3(0, 0, 1), rd = mk3(0, 0, 0);
    PlaneSel sel = makePlaneSel(rd);
    float tmin = 0.f;
    Hit hit{0.f, 0.f, 0.f, kInvalid, kInvalid};
    uint2 spill[kStackSize - kSmemStack];
    // ---- warp state (uniform)
    // Rays per queue fetch: 128 for big waves; a small wave is spread over all resident warps instead (a warp that took 128 of
    // 138 rays would trace them in four rounds while a thousand warps idle: measured 0.29 ms for a 138-ray wave).
    const uint32_t warpsInGrid = gridDim.x * (kSchedBlock / 32);
    const uint32_t share = count / warpsInGrid;
    const uint32_t fetchChunk = share >= kFetchChunk ? kFetchChunk : share >= 32u ? (share & ~31u) : (share > 0u ? share : 1u);
    uint32_t wNext = 0, wEnd = 0;
    bool exhausted = (count == 0);
    unsigned long long stIt[3] = {0, 0, 0}, stLn[3] = {0, 0, 0}, stRf = 0, stRl = 0, stPi = 0, stPl = 0, stOv = 0;
    TraversalStats local{0, 0, 0, 0, 0};

    if (V.instanceCount == 0) { // empty scene: every ray misses
        for (uint32_t i = blockIdx.x * blockDim.x + tid; i < count; i += gridDim.x * blockDim.x) {
            f3 a, b;
            float t0, t1;
            io.load(i, a, b, t0, t1);
            Hit h{t1, 0.f, 0.f, kInvalid, kInvalid};
            io.store(i, h, false);
        }
        return;
    }

    // stack entry = {reference, key}; key = entry-distance bits with the child slot in the low three bits (masked off on pop)
    auto pushDeep = [&](uint32_t ref, uint32_t key) { // any depth: shared memory, then the local-memory spill, then "dropped"
        if (top < smemLimit) stackStore(top, ref, key), top += kEntryStride;
        else if (top < hardLimit) spill[(top - smemLimit) / kEntryStride] = make_uint2(ref, key), top += kEntryStride;
        else {
            if (kStats) ++stOv;
            *V.overflowFlag = 1u; // a dropped entry is a possibly missed hit: the host turns the flag into an error
        }
    };
    // Pops the next live entry into `cur`; an empty stack finishes the ray.  Leaving an instance restores the world-space ray.
    auto popOrFinish = [&](bool occluded) {
        bool done = occluded;
        while (!done) {
            if (inBlas && top == topBase) { // the instance is finished: back to the world-space ray
                o = mk3(worldLoad<0>(world), worldLoad<1>(world), worldLoad<2>(world));
                rd = mk3(worldLoad<6>(world), worldLoad<7>(world), worldLoad<8>(world));
                sel = makePlaneSel(rd), inBlas = false;
            }
            if (top == bottom) {
                done = true;
                break;
            }
            top -= kEntryStride;
            if (kAnyHit && kAnyHitUnordered) { // no distance to cull by: hit.t stays tmax until the ray is occluded, and then it is finished
                cur = top < smemLimit ? stackLoadRef(top) : spill[(top - smemLimit) / kEntryStride].x;
                return;
            }
            uint32_t r, key;
            if (top < smemLimit) stackLoad(top, r, key);
            else r = spill[(top - smemLimit) / kEntryStride].x, key = spill[(top - smemLimit) / kEntryStride].y;
            if (__uint_as_float(key & ~7u) < hit.t) {
                cur = r;
                return;
            }
        }
        io.store(rayIdx, hit, occluded);
        alive = false;
        cur = kNone;
    };

    for (;;) {
        // ------------------------------------------------------------ refill idle lanes
        const unsigned idle = __ballot_sync(full, !alive);
        if (!exhausted && (uint32_t)__popc(idle) >= prm.refillMin) {
            if (wNext == wEnd) { // take the next chunk of the queue
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, fetchChunk);
                base = __shfl_sync(full, base, 0);
                if (base >= count) exhausted = true;
                else wNext = base, wEnd = min(base + fetchChunk, count);
            }
            if (!exhausted) {
                const uint32_t want = (uint32_t)__popc(idle), avail = wEnd - wNext, take = min(want, avail);
                const uint32_t rank = (uint32_t)__popc(idle & ((1u << lane) - 1u));
                if (!alive && rank < take) {
                    rayIdx = wNext + rank;
                    f3 O, D;
                    float tmax;
                    const bool live = io.load(rayIdx, O, D, tmin, tmax);
                    hit.t = tmax, hit.u = hit.v = 0.f, hit.prim = kInvalid, hit.inst = kInvalid;
                    if (live) {
                        const f3 Dn = normalizeRayDir(D);
                        o = O, d = Dn, rd = boxRcp3(Dn), sel = makePlaneSel(rd);
                        worldStore<0>(world, O.x), worldStore<1>(world, O.y), worldStore<2>(world, O.z);
                        worldStore<3>(world, Dn.x), worldStore<4>(world, Dn.y), worldStore<5>(world, Dn.z);
                        worldStore<6>(world, rd.x), worldStore<7>(world, rd.y), worldStore<8>(world, rd.z);
                        inBlas = false, top = bottom, topBase = bottom, cur = V.tlasRoot, alive = true;
                    } else io.store(rayIdx, hit, false);
                }
                wNext += take;
                if (kStats) ++stRf, stRl += take;
            }
        }
        // ------------------------------------------------------------ vote
        const bool leafish = (cur & kLeafBit) != 0;
        const bool isN = alive && !leafish, isT = alive && leafish && inBlas, isI = alive && leafish && !inBlas;
        unsigned mN = __ballot_sync(full, isN), mT = __ballot_sync(full, isT);
        const unsigned mI = __ballot_sync(full, isI);
        if (!(mN | mT | mI)) {
            if (exhausted) break;
            continue; // every lane idle: the refill above runs next round (refillMin <= 32)
        }
        const int cN = __popc(mN), cT = __popc(mT), cI = __popc(mI);
        int phase; // 0 node, 1 triangle, 2 instance
        if (cN && cN + (int)prm.biasN >= cT && cN + (int)prm.biasN >= cI) phase = 0;
        else phase = (cT >= cI) ? 1 : 2;
        if (kStats) ++stPi, stPl += cN + cT + cI;

        if (phase == 0) {
            // ---- node phase: node steps while enough lanes hold a node.  A step = eight slab tests, the nearest child becomes
            //      `cur`, the others are pushed; a lane without a hit child pops; a lane that receives a leaf waits.
            do {
                const bool mine = (mN >> lane) & 1u;
                const bool deepAny = __any_sync(full, mine && top >= deepLimit); // warp-uniform choice of the push flavour
                if (kStats) stIt[0]++, stLn[0] += __popc(mN);
                if (mine) {
                    const uint4* np = reinterpret_cast<const uint4*>((inBlas ? V.blasNodes : V.tlasNodes) + cur);
                    const uint4 hdr = __ldg(np);
                    if (kStats) { local.nodeVisits++; if (!inBlas) local.tlasVisits++; }
                    const NodeFrame F = makeNodeFrame(hdr, o, rd);
                    const bool wide = (hdr.w >> 24) > 4u;
                    const bool anyWide = __any_sync(mN, wide); // mN = exactly the lanes inside this branch
                    if (kAnyHit && kAnyHitUnordered) {
                        // Occlusion query: the order in which the hit children are visited cannot change the answer, so no sort keys,
                        // no nearest-first selection and one-word stack entries: the first hit child (slot order) continues, the others are pushed.
                        uint32_t next = kNone;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            if (half == 1 && !anyWide) break;
                            const bool present = half == 0 || wide;
                            uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0, q2 = q0;
                            if (present) q0 = __ldg(np + 1 + 3 * half), q1 = __ldg(np + 2 + 3 * half), q2 = __ldg(np + 3 + 3 * half);
                            const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                float tn;
                                const bool h = childTestPadded(F, sel, w[3 * k + 1], w[3 * k + 2], tmin, hit.t, tn) && w[3 * k] != kInvalid && present;
                                const bool push = h && next != kNone;
                                if (!deepAny) {
                                    stackStoreRef(top, w[3 * k]);
                                    top += push ? kEntryStride : 0u;
                                } else if (push) pushDeep(w[3 * k], 0u);
                                next = (h && next == kNone) ? w[3 * k] : next;
                            }
                        }
                        if (kStats) {
                            const unsigned long long depth = (top - bottom) / kEntryStride + 1;
                            if (depth > local.maxStack) local.maxStack = depth;
                        }
                        if (next != kNone) cur = next;
                        else popOrFinish(false);
                    } else {
                    uint32_t bestKey[2] = {kNone, kNone}, bestRef[2] = {kNone, kNone};
                    const uint32_t keyMask = ~7u;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (half == 1 && !anyWide) break; // warp-uniform: no lane of this step has more than four children
                        const bool present = half == 0 || wide;
                        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0, q2 = q0;
                        if (present) q0 = __ldg(np + 1 + 3 * half), q1 = __ldg(np + 2 + 3 * half), q2 = __ldg(np + 3 + 3 * half);
                        const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
                        uint32_t key[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float tn;
                            const bool h = childTestPadded(F, sel, w[3 * k + 1], w[3 * k + 2], tmin, hit.t, tn) && w[3 * k] != kInvalid && present;
                            // tn >= tmin >= 0: its bits order like the value; the slot number makes keys unique
                            key[k] = h ? andOr(__float_as_uint(tn), keyMask, (uint32_t)(4 * half + k)) : kNone;
                        }
                        const uint32_t m = min(min(key[0], key[1]), min(key[2], key[3]));
                        uint32_t r = kNone;
                        if (!deepAny) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (key[k] == m) r = w[3 * k];
                                // branch-free push: the entry is always written above the top, the top moves only for a real push
                                stackStore(top, w[3 * k], key[k]);
                                top += (key[k] != m && key[k] != kNone) ? kEntryStride : 0u;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (key[k] == m) r = w[3 * k];
                                else if (key[k] != kNone) pushDeep(w[3 * k], key[k]);
                            }
                        }
                        bestKey[half] = m, bestRef[half] = r;
                    }
                    // nearest of the two halves continues, the other one (if any) is pushed
                    const bool firstWins = bestKey[0] <= bestKey[1];
                    const uint32_t winKey = firstWins ? bestKey[0] : bestKey[1], winRef = firstWins ? bestRef[0] : bestRef[1];
                    const uint32_t loseKey = firstWins ? bestKey[1] : bestKey[0], loseRef = firstWins ? bestRef[1] : bestRef[0];
                    if (!deepAny) {
                        stackStore(top, loseRef, loseKey);
                        top += (loseKey != kNone) ? kEntryStride : 0u;
                    } else if (loseKey != kNone) pushDeep(loseRef, loseKey);
                    if (kStats) {
                        const unsigned long long depth = (top - bottom) / kEntryStride + 1;
                        if (depth > local.maxStack) local.maxStack = depth;
                    }
                    if (winKey != kNone) cur = winRef;
                    else popOrFinish(false);
                    }
                }
                mN = __ballot_sync(full, alive && !(cur & kLeafBit));
            } while ((uint32_t)__popc(mN) >= prm.keepN);
        } else if (phase == 1) {
            // ---- triangle phase: one triangle of the leaf in hand per step; cur = leaf | first << 3 | (triangles left after this one)
            do {
                const bool mine = (mT >> lane) & 1u;
                if (kStats) stIt[1]++, stLn[1] += __popc(mT);
                if (mine) {
                    if (cur == kInvalid) popOrFinish(false); // BLAS root of a hidden instance (never reached: its box is empty)
                    else {
                        const uint32_t first = (cur & 0x7fffffffu) >> 3;
                        const float4* tp = reinterpret_cast<const float4*>(V.tris + first);
                        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                        TriRecord T;
                        T.v0x = a.x, T.v0y = a.y, T.v0z = a.z, T.prim = __float_as_uint(a.w);
                        T.e1x = b.x, T.e1y = b.y, T.e1z = b.z, T.e2x = c.x, T.e2y = c.y, T.e2z = c.z;
                        float t, u, v;
                        bool occluded = false;
                        if (kStats) local.triTests++;
                        if (triangleTest(T, o, d, tmin, hit.t, t, u, v)) {
                            if (kAnyHit) occluded = true;
                            hit.t = t, hit.u = u, hit.v = v, hit.prim = T.prim, hit.inst = curInst;
                        }
                        if ((cur & 7u) && !occluded) cur += 7u; // next record, one fewer left
                        else popOrFinish(occluded);
                    }
                }
                mT = __ballot_sync(full, alive && (cur & kLeafBit) && inBlas);
            } while ((uint32_t)__popc(mT) >= prm.keepT);
        } else {
            if (kStats) stIt[2]++, stLn[2] += cI;
            if (isI) {
                // ---- enter the instance (tiny_bvh.h:2305-2315): the ray goes to instance space, t stays world-space
                const uint32_t ii = cur & 0x7fffffffu;
                const float4* ip = reinterpret_cast<const float4*>(V.inst + ii);
                const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
                const uint4 tail = __ldg(reinterpret_cast<const uint4*>(ip + 4));
                const float T[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
                if (kStats) local.instanceEntries++;
                const f3 O = mk3(worldLoad<0>(world), worldLoad<1>(world), worldLoad<2>(world));
                const f3 Dn = mk3(worldLoad<3>(world), worldLoad<4>(world), worldLoad<5>(world));
                o = xformPoint(O, T);
                d = xformVector(Dn, T);
                rd = boxRcp3(d);
                sel = makePlaneSel(rd);
                inBlas = true, topBase = top;
                curInst = tail.y;
                cur = tail.x;
            }
        }
    }
    if (kStats) {
        for (int o2 = 16; o2; o2 >>= 1) {
            local.nodeVisits += __shfl_xor_sync(full, local.nodeVisits, o2), local.triTests += __shfl_xor_sync(full, local.triTests, o2);
            local.tlasVisits += __shfl_xor_sync(full, local.tlasVisits, o2), local.instanceEntries += __shfl_xor_sync(full, local.instanceEntries, o2);
            local.maxStack = max(local.maxStack, __shfl_xor_sync(full, local.maxStack, o2));
            stOv += __shfl_xor_sync(full, stOv, o2);
        }
        if (lane == 0) {
            atomicAdd(&stats->nodeVisits, local.nodeVisits), atomicAdd(&stats->triTests, local.triTests);
            atomicAdd(&stats->tlasVisits, local.tlasVisits), atomicAdd(&stats->instanceEntries, local.instanceEntries);
            atomicMax(&stats->maxStack, stOv ? (unsigned long long)kStackSize + 1 : local.maxStack);
            if (sched) {
                for (int k = 0; k < 3; ++k) atomicAdd(&sched->iters[k], stIt[k]), atomicAdd(&sched->lanes[k], stLn[k]);
                atomicAdd(&sched->refills, stRf), atomicAdd(&sched->refillLanes, stRl);
                atomicAdd(&sched->popIters, stPi), atomicAdd(&sched->popLanes, stPl);
                atomicAdd(&sched->overflow, stOv);
            }
        }
    }
}

#endif // __CUDACC__

} // namespace gk
