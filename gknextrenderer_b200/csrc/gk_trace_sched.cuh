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
constexpr int kSmemStack = 24;    // stack entries per lane held in shared memory (kStackSize in total)
constexpr uint32_t kFetchChunk = 128; // rays a warp takes from the queue per atomic
constexpr uint32_t kNone = 0xffffffffu;

struct SchedParams {
    uint32_t refillMin;  // refill the warp's idle lanes once this many have finished (1..32)
    uint32_t biasN;      // vote: the node class wins against another class unless that one is larger by more than this many lanes
};

struct SchedStats { // per launch, summed over warps (kStats only)
    unsigned long long iters[3];  // steps run per class N, T, I
    unsigned long long lanes[3];  // lanes that took part
    unsigned long long refills, refillLanes;
    unsigned long long popIters, popLanes;
    unsigned long long overflow;  // stack entries dropped (must stay 0)
};

// Shared-memory stack access by 32-bit shared-space address (one address register, immediate offset for the key half):
// plain C++ indexing made the compiler rebuild the base address and branch around every push.
__device__ __forceinline__ void stackStore(uint32_t addr, uint32_t ref, uint32_t key)
{
    asm volatile("st.shared.u32 [%0], %1;\n\tst.shared.u32 [%0+%3], %2;" ::"r"(addr), "r"(ref), "r"(key), "n"(kSmemStack * kSchedBlock * 4));
}
__device__ __forceinline__ void stackLoad(uint32_t addr, uint32_t& ref, uint32_t& key)
{
    asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+%3];" : "=r"(ref), "=r"(key) : "r"(addr), "n"(kSmemStack * kSchedBlock * 4));
}

// One warp-scheduled traversal over `count` rays of `io`.  cursor: device counter, zero at launch.
// Stack addressing: entry e of the lane lives at word e * kSchedBlock + tid of `sStack` (references) and kSmemStack * kSchedBlock
// words further (keys); the lane keeps `top`, the shared-space byte address of its next free entry, so a push is two
// stores and one add, and "empty" is top == bottom.
template <bool kAnyHit, bool kStats, class RayIO>
__device__ __forceinline__ void traverseScheduled(const SceneView& V, const RayIO& io, uint32_t count, uint32_t* __restrict__ cursor, const SchedParams prm,
                                                  uint32_t* sStack, TraversalStats* stats, SchedStats* sched)
{
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    constexpr uint32_t kEntryStride = kSchedBlock * 4; // bytes between consecutive entries of a lane
    const uint32_t bottom = (uint32_t)__cvta_generic_to_shared(sStack) + tid * 4u;
    const uint32_t smemLimit = bottom + kSmemStack * kEntryStride;       // first entry outside shared memory
    const uint32_t deepLimit = bottom + (kSmemStack - 8) * kEntryStride; // a node step stores at most 8 entries (7 pushes + one dead store)
    const uint32_t hardLimit = bottom + kStackSize * kEntryStride;
    // ---- lane state
    bool alive = false, inBlas = false;
    uint32_t cur = kNone, curInst = 0, rayIdx = 0;
    uint32_t top = bottom, topBase = bottom;
    f3 O = mk3(0, 0, 0), Dn = mk3(0, 0, 1), o = O, d = Dn, rd = mk3(0, 0, 0);
    PlaneSel sel = makePlaneSel(rd);
    float tmin = 0.f;
    Hit hit{0.f, 0.f, 0.f, kInvalid, kInvalid};
    uint2 spill[kStackSize - kSmemStack];
    // ---- warp state (uniform)
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

    for (;;) {
        // ------------------------------------------------------------ refill idle lanes
        const unsigned idle = __ballot_sync(full, !alive);
        if (!exhausted && (uint32_t)__popc(idle) >= prm.refillMin) {
            if (wNext == wEnd) { // take the next chunk of the queue
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, kFetchChunk);
                base = __shfl_sync(full, base, 0);
                if (base >= count) exhausted = true;
                else wNext = base, wEnd = min(base + kFetchChunk, count);
            }
            if (!exhausted) {
                const uint32_t want = (uint32_t)__popc(idle), avail = wEnd - wNext, take = min(want, avail);
                const uint32_t rank = (uint32_t)__popc(idle & ((1u << lane) - 1u));
                if (!alive && rank < take) {
                    rayIdx = wNext + rank;
                    f3 D;
                    float tmax;
                    const bool live = io.load(rayIdx, O, D, tmin, tmax);
                    hit.t = tmax, hit.u = hit.v = 0.f, hit.prim = kInvalid, hit.inst = kInvalid;
                    if (live) {
                        Dn = normalizeRayDir(D);
                        o = O, d = Dn, rd = boxRcp3(Dn), sel = makePlaneSel(rd);
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
        const unsigned mN = __ballot_sync(full, isN), mT = __ballot_sync(full, isT), mI = __ballot_sync(full, isI);
        if (!(mN | mT | mI)) {
            if (exhausted) break;
            continue; // every lane idle: the refill above runs next round (refillMin <= 32)
        }
        const int cN = __popc(mN), cT = __popc(mT), cI = __popc(mI);
        int phase; // 0 node, 1 triangle, 2 instance
        if (cN && cN + (int)prm.biasN >= cT && cN + (int)prm.biasN >= cI) phase = 0;
        else phase = (cT >= cI) ? 1 : 2;
        if (kStats) stIt[phase]++, stLn[phase] += (phase == 0 ? cN : phase == 1 ? cT : cI);

        bool needPop = false, done = false, occluded = false;
        if (phase == 0) {
            // ---- node step: eight slab tests, the nearest child becomes `cur`, the others are pushed
            const bool deepAny = __any_sync(full, isN && top >= deepLimit); // warp-uniform choice of the push flavour
            if (isN) {
                const uint4* np = reinterpret_cast<const uint4*>((inBlas ? V.blasNodes : V.tlasNodes) + cur);
                const uint4 hdr = __ldg(np);
                if (kStats) { local.nodeVisits++; if (!inBlas) local.tlasVisits++; }
                const NodeFrame F = makeNodeFrame(hdr, o, rd);
                const bool wide = (hdr.w >> 24) > 4u;
                const bool anyWide = __any_sync(mN, wide); // mN = exactly the lanes inside this branch
                uint32_t bestKey[2] = {kNone, kNone}, bestRef[2] = {kNone, kNone};
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
                        const bool h = childTest(F, sel, w[3 * k + 1], w[3 * k + 2], tmin, hit.t, tn) && w[3 * k] != kInvalid && present;
                        // tn >= tmin >= 0: its bits order like the value; the slot number makes keys unique
                        key[k] = h ? ((__float_as_uint(tn) & ~7u) | (uint32_t)(4 * half + k)) : kNone;
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
                else needPop = true;
            }
        } else if (phase == 1) {
            if (isT) {
                // ---- one triangle of the leaf in hand: cur = leaf | first << 3 | (triangles left after this one)
                if (cur == kInvalid) needPop = true; // BLAS root of a hidden instance (never reached: its box is empty)
                else {
                    const uint32_t first = (cur & 0x7fffffffu) >> 3;
                    const float4* tp = reinterpret_cast<const float4*>(V.tris + first);
                    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                    TriRecord T;
                    T.v0x = a.x, T.v0y = a.y, T.v0z = a.z, T.prim = __float_as_uint(a.w);
                    T.e1x = b.x, T.e1y = b.y, T.e1z = b.z, T.e2x = c.x, T.e2y = c.y, T.e2z = c.z;
                    float t, u, v;
                    if (kStats) local.triTests++;
                    if (triangleTest(T, o, d, tmin, hit.t, t, u, v)) {
                        if (kAnyHit) done = true, occluded = true;
                        hit.t = t, hit.u = u, hit.v = v, hit.prim = T.prim, hit.inst = curInst;
                    }
                    if (cur & 7u) cur += 7u; // next record, one fewer left
                    else needPop = true;
                }
            }
        } else {
            if (isI) {
                // ---- enter the instance (tiny_bvh.h:2305-2315): the ray goes to instance space, t stays world-space
                const uint32_t ii = cur & 0x7fffffffu;
                const float4* ip = reinterpret_cast<const float4*>(V.inst + ii);
                const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
                const uint4 tail = __ldg(reinterpret_cast<const uint4*>(ip + 4));
                const float T[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
                if (kStats) local.instanceEntries++;
                o = xformPoint(O, T);
                d = xformVector(Dn, T);
                rd = boxRcp3(d);
                sel = makePlaneSel(rd);
                inBlas = true, topBase = top;
                curInst = tail.y;
                cur = tail.x;
            }
        }
        // ------------------------------------------------------------ pop the next live entry
        if (kStats) {
            const unsigned pm = __ballot_sync(full, needPop && !done);
            if (pm) ++stPi, stPl += __popc(pm);
        }
        if (needPop && !done) {
            for (;;) {
                if (inBlas && top == topBase) { // the instance is finished: back to the world-space ray
                    o = O, d = Dn, rd = boxRcp3(Dn), sel = makePlaneSel(rd), inBlas = false;
                }
                if (top == bottom) {
                    done = true;
                    break;
                }
                top -= kEntryStride;
                uint32_t r, key;
                if (top < smemLimit) stackLoad(top, r, key);
                else r = spill[(top - smemLimit) / kEntryStride].x, key = spill[(top - smemLimit) / kEntryStride].y;
                if (__uint_as_float(key & ~7u) < hit.t) {
                    cur = r;
                    break;
                }
            }
        }
        if (done) {
            io.store(rayIdx, hit, occluded);
            alive = false;
            cur = kNone;
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
