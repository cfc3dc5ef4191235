// gk_bvh_build.cu — GPU construction and refit of the acceleration structures.
//
// Replaces, on the device:
//   vkCmdBuildAccelerationStructuresKHR for the BLAS set ..... RayTraceBaseRenderer.cpp:298-362
//   the per-dirty-frame TLAS rebuild ......................... RayTraceBaseRenderer.cpp:176-228,
//                                                             TopLevelAccelerationStructure.cpp:73-111
//   tinybvh BVH::Build / BLASInstance::Update on the CPU side  tiny_bvh.h:1565-1766, 6718-6758
//
// Pipeline (all kernels on the context stream):
//   1. group bounds      one atomic min/max box per model (BLAS forest) or for the scene (TLAS)
//   2. Morton keys       64-bit key = group << 42 | 42-bit Morton code of the box centre
//   3. radix sort        cub::DeviceRadixSort (library call; the one non-hand-written step)
//   4. Karras 2012       binary radix tree over the sorted keys, every model at once: a model's
//                        keys share the group prefix, so the tree contains one subtree per model
//   5. bounds            bottom-up box propagation with arrival counters
//   6. collapse          top-down, level-synchronous: each 8-wide node absorbs the binary nodes
//                        with the largest surface area until it has 8 children
//   7. quantise          child boxes -> 8 bits against the node box (also the refit path)
// Refit (dynamic scenes) re-runs 5 and 7 only.
//
// Bytes per primitive (roofline model): keys+indices 12 B x (2 + 2x radix passes),
// boxes 32 B r/w, binary node 40 B, wide node 128 B / ~3 prims.
#include "gk_context.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace gk {

DevBuf<float4>& sceneTriPositions();

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ void atomicMinFloat(float* addr, float v)
{
    // monotone int mapping of IEEE floats
    if (v >= 0) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomicMaxFloat(float* addr, float v)
{
    if (v >= 0) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void k_init_group_bounds(float4* lo, float4* hi, uint32_t groups)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    lo[g] = make_float4(kFar, kFar, kFar, 0);
    hi[g] = make_float4(-kFar, -kFar, -kFar, 0);
}

__global__ void k_group_bounds(const float4* __restrict__ plo, const float4* __restrict__ phi, const uint32_t* __restrict__ group, uint32_t n,
                               float4* glo, float4* ghi)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = plo[i], b = phi[i];
    if (a.x > b.x) return; // empty box (hidden instance)
    const uint32_t g = group ? group[i] : 0;
    // warp-aggregate when the whole warp targets one group (the common case)
    const unsigned full = __activemask();
    const uint32_t g0 = __shfl_sync(full, g, __ffs(full) - 1);
    if (__all_sync(full, g == g0)) {
        float lx = a.x, ly = a.y, lz = a.z, hx = b.x, hy = b.y, hz = b.z;
        for (int o = 16; o; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(full, lx, o)), ly = fminf(ly, __shfl_xor_sync(full, ly, o)), lz = fminf(lz, __shfl_xor_sync(full, lz, o));
            hx = fmaxf(hx, __shfl_xor_sync(full, hx, o)), hy = fmaxf(hy, __shfl_xor_sync(full, hy, o)), hz = fmaxf(hz, __shfl_xor_sync(full, hz, o));
        }
        if (full == 0xffffffffu) {
            if ((threadIdx.x & 31) == 0) {
                atomicMinFloat(&glo[g].x, lx), atomicMinFloat(&glo[g].y, ly), atomicMinFloat(&glo[g].z, lz);
                atomicMaxFloat(&ghi[g].x, hx), atomicMaxFloat(&ghi[g].y, hy), atomicMaxFloat(&ghi[g].z, hz);
            }
            return;
        }
    }
    atomicMinFloat(&glo[g].x, a.x), atomicMinFloat(&glo[g].y, a.y), atomicMinFloat(&glo[g].z, a.z);
    atomicMaxFloat(&ghi[g].x, b.x), atomicMaxFloat(&ghi[g].y, b.y), atomicMaxFloat(&ghi[g].z, b.z);
}

__device__ __forceinline__ unsigned long long spread14(uint32_t v)
{
    // 14 bits -> every third bit of a 42-bit word
    unsigned long long x = v & 0x3fffu;
    x = (x | (x << 32)) & 0x001f00000000ffffull;
    x = (x | (x << 16)) & 0x001f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ plo, const float4* __restrict__ phi, const uint32_t* __restrict__ group, uint32_t n,
                         const float4* __restrict__ glo, const float4* __restrict__ ghi, unsigned long long* __restrict__ keys, uint32_t* __restrict__ order,
                         int sizeBits)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = plo[i], b = phi[i];
    const uint32_t g = group ? group[i] : 0;
    const float4 gl = glo[g], gh = ghi[g];
    float cx = 0.5f, cy = 0.5f, cz = 0.5f;
    if (a.x <= b.x) {
        const float ex = gh.x - gl.x, ey = gh.y - gl.y, ez = gh.z - gl.z;
        cx = ex > 0 ? ((a.x + b.x) * 0.5f - gl.x) / ex : 0.5f;
        cy = ey > 0 ? ((a.y + b.y) * 0.5f - gl.y) / ey : 0.5f;
        cz = ez > 0 ? ((a.z + b.z) * 0.5f - gl.z) / ez : 0.5f;
    }
    const uint32_t qx = min(16383u, (uint32_t)(fmaxf(cx, 0.f) * 16384.f)), qy = min(16383u, (uint32_t)(fmaxf(cy, 0.f) * 16384.f)),
                   qz = min(16383u, (uint32_t)(fmaxf(cz, 0.f) * 16384.f));
    unsigned long long m = (spread14(qx) << 2) | (spread14(qy) << 1) | spread14(qz);
    if (sizeBits > 0 && a.x <= b.x) {
        // extended Morton code (Vinkler et al. 2017): bits of the box size are woven in between the position
        // bits (one before every second x/y/z triple), so that boxes much larger than their neighbours
        // split off near the top of the tree instead of bloating every level below
        const float ex = gh.x - gl.x, ey = gh.y - gl.y, ez = gh.z - gl.z;
        const float gd = sqrtf(ex * ex + ey * ey + ez * ez);
        const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
        const float rel = gd > 0 ? sqrtf(dx * dx + dy * dy + dz * dz) / gd : 0.f;
        const uint32_t s = min((1u << sizeBits) - 1u, (uint32_t)(fminf(rel, 1.f) * (float)(1u << sizeBits)));
        unsigned long long e = 0;
        int sb = sizeBits - 1;
        for (int t = 0; t < 14; ++t) { // from the most significant triple down
            if ((t & 1) == 0 && sb >= 0) e = (e << 1) | ((s >> sb--) & 1u);
            e = (e << 3) | ((m >> (3 * (13 - t))) & 7ull);
        }
        m = e;
    }
    keys[i] = ((unsigned long long)g << (42 + max(sizeBits, 0))) | m;
    order[i] = i;
}

// ------------------------------------------------------------------ Karras radix tree
__device__ __forceinline__ int deltaKeys(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_radix_tree(const unsigned long long* __restrict__ keys, int n, uint32_t* __restrict__ left, uint32_t* __restrict__ right,
                             uint32_t* __restrict__ parentI, uint32_t* __restrict__ parentL, uint32_t* __restrict__ first, uint32_t* __restrict__ last)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (deltaKeys(keys, n, i, i + 1) - deltaKeys(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = deltaKeys(keys, n, i, i - d);
    int lmax = 2;
    while (deltaKeys(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (deltaKeys(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = deltaKeys(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (deltaKeys(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t L = (lo == gamma) ? (kLeafBit | (uint32_t)gamma) : (uint32_t)gamma;
    const uint32_t R = (hi == gamma + 1) ? (kLeafBit | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    left[i] = L, right[i] = R;
    first[i] = (uint32_t)lo, last[i] = (uint32_t)hi;
    if (L & kLeafBit) parentL[gamma] = (uint32_t)i; else parentI[gamma] = (uint32_t)i;
    if (R & kLeafBit) parentL[gamma + 1] = (uint32_t)i; else parentI[gamma + 1] = (uint32_t)i;
    if (i == 0) parentI[0] = kInvalid;
}

__global__ void k_leaf_boxes(const float4* __restrict__ plo, const float4* __restrict__ phi, const uint32_t* __restrict__ order, uint32_t n,
                             float4* __restrict__ llo, float4* __restrict__ lhi)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t p = order[k];
    llo[k] = plo[p], lhi[k] = phi[p];
}

// Cost model of the wide collapse (node visit = 1): one triangle test, one instance entry.
constexpr float kCostInstance = 1.5f;
constexpr float kRefitGrowthLimit = 1.25f;

__device__ __forceinline__ float boxAreaOrZero(float lx, float ly, float lz, float hx, float hy, float hz)
{
    if (lx > hx) return 0.f; // empty box (hidden instance)
    const float ex = hx - lx, ey = hy - ly, ez = hz - lz;
    return ex * ey + ey * ez + ez * ex;
}

// Bottom-up: boxes of the internal nodes and, when `cost` is given, the collapse cost table of
// Ylitie et al. 2017 ("Efficient incoherent ray traversal on GPUs through compressed wide BVHs", 3.1):
// cost[8*n + i-1] = cheapest SAH cost of representing the subtree of n with at most i slots of
// its parent's wide node (i = 1..7); slot [7] holds the cost of n as a wide node of its own.
__global__ void k_propagate_bounds(uint32_t n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right, const uint32_t* __restrict__ parentI,
                                   const uint32_t* __restrict__ parentL, const float4* llo, const float4* lhi, float4* ilo, float4* ihi, int* flags,
                                   float* cost, const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, uint32_t leafMax, float primCost,
                                   float* __restrict__ areaSum)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || n < 2) return;
    uint32_t cur = parentL[k];
    float myArea = 0.f; // surface area of the internal nodes this thread completes (tree quality measure)
    while (cur != kInvalid) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1) == 0) break; // first arrival: the sibling subtree finishes the job
        const uint32_t L = left[cur], R = right[cur];
        const volatile float4* pl0 = (L & kLeafBit) ? llo + (L & 0x7fffffffu) : ilo + L;
        const volatile float4* ph0 = (L & kLeafBit) ? lhi + (L & 0x7fffffffu) : ihi + L;
        const volatile float4* pl1 = (R & kLeafBit) ? llo + (R & 0x7fffffffu) : ilo + R;
        const volatile float4* ph1 = (R & kLeafBit) ? lhi + (R & 0x7fffffffu) : ihi + R;
        const float l0x = pl0->x, l0y = pl0->y, l0z = pl0->z, h0x = ph0->x, h0y = ph0->y, h0z = ph0->z;
        const float l1x = pl1->x, l1y = pl1->y, l1z = pl1->z, h1x = ph1->x, h1y = ph1->y, h1z = ph1->z;
        const float lx = fminf(l0x, l1x), ly = fminf(l0y, l1y), lz = fminf(l0z, l1z);
        const float hx = fmaxf(h0x, h1x), hy = fmaxf(h0y, h1y), hz = fmaxf(h0z, h1z);
        ilo[cur] = make_float4(lx, ly, lz, 0), ihi[cur] = make_float4(hx, hy, hz, 0);
        myArea += boxAreaOrZero(lx, ly, lz, hx, hy, hz);
        if (cost) {
            float cl[7], cr[7];
            if (L & kLeafBit) {
                const float a = boxAreaOrZero(l0x, l0y, l0z, h0x, h0y, h0z) * primCost;
#pragma unroll
                for (int i = 0; i < 7; ++i) cl[i] = a;
            } else {
                const volatile float* c = cost + 8 * (size_t)L;
#pragma unroll
                for (int i = 0; i < 7; ++i) cl[i] = c[i];
            }
            if (R & kLeafBit) {
                const float a = boxAreaOrZero(l1x, l1y, l1z, h1x, h1y, h1z) * primCost;
#pragma unroll
                for (int i = 0; i < 7; ++i) cr[i] = a;
            } else {
                const volatile float* c = cost + 8 * (size_t)R;
#pragma unroll
                for (int i = 0; i < 7; ++i) cr[i] = c[i];
            }
            float D[9]; // D[j]: children of `cur` spread over j slots
#pragma unroll
            for (int j = 2; j <= 8; ++j) {
                float best = kFar;
#pragma unroll
                for (int a = 1; a < j; ++a) best = fminf(best, cl[a - 1] + cr[j - a - 1]);
                D[j] = best;
            }
            const float A = boxAreaOrZero(lx, ly, lz, hx, hy, hz);
            const uint32_t P = last[cur] - first[cur] + 1;
            const float asNode = D[8] + A;
            const float asLeaf = (leafMax > 0 && P <= leafMax) ? A * (float)P * primCost : kFar;
            float* o = cost + 8 * (size_t)cur;
            float prev = fminf(asLeaf, asNode);
            o[0] = prev;
#pragma unroll
            for (int i = 2; i <= 7; ++i) {
                prev = fminf(D[i], prev);
                o[i - 1] = prev;
            }
            o[7] = asNode;
        }
        cur = parentI[cur];
    }
    if (areaSum && myArea > 0.f) atomicAdd(areaSum, myArea);
}

// ------------------------------------------------------------------ PLOC topology (TLAS)
// Parallel locally-ordered clustering (Meister & Bittner 2018): the clusters, kept in Morton order,
// repeatedly merge with their nearest neighbour (smallest surface area of the union) inside a
// window of +-radius positions (Context::tlasPlocRadius) when the choice is mutual.  Far better trees than the radix tree
// of the same order for boxes of mixed size and overlap (instances); costs ~log n rounds of small
// launches, so it is used when a TLAS is built for keeps, not for per-frame rebuilds of huge ones.

struct PlocClusters {
    uint32_t* ref; // leaf bit | sorted position, or internal node index
    float4* lo;
    float4* hi;
};

__global__ void k_ploc_init(uint32_t n, const float4* __restrict__ llo, const float4* __restrict__ lhi, PlocClusters C)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    C.ref[i] = kLeafBit | i, C.lo[i] = llo[i], C.hi[i] = lhi[i];
}

__global__ void k_ploc_nearest(uint32_t m, PlocClusters C, uint32_t* __restrict__ nn, int radius, const uint32_t* __restrict__ grp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int)m) return;
    const float4 a = C.lo[i], b = C.hi[i];
    float best = 3.0e38f;
    int bestJ = i;
    uint32_t bestH = 0xffffffffu;
    const int j0 = max(0, i - radius), j1 = min((int)m - 1, i + radius);
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        if (grp && grp[j] != grp[i]) continue; // forest: clusters of different groups (models) never merge
        const float4 c = C.lo[j], d = C.hi[j];
        // an empty box (hidden instance: lo > hi) leaves the other box unchanged
        const float area = boxAreaOrZero(fminf(a.x, c.x), fminf(a.y, c.y), fminf(a.z, c.z), fmaxf(b.x, d.x), fmaxf(b.y, d.y), fmaxf(b.z, d.z));
        // equal areas (bricks on a lattice) are ordered by a hash of the PAIR, which both partners compute alike:
        // with "first wins" every cluster of a run would point at its left neighbour and one pair per run would merge
        const uint32_t lo2 = (uint32_t)min(i, j), hi2 = (uint32_t)max(i, j);
        uint32_t h = lo2 * 0x9E3779B1u ^ hi2 * 0x85EBCA77u;
        h ^= h >> 15, h *= 0x2C1B3C6Du, h ^= h >> 12;
        if (area < best || (area == best && h < bestH)) best = area, bestJ = j, bestH = h;
    }
    nn[i] = (uint32_t)bestJ;
}

__global__ void k_ploc_merge(uint32_t m, PlocClusters C, const uint32_t* __restrict__ nn, uint32_t* __restrict__ valid, uint32_t* __restrict__ nodeCounter,
                             uint32_t* __restrict__ left, uint32_t* __restrict__ right, uint32_t* __restrict__ parentI, uint32_t* __restrict__ parentL)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t j = nn[i];
    const bool mutual = j != i && nn[j] == i;
    if (!mutual) {
        valid[i] = 1;
        return;
    }
    if (i > j) {
        valid[i] = 0; // absorbed by its partner
        return;
    }
    const uint32_t id = atomicAdd(nodeCounter, 1u);
    const uint32_t L = C.ref[i], R = C.ref[j];
    left[id] = L, right[id] = R;
    if (L & kLeafBit) parentL[L & 0x7fffffffu] = id; else parentI[L] = id;
    if (R & kLeafBit) parentL[R & 0x7fffffffu] = id; else parentI[R] = id;
    const float4 a = C.lo[i], b = C.hi[i], c = C.lo[j], d = C.hi[j];
    C.ref[i] = id;
    C.lo[i] = make_float4(fminf(a.x, c.x), fminf(a.y, c.y), fminf(a.z, c.z), 0);
    C.hi[i] = make_float4(fmaxf(b.x, d.x), fmaxf(b.y, d.y), fmaxf(b.z, d.z), 0);
    valid[i] = 1;
}

__global__ void k_ploc_compact(uint32_t m, PlocClusters in, PlocClusters out, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ pos, uint32_t* __restrict__ newCount,
                               const uint32_t* __restrict__ grpIn, uint32_t* __restrict__ grpOut)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (valid[i]) {
        const uint32_t p = pos[i];
        out.ref[p] = in.ref[i], out.lo[p] = in.lo[i], out.hi[p] = in.hi[i];
        if (grpIn) grpOut[p] = grpIn[i];
    }
    if (i == m - 1) *newCount = pos[i] + valid[i];
}

__global__ void k_ploc_finish(PlocClusters C, uint32_t* __restrict__ parentI, uint32_t* __restrict__ rootOut)
{
    const uint32_t r = C.ref[0];
    if (!(r & kLeafBit)) parentI[r] = kInvalid;
    *rootOut = r;
}

__global__ void k_ploc_groups(uint32_t n, const unsigned long long* __restrict__ keys, int shift, uint32_t* __restrict__ grp, uint32_t* __restrict__ groupBase)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = (uint32_t)(keys[i] >> shift);
    grp[i] = g;
    if (i == 0 || (uint32_t)(keys[i - 1] >> shift) != g) groupBase[g] = i; // groups are contiguous in the sorted order
}

__global__ void k_ploc_forest_finish(uint32_t m, PlocClusters C, uint32_t* __restrict__ parentI)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t r = C.ref[i];
    if (!(r & kLeafBit)) parentI[r] = kInvalid; // root of its group
}

// A PLOC subtree does not cover a contiguous run of the Morton order, but BLAS leaves are runs of the
// triangle array.  The leaves are therefore renumbered in depth-first order of the finished trees:
// sizes bottom-up, position of a leaf = sum of the left-sibling sizes on its way up, then the key
// range [first,last] of every node bottom-up again.
__global__ void k_subtree_size(uint32_t n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right, const uint32_t* __restrict__ parentI,
                               const uint32_t* __restrict__ parentL, uint32_t* size, int* flags)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t cur = parentL[k];
    while (cur != kInvalid) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1) == 0) return;
        const uint32_t L = left[cur], R = right[cur];
        const volatile uint32_t* vs = size;
        size[cur] = ((L & kLeafBit) ? 1u : vs[L]) + ((R & kLeafBit) ? 1u : vs[R]);
        cur = parentI[cur];
    }
}

__global__ void k_dfs_position(uint32_t n, const uint32_t* __restrict__ grp, const uint32_t* __restrict__ groupBase, const uint32_t* __restrict__ left,
                               const uint32_t* __restrict__ right, const uint32_t* __restrict__ parentI, const uint32_t* __restrict__ parentL,
                               const uint32_t* __restrict__ size, uint32_t* __restrict__ newPos)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t pos = 0, child = kLeafBit | k, cur = parentL[k];
    while (cur != kInvalid) {
        if (right[cur] == child) {
            const uint32_t L = left[cur];
            pos += (L & kLeafBit) ? 1u : size[L];
        }
        child = cur;
        cur = parentI[cur];
    }
    newPos[k] = groupBase[grp[k]] + pos;
}

__global__ void k_dfs_permute(uint32_t n, const uint32_t* __restrict__ newPos, const uint32_t* __restrict__ order, const uint32_t* __restrict__ parentL,
                              uint32_t* __restrict__ order2, uint32_t* __restrict__ parentL2)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t p = newPos[k];
    order2[p] = order[k], parentL2[p] = parentL[k];
}

__global__ void k_dfs_fix_refs(uint32_t nodes, const uint32_t* __restrict__ newPos, uint32_t* __restrict__ left, uint32_t* __restrict__ right)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nodes) return;
    const uint32_t L = left[i], R = right[i];
    if (L & kLeafBit) left[i] = kLeafBit | newPos[L & 0x7fffffffu];
    if (R & kLeafBit) right[i] = kLeafBit | newPos[R & 0x7fffffffu];
}

// internal-node slots PLOC did not use (one per group is spare) get an empty key range so that
// k_group_roots never mistakes them for a root
__global__ void k_mark_unused_nodes(uint32_t from, uint32_t to, uint32_t* __restrict__ first, uint32_t* __restrict__ last)
{
    const uint32_t i = from + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= to) return;
    first[i] = 1u, last[i] = 0u;
}

__global__ void k_subtree_range(uint32_t n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right, const uint32_t* __restrict__ parentI,
                                const uint32_t* __restrict__ parentL, uint32_t* first, uint32_t* last, int* flags)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t cur = parentL[k];
    while (cur != kInvalid) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1) == 0) return;
        const uint32_t L = left[cur], R = right[cur];
        const volatile uint32_t *vf = first, *vl = last;
        first[cur] = (L & kLeafBit) ? (L & 0x7fffffffu) : vf[L];
        last[cur] = (R & kLeafBit) ? (R & 0x7fffffffu) : vl[R];
        cur = parentI[cur];
    }
}

// ------------------------------------------------------------------ per-group roots
__global__ void k_group_roots(const unsigned long long* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ first, const uint32_t* __restrict__ last,
                              uint32_t* __restrict__ groupRoot, int groupShift)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { // single-primitive groups: the leaf is the root
        const unsigned long long g = keys[i] >> groupShift;
        const bool lonelyL = (i == 0) || (keys[i - 1] >> groupShift) != g, lonelyR = (i == n - 1) || (keys[i + 1] >> groupShift) != g;
        if (lonelyL && lonelyR) groupRoot[g] = kLeafBit | i;
    }
    if (i + 1 < n) {
        const uint32_t a = first[i], b = last[i];
        const unsigned long long g = keys[a] >> groupShift;
        if ((keys[b] >> groupShift) == g && (a == 0 || (keys[a - 1] >> groupShift) != g) && (b == n - 1 || (keys[b + 1] >> groupShift) != g)) groupRoot[g] = i;
    }
}

// ------------------------------------------------------------------ triangle records in leaf order
__global__ void k_write_tri_records(const float4* __restrict__ triP, const uint32_t* __restrict__ order, const ModelInfo* __restrict__ models,
                                    const uint32_t* __restrict__ group, uint32_t n, TriRecord* __restrict__ out)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t p = order[k];
    const float4 a = triP[(size_t)p * 3], b = triP[(size_t)p * 3 + 1], c = triP[(size_t)p * 3 + 2];
    TriRecord r;
    r.v0x = a.x, r.v0y = a.y, r.v0z = a.z;
    r.prim = p - models[group[p]].triOffset; // triangle index inside its model
    r.e1x = __fsub_rn(b.x, a.x), r.e1y = __fsub_rn(b.y, a.y), r.e1z = __fsub_rn(b.z, a.z);
    r.e2x = __fsub_rn(c.x, a.x), r.e2y = __fsub_rn(c.y, a.y), r.e2z = __fsub_rn(c.z, a.z);
    r.pad0 = r.pad1 = 0;
    float4* o = reinterpret_cast<float4*>(out + k);
    const float4* s = reinterpret_cast<const float4*>(&r);
    o[0] = s[0], o[1] = s[1], o[2] = s[2];
}

// ------------------------------------------------------------------ collapse to 8-wide
struct Bvh2View {
    const uint32_t *left, *right, *first, *last, *order;
    const float4 *ilo, *ihi, *llo, *lhi;
    const float* cost; // collapse cost table (null: greedy surface-area collapse)
    float primCost;
};

__device__ __forceinline__ float boxArea(float4 lo, float4 hi)
{
    const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
    return ex * ey + ey * ez + ez * ex;
}

// Turns a binary-tree reference into the reference stored in a wide node.
__device__ __forceinline__ uint32_t finalLeafRef(const Bvh2View& B, uint32_t ref, uint32_t leafMax, bool tlas, bool& isLeaf)
{
    if (ref & kLeafBit) {
        isLeaf = true;
        const uint32_t k = ref & 0x7fffffffu;
        return tlas ? (kLeafBit | B.order[k]) : (kLeafBit | (k << 3));
    }
    const uint32_t size = B.last[ref] - B.first[ref] + 1;
    bool asLeaf = !tlas && size <= leafMax;
    if (asLeaf && B.cost) asLeaf = boxArea(B.ilo[ref], B.ihi[ref]) * (float)size * B.primCost <= B.cost[8 * (size_t)ref + 7];
    if (asLeaf) {
        isLeaf = true;
        return kLeafBit | (B.first[ref] << 3) | (size - 1);
    }
    isLeaf = false;
    return ref;
}

// Seeds one collapse task per group whose root needs a wide node; others get their leaf reference.
__global__ void k_collapse_seed(Bvh2View B, const uint32_t* __restrict__ groupRoot, uint32_t groups, uint32_t leafMax, int tlas, uint32_t* __restrict__ rootRef,
                                uint32_t* __restrict__ tasks, uint32_t* __restrict__ counters /* [0]=taskCount [1]=wideCount */)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const uint32_t r = groupRoot[g];
    if (r == kInvalid) {
        rootRef[g] = kInvalid;
        return;
    }
    bool leaf;
    const uint32_t ref = finalLeafRef(B, r, leafMax, tlas != 0, leaf);
    if (leaf) {
        rootRef[g] = ref;
        return;
    }
    const uint32_t w = atomicAdd(&counters[1], 1u);
    const uint32_t t = atomicAdd(&counters[0], 1u);
    tasks[2 * t] = r, tasks[2 * t + 1] = w;
    rootRef[g] = w;
}

// One collapse task: binary node b2 becomes wide node w; its internal children are queued for the next level.
__device__ __forceinline__ void collapseTask(const Bvh2View& B, uint32_t b2, uint32_t w, uint32_t* __restrict__ tasksOut, uint32_t* __restrict__ outCount,
                                             uint32_t* __restrict__ wideCount, WideNode* __restrict__ nodes, uint32_t* __restrict__ nodeSrc, uint32_t leafMax, int tlas,
                                             uint32_t nodeCapacity)
{
    uint32_t slot[8];
    int cnt = 0;
    if (B.cost) {
        // spread the subtree of b2 over the eight slots as the cost table dictates (left to right)
        uint32_t stRef[16];
        int stBudget[16];
        int sp = 0;
        stRef[sp] = b2, stBudget[sp] = 9, ++sp; // 9: "open this node over 8 slots"
        while (sp) {
            --sp;
            const uint32_t m = stRef[sp];
            int budget = stBudget[sp];
            if ((m & kLeafBit) || budget == 1) {
                slot[cnt++] = m;
                continue;
            }
            if (budget == 9) budget = 8;
            else {
                // fewer slots cost the same: hand the budget down (table entries are running minima)
                const float* cm = B.cost + 8 * (size_t)m;
                while (budget > 1 && cm[budget - 1] == cm[budget - 2]) --budget;
                if (budget == 1) {
                    slot[cnt++] = m;
                    continue;
                }
            }
            const uint32_t L = B.left[m], R = B.right[m];
            float cl[7], cr[7];
            if (L & kLeafBit) for (int i = 0; i < 7; ++i) cl[i] = 0.f;
            else for (int i = 0; i < 7; ++i) cl[i] = B.cost[8 * (size_t)L + i];
            if (R & kLeafBit) for (int i = 0; i < 7; ++i) cr[i] = 0.f;
            else for (int i = 0; i < 7; ++i) cr[i] = B.cost[8 * (size_t)R + i];
            int bestA = 1;
            float best = kFar;
            for (int a = 1; a < budget; ++a) {
                const float v = cl[a - 1] + cr[budget - a - 1];
                if (v < best) best = v, bestA = a;
            }
            // right first so that the left subtree is emitted first
            stRef[sp] = R, stBudget[sp] = budget - bestA, ++sp;
            stRef[sp] = L, stBudget[sp] = bestA, ++sp;
        }
    } else {
    cnt = 2;
    slot[0] = B.left[b2], slot[1] = B.right[b2];
    // greedy: open the internal child with the largest box until 8 slots are used
    for (;;) {
        if (cnt == 8) break;
        int best = -1;
        float bestA = -1.f;
        for (int i = 0; i < cnt; ++i) {
            const uint32_t r = slot[i];
            if (r & kLeafBit) continue;
            if (!tlas && (B.last[r] - B.first[r] + 1) <= leafMax) continue;
            const float a = boxArea(B.ilo[r], B.ihi[r]);
            if (a > bestA) bestA = a, best = i;
        }
        if (best < 0) break;
        const uint32_t r = slot[best];
        // keep the key order of the children: left stays, right is inserted after it
        for (int i = cnt; i > best + 1; --i) slot[i] = slot[i - 1];
        slot[best] = B.left[r], slot[best + 1] = B.right[r];
        ++cnt;
    }
    }
    uint32_t refs[8];
    bool leaf[8];
    int internal = 0;
    for (int i = 0; i < cnt; ++i) {
        refs[i] = finalLeafRef(B, slot[i], leafMax, tlas != 0, leaf[i]);
        if (!leaf[i]) ++internal;
    }
    uint32_t base = 0, tbase = 0;
    if (internal) {
        base = atomicAdd(wideCount, (uint32_t)internal);
        tbase = atomicAdd(outCount, (uint32_t)internal);
    }
    if (w >= nodeCapacity || base + internal > nodeCapacity) return; // capacity is an upper bound; never expected
    WideNode n;
    n.ox = n.oy = n.oz = 0.f;
    n.ex = n.ey = n.ez = 127;
    n.count = (uint8_t)cnt;
    n.spare[0] = n.spare[1] = n.spare[2] = n.spare[3] = 0;
    int k = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t src = kInvalid;
        n.c[i].ref = kInvalid;
        if (i < cnt) {
            src = slot[i];
            if (leaf[i]) n.c[i].ref = refs[i];
            else {
                n.c[i].ref = base + k;
                tasksOut[2 * (tbase + k)] = refs[i], tasksOut[2 * (tbase + k) + 1] = base + k;
                ++k;
            }
        }
        nodeSrc[(size_t)w * 8 + i] = src; // binary-tree node each slot was made from (refit re-quantises from these)
        for (int a = 0; a < 3; ++a) n.c[i].qlo[a] = 255, n.c[i].qhi[a] = 0;
        n.c[i].k00 = 0x00, n.c[i].k4b = 0x4B;
    }
    uint4* o = reinterpret_cast<uint4*>(nodes + w);
    const uint4* s = reinterpret_cast<const uint4*>(&n);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = s[i];
}

// The whole top-down collapse in ONE cooperative launch: the levels are separated by grid-wide barriers instead of a host round trip per
// level (a 200 k-instance TLAS has about a dozen).  Three queue counters rotate (this level's input, its output, and the one being cleared
// for the level after), so that no block clears a counter another block still reads.
constexpr int kCollapseBlock = 128;
__global__ void __launch_bounds__(kCollapseBlock) k_collapse_all(Bvh2View B, uint32_t* __restrict__ tasksA, uint32_t* __restrict__ tasksB,
                                                                 uint32_t* __restrict__ counters /* [0],[2],[3] = queue counts, [1] = wideCount */,
                                                                 WideNode* __restrict__ nodes, uint32_t* __restrict__ nodeSrc, uint32_t leafMax, int tlas,
                                                                 uint32_t nodeCapacity)
{
    cg::grid_group grid = cg::this_grid();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    const int slotOf[3] = {0, 2, 3};
    uint32_t* qin = tasksA;
    uint32_t* qout = tasksB;
    for (int level = 0;; ++level) {
        uint32_t* cin = counters + slotOf[level % 3];
        uint32_t* cout = counters + slotOf[(level + 1) % 3];
        const uint32_t tasks = *reinterpret_cast<volatile uint32_t*>(cin);
        if (tasks == 0) break;
        if (gtid == 0) counters[slotOf[(level + 2) % 3]] = 0;
        for (uint32_t ti = gtid; ti < tasks; ti += gsize) collapseTask(B, qin[2 * ti], qin[2 * ti + 1], qout, cout, counters + 1, nodes, nodeSrc, leafMax, tlas, nodeCapacity);
        grid.sync();
        uint32_t* t = qin;
        qin = qout, qout = t;
    }
}

// Per-axis step exponent: the smallest power of two with extent / step <= 250, but never
// finer than 2^-18 of the coordinate magnitude (so that origin - step is a distinct float).
__device__ __forceinline__ int chooseExponent(float lo, float hi)
{
    const float extent = hi - lo;
    int e = -126;
    if (extent > 0.f) {
        const float x = extent / 248.f;
        int ex = ilogbf(x);
        if (ldexpf(1.f, ex) < x) ++ex;
        e = ex;
    }
    const float mag = fmaxf(fabsf(lo), fabsf(hi));
    if (mag > 0.f) e = max(e, ilogbf(mag) - 18);
    return min(max(e, -120), 120);
}

__global__ void k_quantise(Bvh2View B, WideNode* __restrict__ nodes, const uint32_t* __restrict__ nodeSrc, uint32_t count)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= count) return;
    WideNode n;
    {
        uint4* d = reinterpret_cast<uint4*>(&n);
        const uint4* s = reinterpret_cast<const uint4*>(nodes + w);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = s[i];
    }
    float4 lo[8], hi[8];
    float4 nlo = make_float4(kFar, kFar, kFar, 0), nhi = make_float4(-kFar, -kFar, -kFar, 0);
    for (int i = 0; i < (int)n.count; ++i) {
        const uint32_t r = nodeSrc[(size_t)w * 8 + i];
        if (r & kLeafBit) lo[i] = B.llo[r & 0x7fffffffu], hi[i] = B.lhi[r & 0x7fffffffu];
        else lo[i] = B.ilo[r], hi[i] = B.ihi[r];
        if (lo[i].x <= hi[i].x) {
            nlo.x = fminf(nlo.x, lo[i].x), nlo.y = fminf(nlo.y, lo[i].y), nlo.z = fminf(nlo.z, lo[i].z);
            nhi.x = fmaxf(nhi.x, hi[i].x), nhi.y = fmaxf(nhi.y, hi[i].y), nhi.z = fmaxf(nhi.z, hi[i].z);
        }
    }
    if (nlo.x > nhi.x) nlo = nhi = make_float4(0, 0, 0, 0); // every child empty
    const int e[3] = {chooseExponent(nlo.x, nhi.x), chooseExponent(nlo.y, nhi.y), chooseExponent(nlo.z, nhi.z)};
    const float step[3] = {ldexpf(1.f, e[0]), ldexpf(1.f, e[1]), ldexpf(1.f, e[2])};
    const float org[3] = {nlo.x - 2.f * step[0], nlo.y - 2.f * step[1], nlo.z - 2.f * step[2]};
    n.ox = org[0], n.oy = org[1], n.oz = org[2];
    n.ex = (uint8_t)(e[0] + 127), n.ey = (uint8_t)(e[1] + 127), n.ez = (uint8_t)(e[2] + 127);
    for (int i = 0; i < 8; ++i) {
        if (i < (int)n.count && lo[i].x <= hi[i].x) {
            const float l3[3] = {lo[i].x, lo[i].y, lo[i].z}, h3[3] = {hi[i].x, hi[i].y, hi[i].z};
            for (int a = 0; a < 3; ++a) {
                // one whole quantisation step of slack on both sides: covers the decode error of the
                // traversal (MUFU reciprocal, fused t = q*s + b) with a wide margin
                const float ql = floorf((l3[a] - org[a]) / step[a] - 1.0f), qh = ceilf((h3[a] - org[a]) / step[a] + 1.0f);
                n.c[i].qlo[a] = (uint8_t)fminf(fmaxf(ql, 0.f), 255.f);
                n.c[i].qhi[a] = (uint8_t)fminf(fmaxf(qh, 0.f), 255.f);
            }
        } else {
            for (int a = 0; a < 3; ++a) n.c[i].qlo[a] = 255, n.c[i].qhi[a] = 0;
        }
    }
    uint4* o = reinterpret_cast<uint4*>(nodes + w);
    const uint4* s = reinterpret_cast<const uint4*>(&n);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = s[i];
}

// ------------------------------------------------------------------ instances
// BLASInstance::Update + InvertTransform (tiny_bvh.h:6718-6758): cofactor inverse with the
// reference's term order, world box from the 8 corners of the BLAS root box.
__device__ void invert4x4RowMajor(const float* T, float* o)
{
#define M3(a, b, c) __fmul_rn(__fmul_rn(T[a], T[b]), T[c])
#define S6(p0, p1, p2, p3, p4, p5) __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3), p4), p5)
    o[0] = S6(M3(5, 10, 15), -M3(5, 11, 14), -M3(9, 6, 15), M3(9, 7, 14), M3(13, 6, 11), -M3(13, 7, 10));
    o[1] = S6(-M3(1, 10, 15), M3(1, 11, 14), M3(9, 2, 15), -M3(9, 3, 14), -M3(13, 2, 11), M3(13, 3, 10));
    o[2] = S6(M3(1, 6, 15), -M3(1, 7, 14), -M3(5, 2, 15), M3(5, 3, 14), M3(13, 2, 7), -M3(13, 3, 6));
    o[3] = S6(-M3(1, 6, 11), M3(1, 7, 10), M3(5, 2, 11), -M3(5, 3, 10), -M3(9, 2, 7), M3(9, 3, 6));
    o[4] = S6(-M3(4, 10, 15), M3(4, 11, 14), M3(8, 6, 15), -M3(8, 7, 14), -M3(12, 6, 11), M3(12, 7, 10));
    o[5] = S6(M3(0, 10, 15), -M3(0, 11, 14), -M3(8, 2, 15), M3(8, 3, 14), M3(12, 2, 11), -M3(12, 3, 10));
    o[6] = S6(-M3(0, 6, 15), M3(0, 7, 14), M3(4, 2, 15), -M3(4, 3, 14), -M3(12, 2, 7), M3(12, 3, 6));
    o[7] = S6(M3(0, 6, 11), -M3(0, 7, 10), -M3(4, 2, 11), M3(4, 3, 10), M3(8, 2, 7), -M3(8, 3, 6));
    o[8] = S6(M3(4, 9, 15), -M3(4, 11, 13), -M3(8, 5, 15), M3(8, 7, 13), M3(12, 5, 11), -M3(12, 7, 9));
    o[9] = S6(-M3(0, 9, 15), M3(0, 11, 13), M3(8, 1, 15), -M3(8, 3, 13), -M3(12, 1, 11), M3(12, 3, 9));
    o[10] = S6(M3(0, 5, 15), -M3(0, 7, 13), -M3(4, 1, 15), M3(4, 3, 13), M3(12, 1, 7), -M3(12, 3, 5));
    o[11] = S6(-M3(0, 5, 11), M3(0, 7, 9), M3(4, 1, 11), -M3(4, 3, 9), -M3(8, 1, 7), M3(8, 3, 5));
    o[12] = S6(-M3(4, 9, 14), M3(4, 10, 13), M3(8, 5, 14), -M3(8, 6, 13), -M3(12, 5, 10), M3(12, 6, 9));
    o[13] = S6(M3(0, 9, 14), -M3(0, 10, 13), -M3(8, 1, 14), M3(8, 2, 13), M3(12, 1, 10), -M3(12, 2, 9));
    o[14] = S6(-M3(0, 5, 14), M3(0, 6, 13), M3(4, 1, 14), -M3(4, 2, 13), -M3(12, 1, 6), M3(12, 2, 5));
    o[15] = S6(M3(0, 5, 10), -M3(0, 6, 9), -M3(4, 1, 10), M3(4, 2, 9), M3(8, 1, 6), -M3(8, 2, 5));
#undef M3
#undef S6
    const float det = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], o[0]), __fmul_rn(T[1], o[4])), __fmul_rn(T[2], o[8])), __fmul_rn(T[3], o[12]));
    if (det == 0) {
        for (int i = 0; i < 16; ++i) o[i] = (i % 5 == 0) ? 1.f : 0.f; // tinybvh leaves the identity in place
        return;
    }
    const float invdet = __fdiv_rn(1.0f, det);
    for (int i = 0; i < 16; ++i) o[i] = __fmul_rn(o[i], invdet);
}

__global__ void k_update_instances(const GkNodeProxy* __restrict__ nodes, uint32_t count, const ModelInfo* __restrict__ models, uint32_t modelCount,
                                   InstRecord* __restrict__ inst, float4* __restrict__ plo, float4* __restrict__ phi)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const GkNodeProxy& P = nodes[i];
    InstRecord R;
    const uint32_t model = P.modelId / 10;
    const bool visible = P.visible && !P.nort && model < modelCount; // RayTraceBaseRenderer.cpp:188-189
    float T[16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) T[r * 4 + c] = P.worldTS[c * 4 + r]; // glm column-major -> tinybvh row-major
    invert4x4RowMajor(T, R.invT);
    R.node = i;
    R.indexOffset = model < modelCount ? models[model].indexOffset : 0u, R.vertexOffset = model < modelCount ? models[model].vertexOffset : 0u;
    R.blasRoot = visible ? models[model].blasRoot : kInvalid;
    float4 lo = make_float4(kFar, kFar, kFar, 0), hi = make_float4(-kFar, -kFar, -kFar, 0);
    if (visible && R.blasRoot != kInvalid) {
        const ModelInfo& M = models[model];
        for (int j = 0; j < 8; ++j) {
            const f3 p = mk3(j & 1 ? M.bmax[0] : M.bmin[0], j & 2 ? M.bmax[1] : M.bmin[1], j & 4 ? M.bmax[2] : M.bmin[2]);
            const f3 t = xformPoint(p, T);
            lo.x = fminf(lo.x, t.x), lo.y = fminf(lo.y, t.y), lo.z = fminf(lo.z, t.z);
            hi.x = fmaxf(hi.x, t.x), hi.y = fmaxf(hi.y, t.y), hi.z = fmaxf(hi.z, t.z);
        }
    }
    plo[i] = lo, phi[i] = hi;
    float4* o = reinterpret_cast<float4*>(inst + i);
    const float4* s = reinterpret_cast<const float4*>(&R);
#pragma unroll
    for (int k = 0; k < 5; ++k) o[k] = s[k];
}

// The world box of k_update_instances is the box of the eight transformed corners of the BLAS root box
// (what tinybvh's BLASInstance::Update computes).  For a rotated instance it is much larger than the geometry
// (a sphere's world box does not grow under rotation, its corner box does), and every ray that pierces the
// slack pays a useless instance entry.  One warp per rotated instance folds the transformed triangle vertices
// instead.  The TLAS is ours, so only conservativeness matters: the leaf box is padded by a whole quantisation
// step of its parent node, orders of magnitude above the rounding of these transforms.
__global__ void __launch_bounds__(256) k_tight_instance_bounds(const GkNodeProxy* __restrict__ nodes, uint32_t count, const ModelInfo* __restrict__ models, uint32_t modelCount,
                                                               const InstRecord* __restrict__ inst, const TriRecord* __restrict__ tris, uint32_t maxTris,
                                                               float4* __restrict__ plo, float4* __restrict__ phi)
{
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (i >= count) return;
    if (inst[i].blasRoot == kInvalid) return; // hidden / not ray traced
    const GkNodeProxy& P = nodes[i];
    const uint32_t model = P.modelId / 10;
    if (model >= modelCount) return;
    const ModelInfo& M = models[model];
    if (M.triCount == 0 || M.triCount > maxTris) return;
    float T[12]; // rows 0..2 of the row-major world matrix (glm column-major -> transposed)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) T[r * 4 + c] = P.worldTS[c * 4 + r];
    // axis-aligned placement (each row of the 3x3 part has one non-zero): the corner box is already exact
    bool aligned = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int nz = (T[r * 4] != 0.f) + (T[r * 4 + 1] != 0.f) + (T[r * 4 + 2] != 0.f);
        aligned = aligned && nz <= 1;
    }
    if (aligned) return;
    float lx = kFar, ly = kFar, lz = kFar, hx = -kFar, hy = -kFar, hz = -kFar;
    // model m's triangles occupy [triOffset, triOffset + triCount) of the sorted records too (the model index leads the sort key)
    const float4* tp = reinterpret_cast<const float4*>(tris + M.triOffset);
    for (uint32_t t = lane; t < M.triCount; t += 32) {
        const float4 a = __ldg(tp + 3 * t), b = __ldg(tp + 3 * t + 1), c = __ldg(tp + 3 * t + 2);
        const float vx[3] = {a.x, a.x + b.x, a.x + c.x}, vy[3] = {a.y, a.y + b.y, a.y + c.y}, vz[3] = {a.z, a.z + b.z, a.z + c.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float wx = T[0] * vx[k] + T[1] * vy[k] + T[2] * vz[k] + T[3];
            const float wy = T[4] * vx[k] + T[5] * vy[k] + T[6] * vz[k] + T[7];
            const float wz = T[8] * vx[k] + T[9] * vy[k] + T[10] * vz[k] + T[11];
            lx = fminf(lx, wx), ly = fminf(ly, wy), lz = fminf(lz, wz);
            hx = fmaxf(hx, wx), hy = fmaxf(hy, wy), hz = fmaxf(hz, wz);
        }
    }
    for (int o = 16; o; o >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)), ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)), lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
        hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)), hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)), hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
    }
    if (lane == 0) {
        // never larger than the corner box, and a relative safety margin on top of the quantisation padding
        const float4 clo = plo[i], chi = phi[i];
        const float ex = 1e-5f * (hx - lx) + 1e-6f, ey = 1e-5f * (hy - ly) + 1e-6f, ez = 1e-5f * (hz - lz) + 1e-6f;
        plo[i] = make_float4(fmaxf(clo.x, lx - ex), fmaxf(clo.y, ly - ey), fmaxf(clo.z, lz - ez), 0);
        phi[i] = make_float4(fminf(chi.x, hx + ex), fminf(chi.y, hy + ey), fminf(chi.z, hz + ez), 0);
    }
}

__global__ void k_model_bounds_from_groups(ModelInfo* models, uint32_t modelCount, const float4* glo, const float4* ghi, const uint32_t* rootRef)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= modelCount) return;
    models[m].bmin[0] = glo[m].x, models[m].bmin[1] = glo[m].y, models[m].bmin[2] = glo[m].z, models[m].bmin[3] = 0;
    models[m].bmax[0] = ghi[m].x, models[m].bmax[1] = ghi[m].y, models[m].bmax[2] = ghi[m].z, models[m].bmax[3] = 0;
    models[m].blasRoot = rootRef[m];
}

// ------------------------------------------------------------------ host orchestration
void Lbvh::release()
{
    keys.release(), keysAlt.release(), order.release(), orderAlt.release(), left.release(), right.release(), parentI.release(), parentL.release();
    first.release(), last.release(), cost.release(), ilo.release(), ihi.release(), llo.release(), lhi.release(), plo.release(), phi.release(), group.release(), flags.release();
}

static inline unsigned gridFor(size_t n, unsigned block = 256) { return (unsigned)((n + block - 1) / block); }

// Steps 1-5 over T.plo/T.phi/T.group (group may be null => one group).
static GkStatus buildRadixTree(Context& c, Lbvh& T, const uint32_t* group, uint32_t groups, int keyBits, int sizeBits = 0, bool radixTopology = true)
{
    const uint32_t n = T.n;
    cudaStream_t st = c.stream;
    GK_CUDA(T.keys.reserve(n));
    GK_CUDA(T.keysAlt.reserve(n));
    GK_CUDA(T.order.reserve(n));
    GK_CUDA(T.orderAlt.reserve(n));
    GK_CUDA(T.left.reserve(n));
    GK_CUDA(T.right.reserve(n));
    GK_CUDA(T.parentI.reserve(n));
    GK_CUDA(T.parentL.reserve(n));
    GK_CUDA(T.first.reserve(n));
    GK_CUDA(T.last.reserve(n));
    GK_CUDA(T.ilo.reserve(n));
    GK_CUDA(T.ihi.reserve(n));
    GK_CUDA(T.llo.reserve(n));
    GK_CUDA(T.lhi.reserve(n));
    GK_CUDA(T.flags.reserve(n));
    GK_CUDA(c.dGroupLo.reserve(groups));
    GK_CUDA(c.dGroupHi.reserve(groups));
    k_init_group_bounds<<<gridFor(groups), 256, 0, st>>>(c.dGroupLo.p, c.dGroupHi.p, groups);
    k_group_bounds<<<gridFor(n), 256, 0, st>>>(T.plo.p, T.phi.p, group, n, c.dGroupLo.p, c.dGroupHi.p);
    k_morton<<<gridFor(n), 256, 0, st>>>(T.plo.p, T.phi.p, group, n, c.dGroupLo.p, c.dGroupHi.p, T.keysAlt.p, T.orderAlt.p, sizeBits);
    size_t tempBytes = 0;
    GK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, T.keysAlt.p, T.keys.p, T.orderAlt.p, T.order.p, (int)n, 0, keyBits, st));
    GK_CUDA(c.dSortTemp.reserve(tempBytes));
    tempBytes = c.dSortTemp.bytes();
    GK_CUDA(cub::DeviceRadixSort::SortPairs(c.dSortTemp.p, tempBytes, T.keysAlt.p, T.keys.p, T.orderAlt.p, T.order.p, (int)n, 0, keyBits, st));
    if (n > 1 && radixTopology) k_radix_tree<<<gridFor(n - 1), 256, 0, st>>>(T.keys.p, (int)n, T.left.p, T.right.p, T.parentI.p, T.parentL.p, T.first.p, T.last.p);
    GK_CUDA(cudaGetLastError());
    return GK_OK;
}

static GkStatus propagateBounds(Context& c, Lbvh& T, bool withCost, uint32_t leafMax, float primCost, float* dAreaSum = nullptr)
{
    const uint32_t n = T.n;
    cudaStream_t st = c.stream;
    k_leaf_boxes<<<gridFor(n), 256, 0, st>>>(T.plo.p, T.phi.p, T.order.p, n, T.llo.p, T.lhi.p);
    T.costValid = false;
    if (n > 1) {
        if (withCost) GK_CUDA(T.cost.reserve(8 * (size_t)n));
        GK_CUDA(cudaMemsetAsync(T.flags.p, 0, sizeof(int) * n, st));
        k_propagate_bounds<<<gridFor(n), 256, 0, st>>>(n, T.left.p, T.right.p, T.parentI.p, T.parentL.p, T.llo.p, T.lhi.p, T.ilo.p, T.ihi.p, T.flags.p,
                                                       withCost ? T.cost.p : nullptr, T.first.p, T.last.p, leafMax, primCost, dAreaSum);
        T.costValid = withCost;
        T.primCost = primCost;
    }
    GK_CUDA(cudaGetLastError());
    return GK_OK;
}

static Bvh2View viewOf(const Lbvh& T)
{
    Bvh2View B;
    B.left = T.left.p, B.right = T.right.p, B.first = T.first.p, B.last = T.last.p, B.order = T.order.p;
    B.ilo = T.ilo.p, B.ihi = T.ihi.p, B.llo = T.llo.p, B.lhi = T.lhi.p;
    B.cost = T.costValid ? T.cost.p : nullptr, B.primCost = T.primCost;
    return B;
}

// Binary topology of T (left/right/parents) by PLOC over the Morton-sorted leaf boxes; the root reference
// is left in c.dCounters[7].  keys/order must be in place (buildRadixTree without the radix tree).
static GkStatus plocTopology(Context& c, Lbvh& T, int radius, bool forest = false, uint32_t groups = 1, int groupShift = 42)
{
    const uint32_t n = T.n;
    cudaStream_t st = c.stream;
    for (int k = 0; k < 2; ++k) {
        GK_CUDA(c.dPlocRef[k].reserve(n));
        GK_CUDA(c.dPlocLo[k].reserve(n));
        GK_CUDA(c.dPlocHi[k].reserve(n));
    }
    GK_CUDA(c.dPlocNn.reserve(n));
    GK_CUDA(c.dPlocValid.reserve(n));
    GK_CUDA(c.dPlocPos.reserve(n));
    GK_CUDA(c.dCounters.reserve(8));
    uint32_t *grpA = nullptr, *grpB = nullptr;
    if (forest) {
        GK_CUDA(c.dPlocGrp[0].reserve(n));
        GK_CUDA(c.dPlocGrp[1].reserve(n));
        GK_CUDA(c.dPlocLeafGrp.reserve(n));
        GK_CUDA(c.dPlocGroupBase.reserve(groups));
        GK_CUDA(cudaMemsetAsync(T.parentL.p, 0xff, sizeof(uint32_t) * n, st)); // a lone triangle of a model has no parent
        k_ploc_groups<<<gridFor(n), 256, 0, st>>>(n, T.keys.p, groupShift, c.dPlocLeafGrp.p, c.dPlocGroupBase.p);
        GK_CUDA(cudaMemcpyAsync(c.dPlocGrp[0].p, c.dPlocLeafGrp.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st));
        grpA = c.dPlocGrp[0].p, grpB = c.dPlocGrp[1].p;
    }
    uint32_t* nodeCounter = c.dCounters.p + 5;
    uint32_t* newCount = c.dCounters.p + 6;
    uint32_t* rootOut = c.dCounters.p + 7;
    GK_CUDA(cudaMemsetAsync(nodeCounter, 0, 3 * sizeof(uint32_t), st));
    k_leaf_boxes<<<gridFor(n), 256, 0, st>>>(T.plo.p, T.phi.p, T.order.p, n, T.llo.p, T.lhi.p);
    PlocClusters A{c.dPlocRef[0].p, c.dPlocLo[0].p, c.dPlocHi[0].p}, B{c.dPlocRef[1].p, c.dPlocLo[1].p, c.dPlocHi[1].p};
    k_ploc_init<<<gridFor(n), 256, 0, st>>>(n, T.llo.p, T.lhi.p, A);
    size_t tempBytes = 0;
    GK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, c.dPlocValid.p, c.dPlocPos.p, (int)n, st));
    GK_CUDA(c.dSortTemp.reserve(tempBytes));
    uint32_t m = n;
    for (int round = 0; m > 1; ++round) {
        if (round > 4096) {
            setLastError("PLOC did not converge");
            return GK_ERR_CUDA;
        }
        k_ploc_nearest<<<gridFor(m), 256, 0, st>>>(m, A, c.dPlocNn.p, radius, grpA);
        k_ploc_merge<<<gridFor(m), 256, 0, st>>>(m, A, c.dPlocNn.p, c.dPlocValid.p, nodeCounter, T.left.p, T.right.p, T.parentI.p, T.parentL.p);
        size_t tb = c.dSortTemp.bytes();
        GK_CUDA(cub::DeviceScan::ExclusiveSum(c.dSortTemp.p, tb, c.dPlocValid.p, c.dPlocPos.p, (int)m, st));
        k_ploc_compact<<<gridFor(m), 256, 0, st>>>(m, A, B, c.dPlocValid.p, c.dPlocPos.p, newCount, grpA, grpB);
        uint32_t next = 0;
        GK_CUDA(cudaMemcpyAsync(&next, newCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
        if (next == 0 || next > m || (next == m && !forest)) { // the globally closest pair is always mutual, so every round merges at least one pair
            setLastError("PLOC made no progress");
            return GK_ERR_CUDA;
        }
        std::swap(A, B);
        std::swap(grpA, grpB);
        if (next == m) break; // forest: one cluster per group is left
        m = next;
    }
    if (!forest) {
        k_ploc_finish<<<1, 1, 0, st>>>(A, T.parentI.p, rootOut);
        GK_CUDA(cudaGetLastError());
        return GK_OK;
    }
    k_ploc_forest_finish<<<gridFor(m), 256, 0, st>>>(m, A, T.parentI.p);
    // ---- depth-first renumbering of the leaves (see k_subtree_size)
    uint32_t nodesMade = 0;
    GK_CUDA(cudaMemcpyAsync(&nodesMade, nodeCounter, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    uint32_t* size = c.dPlocNn.p;      // scratch of the clustering, free now
    uint32_t* newPos = c.dPlocValid.p;
    GK_CUDA(cudaMemsetAsync(T.flags.p, 0, sizeof(int) * n, st));
    k_subtree_size<<<gridFor(n), 256, 0, st>>>(n, T.left.p, T.right.p, T.parentI.p, T.parentL.p, size, T.flags.p);
    k_dfs_position<<<gridFor(n), 256, 0, st>>>(n, c.dPlocLeafGrp.p, c.dPlocGroupBase.p, T.left.p, T.right.p, T.parentI.p, T.parentL.p, size, newPos);
    k_dfs_permute<<<gridFor(n), 256, 0, st>>>(n, newPos, T.order.p, T.parentL.p, T.orderAlt.p, c.dPlocPos.p);
    if (nodesMade) k_dfs_fix_refs<<<gridFor(nodesMade), 256, 0, st>>>(nodesMade, newPos, T.left.p, T.right.p);
    std::swap(T.order, T.orderAlt);
    GK_CUDA(cudaMemcpyAsync(T.parentL.p, c.dPlocPos.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st));
    GK_CUDA(cudaMemsetAsync(T.flags.p, 0, sizeof(int) * n, st));
    k_subtree_range<<<gridFor(n), 256, 0, st>>>(n, T.left.p, T.right.p, T.parentI.p, T.parentL.p, T.first.p, T.last.p, T.flags.p);
    if (n >= 2 && nodesMade < n - 1) k_mark_unused_nodes<<<gridFor(n - 1 - nodesMade), 256, 0, st>>>(nodesMade, n - 1, T.first.p, T.last.p);
    GK_CUDA(cudaGetLastError());
    return GK_OK;
}

// Step 6+7.  rootRef (device, one per group) receives the wide root reference of each group.
static GkStatus collapse(Context& c, Lbvh& T, uint32_t groups, uint32_t leafMax, bool tlas, DevBuf<WideNode>& nodes, DevBuf<uint32_t>& nodeSrc, uint32_t& nodeCount,
                         uint32_t* dRootRef, int groupShift = 42, const uint32_t* dExplicitRoot = nullptr)
{
    const uint32_t n = T.n;
    cudaStream_t st = c.stream;
    const uint32_t capacity = n + groups + 8; // a wide node has >= 2 children: at most n-1 nodes
    GK_CUDA(nodes.reserve(capacity));
    GK_CUDA(nodeSrc.reserve(8 * (size_t)capacity));
    GK_CUDA(c.dTaskA.reserve(2 * (size_t)capacity));
    GK_CUDA(c.dTaskB.reserve(2 * (size_t)capacity));
    GK_CUDA(c.dCounters.reserve(8));
    GK_CUDA(c.dGroupRoot.reserve(groups));
    GK_CUDA(cudaMemsetAsync(c.dGroupRoot.p, 0xff, sizeof(uint32_t) * groups, st));
    // single tree whose root is known (it may live in dCounters, which is cleared next)
    if (dExplicitRoot) GK_CUDA(cudaMemcpyAsync(c.dGroupRoot.p, dExplicitRoot, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    GK_CUDA(cudaMemsetAsync(c.dCounters.p, 0, sizeof(uint32_t) * 8, st));
    const Bvh2View B = viewOf(T);
    if (!dExplicitRoot) k_group_roots<<<gridFor(n), 256, 0, st>>>(T.keys.p, n, T.first.p, T.last.p, c.dGroupRoot.p, groupShift);
    k_collapse_seed<<<gridFor(groups), 256, 0, st>>>(B, c.dGroupRoot.p, groups, leafMax, tlas ? 1 : 0, dRootRef, c.dTaskA.p, c.dCounters.p);
    uint32_t h[2] = {0, 0};
    {
        static int blocksPerSm = 0; // co-resident blocks of the cooperative launch (same for every device this library targets: sm_100a)
        if (!blocksPerSm) GK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, k_collapse_all, kCollapseBlock, 0));
        if (c.smCount == 0) GK_CUDA(cudaDeviceGetAttribute(&c.smCount, cudaDevAttrMultiProcessorCount, c.device));
        const uint32_t wanted = gridFor(n, kCollapseBlock), resident = (uint32_t)std::max(1, blocksPerSm) * (uint32_t)std::max(1, c.smCount);
        uint32_t* ta = c.dTaskA.p;
        uint32_t* tb = c.dTaskB.p;
        uint32_t* cnt = c.dCounters.p;
        WideNode* np = nodes.p;
        uint32_t* sp = nodeSrc.p;
        uint32_t lm = leafMax, cap = capacity;
        int tl = tlas ? 1 : 0;
        Bvh2View view = B;
        void* args[] = {&view, &ta, &tb, &cnt, &np, &sp, &lm, &tl, &cap};
        GK_CUDA(cudaLaunchCooperativeKernel((const void*)k_collapse_all, dim3(std::min(wanted, resident)), dim3(kCollapseBlock), args, 0, st));
    }
    GK_CUDA(cudaMemcpyAsync(h, c.dCounters.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    nodeCount = h[1];
    if (nodeCount > capacity) {
        setLastError("wide-node pool overflow");
        return GK_ERR_OUT_OF_MEMORY;
    }
    if (nodeCount) k_quantise<<<gridFor(nodeCount, 128), 128, 0, st>>>(B, nodes.p, nodeSrc.p, nodeCount);
    GK_CUDA(cudaGetLastError());
    return GK_OK;
}

GkStatus buildBlasForest(Context& c)
{
    cudaStream_t st = c.stream;
    Lbvh& T = c.blasTree;
    const uint32_t groups = (uint32_t)c.models.size();
    if (groups >= (1u << 22)) {
        setLastError("too many models");
        return GK_ERR_UNSUPPORTED;
    }
    ScopedEvents evs(2);
    cudaEvent_t e0 = evs.e[0], e1 = evs.e[1];
    cudaEventRecord(e0, st);
    // GK_BUILD_LOG=1: wall-clock of each phase (stream drained between them; the total then includes those drains)
    const bool log = getenv("GK_BUILD_LOG") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!log) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gk build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    const bool ploc = c.blasPloc && T.n > 2;
    GkStatus s = buildRadixTree(c, T, T.group.p, groups, 64, 0, !ploc);
    if (s != GK_OK) return s;
    mark("bounds + morton + sort");
    if (ploc) {
        s = plocTopology(c, T, c.blasPlocRadius, true, groups, 42);
        if (s != GK_OK) return s;
        mark("ploc topology");
    }
    GK_CUDA(c.dTris.reserve(T.n));
    k_write_tri_records<<<gridFor(T.n), 256, 0, st>>>(sceneTriPositions().p, T.order.p, c.dModels.p, T.group.p, T.n, c.dTris.p);
    mark("triangle records");
    s = propagateBounds(c, T, c.sahCollapse, c.blasLeafMax, c.costTri);
    if (s != GK_OK) return s;
    mark("bounds + cost table");
    DevBuf<uint32_t> rootRef;
    GK_CUDA(rootRef.reserve(groups));
    s = collapse(c, T, groups, c.blasLeafMax, false, c.dBlasNodes, c.dBlasSrc, c.blasNodeCount, rootRef.p);
    if (s != GK_OK) return s;
    k_model_bounds_from_groups<<<gridFor(groups), 256, 0, st>>>(c.dModels.p, groups, c.dGroupLo.p, c.dGroupHi.p, rootRef.p);
    mark("collapse + quantise");
    cudaEventRecord(e1, st);
    GK_CUDA(cudaMemcpyAsync(c.models.data(), c.dModels.p, sizeof(ModelInfo) * groups, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c.msBlasBuild, e0, e1);
    rootRef.release();
    return GK_OK;
}

__global__ void k_scatter_proxies(const uint32_t* __restrict__ indices, const GkNodeProxy* __restrict__ src, uint32_t changed, uint32_t count, GkNodeProxy* __restrict__ nodes)
{
    // 208-byte records as 13 x 16 bytes: one thread per 16-byte piece
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t rec = t / 13u, piece = t - rec * 13u;
    if (rec >= changed) return;
    const uint32_t dst = indices[rec];
    if (dst >= count) return;
    reinterpret_cast<uint4*>(nodes + dst)[piece] = reinterpret_cast<const uint4*>(src + rec)[piece];
}

static GkStatus updateInstancesOnDevice(Context& c, uint32_t count, bool refit, cudaEvent_t e0, cudaEvent_t e1, const GkNodeProxy* hostNodes);

// Sparse form: only `changed` node proxies travel (Scene::UpdateNodes rewrites all of them every dirty frame,
// Scene.cpp:464-511; MagicaLego moves a handful of bricks per frame, MagicaLegoGameInstance.cpp:746-807); the rest of the
// array stays on the device.
GkStatus updateInstancesSparse(Context& c, const uint32_t* indices, const GkNodeProxy* proxies, uint32_t changed, bool refit)
{
    cudaStream_t st = c.stream;
    if (!c.haveScene || !c.haveInstances) {
        setLastError("gk_update_instances_sparse: a full gk_update_instances must come first");
        return GK_ERR_NOT_READY;
    }
    if (changed && (!indices || !proxies)) {
        setLastError("gk_update_instances_sparse: null argument");
        return GK_ERR_INVALID_ARGUMENT;
    }
    for (uint32_t k = 0; k < changed; ++k)
        if (indices[k] >= c.nodeCount) {
            setLastError("gk_update_instances_sparse: proxy index beyond the uploaded array");
            return GK_ERR_INVALID_ARGUMENT;
        }
    ScopedEvents evs(2);
    cudaEventRecord(evs.e[0], st);
    if (changed) {
        GK_CUDA(c.dSparseIdx.reserve(changed));
        GK_CUDA(c.dSparseNodes.reserve(changed));
        GK_CUDA(cudaMemcpyAsync(c.dSparseIdx.p, indices, sizeof(uint32_t) * changed, cudaMemcpyHostToDevice, st));
        GK_CUDA(cudaMemcpyAsync(c.dSparseNodes.p, proxies, sizeof(GkNodeProxy) * changed, cudaMemcpyHostToDevice, st));
        k_scatter_proxies<<<gridFor((size_t)changed * 13, 256), 256, 0, st>>>(c.dSparseIdx.p, c.dSparseNodes.p, changed, c.nodeCount, c.dNodes.p);
    }
    return updateInstancesOnDevice(c, c.nodeCount, refit, evs.e[0], evs.e[1], nullptr);
}

GkStatus updateInstances(Context& c, const GkNodeProxy* nodes, uint32_t count, bool refit)
{
    cudaStream_t st = c.stream;
    if (!c.haveScene) {
        setLastError("gk_update_instances: no scene uploaded");
        return GK_ERR_NOT_READY;
    }
    if (count == 0 || !nodes) {
        setLastError("gk_update_instances: empty instance list");
        return GK_ERR_INVALID_ARGUMENT;
    }
    ScopedEvents evs(2);
    cudaEventRecord(evs.e[0], st);
    GK_CUDA(c.dNodes.reserve(count));
    GK_CUDA(cudaMemcpyAsync(c.dNodes.p, nodes, sizeof(GkNodeProxy) * count, cudaMemcpyHostToDevice, st));
    return updateInstancesOnDevice(c, count, refit, evs.e[0], evs.e[1], nodes);
}

constexpr uint32_t kRefitBackoff = 16;
// Instance records, world boxes and the TLAS (refit or build) from the node proxies in c.dNodes.
static GkStatus updateInstancesOnDevice(Context& c, uint32_t count, bool refit, cudaEvent_t e0, cudaEvent_t e1, const GkNodeProxy* nodes)
{
    cudaStream_t st = c.stream;
    Lbvh& T = c.tlasTree;
    if (refit && (!c.haveInstances || count != T.n)) refit = false;
    // A scene whose refits keep failing the growth guard (C3: 1 % of 200 k bricks teleport every frame) stops paying for the attempt
    // (bound propagation + one host round trip for the area): after two rejections in a row the next kRefitBackoff updates rebuild directly.
    bool guardRebuild = false;
    if (refit && c.refitRejectedInARow >= 2 && c.refitBackoffLeft > 0) {
        c.refitBackoffLeft--;
        refit = false, guardRebuild = true;
    }
    c.nodeCount = count;
    T.n = count;
    GK_CUDA(T.plo.reserve(count));
    GK_CUDA(T.phi.reserve(count));
    GK_CUDA(c.dInst.reserve(count));
    k_update_instances<<<gridFor(count, 128), 128, 0, st>>>(c.dNodes.p, count, c.dModels.p, (uint32_t)c.models.size(), c.dInst.p, T.plo.p, T.phi.p);
    if (c.tightInstanceBounds)
        k_tight_instance_bounds<<<gridFor((size_t)count * 32, 256), 256, 0, st>>>(c.dNodes.p, count, c.dModels.p, (uint32_t)c.models.size(), c.dInst.p, c.dTris.p,
                                                                                 c.tightBoundsMaxTris, T.plo.p, T.phi.p);
    GkStatus s;
    GK_CUDA(c.dCounters.reserve(8));
    float* dArea = reinterpret_cast<float*>(c.dCounters.p + 4);
    float area = 0.f;
    if (refit) {
        // Refit = new boxes on the old topology.  It is kept only while the tree stays good: the summed
        // surface area of the internal nodes may grow to kRefitGrowthLimit x its value at the last build
        // (instances that jump across the scene inflate every ancestor; then a rebuild is cheaper than
        // tracing through the bloated tree).
        GK_CUDA(cudaMemsetAsync(dArea, 0, sizeof(float), st));
        s = propagateBounds(c, T, false, 0, kCostInstance, dArea);
        if (s != GK_OK) return s;
        GK_CUDA(cudaMemcpyAsync(&area, dArea, sizeof(float), cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
        if (area <= kRefitGrowthLimit * c.tlasAreaAtBuild) {
            if (c.tlasNodeCount) k_quantise<<<gridFor(c.tlasNodeCount, 128), 128, 0, st>>>(viewOf(T), c.dTlasNodes.p, c.dTlasSrc.p, c.tlasNodeCount);
            c.refitRejectedInARow = 0;
        } else {
            refit = false;
            guardRebuild = true;
            c.refitRejected++;
            if (++c.refitRejectedInARow >= 2) c.refitBackoffLeft = kRefitBackoff;
        }
    }
    if (!refit) {
        // PLOC when the tree is built to last (first build, instance count changed, small scenes); a rebuild forced by the
        // refit guard on a huge TLAS happens every frame (C3) and keeps the cheaper radix tree
        const bool ploc = c.tlasPloc && count > 2 && count <= (guardRebuild ? c.tlasPlocMaxRebuild : c.tlasPlocMax);
        s = buildRadixTree(c, T, nullptr, 1, 42 + c.tlasSizeBits, c.tlasSizeBits, !ploc);
        if (s != GK_OK) return s;
        if (ploc) {
            GK_CUDA(cudaMemsetAsync(T.first.p, 0, sizeof(uint32_t) * count, st));
            GK_CUDA(cudaMemsetAsync(T.last.p, 0, sizeof(uint32_t) * count, st));
            s = plocTopology(c, T, c.tlasPlocRadius);
            if (s != GK_OK) return s;
        }
        GK_CUDA(cudaMemsetAsync(dArea, 0, sizeof(float), st));
        s = propagateBounds(c, T, c.sahCollapse, 0, kCostInstance, dArea);
        if (s != GK_OK) return s;
        GK_CUDA(cudaMemcpyAsync(&c.tlasAreaAtBuild, dArea, sizeof(float), cudaMemcpyDeviceToHost, st));
        GK_CUDA(c.dRootRef.reserve(1));
        s = collapse(c, T, 1, 1, true, c.dTlasNodes, c.dTlasSrc, c.tlasNodeCount, c.dRootRef.p, 42 + c.tlasSizeBits, ploc ? c.dCounters.p + 7 : nullptr);
        if (s != GK_OK) return s;
        GK_CUDA(cudaMemcpyAsync(&c.tlasRoot, c.dRootRef.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
    }
    cudaEventRecord(e1, st);
    GK_CUDA(cudaGetLastError());
    // instanced triangle count (host side, from the proxies the caller handed over; a sparse update keeps the last figure)
    if (nodes) {
        uint64_t inst = 0;
        for (uint32_t i = 0; i < count; ++i) {
            const uint32_t m = nodes[i].modelId / 10;
            if (nodes[i].visible && !nodes[i].nort && m < c.models.size()) inst += c.models[m].triCount;
        }
        c.instancedTris = inst;
    }
    GK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (refit) c.msRefit = ms; else c.msTlasBuild = ms;
    c.stats.msBvh = ms;
    c.haveInstances = true;
    return GK_OK;
}

} // namespace gk
