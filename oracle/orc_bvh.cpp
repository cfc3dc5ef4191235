// orc_bvh.cpp — see orc_bvh.h.  TEST INFRASTRUCTURE ONLY.
#include "orc_bvh.h"
#include <cstdlib>

namespace orc {

static inline int clampi(int x, int a, int b) { return x > a ? (x < b ? x : b) : a; } // tiny_bvh.h:344

float safeRcp(float x) // tiny_bvh.h:329
{
    if (x > 1e-12f) return 1.0f / x;
    if (x < -1e-12f) return 1.0f / x;
    return kFar;
}

RayQ makeRay(f3 origin, f3 dir, float tmin, float tmax) // tiny_bvh.h:562-567 and :391-395
{
    RayQ r;
    float l = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    float rl = (l == 0) ? 0.0f : (1.0f / l);
    r.O = origin;
    r.D = f3(dir.x * rl, dir.y * rl, dir.z * rl);
    r.rD = f3(safeRcp(r.D.x), safeRcp(r.D.y), safeRcp(r.D.z));
    r.tmin = tmin;
    r.hit.t = tmax;
    r.hit.u = r.hit.v = 0;
    r.hit.prim = r.hit.inst = 0;
    r.instIdx = 0;
    return r;
}

f3 xformPoint(f3 v, const float* T) // tiny_bvh.h:396-404
{
    f3 res(T[0] * v.x + T[1] * v.y + T[2] * v.z + T[3],
           T[4] * v.x + T[5] * v.y + T[6] * v.z + T[7],
           T[8] * v.x + T[9] * v.y + T[10] * v.z + T[11]);
    float w = T[12] * v.x + T[13] * v.y + T[14] * v.z + T[15];
    if (w == 1) return res;
    return res * (1.f / w);
}

f3 xformVector(f3 v, const float* T) // tiny_bvh.h:405-409
{
    return f3(T[0] * v.x + T[1] * v.y + T[2] * v.z,
              T[4] * v.x + T[5] * v.y + T[6] * v.z,
              T[8] * v.x + T[9] * v.y + T[10] * v.z);
}

// Cofactor expansion with the term order of tiny_bvh.h:6737-6756 (the MESA gluInvertMatrix
// formula).  The order matters: the result feeds the instance-space ray transform, so it is
// part of the hit-id arithmetic.  Written as a table of signed triple products:
// out[k] = sum_j sign * T[a]*T[b]*T[c].
void invert4x4RowMajor(const float* T, float* o)
{
    o[0] = T[5] * T[10] * T[15] - T[5] * T[11] * T[14] - T[9] * T[6] * T[15] + T[9] * T[7] * T[14] + T[13] * T[6] * T[11] - T[13] * T[7] * T[10];
    o[1] = -T[1] * T[10] * T[15] + T[1] * T[11] * T[14] + T[9] * T[2] * T[15] - T[9] * T[3] * T[14] - T[13] * T[2] * T[11] + T[13] * T[3] * T[10];
    o[2] = T[1] * T[6] * T[15] - T[1] * T[7] * T[14] - T[5] * T[2] * T[15] + T[5] * T[3] * T[14] + T[13] * T[2] * T[7] - T[13] * T[3] * T[6];
    o[3] = -T[1] * T[6] * T[11] + T[1] * T[7] * T[10] + T[5] * T[2] * T[11] - T[5] * T[3] * T[10] - T[9] * T[2] * T[7] + T[9] * T[3] * T[6];
    o[4] = -T[4] * T[10] * T[15] + T[4] * T[11] * T[14] + T[8] * T[6] * T[15] - T[8] * T[7] * T[14] - T[12] * T[6] * T[11] + T[12] * T[7] * T[10];
    o[5] = T[0] * T[10] * T[15] - T[0] * T[11] * T[14] - T[8] * T[2] * T[15] + T[8] * T[3] * T[14] + T[12] * T[2] * T[11] - T[12] * T[3] * T[10];
    o[6] = -T[0] * T[6] * T[15] + T[0] * T[7] * T[14] + T[4] * T[2] * T[15] - T[4] * T[3] * T[14] - T[12] * T[2] * T[7] + T[12] * T[3] * T[6];
    o[7] = T[0] * T[6] * T[11] - T[0] * T[7] * T[10] - T[4] * T[2] * T[11] + T[4] * T[3] * T[10] + T[8] * T[2] * T[7] - T[8] * T[3] * T[6];
    o[8] = T[4] * T[9] * T[15] - T[4] * T[11] * T[13] - T[8] * T[5] * T[15] + T[8] * T[7] * T[13] + T[12] * T[5] * T[11] - T[12] * T[7] * T[9];
    o[9] = -T[0] * T[9] * T[15] + T[0] * T[11] * T[13] + T[8] * T[1] * T[15] - T[8] * T[3] * T[13] - T[12] * T[1] * T[11] + T[12] * T[3] * T[9];
    o[10] = T[0] * T[5] * T[15] - T[0] * T[7] * T[13] - T[4] * T[1] * T[15] + T[4] * T[3] * T[13] + T[12] * T[1] * T[7] - T[12] * T[3] * T[5];
    o[11] = -T[0] * T[5] * T[11] + T[0] * T[7] * T[9] + T[4] * T[1] * T[11] - T[4] * T[3] * T[9] - T[8] * T[1] * T[7] + T[8] * T[3] * T[5];
    o[12] = -T[4] * T[9] * T[14] + T[4] * T[10] * T[13] + T[8] * T[5] * T[14] - T[8] * T[6] * T[13] - T[12] * T[5] * T[10] + T[12] * T[6] * T[9];
    o[13] = T[0] * T[9] * T[14] - T[0] * T[10] * T[13] - T[8] * T[1] * T[14] + T[8] * T[2] * T[13] + T[12] * T[1] * T[10] - T[12] * T[2] * T[9];
    o[14] = -T[0] * T[5] * T[14] + T[0] * T[6] * T[13] + T[4] * T[1] * T[14] - T[4] * T[2] * T[13] - T[12] * T[1] * T[6] + T[12] * T[2] * T[5];
    o[15] = T[0] * T[5] * T[10] - T[0] * T[6] * T[9] - T[4] * T[1] * T[10] + T[4] * T[2] * T[9] + T[8] * T[1] * T[6] - T[8] * T[2] * T[5];
    const float det = T[0] * o[0] + T[1] * o[4] + T[2] * o[8] + T[3] * o[12];
    if (det == 0) return;
    const float invdet = 1.0f / det;
    for (int i = 0; i < 16; i++) o[i] *= invdet;
}

static inline float halfArea(f3 e) { return e.x < -kFar ? 0 : (e.x * e.y + e.y * e.z + e.z * e.x); } // tiny_bvh.h:276

// Binned-SAH object-split build over a list of boxes (tiny_bvh.h:1674-1766).  Both the
// triangle BLAS and the instance TLAS go through this; tinybvh shares the same routine.
void Bvh2::buildOverBoxes(const std::vector<f3>& lo, const std::vector<f3>& hi)
{
    const uint32_t n = (uint32_t)lo.size();
    constexpr int B = 8; // BVHBINS
    nodes.assign(n * 2 < 2 ? 2 : n * 2, Node2{});
    prim.resize(n);
    Node2& root = nodes[0];
    root.leftFirst = 0;
    root.count = n;
    root.bmin = f3(kFar);
    root.bmax = f3(-kFar);
    for (uint32_t i = 0; i < n; ++i) {
        prim[i] = i;
        root.bmin = vmin(root.bmin, lo[i]);
        root.bmax = vmax(root.bmax, hi[i]);
    }
    uint32_t nextFree = 2; // slot 1 stays empty so that sibling pairs share a cache line
    std::vector<uint32_t> todo;
    uint32_t cur = 0;
    const f3 minDim = (root.bmax - root.bmin) * 1e-20f;
    f3 keepLMin, keepLMax, keepRMin, keepRMax;
    for (;;) {
        for (;;) {
            Node2& nd = nodes[cur];
            f3 binLo[3][B], binHi[3][B];
            uint32_t cnt[3][B];
            for (int a = 0; a < 3; ++a)
                for (int i = 0; i < B; ++i) binLo[a][i] = f3(kFar), binHi[a][i] = f3(-kFar), cnt[a][i] = 0;
            const f3 ext = nd.bmax - nd.bmin;
            const f3 scale((float)B / ext.x, (float)B / ext.y, (float)B / ext.z);
            const f3 base = nd.bmin;
            for (uint32_t i = 0; i < nd.count; ++i) {
                const uint32_t f = prim[nd.leftFirst + i];
                const f3 c = ((lo[f] + hi[f]) * 0.5f - base) * scale;
                int bx = clampi((int32_t)c.x, 0, B - 1), by = clampi((int32_t)c.y, 0, B - 1), bz = clampi((int32_t)c.z, 0, B - 1);
                binLo[0][bx] = vmin(binLo[0][bx], lo[f]), binHi[0][bx] = vmax(binHi[0][bx], hi[f]), cnt[0][bx]++;
                binLo[1][by] = vmin(binLo[1][by], lo[f]), binHi[1][by] = vmax(binHi[1][by], hi[f]), cnt[1][by]++;
                binLo[2][bz] = vmin(binLo[2][bz], lo[f]), binHi[2][bz] = vmax(binHi[2][bz], hi[f]), cnt[2][bz]++;
            }
            float best = kFar;
            const float rArea = 1.0f / (ext.x * ext.y + ext.y * ext.z + ext.z * ext.x);
            uint32_t bestAxis = 0, bestPos = 0;
            for (int a = 0; a < 3; ++a) {
                if (!((nd.bmax[a] - nd.bmin[a]) > minDim[a])) continue;
                f3 lLo[B - 1], lHi[B - 1], rLo[B - 1], rHi[B - 1];
                float aL[B - 1], aR[B - 1];
                f3 l1(kFar), l2(-kFar), r1(kFar), r2(-kFar);
                uint32_t nL = 0, nR = 0;
                for (int i = 0; i < B - 1; ++i) {
                    lLo[i] = l1 = vmin(l1, binLo[a][i]);
                    rLo[B - 2 - i] = r1 = vmin(r1, binLo[a][B - 1 - i]);
                    lHi[i] = l2 = vmax(l2, binHi[a][i]);
                    rHi[B - 2 - i] = r2 = vmax(r2, binHi[a][B - 1 - i]);
                    nL += cnt[a][i], nR += cnt[a][B - 1 - i];
                    aL[i] = nL == 0 ? kFar : (halfArea(l2 - l1) * (float)nL);
                    aR[B - 2 - i] = nR == 0 ? kFar : (halfArea(r2 - r1) * (float)nR);
                }
                for (int i = 0; i < B - 1; ++i) {
                    const float C = 1 /*C_TRAV*/ + rArea * 1 /*C_INT*/ * (aL[i] + aR[i]);
                    if (C < best) {
                        best = C, bestAxis = (uint32_t)a, bestPos = (uint32_t)i;
                        keepLMin = lLo[i], keepRMin = rLo[i], keepLMax = lHi[i], keepRMax = rHi[i];
                    }
                }
            }
            if (best >= (float)nd.count) break; // a leaf is cheaper
            uint32_t j = nd.leftFirst + nd.count, src = nd.leftFirst;
            const float sc = scale[(int)bestAxis], b0 = base[(int)bestAxis];
            for (uint32_t i = 0; i < nd.count; ++i) {
                const uint32_t f = prim[src];
                int32_t bi = (uint32_t)(((lo[f][(int)bestAxis] + hi[f][(int)bestAxis]) * 0.5f - b0) * sc);
                bi = clampi(bi, 0, B - 1);
                if ((uint32_t)bi <= bestPos) src++;
                else {
                    uint32_t t = prim[src];
                    prim[src] = prim[--j];
                    prim[j] = t;
                }
            }
            const uint32_t nl = src - nd.leftFirst, nr = nd.count - nl;
            if (nl == 0 || nr == 0) break;
            const uint32_t L = nextFree++, R = nextFree++;
            nodes[L].bmin = keepLMin, nodes[L].bmax = keepLMax, nodes[L].leftFirst = nd.leftFirst, nodes[L].count = nl;
            nodes[R].bmin = keepRMin, nodes[R].bmax = keepRMax, nodes[R].leftFirst = j, nodes[R].count = nr;
            nd.leftFirst = L, nd.count = 0;
            todo.push_back(R);
            cur = L;
        }
        if (todo.empty()) break;
        cur = todo.back();
        todo.pop_back();
    }
    bmin = nodes[0].bmin, bmax = nodes[0].bmax;
    nodes.resize(nextFree);
}

void Blas::build() // tiny_bvh.h:1636-1650 (fragment boxes of de-indexed triangles)
{
    const size_t n = tri.size() / 3;
    std::vector<f3> lo(n), hi(n);
    for (size_t i = 0; i < n; ++i) {
        const f3 a = tri[3 * i].xyz(), b = tri[3 * i + 1].xyz(), c = tri[3 * i + 2].xyz();
        lo[i] = vmin(a, vmin(b, c));
        hi[i] = vmax(a, vmax(b, c));
    }
    bvh.buildOverBoxes(lo, hi);
}

static inline float slab(const RayQ& r, const Node2& n) // tiny_bvh.h:6920-6932
{
    float tx1 = (n.bmin.x - r.O.x) * r.rD.x, tx2 = (n.bmax.x - r.O.x) * r.rD.x;
    float tmin = fminf_(tx1, tx2), tmax = fmaxf_(tx1, tx2);
    float ty1 = (n.bmin.y - r.O.y) * r.rD.y, ty2 = (n.bmax.y - r.O.y) * r.rD.y;
    tmin = fmaxf_(tmin, fminf_(ty1, ty2));
    tmax = fminf_(tmax, fmaxf_(ty1, ty2));
    float tz1 = (n.bmin.z - r.O.z) * r.rD.z, tz2 = (n.bmax.z - r.O.z) * r.rD.z;
    tmin = fmaxf_(tmin, fminf_(tz1, tz2));
    tmax = fminf_(tmax, fmaxf_(tz1, tz2));
    if (tmax >= tmin && tmin < r.hit.t && tmax >= 0) return tmin;
    return kFar;
}

// Möller–Trumbore, tiny_bvh.h:6815-6843.  Returns true when the hit record was shortened.
static inline bool triTest(RayQ& r, const f4* tri, uint32_t idx)
{
    const f3 v0 = tri[idx * 3].xyz();
    const f3 e1 = tri[idx * 3 + 1].xyz() - v0;
    const f3 e2 = tri[idx * 3 + 2].xyz() - v0;
    const f3 h = cross(r.D, e2);
    const float a = dot(e1, h);
    if (fabsf(a) < 0.0000001f) return false;
    const float f = 1 / a;
    const f3 s = r.O - v0;
    const float u = f * dot(s, h);
    if (u < 0 || u > 1) return false;
    const f3 q = cross(s, e1);
    const float v = f * dot(r.D, q);
    if (v < 0 || u + v > 1) return false;
    const float t = f * dot(e2, q);
    if (t > r.tmin && t < r.hit.t) {
        r.hit.t = t, r.hit.u = u, r.hit.v = v;
        r.hit.prim = idx, r.hit.inst = r.instIdx;
        return true;
    }
    return false;
}

void Blas::intersect(RayQ& r, uint64_t* nodeVisits, uint64_t* triTests) const // tiny_bvh.h:2245-2292
{
    const Node2* nodes = bvh.nodes.data();
    const Node2* node = nodes;
    const Node2* stack[64];
    int sp = 0;
    for (;;) {
        if (nodeVisits) ++*nodeVisits;
        if (node->count) {
            for (uint32_t i = 0; i < node->count; ++i) {
                if (triTests) ++*triTests;
                triTest(r, tri.data(), bvh.prim[node->leftFirst + i]);
            }
            if (!sp) break;
            node = stack[--sp];
            continue;
        }
        const Node2* c1 = nodes + node->leftFirst;
        const Node2* c2 = c1 + 1;
        float d1 = slab(r, *c1), d2 = slab(r, *c2);
        if (d1 > d2) {
            float t = d1; d1 = d2; d2 = t;
            const Node2* p = c1; c1 = c2; c2 = p;
        }
        if (d1 == kFar) {
            if (!sp) break;
            node = stack[--sp];
        } else {
            node = c1;
            if (d2 != kFar) stack[sp++] = c2;
        }
    }
}

bool Blas::occluded(const RayQ& r0) const
{
    // any-hit: same traversal, stop at the first accepted triangle
    RayQ r = r0;
    const Node2* nodes = bvh.nodes.data();
    const Node2* node = nodes;
    const Node2* stack[64];
    int sp = 0;
    for (;;) {
        if (node->count) {
            for (uint32_t i = 0; i < node->count; ++i)
                if (triTest(r, tri.data(), bvh.prim[node->leftFirst + i])) return true;
            if (!sp) break;
            node = stack[--sp];
            continue;
        }
        const Node2* c1 = nodes + node->leftFirst;
        const Node2* c2 = c1 + 1;
        float d1 = slab(r, *c1), d2 = slab(r, *c2);
        if (d1 > d2) {
            float t = d1; d1 = d2; d2 = t;
            const Node2* p = c1; c1 = c2; c2 = p;
        }
        if (d1 == kFar) {
            if (!sp) break;
            node = stack[--sp];
        } else {
            node = c1;
            if (d2 != kFar) stack[sp++] = c2;
        }
    }
    return false;
}

void Instance::update(const Blas& b) // tiny_bvh.h:6718-6732
{
    invert4x4RowMajor(T, invT);
    bmin = f3(kFar), bmax = f3(-kFar);
    for (int j = 0; j < 8; ++j) {
        const f3 p(j & 1 ? b.bvh.bmax.x : b.bvh.bmin.x, j & 2 ? b.bvh.bmax.y : b.bvh.bmin.y, j & 4 ? b.bvh.bmax.z : b.bvh.bmin.z);
        const f3 t = xformPoint(p, T);
        bmin = vmin(bmin, t), bmax = vmax(bmax, t);
    }
}

void Tlas::build(const std::vector<Blas>& blas) // tiny_bvh.h:1565-1603
{
    std::vector<f3> lo(inst.size()), hi(inst.size());
    for (size_t i = 0; i < inst.size(); ++i) {
        for (int k = 0; k < 16; ++k) inst[i].invT[k] = (k % 5 == 0) ? 1.f : 0.f;
        inst[i].update(blas[inst[i].blas]);
        lo[i] = inst[i].bmin, hi[i] = inst[i].bmax;
    }
    bvh.buildOverBoxes(lo, hi);
}

void Tlas::intersect(const std::vector<Blas>& blas, RayQ& r, uint64_t* nodeVisits, uint64_t* triTests) const // tiny_bvh.h:2294-2353
{
    const Node2* nodes = bvh.nodes.data();
    const Node2* node = nodes;
    const Node2* stack[64];
    int sp = 0;
    for (;;) {
        if (nodeVisits) ++*nodeVisits;
        if (node->count) {
            for (uint32_t i = 0; i < node->count; ++i) {
                const uint32_t ii = bvh.prim[node->leftFirst + i];
                const Instance& in = inst[ii];
                RayQ tmp;
                tmp.O = xformPoint(r.O, in.invT);
                tmp.D = xformVector(r.D, in.invT); // not re-normalised: t stays in world units
                tmp.instIdx = ii;
                tmp.hit = r.hit;
                tmp.tmin = r.tmin;
                tmp.rD = f3(safeRcp(tmp.D.x), safeRcp(tmp.D.y), safeRcp(tmp.D.z));
                blas[in.blas].intersect(tmp, nodeVisits, triTests);
                r.hit = tmp.hit;
            }
            if (!sp) break;
            node = stack[--sp];
            continue;
        }
        const Node2* c1 = nodes + node->leftFirst;
        const Node2* c2 = c1 + 1;
        float d1 = slab(r, *c1), d2 = slab(r, *c2);
        if (d1 > d2) {
            float t = d1; d1 = d2; d2 = t;
            const Node2* p = c1; c1 = c2; c2 = p;
        }
        if (d1 == kFar) {
            if (!sp) break;
            node = stack[--sp];
        } else {
            node = c1;
            if (d2 != kFar) stack[sp++] = c2;
        }
    }
}

bool Tlas::occluded(const std::vector<Blas>& blas, const RayQ& r) const // tiny_bvh.h:2398-2450
{
    const Node2* nodes = bvh.nodes.data();
    const Node2* node = nodes;
    const Node2* stack[64];
    int sp = 0;
    for (;;) {
        if (node->count) {
            for (uint32_t i = 0; i < node->count; ++i) {
                const Instance& in = inst[bvh.prim[node->leftFirst + i]];
                RayQ tmp;
                tmp.O = xformPoint(r.O, in.invT);
                tmp.D = xformVector(r.D, in.invT);
                tmp.hit = r.hit;
                tmp.tmin = r.tmin;
                tmp.instIdx = 0;
                tmp.rD = f3(safeRcp(tmp.D.x), safeRcp(tmp.D.y), safeRcp(tmp.D.z));
                if (blas[in.blas].occluded(tmp)) return true;
            }
            if (!sp) break;
            node = stack[--sp];
            continue;
        }
        const Node2* c1 = nodes + node->leftFirst;
        const Node2* c2 = c1 + 1;
        float d1 = slab(r, *c1), d2 = slab(r, *c2);
        if (d1 > d2) {
            float t = d1; d1 = d2; d2 = t;
            const Node2* p = c1; c1 = c2; c2 = p;
        }
        if (d1 == kFar) {
            if (!sp) break;
            node = stack[--sp];
        } else {
            node = c1;
            if (d2 != kFar) stack[sp++] = c2;
        }
    }
    return false;
}

} // namespace orc
