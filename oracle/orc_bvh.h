// orc_bvh.h — CPU restatement of the reference's CPU ray query.
//
// TEST INFRASTRUCTURE ONLY (see orc_math.h).  This restates, in our own words,
// the parts of tinybvh 1.3.8 (vendored by the reference at
// src/ThirdParty/tinybvh/tiny_bvh.h) that FCPUAccelerationStructure uses
// (src/Assets/CPUAccelerationStructure.cpp:171-307):
//
//   binned-SAH object-split builder ........ tiny_bvh.h:1674-1766 (8 bins, C_INT=C_TRAV=1)
//   triangle fragments / root bounds ....... tiny_bvh.h:1605-1673
//   TLAS over instance boxes ............... tiny_bvh.h:1565-1603
//   instance world box + matrix inverse .... tiny_bvh.h:6718-6758
//   ordered 2-ary traversal (BLAS, TLAS) ... tiny_bvh.h:2245-2353
//   Möller–Trumbore triangle test .......... tiny_bvh.h:6815-6843
//   slab test .............................. tiny_bvh.h:6920-6932
//   ray setup (normalise, safe reciprocal) . tiny_bvh.h:329, 562-567
//
// It is validated bit-for-bit (t, u, v, prim, inst) against the real tinybvh
// compiled from the reference tree (oracle/ref_glue.cpp -> oracle/_ref/) by
// tests/test_oracle_vs_ref.py.  The only extension is `tmin`: tinybvh accepts
// t > 0; the shaders' RayQuery uses TMin = 1e-3 (Shading.slang:665,712), so the
// query takes a lower bound, with tmin = 0 reproducing tinybvh exactly.
#pragma once
#include "orc_math.h"
#include <vector>

namespace orc {

constexpr float kFar = 1e30f; // BVH_FAR, tiny_bvh.h:129

struct Node2 { // 32-byte two-child node, tiny_bvh.h "Wald" layout
    f3 bmin;
    uint32_t leftFirst;
    f3 bmax;
    uint32_t count; // >0: leaf over prim[leftFirst .. leftFirst+count)
};

struct Hit {
    float t, u, v;
    uint32_t prim, inst;
};

struct RayQ {
    f3 O, D, rD;
    float tmin;
    Hit hit;
    uint32_t instIdx;
};

struct Bvh2 {
    std::vector<Node2> nodes; // node 1 is unused (as in tinybvh)
    std::vector<uint32_t> prim;
    f3 bmin, bmax;
    // statistics of the last traversal (for the roofline's bytes/ray model)
    void buildOverBoxes(const std::vector<f3>& lo, const std::vector<f3>& hi);
};

struct Blas {
    std::vector<f4> tri; // 3 vertices per triangle, de-indexed (CPUAccelerationStructure.cpp:198-200)
    Bvh2 bvh;
    void build();
    void intersect(RayQ& r, uint64_t* nodeVisits, uint64_t* triTests) const;
    bool occluded(const RayQ& r) const;
};

struct Instance {
    float T[16];    // row-major world transform (transpose of glm's, CPUAccelerationStructure.cpp:249)
    float invT[16]; // row-major inverse
    f3 bmin, bmax;
    uint32_t blas;
    void update(const Blas& b);
};

struct Tlas {
    std::vector<Instance> inst;
    Bvh2 bvh;
    void build(const std::vector<Blas>& blas);
    // closest hit; r.hit.t must be preset to tmax
    void intersect(const std::vector<Blas>& blas, RayQ& r, uint64_t* nodeVisits = nullptr, uint64_t* triTests = nullptr) const;
    bool occluded(const std::vector<Blas>& blas, const RayQ& r) const;
};

float safeRcp(float x);
RayQ makeRay(f3 origin, f3 dir, float tmin, float tmax);
void invert4x4RowMajor(const float* T, float* out);
f3 xformPoint(f3 v, const float* T);
f3 xformVector(f3 v, const float* T);

} // namespace orc
