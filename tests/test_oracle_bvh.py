"""Pins the oracle's BVH restatement (oracle/orc_bvh.cpp):
  1. against the committed golden vectors produced by the REAL tinybvh (tests/tools/make_golden.py),
  2. against the real tinybvh itself (oracle/_ref) when that library is present, bit for bit,
     including the built trees node by node,
  3. against a brute-force closest hit that uses no BVH at all.
CPU only."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import gknextrenderer_b200 as gk
import oracle_lib as ol

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _scene_hash(eng):
    d = eng.scene_desc().contents
    h = hashlib.sha256()
    for m in range(d.modelCount):
        md = d.models[m]
        h.update(C.string_at(md.vertices, md.vertexCount * 52))
        h.update(C.string_at(md.indices, md.indexCount * 4))
    nodes, n = eng.update_nodes()
    h.update(C.string_at(nodes, n * 208))
    return h.hexdigest()


def _engine(name):
    eng = gk.Engine("cornell") if name == "cornell" else gk.Engine("room", 20000, 99)
    eng.set(TAA=0)
    return eng


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name,fixture", [("cornell", "cornell_tinybvh.npz"), ("room20k", "room20k_tinybvh.npz")])
def test_oracle_reproduces_tinybvh_golden(built, name, fixture):
    g = np.load(os.path.join(GOLDEN, fixture))
    eng = _engine(name)
    if _scene_hash(eng) != str(g["scene_sha256"]):
        pytest.skip("host libm produced a different scene than the one the fixture was generated on")
    nodes, n = eng.update_nodes()
    orc = ol.OracleScene(eng.scene_desc(), nodes, n)
    tuv, ids = orc.intersect(g["rays"], threads=4)
    assert np.array_equal(ids, g["ids"]), "hit (triangle, instance) ids differ from tinybvh"
    assert np.array_equal(_bits(tuv), _bits(g["tuv"])), "t/u/v differ bitwise from tinybvh"
    assert "tinybvh" in str(g["source"])


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref (real tinybvh) not built")
@pytest.mark.parametrize("name", ["cornell", "room20k"])
def test_oracle_matches_real_tinybvh_bitwise(built, name):
    eng = _engine(name)
    nodes, n = eng.update_nodes()
    desc = eng.scene_desc()
    orc, ref = ol.OracleScene(desc, nodes, n), ol.OracleScene(desc, nodes, n, use_ref=True)
    # identical trees: node count and every 32-byte node
    for m in range(desc.contents.modelCount):
        a, b = orc.lib.orc_blas_node_count(orc.h, m), ref.lib.ref_blas_node_count(ref.h, m)
        assert a == b
        na, nb = np.zeros((a, 8), np.uint32), np.zeros((b, 8), np.uint32)
        orc.lib.orc_blas_nodes(orc.h, m, ol.ptr(na))
        ref.lib.ref_blas_nodes(ref.h, m, ol.ptr(nb))
        keep = np.ones(a, bool)
        keep[1] = False  # slot 1 is an unused alignment filler in both
        assert np.array_equal(na[keep], nb[keep]), f"BLAS {m} differs"
    a, b = orc.lib.orc_tlas_node_count(orc.h), ref.lib.ref_tlas_node_count(ref.h)
    assert a == b
    na, nb = np.zeros((a, 8), np.uint32), np.zeros((b, 8), np.uint32)
    orc.lib.orc_tlas_nodes(orc.h, ol.ptr(na))
    ref.lib.ref_tlas_nodes(ref.h, ol.ptr(nb))
    keep = np.ones(a, bool)
    if a > 1:
        keep[1] = False
    assert np.array_equal(na[keep], nb[keep]), "TLAS differs"
    # identical query results on coherent and incoherent rays
    rng = np.random.default_rng(7)
    W, H = (320, 180)
    rays = ol.primary_rays(eng.ubo(W, H), W, H)
    extra = np.zeros((20000, 8), np.float32)
    extra[:, 0:3] = rng.uniform(-6, 6, (20000, 3))
    d = rng.normal(size=(20000, 3))
    extra[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    extra[:, 7] = 1000.0
    rays = np.concatenate([rays, extra])
    t1, i1 = orc.intersect(rays, threads=4)
    t2, i2 = ref.intersect(rays, threads=4)
    assert np.array_equal(i1, i2)
    assert np.array_equal(_bits(t1), _bits(t2))


def test_oracle_bvh_agrees_with_bruteforce(built):
    eng = _engine("cornell")
    nodes, n = eng.update_nodes()
    orc = ol.OracleScene(eng.scene_desc(), nodes, n)
    rays = ol.primary_rays(eng.ubo(160, 90), 160, 90)
    t1, i1 = orc.intersect(rays)
    t2, i2 = orc.intersect_bruteforce(rays)
    # same nearest distance everywhere; ids may differ only on exact-t ties
    assert np.array_equal(_bits(t1[:, 0]), _bits(t2[:, 0]))
    differ = (i1 != i2).any(axis=1)
    assert differ.sum() <= 4, f"{differ.sum()} id differences between BVH and brute force"


def test_oracle_tmin_and_miss_conventions(built):
    eng = _engine("cornell")
    nodes, n = eng.update_nodes()
    orc = ol.OracleScene(eng.scene_desc(), nodes, n)
    # a ray from inside the box towards the back wall, then the same ray with tmin beyond the wall
    r = np.array([[0, 2.7, 0, 0.0, 0, 0, -1, 1000.0], [0, 2.7, 0, 5.0, 0, 0, -1, 1000.0], [0, 50, 0, 0, 0, 1, 0, 1000.0]], np.float32)
    tuv, ids = orc.intersect(r)
    assert ids[0, 1] == 0 and abs(tuv[0, 0] - 2.775) < 1e-5
    assert ids[1, 1] == 0xFFFFFFFF and tuv[1, 0] == 1000.0  # nothing beyond the wall
    assert ids[2, 1] == 0xFFFFFFFF and ids[2, 0] == 0xFFFFFFFF


def test_oracle_raycast_record(built):
    """RayCastInCPU mirror: hit point, untransformed-length normal (world * n), node id, T."""
    eng = _engine("cornell")
    nodes, n = eng.update_nodes()
    orc = ol.OracleScene(eng.scene_desc(), nodes, n)
    od = np.array([[0, 2.78, 10.78, 0, 0, -1], [0, 2.78, 10.78, 0, 1, 0]], np.float32)
    res = orc.raycast(od)
    assert res[0].Hitted == 1 and res[0].InstanceId == 0
    assert abs(res[0].T - (10.78 + 2.775)) < 1e-4
    assert np.allclose(res[0].Normal[:3], [0, 0, 1], atol=1e-6)
    assert np.allclose(res[0].HitPoint[:3], [0, 2.78, -2.775], atol=1e-4)
    assert res[1].Hitted == 0
