"""Generates tests/golden/*.npz by running the REAL tinybvh (oracle/_ref, compiled from
/root/reference/src/ThirdParty/tinybvh/tiny_bvh.h) on rays of the built-in scenes.
Run in the build container (the reference tree must be present):  python tests/tools/make_golden.py
The fixtures pin oracle/orc_bvh.cpp wherever the reference tree is absent (e.g. the GPU box)."""
import ctypes as C
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import gknextrenderer_b200 as gk  # noqa: E402
import oracle_lib as ol  # noqa: E402


def scene_hash(eng):
    d = eng.scene_desc().contents
    h = hashlib.sha256()
    for m in range(d.modelCount):
        md = d.models[m]
        h.update(C.string_at(md.vertices, md.vertexCount * 52))
        h.update(C.string_at(md.indices, md.indexCount * 4))
    nodes, n = eng.update_nodes()
    h.update(C.string_at(nodes, n * 208))
    return h.hexdigest()


def random_rays(rng, n, lo, hi, tmax=1000.0):
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros((n, 8), np.float32)
    r[:, 0:3], r[:, 4:7], r[:, 7] = o, d, tmax
    return r


def main():
    assert ol.have_ref(), "oracle/_ref/libtinybvh_ref.so missing: run make -C oracle with /root/reference present"
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(20261017)
    # 1. Cornell box: primary rays of the 640x360 frame (every 16th) + incoherent rays inside the box
    eng = gk.Engine("cornell")
    eng.set(TAA=0)
    hsh = scene_hash(eng)
    nodes, n = eng.update_nodes()
    ref = ol.OracleScene(eng.scene_desc(), nodes, n, use_ref=True)
    rays = np.concatenate([ol.primary_rays(eng.ubo(640, 360), 640, 360)[::16], random_rays(rng, 8192, (-2.7, 0.05, -2.7), (2.7, 5.5, 2.7))])
    tuv, ids = ref.intersect(rays)
    np.savez_compressed(os.path.join(out, "cornell_tinybvh.npz"), rays=rays, tuv=tuv, ids=ids, scene_sha256=hsh, source=ref.lib.ref_version().decode())
    print("cornell", len(rays), "hits", int((ids[:, 1] != 0xffffffff).sum()))
    # 2. small procedural room (many instances, rotations, non-uniform scales)
    eng = gk.Engine("room", 20000, 99)
    eng.set(TAA=0)
    hsh = scene_hash(eng)
    nodes, n = eng.update_nodes()
    ref = ol.OracleScene(eng.scene_desc(), nodes, n, use_ref=True)
    rays = np.concatenate([ol.primary_rays(eng.ubo(320, 180), 320, 180)[::8], random_rays(rng, 8192, (-9.5, 0.1, -9.5), (9.5, 3.9, 9.5))])
    tuv, ids = ref.intersect(rays)
    np.savez_compressed(os.path.join(out, "room20k_tinybvh.npz"), rays=rays, tuv=tuv, ids=ids, scene_sha256=hsh, source=ref.lib.ref_version().decode())
    print("room20k", len(rays), "hits", int((ids[:, 1] != 0xffffffff).sum()), "instances", n)


if __name__ == "__main__":
    main()
