"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, the POD layouts match the reference's, the backend refuses to run without a GPU
(no CPU fallback), and the host mirror builds the reference's Cornell box."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gknextrenderer_b200 as gk
from gknextrenderer_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_are_exported(built):
    hdr = open(os.path.join(ROOT, "include", "gknext_cuda.h")).read()
    declared = set(re.findall(r"\b(gk_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(N.CUDA_LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in gknext_cuda.h but not exported"
    assert declared == set(N.CUDA_API), declared ^ set(N.CUDA_API)
    assert gk.cuda_lib().gk_abi_version() == 1


def test_compositor_header_symbols_are_exported(built):
    """include/gknext_compositor.h (multi-GPU compositor over NCCL, lib/libgknext_comp.so): every declared symbol is exported and bound;
    no collective is called here (that needs GPUs) - only the argument checks that run before any NCCL call."""
    hdr = open(os.path.join(ROOT, "include", "gknext_compositor.h")).read()
    declared = set(re.findall(r"\b(gkc_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 10
    lib = C.CDLL(N.COMP_LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in gknext_compositor.h but not exported"
    assert declared == set(N.COMP_API), declared ^ set(N.COMP_API)
    comp = gk.comp_lib()
    out = C.c_void_p()
    ident = (C.c_ubyte * N.GKC_UNIQUE_ID_BYTES)()
    assert comp.gkc_create(None, 0, 2, ident, N.GKC_UNIQUE_ID_BYTES, C.byref(out)) < 0 and b"invalid argument" in comp.gkc_last_error()
    assert comp.gkc_get_unique_id(ident, 8) < 0
    assert comp.gkc_composite_frame(None) < 0 and comp.gkc_world(None) == 0 and comp.gkc_rank(None) == -1
    comp.gkc_destroy(None)


def test_pod_layouts_match_reference():
    U = N.GkUniformBufferObject
    assert C.sizeof(U) == 784
    offs = {f: getattr(U, f).offset for f in ("ViewportRect", "SunViewProjection", "Aperture", "TotalFrames", "NumberOfSamples", "BFSize", "SuperResolution")}
    # offsets probed from the reference's C++ view of BasicTypes.slang
    assert offs == {"ViewportRect": 512, "SunViewProjection": 576, "Aperture": 640, "TotalFrames": 668, "NumberOfSamples": 676, "BFSize": 756, "SuperResolution": 776}
    assert C.sizeof(N.GkNodeProxy) == 208 and N.GkNodeProxy.worldTS.offset == 16 and N.GkNodeProxy.matId.offset == 144
    assert C.sizeof(N.GkMaterial) == 64 and N.GkMaterial.MaterialModel.offset == 36
    assert C.sizeof(N.GkVertex) == 52 and C.sizeof(N.GkLightObject) == 80 and C.sizeof(N.GkRayCastResult) == 48


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback_without_gpu(built):
    with pytest.raises(gk.GkError) as e:
        gk.Renderer(64, 64)
    assert e.value.status in (N.GK_ERR_CUDA, N.GK_ERR_UNSUPPORTED)
    assert "no CUDA device" in str(e.value) or "sm_100a" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing in the package may import, link or load it."""
    pkg = os.path.join(ROOT, "gknextrenderer_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_lib" not in text and "orc_" not in text, f"{f} touches the oracle"


def test_cornell_scene_matches_reference_definition(built):
    eng = gk.Engine("cornell")
    d = eng.scene_desc().contents
    assert d.modelCount == 3 and d.materialCount == 6 and d.lightCount == 1
    tris = [d.models[m].indexCount // 3 for m in range(3)]
    assert tris == [12, 12, 1024]  # CornellBox.cpp, CreateBox, CreateSphere(32x16)
    assert eng.triangles() == 1048
    mats = [d.materials[i] for i in range(6)]
    assert [m.MaterialModel for m in mats] == [0, 0, 0, 4, 0, 5]
    assert np.allclose(mats[0].Diffuse[:3], [0.65, 0.05, 0.05]) and np.allclose(mats[3].Diffuse[:3], [2000, 2000, 2000])
    assert mats[5].Fuzziness == pytest.approx(0.01) and mats[5].RefractionIndex == pytest.approx(1.45)
    nodes, n = eng.update_nodes()
    assert n == 3
    assert [nodes[i].modelId for i in range(3)] == [0, 20, 10]  # model*10 + section (Scene.cpp:494)
    assert [nodes[i].instanceId for i in range(3)] == [0, 1, 2]
    w = np.array(nodes[1].worldTS[:], np.float32).reshape(4, 4).T
    assert np.allclose(w[:3, 3], [1.30, 1.01, 0.80])
    assert np.allclose(w[:3, :3] @ w[:3, :3].T, np.eye(3), atol=1e-6)  # pure rotation
    w2 = np.array(nodes[2].worldTS[:], np.float32).reshape(4, 4).T
    assert np.allclose(np.linalg.norm(w2[:3, :3], axis=0), [1, 2, 1], atol=1e-6)
    # light quad of the Cornell box (CornellBox.cpp:99-119)
    L = d.lights[0]
    assert L.lightMatIdx == 3 and L.normal_area[1] == -1.0
    assert L.normal_area[3] == pytest.approx((393 - 163) / 555 * 5.55 * (432 - 202) / 555 * 5.55, rel=1e-5)


def test_ubo_fill_follows_engine(built):
    eng = gk.Engine("cornell")
    eng.set(TAA=0, NumberOfSamples=8, NumberOfBounces=4, TemporalFrames=16, Denoiser=0, ProgressiveRender=1)
    u = eng.ubo(640, 360)
    P = np.array(u.Projection[:], np.float32).reshape(4, 4).T
    t = np.tan(np.radians(40.0) / 2)
    assert P[0, 0] == pytest.approx(1 / (640 / 360 * t), rel=1e-6)
    assert P[1, 1] == pytest.approx(-1 / t, rel=1e-6)  # Vulkan y flip (Engine.cpp:678)
    assert P[2, 2] == pytest.approx(10000.0 / (0.1 - 10000.0), rel=1e-6) and P[3, 2] == -1
    MV, MVI = np.array(u.ModelView[:], np.float64).reshape(4, 4).T, np.array(u.ModelViewInverse[:], np.float64).reshape(4, 4).T
    assert np.allclose(MV @ MVI, np.eye(4), atol=1e-5)
    PI = np.array(u.ProjectionInverse[:], np.float64).reshape(4, 4).T
    assert np.allclose(P.astype(np.float64) @ PI, np.eye(4), atol=1e-4)
    assert np.allclose(MVI[:3, 3], [0, 2.78, 10.78], atol=1e-5)
    assert u.TemporalFrames == 1024 // 16 and u.ProgressiveRender == 1 and u.BFSize == 0
    assert u.HasSky == 0 and u.HasSun == 0 and u.NumberOfSamples == 8 and u.NumberOfBounces == 4
    assert tuple(u.ViewportRect) == (0, 0, 640, 360)
    eng.set(ProgressiveRender=0, Denoiser=1, TAA=1)
    u2 = eng.ubo(640, 360)
    assert u2.TemporalFrames == 16 and u2.BFSize == 5 and u2.BFSigma == 2.0 and u2.BFSigmaLum == 3.0
    P2 = np.array(u2.Projection[:], np.float32).reshape(4, 4).T
    # Halton(2,3) jitter of frame 0 is (0.5, 1/3) - 0.5 (Engine.cpp:695-702)
    assert P2[0, 2] == 0 and P2[1, 2] == pytest.approx((1.0 / 3.0 - 0.5) / 360 * 2.0, rel=1e-5)
    o, d = eng.screen_ray(320, 180, 640, 360)
    assert np.allclose(o, [0, 2.78, 10.78], atol=1e-5) and np.allclose(d / np.linalg.norm(d), [0, 0, -1], atol=1e-3)  # prevUBO_ carries the TAA jitter


def test_procedural_scenes_have_the_named_size(built):
    eng = gk.Engine("room", 100000, 1234)
    assert 100000 <= eng.triangles(True) < 103000
    nodes, n = eng.update_nodes()
    assert n > 50
    eng = gk.Engine("bricks", 2000, 42)
    nodes, n = eng.update_nodes()
    assert n == 2001
    before = np.array([nodes[i].worldTS[12] for i in range(n)])
    eng.step_scene(1)
    nodes, n = eng.update_nodes()
    after = np.array([nodes[i].worldTS[12] for i in range(n)])
    assert 1 <= (before != after).sum() <= 20  # 1 % of the bricks move per frame
    eng = gk.Engine("city", 3, 4, 7, 4)
    assert eng.triangles() == 3 * 12 * 16 + 12 and eng.update_nodes()[1] == 17


def _proxy_bytes(nodes, n):
    import ctypes
    return np.frombuffer(ctypes.string_at(nodes, n * ctypes.sizeof(gk.GkNodeProxy)), np.uint8).reshape(n, -1).copy()


def test_incremental_update_nodes_equals_the_full_pass(built):
    """Scene::MarkNodeDirty: re-evaluating only the marked and the still-settling nodes must leave the proxy array exactly as
    the full loop of the reference (Scene.cpp:464-511) writes it, and ChangedProxies() must name every record that differs."""
    inc, full = gk.Engine("bricks", 3000, 42), gk.Engine("bricks", 3000, 42)
    for e in (inc, full):
        e.update_nodes()
    assert inc.changed_proxies() is None  # the first pass is a full one ...
    inc.update_nodes(); full.mark_dirty(); full.update_nodes()
    assert inc.changed_proxies().size == 3001  # ... and every node settles from its placeholder previous transform in the second
    prev = _proxy_bytes(*inc.update_nodes())
    for frame in range(1, 7):
        inc.step_scene(frame)
        full.step_scene(frame); full.mark_dirty()
        a, b = _proxy_bytes(*inc.update_nodes()), _proxy_bytes(*full.update_nodes())
        assert np.array_equal(a, b), f"frame {frame}"
        changed = inc.changed_proxies()
        assert changed is not None and full.changed_proxies() is None
        differs = np.flatnonzero((a != prev).any(axis=1))
        assert set(differs) <= set(changed.tolist()) and 0 < changed.size <= 2 * 30 + 2
        prev = a
    # nothing marked: a frame without motion settles the last movers, then the scene is clean
    a = _proxy_bytes(*inc.update_nodes()); full.mark_dirty(); b = _proxy_bytes(*full.update_nodes())
    assert np.array_equal(a, b)
    inc.set_node_translation(5, 1.0, 0.5, 1.0); full.set_node_translation(5, 1.0, 0.5, 1.0); full.mark_dirty()
    a, b = _proxy_bytes(*inc.update_nodes()), _proxy_bytes(*full.update_nodes())
    assert np.array_equal(a, b) and 5 in inc.changed_proxies()
