"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when present, the real tinybvh
glue (oracle/_ref/libtinybvh_ref.so).  TEST INFRASTRUCTURE: imported by tests/, the smoke
check and bench.py's CPU-baseline legs only — never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libtinybvh_ref.so")

_P = C.c_void_p


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def load_oracle():
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    lib = C.CDLL(ORACLE_SO)
    lib.orc_scene_create.restype = _P
    lib.orc_scene_create.argtypes = [_P, _P, C.c_uint32]
    lib.orc_scene_set_nodes.argtypes = [_P, _P, C.c_uint32]
    lib.orc_scene_destroy.argtypes = [_P]
    lib.orc_intersect.restype = C.c_double
    lib.orc_intersect.argtypes = [_P, _P, C.c_uint32, _P, _P, C.c_int, _P]
    lib.orc_intersect_bruteforce.argtypes = [_P, _P, C.c_uint32, _P, _P]
    lib.orc_intersect_bruteforce_mt.argtypes = [_P, _P, C.c_uint32, _P, _P, C.c_int]
    lib.orc_raycast.argtypes = [_P, _P, C.c_uint32, _P]
    lib.orc_raycast_task.argtypes = [_P, _P, C.c_uint32]
    lib.orc_bake_probes.argtypes = [_P, _P, _P, _P, C.c_uint32, C.c_uint32, C.c_int]
    lib.orc_blas_node_count.restype = C.c_uint32
    lib.orc_blas_node_count.argtypes = [_P, C.c_uint32]
    lib.orc_blas_nodes.argtypes = [_P, C.c_uint32, _P]
    lib.orc_tlas_node_count.restype = C.c_uint32
    lib.orc_tlas_node_count.argtypes = [_P]
    lib.orc_tlas_nodes.argtypes = [_P, _P]
    lib.orc_render.argtypes = [_P, _P, C.c_uint32, C.c_uint32, _P, _P] + [_P] * 9 + [C.c_int]
    lib.orc_reproject.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_int, C.c_int] + [_P] * 7
    lib.orc_denoise_jbf.argtypes = [_P, C.c_uint32, C.c_uint32] + [_P] * 7
    lib.orc_float_to_half.restype = C.c_uint16
    lib.orc_float_to_half.argtypes = [C.c_float]
    lib.orc_half_to_float.restype = C.c_float
    lib.orc_half_to_float.argtypes = [C.c_uint16]
    lib.orc_glm_to_half.restype = C.c_uint16
    lib.orc_glm_to_half.argtypes = [C.c_float]
    return lib


def have_ref():
    return os.path.exists(REF_SO)


def load_ref():
    lib = C.CDLL(REF_SO)
    lib.ref_scene_create.restype = _P
    lib.ref_scene_create.argtypes = [_P, _P, C.c_uint32]
    lib.ref_scene_destroy.argtypes = [_P]
    lib.ref_intersect.restype = C.c_double
    lib.ref_intersect.argtypes = [_P, _P, C.c_uint32, _P, _P, C.c_int]
    lib.ref_blas_node_count.restype = C.c_uint32
    lib.ref_blas_node_count.argtypes = [_P, C.c_uint32]
    lib.ref_blas_nodes.argtypes = [_P, C.c_uint32, _P]
    lib.ref_tlas_node_count.restype = C.c_uint32
    lib.ref_tlas_node_count.argtypes = [_P]
    lib.ref_tlas_nodes.argtypes = [_P, _P]
    lib.ref_version.restype = C.c_char_p
    return lib


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleScene:
    """Oracle-side scene built from the same GkSceneDesc / NodeProxy arrays the product gets."""

    def __init__(self, desc, nodes, count, use_ref=False):
        self.use_ref = use_ref
        self.lib = load_ref() if use_ref else load_oracle()
        create = self.lib.ref_scene_create if use_ref else self.lib.orc_scene_create
        self.h = create(C.cast(desc, _P), C.cast(nodes, _P), count)

    def close(self):
        if self.h:
            (self.lib.ref_scene_destroy if self.use_ref else self.lib.orc_scene_destroy)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_nodes(self, nodes, count):
        assert not self.use_ref
        self.lib.orc_scene_set_nodes(self.h, C.cast(nodes, _P), count)

    def intersect(self, rays: np.ndarray, threads=1, stats=False):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        tuv = np.empty((n, 3), np.float32)
        ids = np.empty((n, 2), np.uint32)
        st = np.zeros(2, np.uint64) if stats else None
        if self.use_ref:
            secs = self.lib.ref_intersect(self.h, ptr(rays), n, ptr(tuv), ptr(ids), threads)
        else:
            secs = self.lib.orc_intersect(self.h, ptr(rays), n, ptr(tuv), ptr(ids), threads, ptr(st))
        self.last_seconds = secs
        self.last_stats = st
        return tuv, ids

    def intersect_bruteforce(self, rays: np.ndarray, threads=None):
        """Exhaustive closest hit (no BVH): every triangle of every instance, the reference's per-triangle arithmetic."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        tuv = np.empty((n, 3), np.float32)
        ids = np.empty((n, 2), np.uint32)
        self.lib.orc_intersect_bruteforce_mt(self.h, ptr(rays), n, ptr(tuv), ptr(ids), threads or (os.cpu_count() or 1))
        return tuv, ids

    def raycast(self, origin_dir: np.ndarray):
        from gknextrenderer_b200._native import GkRayCastResult
        od = np.ascontiguousarray(origin_dir, np.float32)
        out = (GkRayCastResult * od.shape[0])()
        self.lib.orc_raycast(self.h, ptr(od), od.shape[0], out)
        return out

    def raycast_task(self, io: np.ndarray):
        """Task.RayCast.comp.slang on an (n, 24) 4-byte view of RayCastIO records, in place."""
        assert io.dtype.itemsize == 4 and io.shape[1] == 24 and io.flags["C_CONTIGUOUS"]
        self.lib.orc_raycast_task(self.h, io.ctypes.data_as(C.c_void_p), io.shape[0])
        return io

    def bake_probes(self, ubo, cubes: np.ndarray, voxels: np.ndarray, first: int, count: int, threads=8):
        """Probe baker restatement, in place on cubes (N, 14) / voxels (N, 4) uint32."""
        assert cubes.dtype == np.uint32 and voxels.dtype == np.uint32 and cubes.flags["C_CONTIGUOUS"] and voxels.flags["C_CONTIGUOUS"]
        self.lib.orc_bake_probes(self.h, C.cast(C.byref(ubo), _P), ptr(cubes), ptr(voxels), first, count, threads)

    def render(self, ubo, width, height, threads=8, cubes=None, voxels=None):
        px = width * height
        o = {
            "diffuse": np.zeros((height, width, 4), np.float32), "spec": np.zeros((height, width, 4), np.float32),
            "albedo": np.zeros((height, width, 4), np.float32), "normal": np.zeros((height, width, 4), np.float32),
            "motion": np.zeros((height, width, 2), np.float32), "depth": np.zeros((height, width), np.float32),
            "objectId": np.zeros((height, width), np.uint32), "primIds": np.zeros((height, width, 2), np.uint32),
            "rayCount": np.zeros((height, width), np.uint32),
        }
        assert px == o["depth"].size
        self.lib.orc_render(self.h, C.cast(C.byref(ubo), _P), width, height, ptr(cubes) if cubes is not None else None,
                            ptr(voxels) if voxels is not None else None, ptr(o["diffuse"]), ptr(o["spec"]), ptr(o["albedo"]),
                            ptr(o["normal"]), ptr(o["motion"]), ptr(o["depth"]), ptr(o["objectId"]), ptr(o["primIds"]), ptr(o["rayCount"]), threads)
        return o


def primary_rays(ubo, width, height, tmax=2000.0) -> np.ndarray:
    """Camera rays of Shading.slang:292-298 in float32 numpy, same association as the oracle.
    Only used to build *inputs* (ray buffers) that both sides then consume identically."""
    f = np.float32
    MVI = np.array(ubo.ModelViewInverse[:], f).reshape(4, 4).T  # [row, col]
    PI = np.array(ubo.ProjectionInverse[:], f).reshape(4, 4).T
    xs = (np.arange(width, dtype=f) / f(width)) * f(2) - f(1)
    ys = (np.arange(height, dtype=f) / f(height)) * f(2) - f(1)
    ux, uy = np.meshgrid(xs, ys)

    def mul(M, v):  # (c0*x + c1*y) + (c2*z + c3*w)
        return [(M[r, 0] * v[0] + M[r, 1] * v[1]) + (M[r, 2] * v[2] + M[r, 3] * v[3]) for r in range(4)]

    one = np.ones_like(ux)
    t = mul(PI, [ux, uy, one, one])
    l = np.sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2])
    rl = f(1) / l
    tn = [t[0] * rl, t[1] * rl, t[2] * rl]
    d = mul(MVI, [tn[0], tn[1], tn[2], np.zeros_like(ux)])
    l = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    rl = f(1) / l
    d = [d[0] * rl, d[1] * rl, d[2] * rl]
    o = mul(MVI, [f(0) * one, f(0) * one, f(0) * one, one])
    rays = np.empty((height, width, 8), f)
    rays[..., 0], rays[..., 1], rays[..., 2], rays[..., 3] = o[0], o[1], o[2], 0.0
    rays[..., 4], rays[..., 5], rays[..., 6], rays[..., 7] = d[0], d[1], d[2], tmax
    return rays.reshape(-1, 8)
