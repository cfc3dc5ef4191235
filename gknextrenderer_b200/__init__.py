"""gknextrenderer_b200 — B200 (sm_100a) CUDA backend for gkNextRenderer's path-tracing hot path.

Layers:
  csrc/   hand-written CUDA kernels + the C ABI of include/gknext_cuda.h  -> lib/libgknext_cuda.so
  host/   C++ mirror of the reference's Assets::* / engine / LogicRendererBase interface,
          driving the C ABI the way the reference's host code would      -> lib/libgknext_host.so
  this package: thin ctypes wrappers used by tests/, bench.py and __graft_entry__.py.

Nothing here computes on the CPU: every render / intersect call goes through the C ABI into
CUDA kernels and fails loudly when the extension or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from ._native import GkUniformBufferObject, GkNodeProxy, GkSceneDesc, GkFrameStats, GkBvhInfo, GkConfig, GkRayCastResult, PLANES  # noqa: F401

__all__ = ["Engine", "Renderer", "GkError", "PLANES", "plane_dtype"]


class GkError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"[gk status {status}] {message}")
        self.status = status


_cuda = None
_host = None


def cuda_lib():
    global _cuda
    if _cuda is None:
        _cuda = N.load_cuda()
    return _cuda


def host_lib():
    global _host
    if _host is None:
        _host = N.load_host()
    return _host


_comp = None


def comp_lib():
    """lib/libgknext_comp.so: the multi-GPU compositor over NCCL (include/gknext_compositor.h)."""
    global _comp
    if _comp is None:
        _comp = N.load_comp()
    return _comp


def plane_dtype(name: str):
    """(numpy dtype, channels) of a plane as gk_readback returns it."""
    if name in ("OBJECT_ID0", "OBJECT_ID1", "RAY_COUNT"):
        return np.uint32, 1
    if name in ("DEPTH", "PRIMARY_T"):
        return np.float32, 1
    if name == "MOTION":
        return np.float32, 2
    if name == "PRIMARY_IDS":
        return np.uint32, 2
    if name in ("RADIANCE_DIFFUSE_F32", "RADIANCE_SPECULAR_F32"):
        return np.float32, 4
    return np.float16, 4


class Engine:
    """Host mirror: scene (Assets::Scene), user settings and per-frame UBO (NextEngine)."""

    def __init__(self, scene: str = "cornell", p0: int = 0, p1: int = 0, p2: int = 0, p3: int = 0):
        self.lib = host_lib()
        self.h = self.lib.gkh_engine_create(scene.encode(), p0, p1, p2, p3)
        if not self.h:
            raise GkError(-1, self.lib.gkh_last_error().decode())
        self.scene_name = scene

    def close(self):
        if self.h:
            self.lib.gkh_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, **settings):
        for k, v in settings.items():
            if self.lib.gkh_set_setting(self.h, k.encode(), float(v)) != 0:
                raise GkError(-1, self.lib.gkh_last_error().decode())
        return self

    def look_at(self, eye, center, up=(0, 1, 0), fov=40.0):
        a = (C.c_float * 3)(*eye)
        b = (C.c_float * 3)(*center)
        c = (C.c_float * 3)(*up)
        self.lib.gkh_set_camera_lookat(self.h, a, b, c, float(fov))

    def scene_desc(self):
        return self.lib.gkh_scene_desc(self.h)

    def triangles(self, instanced=False) -> int:
        return int(self.lib.gkh_scene_triangles(self.h, 1 if instanced else 0))

    def update_nodes(self):
        """Scene::UpdateNodes — returns (pointer to NodeProxy[], count)."""
        n = self.lib.gkh_update_nodes(self.h)
        return self.lib.gkh_node_proxies(self.h), int(n)

    def changed_proxies(self):
        """Indices of the proxies the last update_nodes() rewrote, or None when it was a full pass."""
        ptr = C.POINTER(C.c_uint32)()
        n = self.lib.gkh_changed_proxies(self.h, C.byref(ptr))
        if n < 0:
            return None
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, np.uint32)

    def mark_dirty(self):
        self.lib.gkh_mark_dirty(self.h)

    def step_scene(self, frame: int):
        self.lib.gkh_scene_step(self.h, frame)

    def set_node_translation(self, node, x, y, z):
        self.lib.gkh_set_node_translation(self.h, node, x, y, z)

    def ubo(self, width, height) -> GkUniformBufferObject:
        u = GkUniformBufferObject()
        self.lib.gkh_get_ubo(self.h, width, height, C.byref(u))
        return u

    def advance_frame(self):
        self.lib.gkh_advance_frame(self.h)

    def screen_ray(self, x, y, width, height):
        o = (C.c_float * 3)()
        d = (C.c_float * 3)()
        self.lib.gkh_screen_ray(self.h, x, y, width, height, o, d)
        return np.array(o[:], np.float32), np.array(d[:], np.float32)


class Renderer:
    """One GkContext: the CUDA path tracer behind the C ABI."""

    def __init__(self, width, height, device=-1, tile_index=0, tile_count=1, tile_rows=16, trace_all_rows=False):
        self.lib = cuda_lib()
        cfg = GkConfig()
        cfg.device, cfg.width, cfg.height = device, width, height
        cfg.tileIndex, cfg.tileCount, cfg.tileRows = tile_index, tile_count, tile_rows
        cfg.flags = 1 if trace_all_rows else 0  # GK_CFG_TRACE_ALL_ROWS
        h = C.c_void_p()
        self._check(self.lib.gk_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.width, self.height = width, height

    def _check(self, status):
        if status != N.GK_OK:
            raise GkError(status, self.lib.gk_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.gk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- scene ---
    def upload_scene(self, desc):
        self._check(self.lib.gk_upload_scene(self.h, desc))

    def update_instances(self, nodes, count, refit=False):
        self._check(self.lib.gk_update_instances(self.h, nodes, count, 1 if refit else 0))

    def update_instances_sparse(self, indices, proxies, refit=True):
        """gk_update_instances_sparse: `proxies[k]` replaces record `indices[k]` of the array already on the device."""
        indices = np.ascontiguousarray(indices, np.uint32)
        assert len(proxies) == indices.size
        self._check(self.lib.gk_update_instances_sparse(self.h, indices.ctypes.data_as(C.c_void_p), C.cast(proxies, C.c_void_p), indices.size, 1 if refit else 0))

    def load(self, engine: Engine, refit=False):
        self.upload_scene(engine.scene_desc())
        nodes, n = engine.update_nodes()
        self.update_instances(nodes, n, refit)

    def set_probes(self, cubes: np.ndarray, voxels: np.ndarray):
        """Ambient-cube grid (192 x 48 x 192 probes): cubes (N, 14) uint32, voxels (N, 4) uint32."""
        cubes = np.ascontiguousarray(cubes, np.uint32)
        voxels = np.ascontiguousarray(voxels, np.uint32)
        assert cubes.shape[1] == 14 and voxels.shape[1] == 4 and cubes.shape[0] == voxels.shape[0]
        self._check(self.lib.gk_set_probes(self.h, cubes.ctypes.data_as(C.c_void_p), voxels.ctypes.data_as(C.c_void_p), cubes.shape[0]))

    def bake_probes(self, first: int, count: int):
        """Probe baker (Bake.HwAmbientCube / FGpuProbeGenerator::Render) on probes [first, first + count)."""
        self._check(self.lib.gk_bake_probes(self.h, first, count))

    def get_probes(self):
        """(cubes (N, 14) uint32, voxels (N, 4) uint32) of the 192 x 48 x 192 grid."""
        n = 192 * 192 * 48
        cubes, voxels = np.empty((n, 14), np.uint32), np.empty((n, 4), np.uint32)
        self._check(self.lib.gk_get_probes(self.h, cubes.ctypes.data_as(C.c_void_p), voxels.ctypes.data_as(C.c_void_p), n))
        return cubes, voxels

    def set_ubo(self, ubo):
        self._check(self.lib.gk_set_ubo(self.h, C.byref(ubo)))

    # --- frames ---
    def render_frame(self):
        self._check(self.lib.gk_render_frame(self.h))

    def trace_frame(self):
        self._check(self.lib.gk_trace_frame(self.h))

    def filter_frame(self):
        self._check(self.lib.gk_filter_frame(self.h))

    def synchronize(self):
        self._check(self.lib.gk_synchronize(self.h))

    def stats(self) -> GkFrameStats:
        s = GkFrameStats()
        self._check(self.lib.gk_get_stats(self.h, C.byref(s)))
        return s

    def bvh_info(self) -> GkBvhInfo:
        s = GkBvhInfo()
        self._check(self.lib.gk_get_bvh_info(self.h, C.byref(s)))
        return s

    def set_option(self, name: str, value: float):
        """Tuning hook (gk_set_option): e.g. trace_variant, sched_refill_min, coop_threshold."""
        self._check(self.lib.gk_set_option(self.h, name.encode(), float(value)))

    def measure_read_bandwidth(self, nbytes: int, reps: int = 8) -> float:
        """GB/s of the library's streaming-read probe over a scratch buffer of nbytes (L2-resident when small)."""
        out = C.c_float()
        self._check(self.lib.gk_measure_read_bandwidth(self.h, nbytes, reps, C.byref(out)))
        return float(out.value)

    def set_traversal_stats(self, on: bool):
        self._check(self.lib.gk_set_traversal_stats(self.h, 1 if on else 0))

    # --- planes ---
    def plane_bytes(self, name):
        return int(self.lib.gk_plane_bytes(self.h, PLANES[name]))

    def readback(self, name, out: np.ndarray | None = None) -> np.ndarray:
        dt, ch = plane_dtype(name)
        shape = (self.height, self.width, ch) if ch > 1 else (self.height, self.width)
        if out is None:
            out = np.empty(shape, dt)
        assert out.nbytes == self.plane_bytes(name), (out.nbytes, self.plane_bytes(name))
        self._check(self.lib.gk_readback(self.h, PLANES[name], out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def readback_async(self, name, dst_ptr: int, nbytes: int):
        """Queue a device->host copy of a plane behind the submitted work (dst should be pinned memory)."""
        self._check(self.lib.gk_readback_async(self.h, PLANES[name], C.c_void_p(dst_ptr), nbytes))

    def readback_wait(self):
        self._check(self.lib.gk_readback_wait(self.h))

    def upload_plane(self, name, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes == self.plane_bytes(name), (arr.nbytes, self.plane_bytes(name))
        self._check(self.lib.gk_upload_plane(self.h, PLANES[name], arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def plane_device_ptr(self, name) -> int:
        return int(self.lib.gk_plane_device(self.h, PLANES[name]) or 0)

    def exchange_bytes(self) -> int:
        return int(self.lib.gk_exchange_bytes(self.h))

    def exchange_pack(self, d_staging: int):
        self._check(self.lib.gk_exchange_pack(self.h, C.c_void_p(d_staging)))

    def exchange_unpack(self, d_all: int):
        self._check(self.lib.gk_exchange_unpack(self.h, C.c_void_p(d_all)))

    def exchange_ipc_handles(self) -> bytes:
        buf = C.create_string_buffer(N.GK_EXCHANGE_IPC_BYTES)
        self._check(self.lib.gk_exchange_ipc_handles(self.h, buf, N.GK_EXCHANGE_IPC_BYTES))
        return buf.raw

    def exchange_open_peers(self, handles_all: bytes, world: int):
        self._check(self.lib.gk_exchange_open_peers(self.h, handles_all, world))

    def filter_frame_owned(self):
        self._check(self.lib.gk_filter_frame_owned(self.h))

    def exchange_push_final(self, dst_rank: int = -1):
        self._check(self.lib.gk_exchange_push_final(self.h, dst_rank))

    def frame_shard_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.lib.gk_frame_shard_handle(self.h, buf, 64))
        return buf.raw

    def frame_shard_open(self, handles_all: bytes, world: int):
        self._check(self.lib.gk_frame_shard_open(self.h, handles_all, world))

    def frame_shard_push(self):
        self._check(self.lib.gk_frame_shard_push(self.h))

    def frame_shard_accumulate(self):
        self._check(self.lib.gk_frame_shard_accumulate(self.h))

    def exchange_push(self):
        self._check(self.lib.gk_exchange_push(self.h))

    def exchange_close_peers(self):
        self._check(self.lib.gk_exchange_close_peers(self.h))

    def stream(self) -> int:
        return int(self.lib.gk_stream(self.h) or 0)

    # --- queries ---
    def intersect(self, rays: np.ndarray):
        """rays: (n, 8) float32 {O.xyz, tmin, D.xyz, tmax} -> (tuv (n,3) f32, ids (n,2) u32)."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        tuv = np.empty((n, 3), np.float32)
        ids = np.empty((n, 2), np.uint32)
        self._check(self.lib.gk_intersect(self.h, rays.ctypes.data_as(C.c_void_p), n, tuv.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p)))
        return tuv, ids

    def intersect_device(self, d_rays: int, n: int, d_tuv: int, d_ids: int, any_hit=False):
        self._check(self.lib.gk_intersect_device(self.h, C.c_void_p(d_rays), n, C.c_void_p(d_tuv), C.c_void_p(d_ids), 1 if any_hit else 0))

    def raycast(self, origin_dir: np.ndarray):
        od = np.ascontiguousarray(origin_dir, np.float32)
        n = od.shape[0]
        out = (GkRayCastResult * n)()
        self._check(self.lib.gk_raycast(self.h, od.ctypes.data_as(C.c_void_p), n, out))
        return out

    def raycast_task(self, io: np.ndarray):
        """GPU ray-cast task (Task.RayCast.comp.slang) on an (n, 24) float32/uint32 view of RayCastIO records, in place."""
        assert io.dtype.itemsize == 4 and io.shape[1] == 24 and io.flags["C_CONTIGUOUS"]
        self._check(self.lib.gk_raycast_task(self.h, io.ctypes.data_as(C.c_void_p), io.shape[0]))
        return io

    def set_ray_capture(self, wave: int):
        self._check(self.lib.gk_set_ray_capture(self.h, wave))

    def captured_rays(self, capacity: int) -> np.ndarray:
        buf = np.empty((capacity, 8), np.float32)
        cnt = C.c_uint32()
        self._check(self.lib.gk_get_captured_rays(self.h, buf.ctypes.data_as(C.c_void_p), capacity, C.byref(cnt)))
        return buf[: min(capacity, cnt.value)]
