"""ctypes bindings for the two in-tree shared libraries.

* ``lib/libgknext_cuda.so`` — the C ABI of ``include/gknext_cuda.h`` (CUDA kernels, sm_100a).
* ``lib/libgknext_host.so`` — the C++ host mirror of the reference's scene/engine/renderer
  interface (``host/``), flattened to C for this harness.

There is no fallback: if the CUDA library is missing this module raises at import, and every
entry point that computes fails with GK_ERR_CUDA when no B200-class device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
CUDA_LIB_PATH = os.path.join(LIB_DIR, "libgknext_cuda.so")
HOST_LIB_PATH = os.path.join(LIB_DIR, "libgknext_host.so")
COMP_LIB_PATH = os.path.join(LIB_DIR, "libgknext_comp.so")
GKC_UNIQUE_ID_BYTES = 128

GK_EXCHANGE_IPC_BYTES = 8 * 64
GK_OK = 0
GK_ERR_INVALID_ARGUMENT = -1
GK_ERR_CUDA = -2
GK_ERR_OUT_OF_MEMORY = -3
GK_ERR_NOT_READY = -4
GK_ERR_UNSUPPORTED = -5

# GkPlane
PLANES = {
    "OUTPUT_DIFFUSE": 0, "OUTPUT_SPECULAR": 1, "ALBEDO": 2, "NORMAL": 3, "OBJECT_ID0": 4, "OBJECT_ID1": 5,
    "MOTION": 6, "DEPTH": 7, "ACCUM_DIFFUSE": 8, "ACCUM_SPECULAR": 9, "ACCUM_ALBEDO": 10,
    "HISTORY_DIFFUSE": 11, "HISTORY_SPECULAR": 12, "HISTORY_ALBEDO": 13, "DENOISED": 14,
    "RADIANCE_DIFFUSE_F32": 15, "RADIANCE_SPECULAR_F32": 16, "PRIMARY_IDS": 17, "PRIMARY_T": 18, "RAY_COUNT": 19,
}


class GkUniformBufferObject(C.Structure):
    _fields_ = (
        [(n, C.c_float * 16) for n in ("ModelView", "Projection", "ModelViewInverse", "ProjectionInverse", "ViewProjection",
                                       "PrevViewProjection", "ViewProjectionUnJit", "PrevViewProjectionUnJit")]
        + [(n, C.c_float * 4) for n in ("ViewportRect", "SunDirection", "SunColor", "BackGroundColor")]
        + [("SunViewProjection", C.c_float * 16)]
        + [("Aperture", C.c_float), ("FocusDistance", C.c_float), ("SkyRotation", C.c_float), ("HeatmapScale", C.c_float),
           ("PaperWhiteNit", C.c_float), ("SkyIntensity", C.c_float), ("SkyIdx", C.c_uint32), ("TotalFrames", C.c_uint32),
           ("MaxNumberOfBounces", C.c_uint32), ("NumberOfSamples", C.c_uint32), ("NumberOfBounces", C.c_uint32), ("RandomSeed", C.c_uint32),
           ("LightCount", C.c_uint32), ("HasSky", C.c_uint32), ("ShowHeatmap", C.c_uint32), ("UseCheckerBoard", C.c_uint32),
           ("TemporalFrames", C.c_uint32), ("HasSun", C.c_uint32), ("HDR", C.c_uint32), ("AdaptiveSample", C.c_uint32),
           ("AdaptiveVariance", C.c_float), ("AdaptiveSteps", C.c_uint32), ("TAA", C.c_uint32), ("SelectedId", C.c_uint32),
           ("ShowEdge", C.c_uint32), ("ProgressiveRender", C.c_uint32), ("BFSigma", C.c_float), ("BFSigmaLum", C.c_float),
           ("BFSigmaNormal", C.c_float), ("BFSize", C.c_uint32), ("FastGather", C.c_uint32), ("FastInterpole", C.c_uint32),
           ("DebugDraw_Lighting", C.c_uint32), ("DisableSpatialReuse", C.c_uint32), ("SuperResolution", C.c_uint32), ("_pad", C.c_uint32)]
    )


assert C.sizeof(GkUniformBufferObject) == 784


class GkNodeProxy(C.Structure):
    _fields_ = [("instanceId", C.c_uint32), ("modelId", C.c_uint32), ("visible", C.c_uint32), ("nort", C.c_uint32),
                ("worldTS", C.c_float * 16), ("combinedPrevTS", C.c_float * 16), ("matId", C.c_uint32 * 16)]


assert C.sizeof(GkNodeProxy) == 208


class GkVertex(C.Structure):
    _fields_ = [("Position", C.c_float * 3), ("Normal", C.c_float * 3), ("Tangent", C.c_float * 4), ("TexCoord", C.c_float * 2),
                ("MaterialIndex", C.c_uint32)]


assert C.sizeof(GkVertex) == 52


class GkMaterial(C.Structure):
    _fields_ = [("Diffuse", C.c_float * 4), ("DiffuseTextureId", C.c_int32), ("MRATextureId", C.c_int32), ("NormalTextureId", C.c_int32),
                ("Fuzziness", C.c_float), ("RefractionIndex", C.c_float), ("MaterialModel", C.c_uint32), ("Metalness", C.c_float),
                ("RefractionIndex2", C.c_float), ("NormalTextureScale", C.c_float), ("Reserverd2", C.c_float), ("_pad", C.c_uint32 * 2)]


assert C.sizeof(GkMaterial) == 64


class GkLightObject(C.Structure):
    _fields_ = [("p0", C.c_float * 4), ("p1", C.c_float * 4), ("p3", C.c_float * 4), ("normal_area", C.c_float * 4),
                ("lightMatIdx", C.c_uint32), ("reserved", C.c_uint32 * 3)]


assert C.sizeof(GkLightObject) == 80


class GkModelDesc(C.Structure):
    _fields_ = [("vertices", C.POINTER(GkVertex)), ("indices", C.POINTER(C.c_uint32)), ("vertexCount", C.c_uint32), ("indexCount", C.c_uint32)]


class GkSceneDesc(C.Structure):
    _fields_ = [("models", C.POINTER(GkModelDesc)), ("materials", C.POINTER(GkMaterial)), ("lights", C.POINTER(GkLightObject)),
                ("modelCount", C.c_uint32), ("materialCount", C.c_uint32), ("lightCount", C.c_uint32), ("reserved", C.c_uint32)]


class GkRayCastResult(C.Structure):
    _fields_ = [("HitPoint", C.c_float * 4), ("Normal", C.c_float * 4), ("T", C.c_float), ("InstanceId", C.c_uint32),
                ("MaterialId", C.c_uint32), ("Hitted", C.c_uint32)]


assert C.sizeof(GkRayCastResult) == 48


class GkRayCastIn(C.Structure):
    _fields_ = [("Origin", C.c_float * 4), ("Direction", C.c_float * 4), ("TMin", C.c_float), ("TMax", C.c_float), ("Reversed0", C.c_float), ("Reversed1", C.c_float)]


class GkRayCastIO(C.Structure):
    _fields_ = [("Context", GkRayCastIn), ("Result", GkRayCastResult)]


assert C.sizeof(GkRayCastIO) == 96


class GkConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32), ("tileIndex", C.c_uint32), ("tileCount", C.c_uint32),
                ("tileRows", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32 * 6)]


class GkFrameStats(C.Structure):
    _fields_ = [("primaryRays", C.c_uint64), ("extensionRays", C.c_uint64), ("shadowRays", C.c_uint64), ("waves", C.c_uint32),
                ("launches", C.c_uint32), ("msTotal", C.c_float), ("msBvh", C.c_float), ("msGenerate", C.c_float), ("msExtend", C.c_float),
                ("msShade", C.c_float), ("msShadow", C.c_float), ("msAccumulate", C.c_float), ("msReproject", C.c_float),
                ("msDenoise", C.c_float), ("nodeVisits", C.c_uint64), ("triTests", C.c_uint64),
                ("tlasVisits", C.c_uint64), ("instanceEntries", C.c_uint64), ("msTail", C.c_float), ("tailPaths", C.c_uint32),
                ("tailExtensionRays", C.c_uint64), ("tailShadowRays", C.c_uint64), ("maxStack", C.c_uint32), ("msTrace", C.c_float),
                ("schedIters", C.c_uint64 * 3), ("schedLanes", C.c_uint64 * 3), ("schedRefills", C.c_uint64), ("schedRefillLanes", C.c_uint64),
                ("schedPopIters", C.c_uint64), ("schedPopLanes", C.c_uint64)]


class GkBvhInfo(C.Structure):
    _fields_ = [("blasCount", C.c_uint32), ("instanceCount", C.c_uint32), ("triangleCount", C.c_uint64), ("instancedTriangles", C.c_uint64),
                ("blasNodes2", C.c_uint32), ("blasNodes8", C.c_uint32), ("tlasNodes2", C.c_uint32), ("tlasNodes8", C.c_uint32),
                ("bytesGeometry", C.c_uint64), ("bytesBvh", C.c_uint64), ("msBlasBuild", C.c_float), ("msTlasBuild", C.c_float),
                ("msRefit", C.c_float), ("refitsRejected", C.c_uint32), ("tlasAreaAtBuild", C.c_float)]


# every symbol include/gknext_cuda.h declares: name -> (restype, argtypes)
_P = C.c_void_p
CUDA_API = {
    "gk_abi_version": (C.c_int, []),
    "gk_last_error": (C.c_char_p, []),
    "gk_create": (C.c_int, [C.POINTER(GkConfig), C.POINTER(_P)]),
    "gk_destroy": (None, [_P]),
    "gk_resize": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "gk_upload_scene": (C.c_int, [_P, C.POINTER(GkSceneDesc)]),
    "gk_update_materials": (C.c_int, [_P, C.POINTER(GkMaterial), C.c_uint32]),
    "gk_update_instances": (C.c_int, [_P, C.POINTER(GkNodeProxy), C.c_uint32, C.c_int]),
    "gk_update_instances_sparse": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_int]),
    "gk_set_probes": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "gk_bake_probes": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "gk_get_probes": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "gk_set_ubo": (C.c_int, [_P, C.POINTER(GkUniformBufferObject)]),
    "gk_render_frame": (C.c_int, [_P]),
    "gk_trace_frame": (C.c_int, [_P]),
    "gk_filter_frame": (C.c_int, [_P]),
    "gk_raycast": (C.c_int, [_P, _P, C.c_uint32, _P]),
    "gk_raycast_task": (C.c_int, [_P, _P, C.c_uint32]),
    "gk_intersect": (C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    "gk_intersect_device": (C.c_int, [_P, _P, C.c_uint32, _P, _P, C.c_int]),
    "gk_plane_bytes": (C.c_size_t, [_P, C.c_int]),
    "gk_readback": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "gk_readback_async": (C.c_int, [_P, C.c_int, C.c_void_p, C.c_size_t]),
    "gk_readback_wait": (C.c_int, [_P]),
    "gk_upload_plane": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "gk_plane_device": (_P, [_P, C.c_int]),
    "gk_exchange_bytes": (C.c_size_t, [_P]),
    "gk_exchange_pack": (C.c_int, [_P, _P]),
    "gk_exchange_unpack": (C.c_int, [_P, _P]),
    "gk_exchange_ipc_handles": (C.c_int, [_P, C.c_void_p, C.c_size_t]),
    "gk_exchange_open_peers": (C.c_int, [_P, C.c_void_p, C.c_uint32]),
    "gk_exchange_push": (C.c_int, [_P]),
    "gk_exchange_close_peers": (C.c_int, [_P]),
    "gk_frame_shard_handle": (C.c_int, [_P, C.c_void_p, C.c_size_t]),
    "gk_frame_shard_open": (C.c_int, [_P, C.c_void_p, C.c_uint32]),
    "gk_frame_shard_push": (C.c_int, [_P]),
    "gk_frame_shard_accumulate": (C.c_int, [_P]),
    "gk_filter_frame_owned": (C.c_int, [_P]),
    "gk_exchange_push_final": (C.c_int, [_P, C.c_int]),
    "gk_host_alloc": (C.c_void_p, [C.c_size_t]),
    "gk_host_free": (None, [C.c_void_p]),
    "gk_synchronize": (C.c_int, [_P]),
    "gk_get_stats": (C.c_int, [_P, C.POINTER(GkFrameStats)]),
    "gk_get_bvh_info": (C.c_int, [_P, C.POINTER(GkBvhInfo)]),
    "gk_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "gk_measure_read_bandwidth": (C.c_int, [_P, C.c_size_t, C.c_int, C.POINTER(C.c_float)]),
    "gk_set_traversal_stats": (C.c_int, [_P, C.c_int]),
    "gk_stream": (_P, [_P]),
    "gk_set_ray_capture": (C.c_int, [_P, C.c_int]),
    "gk_get_captured_rays": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
}

HOST_API = {
    "gkh_last_error": (C.c_char_p, []),
    "gkh_engine_create": (_P, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "gkh_engine_destroy": (None, [_P]),
    "gkh_scene_desc": (C.POINTER(GkSceneDesc), [_P]),
    "gkh_scene_triangles": (C.c_uint64, [_P, C.c_int]),
    "gkh_update_nodes": (C.c_uint32, [_P]),
    "gkh_node_proxies": (C.POINTER(GkNodeProxy), [_P]),
    "gkh_mark_dirty": (None, [_P]),
    "gkh_scene_step": (None, [_P, C.c_uint32]),
    "gkh_changed_proxies": (C.c_int64, [_P, C.POINTER(C.POINTER(C.c_uint32))]),
    "gkh_set_node_translation": (C.c_int, [_P, C.c_uint32, C.c_float, C.c_float, C.c_float]),
    "gkh_set_setting": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "gkh_set_camera_lookat": (C.c_int, [_P, _P, _P, _P, C.c_float]),
    "gkh_get_ubo": (None, [_P, C.c_uint32, C.c_uint32, C.POINTER(GkUniformBufferObject)]),
    "gkh_advance_frame": (None, [_P]),
    "gkh_screen_ray": (None, [_P, C.c_float, C.c_float, C.c_uint32, C.c_uint32, _P, _P]),
    "gkh_renderer_create": (_P, [_P, C.c_int]),
    "gkh_renderer_destroy": (None, [_P]),
    "gkh_renderer_set_trace_all_rows": (C.c_int, [_P, C.c_int]),
    "gkh_renderer_set_tile": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32]),
    "gkh_renderer_create_swapchain": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "gkh_renderer_delete_swapchain": (C.c_int, [_P]),
    "gkh_renderer_post_load_scene": (C.c_int, [_P]),
    "gkh_renderer_before_next_frame": (C.c_int, [_P]),
    "gkh_renderer_instance_bytes_uploaded": (C.c_uint64, [_P]),
    "gkh_renderer_render": (C.c_int, [_P]),
    "gkh_renderer_context": (_P, [_P]),
}


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


def load_cuda():
    if not os.path.exists(CUDA_LIB_PATH):
        raise RuntimeError(
            f"{CUDA_LIB_PATH} is missing: build it with gknextrenderer_b200/build.sh (or __graft_entry__.build()). "
            "There is no CPU fallback for the path-tracing hot path.")
    return _bind(C.CDLL(CUDA_LIB_PATH, mode=C.RTLD_GLOBAL), CUDA_API)


# every symbol include/gknext_compositor.h declares
COMP_API = {
    "gkc_last_error": (C.c_char_p, []),
    "gkc_get_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "gkc_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(_P)]),
    "gkc_enable_frame_sharding": (C.c_int, [_P]),
    "gkc_barrier": (C.c_int, [_P]),
    "gkc_composite_frame": (C.c_int, [_P]),
    "gkc_composite_final": (C.c_int, [_P, C.c_int]),
    "gkc_composite_frame_shard": (C.c_int, [_P, C.c_int]),
    "gkc_world": (C.c_int, [_P]),
    "gkc_rank": (C.c_int, [_P]),
    "gkc_destroy": (None, [_P]),
}


def load_comp():
    load_cuda()
    if not os.path.exists(COMP_LIB_PATH):
        raise RuntimeError(f"{COMP_LIB_PATH} is missing: build it with gknextrenderer_b200/build.sh")
    return _bind(C.CDLL(COMP_LIB_PATH), COMP_API)


def load_host():
    load_cuda()
    if not os.path.exists(HOST_LIB_PATH):
        raise RuntimeError(f"{HOST_LIB_PATH} is missing: build it with gknextrenderer_b200/build.sh")
    return _bind(C.CDLL(HOST_LIB_PATH), HOST_API)
