"""Multi-GPU frame compositor: one process per GPU, scene and BVH replicated,
the image partitioned by interleaved row tiles, ONE exchange step at frame end.

The reference has no multi-GPU path; this is the only collective the hot path needs.  Tile
mode: rank r traces the rows with (row // tile_rows) % world == r (GkConfig.tileIndex/Count),
then every rank all-gathers the integrator's output planes so that each holds the full
G-buffer and can run the spatial/temporal filters with their halos locally.  Because a pixel's
random sequence depends only on (x, y, frame), the union of the tiles is bit-identical to the
single-GPU frame.

torch is used only as plumbing here: `torch.distributed` (NCCL over NVLink on the GPU box,
gloo in the CPU tests) moving bytes between the planes the CUDA library owns.
"""
from __future__ import annotations

import numpy as np

# planes the path tracer writes and the filters read (name, bytes per pixel)
EXCHANGE_PLANES = [
    ("OUTPUT_DIFFUSE", 8), ("OUTPUT_SPECULAR", 8), ("ALBEDO", 8), ("NORMAL", 8), ("OBJECT_ID0", 4), ("MOTION", 8),
]
PARITY_PLANES = [("RADIANCE_DIFFUSE_F32", 16), ("RADIANCE_SPECULAR_F32", 16), ("PRIMARY_IDS", 8), ("PRIMARY_T", 4), ("RAY_COUNT", 4), ("DEPTH", 4)]


def owned_rows(height: int, tile_rows: int, rank: int, world: int) -> np.ndarray:
    rows = np.arange(height)
    return rows[(rows // tile_rows) % world == rank]


def row_blocks(height: int, tile_rows: int, rank: int, world: int):
    """[(first_row, row_count)] of the contiguous row blocks rank owns, in image order."""
    out = []
    b = rank
    while b * tile_rows < height:
        r0 = b * tile_rows
        out.append((r0, min(tile_rows, height - r0)))
        b += world
    return out


def padded_blocks(height: int, tile_rows: int, world: int) -> int:
    """Row blocks per rank after padding so every rank contributes the same byte count."""
    total = (height + tile_rows - 1) // tile_rows
    return (total + world - 1) // world


def pack_owned(plane2d, height, tile_rows, rank, world):
    """Gather this rank's row blocks of a (H, row_bytes) uint8 plane view into a dense
    (padded_blocks*tile_rows, row_bytes) send buffer (torch tensor in, torch tensor out)."""
    import torch
    nb = padded_blocks(height, tile_rows, world)
    send = torch.zeros((nb * tile_rows, plane2d.shape[1]), dtype=plane2d.dtype, device=plane2d.device)
    for i, (r0, cnt) in enumerate(row_blocks(height, tile_rows, rank, world)):
        send[i * tile_rows: i * tile_rows + cnt] = plane2d[r0: r0 + cnt]
    return send


def unpack_all(gathered, plane2d, height, tile_rows, world):
    """Scatter the all-gathered (world, padded_blocks*tile_rows, row_bytes) buffer back into the
    full plane."""
    for rank in range(world):
        for i, (r0, cnt) in enumerate(row_blocks(height, tile_rows, rank, world)):
            plane2d[r0: r0 + cnt] = gathered[rank, i * tile_rows: i * tile_rows + cnt]


def all_gather_plane(plane2d, height, tile_rows, rank, world, group=None):
    """In-place: after the call every rank's `plane2d` holds all rows."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return
    send = pack_owned(plane2d, height, tile_rows, rank, world)
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(recv, send, group=group)
    else:  # gloo (CPU tests)
        dist.all_gather([recv[i] for i in range(world)], send, group=group)
    unpack_all(recv, plane2d, height, tile_rows, world)


def device_plane_tensor(renderer, name: str, bytes_per_pixel: int):
    """Zero-copy torch view (H, W*bytes_per_pixel) uint8 of a plane owned by the CUDA library."""
    import torch
    ptr = renderer.plane_device_ptr(name)
    nbytes = renderer.height * renderer.width * bytes_per_pixel

    class _Wrap:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}

    t = torch.as_tensor(_Wrap(), device=f"cuda:{torch.cuda.current_device()}")
    return t.view(renderer.height, renderer.width * bytes_per_pixel)


_staging = {}
_p2p = {}


_nativeComp = {}  # context handle -> GkCompositor* (lib/libgknext_comp.so)


def _hkey(renderer):
    return renderer.h.value if hasattr(renderer.h, "value") else int(renderer.h)


def _gkc(status, lib):
    if status != 0:
        raise RuntimeError(f"compositor: {lib.gkc_last_error().decode()} (status {status})")


def enable_native(renderer, rank: int, world: int, group=None) -> bool:
    """The compositor of include/gknext_compositor.h: its own NCCL communicator on the library stream, created from a unique id
    that rank 0 draws and torch.distributed merely carries to the other ranks.  After this, composite_frame / composite_final /
    composite_frame_shard are ONE C call each (barrier, peer-to-peer push, barrier - all enqueued from C++)."""
    import ctypes as C
    import torch.distributed as dist
    from . import _native as N, comp_lib
    lib = comp_lib()
    ident = (C.c_ubyte * N.GKC_UNIQUE_ID_BYTES)()
    if rank == 0:
        _gkc(lib.gkc_get_unique_id(ident, N.GKC_UNIQUE_ID_BYTES), lib)
    box = [bytes(ident)]
    dist.broadcast_object_list(box, src=0, group=group)
    ident = (C.c_ubyte * N.GKC_UNIQUE_ID_BYTES).from_buffer_copy(box[0])
    comp = C.c_void_p()
    status = lib.gkc_create(renderer.h, rank, world, ident, N.GKC_UNIQUE_ID_BYTES, C.byref(comp))
    if status != 0:
        print(f"[compositor] rank {rank}: native compositor unavailable ({lib.gkc_last_error().decode()})", flush=True)
        return False
    _nativeComp[_hkey(renderer)] = comp
    return True


def enable_peer_exchange(renderer, rank: int, world: int, group=None) -> bool:
    """Map the exchange planes of all ranks into this process (CUDA IPC) so that composite_frame can
    push rows straight into the peers.  Returns False (and keeps the NCCL all-gather path) when the
    handles cannot be opened, e.g. no peer access between the devices.
    GK_COMPOSITOR=native (default) drives the exchange from lib/libgknext_comp.so (C++ over NCCL); GK_COMPOSITOR=torch keeps the
    choreography in this file (torch.distributed barriers around the same push kernels)."""
    import os
    import torch
    import torch.distributed as dist
    if world == 1:
        return False
    from . import _native as N
    if os.environ.get("GK_COMPOSITOR", "native") == "native" and dist.get_backend(group) == "nccl":
        # every rank must take the same path: the native driver is used only if it came up on all of them
        try:
            ok = enable_native(renderer, rank, world, group)
        except (OSError, RuntimeError, AttributeError) as e:  # library missing / not loadable on this box
            print(f"[compositor] rank {rank}: lib/libgknext_comp.so unavailable ({e}); torch.distributed drives the exchange", flush=True)
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=torch.device(f"cuda:{torch.cuda.current_device()}"))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()):
            return True
        if ok:  # came up here but not everywhere
            from . import comp_lib
            comp_lib().gkc_destroy(_nativeComp.pop(_hkey(renderer)))
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    mine = torch.frombuffer(bytearray(renderer.exchange_ipc_handles()), dtype=torch.uint8).to(dev)
    every = torch.empty(world * N.GK_EXCHANGE_IPC_BYTES, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(every, mine, group=group)
    ok = True
    try:
        renderer.exchange_open_peers(bytes(every.cpu().numpy().tobytes()), world)
    except Exception as e:  # noqa: BLE001
        print(f"[compositor] rank {rank}: peer mapping unavailable ({e}); using the NCCL all-gather exchange", flush=True)
        ok = False
    # all ranks must take the same path
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    ok = bool(int(flag.item()))
    key = renderer.h.value if hasattr(renderer.h, "value") else int(renderer.h)
    if ok:
        _p2p[key] = (torch.zeros(1, dtype=torch.int32, device=dev), torch.cuda.ExternalStream(renderer.stream(), device=dev))
    else:
        _p2p.pop(key, None)
    return ok


def composite_frame(renderer, rank: int, world: int, tile_rows: int, planes=None, group=None):
    """Frame-end exchange for tile mode.  The CUDA library packs the rows this rank owns of the six
    integrator planes into one staging buffer (one kernel), NCCL all-gathers the staging buffers
    over NVLink, and the library scatters every rank's rows into the full planes (one kernel).
    All three steps are ordered on the library's stream."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return 0
    hkey = renderer.h.value if hasattr(renderer.h, "value") else int(renderer.h)
    if hkey in _nativeComp:
        from . import comp_lib
        _gkc(comp_lib().gkc_composite_frame(_nativeComp[hkey]), comp_lib())
        return renderer.exchange_bytes()
    if hkey in _p2p:
        # peer-to-peer: barrier (every rank is done reading last frame's planes) -> one kernel stores the
        # owned rows into all peers over NVLink -> barrier (all rows have landed).  The barriers are
        # 4-byte all-reduces ordered on the library's stream.
        token, stream = _p2p[hkey]
        renderer.readback_wait()  # peers are about to overwrite planes an asynchronous read-back may still be reading
        with torch.cuda.stream(stream):
            dist.all_reduce(token, group=group)
        renderer.exchange_push()
        with torch.cuda.stream(stream):
            dist.all_reduce(token, group=group)
        return renderer.exchange_bytes()
    nbytes = renderer.exchange_bytes()
    key = (renderer.h.value if hasattr(renderer.h, "value") else int(renderer.h), nbytes, world)
    if key not in _staging:
        dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        _staging.clear()
        _staging[key] = (torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(world * nbytes, dtype=torch.uint8, device=dev),
                         torch.cuda.ExternalStream(renderer.stream(), device=dev))
    send, recv, stream = _staging[key]
    renderer.exchange_pack(send.data_ptr())
    with torch.cuda.stream(stream):
        dist.all_gather_into_tensor(recv, send, group=group)
    renderer.exchange_unpack(recv.data_ptr())
    return nbytes


def composite_final(renderer, rank: int, world: int, dst_rank: int = -1, group=None):
    """Progressive frames without the denoiser are filtered per pixel on the rank that traced them
    (Renderer.filter_frame_owned), so only the finished image travels: the owned rows of rtDenoised
    (8 B/pixel instead of 44) are stored into rank `dst_rank` (-1: all ranks) between two stream barriers.
    Needs enable_peer_exchange()."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return 0
    hkey = renderer.h.value if hasattr(renderer.h, "value") else int(renderer.h)
    if hkey in _nativeComp:
        from . import comp_lib
        _gkc(comp_lib().gkc_composite_final(_nativeComp[hkey], dst_rank), comp_lib())
        return renderer.plane_bytes("DENOISED") // world
    if hkey not in _p2p:
        raise RuntimeError("composite_final needs enable_peer_exchange() to have succeeded")
    token, stream = _p2p[hkey]
    renderer.readback_wait()
    with torch.cuda.stream(stream):
        dist.all_reduce(token, group=group)
    renderer.exchange_push_final(dst_rank)
    with torch.cuda.stream(stream):
        dist.all_reduce(token, group=group)
    return renderer.plane_bytes("DENOISED") // world


_shard = set()


def enable_frame_sharding(renderer, rank: int, world: int, group=None) -> bool:
    """Frame-sharded progressive rendering: map the gather buffers of all ranks (CUDA IPC).  The renderer
    must have been created with trace_all_rows=True; enable_peer_exchange() must have succeeded."""
    import torch
    import torch.distributed as dist
    hkey = renderer.h.value if hasattr(renderer.h, "value") else int(renderer.h)
    if world > 1 and hkey in _nativeComp:
        from . import comp_lib
        _gkc(comp_lib().gkc_enable_frame_sharding(_nativeComp[hkey]), comp_lib())
        _shard.add(hkey)
        return True
    if world == 1 or hkey not in _p2p:
        return False
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    mine = torch.frombuffer(bytearray(renderer.frame_shard_handle()), dtype=torch.uint8).to(dev)
    every = torch.empty(world * 64, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(every, mine, group=group)
    renderer.frame_shard_open(bytes(every.cpu().numpy().tobytes()), world)
    dist.barrier(group=group)
    _shard.add(hkey)
    return True


def composite_frame_shard(renderer, rank: int, world: int, dst_rank: int = 0, group=None):
    """Super-step of frame-sharded progressive rendering, after every rank traced ITS frame (gk_trace_frame with
    TotalFrames = f0 + rank): rows go to their owners, the owners accumulate the `world` frames in order and
    compose, the finished rows go to the presenting rank.  Three stream barriers per super-step."""
    import torch
    import torch.distributed as dist
    hkey = renderer.h.value if hasattr(renderer.h, "value") else int(renderer.h)
    if hkey not in _shard:
        raise RuntimeError("composite_frame_shard needs enable_frame_sharding()")
    if hkey in _nativeComp:
        from . import comp_lib
        _gkc(comp_lib().gkc_composite_frame_shard(_nativeComp[hkey], dst_rank), comp_lib())
        return 3 * 8 * renderer.width * renderer.height * (world - 1) // world
    token, stream = _p2p[hkey]
    renderer.readback_wait()
    with torch.cuda.stream(stream):
        dist.all_reduce(token, group=group)  # every rank has consumed the gather buffers / rtDenoised of the last super-step
    renderer.frame_shard_push()
    with torch.cuda.stream(stream):
        dist.all_reduce(token, group=group)  # all rows have landed
    renderer.frame_shard_accumulate()
    renderer.exchange_push_final(dst_rank)
    with torch.cuda.stream(stream):
        dist.all_reduce(token, group=group)  # the presenting rank holds the whole image
    return 3 * 8 * renderer.width * renderer.height * (world - 1) // world


def release(renderer, group=None):
    """Collective tear-down of the peer mappings of `renderer` (every rank calls it): unmap the peers' planes and gather
    buffers, then wait for all ranks, so that no rank frees (gk_resize / gk_destroy) memory another rank still has mapped.
    The exchange has to be enabled again afterwards."""
    import torch.distributed as dist
    hkey = renderer.h.value if hasattr(renderer.h, "value") else int(renderer.h)
    if hkey in _nativeComp:  # gkc_destroy is the collective tear-down (close peers, barrier, communicator)
        from . import comp_lib
        comp_lib().gkc_destroy(_nativeComp.pop(hkey))
        _shard.discard(hkey)
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=group)
        return
    had = hkey in _p2p or hkey in _shard
    _p2p.pop(hkey, None)
    _shard.discard(hkey)
    for k in [k for k in _staging if k[0] == hkey]:
        _staging.pop(k, None)
    if had:
        renderer.exchange_close_peers()
    if dist.is_available() and dist.is_initialized():
        dist.barrier(group=group)
