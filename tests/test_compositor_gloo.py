"""Multi-GPU host logic on CPU: the row-tile partition and the frame-end all-gather, run with
world_size 2 and 3 over gloo (the GPU box runs the same code over NCCL)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gknextrenderer_b200 import compositor as comp


def test_row_partition_covers_image_exactly_once():
    for H, tr, world in [(360, 16, 2), (1080, 16, 8), (37, 5, 3), (90, 16, 4), (8, 16, 2)]:
        seen = np.zeros(H, int)
        for r in range(world):
            rows = comp.owned_rows(H, tr, r, world)
            seen[rows] += 1
            blocks = comp.row_blocks(H, tr, r, world)
            assert sum(c for _, c in blocks) == len(rows)
            assert all((r0 // tr) % world == r for r0, _ in blocks)
            assert len(blocks) <= comp.padded_blocks(H, tr, world)
        assert (seen == 1).all()


def _worker(rank, world, port, H, W, tr, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        results = {}
        for name, bpp in comp.EXCHANGE_PLANES:
            plane = torch.zeros((H, W * bpp), dtype=torch.uint8)
            rows = comp.owned_rows(H, tr, rank, world)
            # every owned row carries a value that encodes (rank, row); the rest stays zero
            for r in rows:
                plane[r] = (rank * 37 + r * 11 + len(name)) % 251 + 1
            comp.all_gather_plane(plane, H, tr, rank, world)
            results[name] = plane.numpy().copy()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **results)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,tr", [(2, 90, 16), (3, 37, 5)])
def test_frame_end_all_gather_over_gloo(tmp_path, world, H, tr):
    W = 24
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, H, W, tr, str(tmp_path)), nprocs=world, join=True)
    ref = None
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        for name, bpp in comp.EXCHANGE_PLANES:
            exp = np.zeros((H, W * bpp), np.uint8)
            for r in range(world):
                for row in comp.owned_rows(H, tr, r, world):
                    exp[row] = (r * 37 + row * 11 + len(name)) % 251 + 1
            assert np.array_equal(got[name], exp), f"rank {rank} plane {name}"
        if ref is None:
            ref = {k: got[k] for k in got.files}
        else:
            assert all(np.array_equal(ref[k], got[k]) for k in ref)  # every rank ends with the same full frame


def test_frame_sharded_accumulation_equals_sequential_progressive_frames():
    """Host-logic model of gk_frame_shard_push/accumulate (csrc/gk_exchange.cu): `world` frames traced by different
    ranks, rows gathered on their owners and folded into the history in frame order with a round to RGBA16F after
    every step, must equal `world` consecutive progressive reprojection passes of the oracle (ReProject:76-82)."""
    import ctypes as C
    import os
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import oracle_lib as ol
    from gknextrenderer_b200 import GkUniformBufferObject
    lib = ol.load_oracle()
    W, H, world, tile_rows = 64, 40, 4, 8
    rng = np.random.default_rng(7)
    u = GkUniformBufferObject()
    u.ViewportRect[:] = [0, 0, W, H]
    u.ProgressiveRender, u.TemporalFrames, u.TotalFrames = 1, 16, 3
    frames = [np.exp(rng.normal(size=(H, W, 4))).astype(np.float16) for _ in range(world)]
    history = np.exp(rng.normal(size=(H, W, 4))).astype(np.float16)
    zeros2, zeros1, zeros4 = np.zeros((H, W, 2), np.float32), np.zeros((H, W), np.uint32), np.zeros((H, W, 4), np.uint16)
    # sequential: one oracle reprojection pass per frame
    seq = history.view(np.uint16).copy()
    for s in range(world):
        out = np.zeros((H, W, 4), np.uint16)
        lib.orc_reproject(C.byref(u), W, H, 0, 1, ol.ptr(frames[s].view(np.uint16)), ol.ptr(seq), ol.ptr(zeros2), ol.ptr(zeros1), ol.ptr(zeros1), ol.ptr(zeros4), ol.ptr(out))
        seq = out
    # sharded: owner of a row folds the `world` sources of that row (same lerp, same rounding)
    keep = np.float32(1.0) / np.float32(u.TemporalFrames)
    sharded = np.zeros((H, W, 4), np.float16)
    for rank in range(world):
        rows = comp.owned_rows(H, tile_rows, rank, world)
        h = history[rows]
        for s in range(world):  # gather slot s = the frame traced by rank s
            hf, cf = h.astype(np.float32), frames[s][rows].astype(np.float32)
            mixed = hf * (np.float32(1.0) - keep) + cf * keep
            mixed[..., 3] = 1.0  # the pass writes alpha = 1
            h = mixed.astype(np.float16)
        sharded[rows] = h
    assert np.array_equal(sharded.view(np.uint16), seq)
