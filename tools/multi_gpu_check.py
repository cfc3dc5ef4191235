"""torchrun --nproc-per-node N tools/multi_gpu_check.py
Tile-mode multi-GPU frame (trace own row tiles -> NCCL all-gather of the integrator planes ->
filters on every rank) compared bit for bit with a single-GPU frame rendered on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gknextrenderer_b200 as gk  # noqa: E402
from gknextrenderer_b200 import compositor as comp  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    W, H, TR = 640, 360, 16
    eng = gk.Engine("room", 200000, 5)
    frames_mode = os.environ.get("GK_CHECK_MODE", "temporal") == "frames"
    progressive = os.environ.get("GK_CHECK_MODE", "temporal") == "progressive" or frames_mode
    if progressive:  # the reference's benchmark state: per-pixel filters, only the final image is exchanged
        eng.set(TAA=0, NumberOfSamples=1, NumberOfBounces=4, Denoiser=0, ProgressiveRender=1)
    else:
        eng.set(TAA=1, NumberOfSamples=1, NumberOfBounces=4, Denoiser=1, TemporalFrames=8)
    r = gk.Renderer(W, H, device=local, tile_index=rank, tile_count=world, tile_rows=TR, trace_all_rows=frames_mode)
    eng.update_nodes()
    nodes, n = eng.update_nodes()  # steady-state proxies (second tick), shared by both contexts
    r.upload_scene(eng.scene_desc())
    r.update_instances(nodes, n)
    ref = None
    if rank == 0:
        ref = gk.Renderer(W, H, device=local)
        ref.upload_scene(eng.scene_desc())
        ref.update_instances(nodes, n)
    mode = "nccl all-gather"
    if os.environ.get("GK_EXCHANGE", "p2p") == "p2p" and comp.enable_peer_exchange(r, rank, world):
        mode = "peer-to-peer push" + (" (native compositor)" if comp._hkey(r) in comp._nativeComp else " (torch barriers)")
    if frames_mode:
        assert mode.startswith("peer-to-peer push") and comp.enable_frame_sharding(r, rank, world)
        mode = "frame-sharded: every rank traces its own frame, rows accumulate on their owners"
    ok = True
    for frame in range(4):
        if frames_mode:
            # super-step: rank k traces frame number frame*world + k of the single-GPU sequence
            for k in range(world):
                ubo_k = eng.ubo(W, H)
                if k == rank:
                    r.set_ubo(ubo_k)
                    r.trace_frame()
                if rank == 0:
                    ref.set_ubo(ubo_k)
                    ref.render_frame()
                eng.advance_frame()
            moved = comp.composite_frame_shard(r, rank, world, 0)
            out = r.readback("DENOISED")
            if rank == 0:
                exp = ref.readback("DENOISED")
                same = np.array_equal(out.view(np.uint16), exp.view(np.uint16))
                print(f"super-step {frame} ({world} frames): multi-GPU == single-GPU final image: {same}, bytes exchanged per rank: {moved} ({mode})")
                ok = ok and same
            continue
        ubo = eng.ubo(W, H)
        r.set_ubo(ubo)
        r.trace_frame()
        if progressive:
            r.filter_frame_owned()
            moved = comp.composite_final(r, rank, world, -1)
        else:
            moved = comp.composite_frame(r, rank, world, TR)
            r.filter_frame()
        out = r.readback("DENOISED")
        if rank == 0:
            ref.set_ubo(ubo)
            ref.render_frame()
            exp = ref.readback("DENOISED")
            same = np.array_equal(out.view(np.uint16), exp.view(np.uint16))
            ids_same = progressive or np.array_equal(r.readback("OBJECT_ID0"), ref.readback("OBJECT_ID0"))
            print(f"frame {frame}: multi-GPU == single-GPU final image: {same}, object ids: {ids_same}, bytes exchanged per rank: {moved} ({mode})")
            ok = ok and same and ids_same
        eng.advance_frame()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
