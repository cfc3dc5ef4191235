#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload room|roomu|cornell|city|city400|bricks]

A "step" is one frame of the hot path on synthetic input: per-frame UBO -> wavefront path tracer
(generate / extend / shade / shadow / accumulate) -> fused temporal reprojection -> joint bilateral
denoise + tonemap.  Default workload = BASELINE.json configs[1]: procedural 1M-triangle room,
1920x1080, 1 spp, 4 bounces, temporal 16, JBF size 5.

Prints ONE JSON line (rank 0).  `value` = rays traced by all ranks / device time of the K timed
frames (scene, BVH and frame state resident in HBM); `e2e` = the same through the reference-facing
renderer interface with the UBO coming from host memory and the final RGBA16F image read back to
pinned host memory every frame.  `cpu_baseline` = the reference's CPU ray query (real tinybvh,
oracle/_ref) on the very rays the GPU traced, all host threads.  `--impl reference` makes that the
timed arm.  Multi-GPU (torchrun): image rows are interleaved across ranks in 16-row tiles, scene
and BVH replicated, one NCCL all-gather of the integrator planes at frame end, filters replicated.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene, scene args, width, height, settings)
    "room": ("room", (1000000, 1234), 1920, 1080, dict(NumberOfSamples=1, NumberOfBounces=4, TemporalFrames=16, Denoiser=1, TAA=1)),
    # the same room with unique geometry: 413 objects, each its own model / BLAS, nothing instanced (50 MB of triangle records + BVH)
    "roomu": ("roomu", (1000000, 1234), 1920, 1080, dict(NumberOfSamples=1, NumberOfBounces=4, TemporalFrames=16, Denoiser=1, TAA=1)),
    "cornell": ("cornell", (), 640, 360, dict(NumberOfSamples=8, NumberOfBounces=4, TemporalFrames=16, Denoiser=0, TAA=0, ProgressiveRender=1)),
    "bricks": ("bricks", (200000, 42), 1920, 1080, dict(NumberOfSamples=1, NumberOfBounces=4, TemporalFrames=16, Denoiser=1, TAA=1)),
    # progressive accumulation, denoiser off: the state gkNextBenchmark runs in (gkNextBenchmark.cpp:16-31); one step = 1 of the 64 spp
    "city": ("city", (40, 100, 7, 46), 3840, 2160, dict(NumberOfSamples=1, NumberOfBounces=4, TemporalFrames=16, Denoiser=0, TAA=0, ProgressiveRender=1)),
    # the city with 400 building variants instead of 40: 10.2 M UNIQUE triangles (1.3 GB of triangle records + BVH), 254 M instanced -
    # BASELINE.json configs[4]'s "100M-triangle scene" class; frame-sharded at N > 1 with --shard frames
    "city400": ("city", (400, 100, 7, 46), 3840, 2160, dict(NumberOfSamples=1, NumberOfBounces=4, TemporalFrames=16, Denoiser=0, TAA=0, ProgressiveRender=1)),
}
WORKLOAD_NAMES = {
    "room": "C2: procedural 1M-triangle room (1280 instances of 6 meshes), 1920x1080, 1 spp, 4 bounces, sun+sky, reproject + JBF",
    "roomu": "C2u: procedural 1M-UNIQUE-triangle room (413 objects, one BLAS each, no instancing), 1920x1080, 1 spp, 4 bounces, sun+sky, reproject + JBF",
    "cornell": "C1: built-in Cornell box, 640x360, 8 spp, 4 bounces",
    "bricks": "C3: 200k instanced bricks, per-frame TLAS refit, 1920x1080, 1 spp, 4 bounces",
    "city": "C4: 10M-triangle instanced city, 3840x2160, progressive (1 of 64 spp per step), 4 bounces, denoiser off",
    "city400": "C4u: city with 400 building variants (10.2 M unique / 254 M instanced triangles), 3840x2160, progressive (1 of 64 spp per step), 4 bounces, denoiser off",
}
TILE_ROWS = 16  # rows per tile of the multi-GPU partition; main() shrinks it so that every rank gets >= 32 tiles
METRIC = "path-tracing throughput (rays traced per second, whole frame incl. reproject + denoise) and ms/frame at the named resolution"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="room", choices=list(WORKLOADS))
    ap.add_argument("--shard", default="tiles", choices=["tiles", "frames"],
                    help="multi-GPU partition: image tiles of one frame (strong scaling) or, for progressive workloads without the denoiser, "
                         "one whole frame per rank and step (a step is then `gpus` frames: weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs (the clocks line of
    B200_PROFILING.md): an NVML polling thread (every 10 ms), nvidia-smi -lms as the fallback."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        import threading
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.thread, self.p, self.f = None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self._start_smi(gpu_index)

    def _poll(self):
        nv = self.nv
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                if get_reasons:
                    mask = int(get_reasons(self.h))
                    for label, a, b in names:
                        bit = getattr(nv, a, None) or getattr(nv, b, None)
                        if bit and (mask & int(bit)):
                            self.reasons.add(label)
            except Exception:
                pass
            time.sleep(0.01)

    def _start_smi(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.sm:
                out = {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}
        return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        scene_name, scene_args, W, H, settings = WORKLOADS[args.workload]
        return reference_arm(args, scene_name, scene_args, W, H, settings)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    line = run(args, args.workload, args.shard, args.steps, args.warmup, rank, world, local_rank, secondary=False)
    if world > 1 and args.workload == "room" and not os.environ.get("GK_BENCH_NO_SECONDARY"):
        # BASELINE.json configs[3] (C4: 10M-triangle city, 3840x2160, progressive) on the same box in the same run: the
        # north_star's ">= 7x at 8 GPUs on a 10M-triangle scene" is a statement about this workload, so its multi-GPU numbers
        # are produced here, where the driver runs the scaling bench.  One GPU first (rank 0 alone), then all ranks, image tiles
        # of one frame (strong scaling) and frame sharding (weak scaling: a step is `world` progressive frames).
        sec = []
        solo = run(args, "city", "tiles", 6, 3, rank, world, local_rank, secondary=True, solo=True)
        for shard in ("tiles", "frames"):
            res = run(args, "city", shard, 6, 3, rank, world, local_rank, secondary=True)
            if rank == 0 and res is not None and solo is not None:
                res["one_gpu_same_run"] = {"value": solo["value"], "ms_per_step": solo["ms_per_step"]}
                res["speedup_vs_one_gpu_same_run"] = round(res["value"] / solo["value"], 3)
                sec.append(res)
        if rank == 0:
            line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run(args, workload, shard, steps, warmup, rank, world, local_rank, secondary=False, solo=False):
    """One workload through the timed regions; returns the result line on rank 0 (None elsewhere).  solo: rank 0 renders alone
    (the 1-GPU figure of a multi-rank run), the other ranks wait at the closing barrier."""
    import torch
    import torch.distributed as dist
    import gknextrenderer_b200 as gk
    from gknextrenderer_b200 import compositor as comp

    real_world = world
    if solo:
        if rank != 0:
            dist.barrier()
            return None
        world = 1
    args = argparse.Namespace(**vars(args))
    args.workload, args.shard, args.steps, args.warmup = workload, shard, steps, warmup
    scene_name, scene_args, W, H, settings = WORKLOADS[args.workload]
    global TILE_ROWS
    TILE_ROWS = 16
    if "GK_TILE_ROWS" in os.environ:
        TILE_ROWS = int(os.environ["GK_TILE_ROWS"])
    else:  # interleave finely enough that the rounding of tiles per rank stays below ~3 % of a rank's work
        while TILE_ROWS > 2 and H // (TILE_ROWS * world) < 32:
            TILE_ROWS //= 2

    verbose = bool(os.environ.get("GK_BENCH_VERBOSE"))

    def stage(msg):
        if verbose:
            print(f"[bench rank {rank} {time.perf_counter():.3f}] {msg}", file=sys.stderr, flush=True)

    # ---- the reference-facing path: Assets::Scene -> LogicRendererBase-shaped renderer -> C ABI
    eng = gk.Engine(scene_name, *scene_args)
    eng.set(**settings)
    host = gk.host_lib()
    hr = host.gkh_renderer_create(eng.h, local_rank)
    frame_shard = args.shard == "frames" and world > 1
    if frame_shard:
        assert settings.get("ProgressiveRender", 0) == 1 and settings.get("Denoiser", 1) == 0, "--shard frames needs a progressive workload without the denoiser (city, cornell)"
        assert host.gkh_renderer_set_trace_all_rows(hr, 1) == 0
    assert host.gkh_renderer_set_tile(hr, rank, world, TILE_ROWS) == 0

    def chk(rc):
        if rc != 0:
            raise RuntimeError(host.gkh_last_error().decode())

    chk(host.gkh_renderer_create_swapchain(hr, W, H))
    chk(host.gkh_renderer_post_load_scene(hr))
    ctx = C.c_void_p(host.gkh_renderer_context(hr))
    r = gk.Renderer.__new__(gk.Renderer)  # wrap the context the host renderer owns
    r.lib, r.h, r.width, r.height = gk.cuda_lib(), ctx, W, H
    info_before = None
    stream = torch.cuda.ExternalStream(r.stream(), device=torch.device(f"cuda:{local_rank}"))
    dynamic = scene_name == "bricks"

    exchange_mode = "none"
    if world > 1:
        exchange_mode = "nccl all-gather"
        if os.environ.get("GK_EXCHANGE", "p2p") == "p2p" and comp.enable_peer_exchange(r, rank, world):
            exchange_mode = "peer-to-peer push over NVLink (CUDA IPC) between two 4-byte all-reduce barriers"
            if comp._hkey(r) in comp._nativeComp:
                exchange_mode += "; sequenced in C++ by lib/libgknext_comp.so on its own NCCL communicator (include/gknext_compositor.h)"

    local_filters = world > 1 and exchange_mode.startswith("peer") and settings.get("ProgressiveRender", 0) == 1 and settings.get("Denoiser", 1) == 0
    if local_filters:
        exchange_mode = "filters on owned rows; final image rows pushed peer-to-peer over NVLink to rank 0 (the presenting rank) between two 4-byte all-reduce barriers"

    if frame_shard:
        assert exchange_mode.startswith("peer") or exchange_mode.startswith("filters"), "frame sharding needs the peer mapping"
        assert comp.enable_frame_sharding(r, rank, world)
        local_filters = False
        exchange_mode = ("frame sharding: rank k traces the whole frame f0+k; rows of the three source planes go to their owning rank (NVLink stores), which "
                         "accumulates the frames in order and composes; finished rows go to rank 0; three 4-byte all-reduce barriers per step")

    def frame(step_index, exchange=True):
        if dynamic:
            eng.step_scene(step_index)
        chk(host.gkh_renderer_before_next_frame(hr))  # Scene::UpdateNodes -> instances -> TLAS
        if world == 1:
            chk(host.gkh_renderer_render(hr))  # UBO fill + gk_set_ubo + gk_render_frame
            st = r.stats()
            return st.primaryRays + st.extensionRays + st.shadowRays, st.launches, st
        if frame_shard:  # a step = `world` consecutive frames of the progressive sequence, one per rank
            ubo = None
            for k in range(world):
                if k == rank:
                    ubo = eng.ubo(W, H)
                eng.advance_frame()
        else:
            ubo = eng.ubo(W, H)
        r.set_ubo(ubo)
        r.trace_frame()
        st = r.stats()
        rays, launches = st.primaryRays + st.extensionRays + st.shadowRays, st.launches
        if frame_shard:
            if exchange:
                xa, xb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                xa.record(stream)
                comp.composite_frame_shard(r, rank, world, 0)
                xb.record(stream)
                xchg_events.append((xa, xb))
            return rays, launches + 4, st
        if local_filters:
            # progressive, no denoiser: per-pixel filters on the owned rows, then only the final image travels
            r.filter_frame_owned()
            if exchange:
                xa, xb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                xa.record(stream)
                comp.composite_final(r, rank, world, 0)  # the presenting rank
                xb.record(stream)
                xchg_events.append((xa, xb))
            eng.advance_frame()
            return rays, launches + 1, st
        if exchange:
            xa, xb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            xa.record(stream)
            comp.composite_frame(r, rank, world, TILE_ROWS)
            xb.record(stream)
            xchg_events.append((xa, xb))
        r.filter_frame()
        eng.advance_frame()
        return rays, launches + 2, st

    xchg_events = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    stage("scene loaded, warm-up")
    n_warm = max(3, args.warmup)
    for i in range(n_warm):
        frame(i)
    info = r.bvh_info()

    # ---- timed: device-resident inputs
    stage("timed region (device-resident)")
    sampler = ClockSampler(local_rank) if rank == 0 else None  # started before the barrier: NVML start-up must not delay rank 0 inside the timed region
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = dict(rays=0, launches=0, ext=0.0, shd=0.0, shade=0.0, gen=0.0, acc=0.0, rep=0.0, jbf=0.0, bvh=0.0, eray=0, sray=0, pray=0, waves=0, tail=0.0, tail_eray=0, tail_paths=0, ext_launches=0, trace=0.0, trace_kernels=0.0, tail_sray=0)
    xchg_events.clear()
    e0.record(stream)
    t_wall = time.perf_counter()
    for i in range(args.steps):
        rays, launches, st = frame(n_warm + i)
        agg["rays"] += rays; agg["launches"] += launches
        agg["ext"] += st.msExtend; agg["shd"] += st.msShadow; agg["shade"] += st.msShade; agg["gen"] += st.msGenerate; agg["acc"] += st.msAccumulate
        agg["rep"] += st.msReproject; agg["jbf"] += st.msDenoise; agg["bvh"] += st.msBvh if dynamic else 0.0
        agg["eray"] += st.extensionRays; agg["sray"] += st.shadowRays; agg["pray"] += st.primaryRays; agg["waves"] += st.waves
        agg["trace"] += st.msTotal; agg["trace_kernels"] += st.msTrace; agg["tail_sray"] += st.tailShadowRays
        agg["tail"] += st.msTail; agg["tail_eray"] += st.tailExtensionRays; agg["tail_paths"] += st.tailPaths; agg["ext_launches"] += st.waves - (1 if st.tailPaths else 0)
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    dev_ms = e0.elapsed_time(e1)
    agg["xchg"] = sum(a.elapsed_time(b) for a, b in xchg_events)  # pack + all-gather (incl. waiting for the slowest rank) + unpack
    xchg_events.clear()
    clocks = sampler.stop() if sampler else None

    # ---- timed: end to end through the renderer interface, host UBO in, final image out
    # the final image of every step is read back into one of two pinned buffers; the copy runs on the
    # library's copy stream and overlaps the next step (gk_readback_async), the last one is waited for
    stage("timed region (end to end)")
    final_host = [torch.empty((H, W, 4), dtype=torch.float16).pin_memory() for _ in range(2)]
    fin_bytes = r.plane_bytes("DENOISED")
    for i in range(2):
        frame(n_warm + args.steps + i)
        if rank == 0:
            r.readback_async("DENOISED", final_host[i & 1].data_ptr(), fin_bytes)
    r.readback_wait()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    rays_e2e = 0
    uploaded0 = int(host.gkh_renderer_instance_bytes_uploaded(hr))
    for i in range(args.steps):
        rr, _, _ = frame(n_warm + args.steps + 2 + i)
        rays_e2e += rr
        if rank == 0:  # one presenting rank reads the finished image back (the reference presents from one device)
            r.readback_async("DENOISED", final_host[i & 1].data_ptr(), fin_bytes)
    r.readback_wait()
    e3.record(stream)
    barrier()
    e2e_ms = e2.elapsed_time(e3)
    # per step: the UBO (784 B) plus what BeforeNextFrame sent for the instances, counted by the host mirror (the changed proxies and their indices)
    h2d = 784 + (int(host.gkh_renderer_instance_bytes_uploaded(hr)) - uploaded0) / args.steps
    d2h = fin_bytes

    stage("reduce over ranks")
    # ---- reduce over ranks: max time, summed rays
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        c = torch.tensor([agg["rays"], rays_e2e, agg["launches"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_rays, total_rays_e2e, total_launches = float(c[0]), float(c[1]), int(c[2])
        # load balance evidence: device time of the path tracer (generate .. accumulate) and rays per rank and step
        mine = torch.tensor([agg["trace"] / args.steps, agg["rays"] / args.steps], device="cuda", dtype=torch.float64)
        every = torch.empty(2 * world, device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(every, mine)
        per_rank = {"trace_ms": [round(float(v), 3) for v in every[0::2]], "rays": [int(v) for v in every[1::2]]}
    else:
        per_rank = None
        total_rays, total_rays_e2e, total_launches = float(agg["rays"]), float(rays_e2e), int(agg["launches"])

    def release():
        if world > 1:
            comp.release(r)  # collective: unmap the peers before anyone frees its planes
        r.h = None  # the context belongs to the host renderer
        host.gkh_renderer_destroy(hr)
        eng.close()

    if world > 1:
        dist.barrier()
    if rank != 0:
        release()
        return None
    if secondary:
        value = total_rays / (dev_ms * 1e-3) / 1e6
        res = {"workload": WORKLOAD_NAMES[args.workload], "shard": "frames" if frame_shard else "tiles", "n_gpus": world, "value": round(value, 2), "unit": "Mrays/s",
               "ms_per_step": round(dev_ms / args.steps, 4), "steps": args.steps, "warmup": n_warm, "scaling": "weak" if frame_shard else "strong",
               "e2e": {"value": round(total_rays_e2e / (e2e_ms * 1e-3) / 1e6, 2), "ms_per_step": round(e2e_ms / args.steps, 4)},
               "rays_per_step": round(total_rays / args.steps, 0), "per_rank_trace_ms": per_rank["trace_ms"] if per_rank else None, "exchange": exchange_mode,
               "breakdown_ms_per_step": {k: round(agg[k] / args.steps, 4) for k in ("trace_kernels", "shade", "tail", "xchg", "rep", "jbf")}}
        release()
        if solo and real_world > 1:
            dist.barrier()
        return res

    stage("roofline probes (rank 0)")
    # ---- roofline of the dominant kernel (traversal).  SURVEY.md 8(d): L2 / SM-issue bound; algorithmic bytes per ray =
    #      48 (ray in + hit out) + 80 x node visits + 48 x triangle tests, the visits counted by an instrumented frame.
    r.set_traversal_stats(True)
    rays1, _, st1 = frame(10_000, exchange=False)  # rank-local: the other ranks have left, no collective here
    r.set_traversal_stats(False)
    traced = st1.extensionRays + st1.shadowRays + st1.primaryRays
    nodes_per_ray = st1.nodeVisits / max(traced, 1)
    tris_per_ray = st1.triTests / max(traced, 1)
    bytes_per_ray = 48.0 + 80.0 * nodes_per_ray + 48.0 * tris_per_ray
    bytes_per_ray_loaded = 48.0 + 112.0 * nodes_per_ray + 48.0 * tris_per_ray  # what the kernel asks L1 for: 16 B header + 8 x 12 B children of the 128-B node line
    sched_it, sched_ln = list(st1.schedIters), list(st1.schedLanes)
    # rays traced by the traversal launches of the waves (closest hit + any hit) and the time those launches take: extend and
    # shadow kernel of a wave run concurrently, so they are timed as one span
    ext_rays = agg["eray"] + agg["pray"] - agg["tail_eray"] + agg["sray"] - agg["tail_sray"]
    ext_ms = agg["trace_kernels"]
    l2_peak = r.measure_read_bandwidth(48 << 20, 16)       # the library's own streaming-read kernel over an L2-resident 48 MB buffer
    hbm_read = r.measure_read_bandwidth(4 << 30, 1)        # the same kernel over 4 GB: HBM read bandwidth, for reference
    achieved = ext_rays * bytes_per_ray / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(args.workload, {})
    except Exception:
        pass
    px = W * H
    filt_ms = (agg["rep"] + agg["jbf"]) / args.steps
    launches_per_step = max(agg["ext_launches"], 1) / args.steps
    roofline = {"kernel": "k_trace_sched / k_trace (extend + shadow launches of the waves: traversal of the 8-wide quantised two-level BVH; the two run concurrently and are timed as one span per wave)",
                "bound": "l2", "achieved": round(achieved, 1), "peak": round(l2_peak, 1), "unit": "GB/s",
                "frac": round(achieved / l2_peak, 4) if l2_peak else None,
                "traffic": traffic.get("dram_bytes_per_wave_avg"), "traffic_source": traffic.get("source"),
                "algorithmic_bytes_per_wave": round(bytes_per_ray * ext_rays / max(agg["ext_launches"], 1), 0),
                "note": "achieved = SURVEY 8(d) algorithmic bytes (48 B ray/hit + 80 B per node visit + 48 B per triangle test, visits counted by an instrumented frame) x rays / kernel time; "
                        "peak = L2 read bandwidth measured in this run by the library's streaming-read kernel (gk_measure_read_bandwidth, 48 MB buffer, L1 bypassed). "
                        "The node/triangle bytes are served by L1 (hit rate 83-86 %, profiles/) and L2, only the compulsory ray/hit records reach HBM (traffic): the binding resource is "
                        "instruction issue (see issue), as SURVEY 8(d) anticipated (L2 / SM-issue bound).",
                "model": {"bytes_per_ray": round(bytes_per_ray, 1), "bytes_per_ray_loaded_from_l1": round(bytes_per_ray_loaded, 1), "node_visits_per_ray": round(nodes_per_ray, 2),
                          "tri_tests_per_ray": round(tris_per_ray, 2), "instance_entries_per_ray": round(st1.instanceEntries / max(traced, 1), 2),
                          "rays_per_wave_avg": round(ext_rays / max(agg["ext_launches"], 1), 0), "waves_per_step": round(launches_per_step, 1), "ms_per_step": round(ext_ms / args.steps, 4),
                          "Grays_per_s_in_kernel": round(ext_rays / (ext_ms * 1e-3) / 1e9, 4) if ext_ms > 0 else None},
                "issue": {"lanes_per_step_node_tri_instance": [round(sched_ln[k] / max(1, sched_it[k]), 2) for k in range(3)],
                          "steps_per_frame_node_tri_instance": [int(v) for v in sched_it],
                          "alive_lanes_at_vote": round(st1.schedPopLanes / max(1, st1.schedPopIters), 2),
                          "source": "counted live by the scheduled kernel in the instrumented frame (gk_set_traversal_stats): lanes taking part in a step of the 32 of a warp",
                          "ncu": traffic.get("issue")},
                "hbm": {"bound": "hbm", "achieved": round(traffic["dram_bytes_per_wave_avg"] / (ext_ms / max(agg["ext_launches"], 1) * 1e-3) / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(traffic["dram_bytes_per_wave_avg"] / (ext_ms / max(agg["ext_launches"], 1) * 1e-3) / 1e9 / hbm_peak, 4),
                        "note": "ncu-measured DRAM bytes of the same launches against the HBM peak: nowhere near the HBM bound"} if traffic.get("dram_bytes_per_wave_avg") and ext_ms > 0 else None,
                "peak_source": "measured in this run: gk_measure_read_bandwidth over an L2-resident 48 MB buffer", "hbm_read_gbs_same_probe": round(hbm_read, 1)}
    roofline_filters = {"kernel": "k_reproject + k_denoise_jbf", "bound": "hbm", "achieved": round((96 + 48) * px / (filt_ms * 1e-3) / 1e9, 1) if filt_ms > 0 else None,
                        "peak": hbm_peak, "unit": "GB/s", "frac": round((96 + 48) * px / (filt_ms * 1e-3) / 1e9 / hbm_peak, 4) if filt_ms > 0 else None,
                        "bytes_per_pixel": {"reproject": 96, "denoise": 48, "source": "SURVEY.md 8(d)"}, "ms_per_step": round(filt_ms, 4),
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}

    # ---- CPU baseline: the reference's tinybvh on the rays the GPU traced (bounded sample)
    cpu, parity = None, None
    if not args.no_cpu_baseline and world == 1:
        cpu, parity = cpu_baseline(eng, r, frame, W, H)

    # ---- the BVH build once more with every buffer already allocated (the first build of a process pays the cudaMallocs inside its
    # timed span; a scene reload or a streamed-in model does not), and its roofline: bytes per primitive as gk_bvh_build.cu's header
    # counts them (64-bit key + index through the 8 radix passes, boxes, binary node, cost table, share of a 128-byte wide node).
    roofline_build = None
    if world == 1:
        r.upload_scene(eng.scene_desc())
        warm = r.bvh_info()
        eng.mark_dirty()
        nodes, n = eng.update_nodes()
        r.update_instances(nodes, n, False)
        warm_tlas = r.bvh_info()
        prims = int(warm.triangleCount)
        bytes_per_prim = 12 * (2 + 2 * 8) + 2 * 32 + 40 + 32 + 128.0 * warm.blasNodes8 / max(prims, 1)
        gbs = prims * bytes_per_prim / (warm.msBlasBuild * 1e-3) / 1e9 if warm.msBlasBuild > 0 else 0.0
        roofline_build = {"kernel": "BLAS forest build (k_morton, cub radix sort, k_radix_tree, k_leaf_boxes + k_propagate_bounds with the SAH cost table, k_collapse_all, k_quantise)",
                          "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 4),
                          "primitives": prims, "bytes_per_primitive": round(bytes_per_prim, 1), "blas_build_warm_ms": round(warm.msBlasBuild, 3),
                          "Mprims_per_s": round(prims / (warm.msBlasBuild * 1e-3) / 1e6, 1) if warm.msBlasBuild > 0 else None,
                          "tlas_build_warm_ms": round(warm_tlas.msTlasBuild, 3), "instances": int(warm_tlas.instanceCount),
                          "note": "launch- and dependency-bound, not bandwidth-bound: ~25 short kernels and one cooperative launch with a grid barrier per tree level"}

    value = total_rays / (dev_ms * 1e-3) / 1e6
    line = {
        "metric": METRIC,
        "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak" if frame_shard else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "width": W, "height": H, "spp": settings["NumberOfSamples"], "bounces": settings["NumberOfBounces"],
                   "triangles_instanced": int(info.instancedTriangles), "triangles_unique": int(info.triangleCount), "instances": int(info.instanceCount),
                   "partition": (f"one whole frame per rank and step ({world} frames per step), rows owned in {TILE_ROWS}-row tiles for accumulation, scene+BVH replicated" if frame_shard
                                 else f"{TILE_ROWS}-row tiles interleaved over {world} rank(s), scene+BVH replicated"), "exchange": exchange_mode,
                   "l2_policy": "per-frame working set (path state + queues + planes, >500 MB at 1080p) exceeds the 126 MB L2; no explicit flush"},
        "rays_per_step": round(total_rays / args.steps, 0), "gpu_launches": total_launches,
        "e2e": {"value": round(total_rays_e2e / (e2e_ms * 1e-3) / 1e6, 2), "unit": "Mrays/s", "ms_per_step": round(e2e_ms / args.steps, 4), "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "api": "CudaPathTracingRenderer::BeforeNextFrame + Render (host mirror of LogicRendererBase) + gk_readback_async(rtDenoised) to pinned memory, double-buffered"},
        "breakdown_ms_per_step": {k: round(agg[k] / args.steps, 4) for k in ("gen", "ext", "shd", "trace_kernels", "shade", "tail", "acc", "xchg", "rep", "jbf", "bvh")},
        "per_rank": per_rank,
        "tail_paths_per_step": round(agg["tail_paths"] / args.steps, 0),
        "host_gap_ms_per_step": round((wall_ms - dev_ms) / args.steps, 4),
        "bvh": {"blas_build_ms": round(info.msBlasBuild, 3), "blas_build_note": "first build of the process: includes the device allocations (see roofline_build for the warm figure)", "tlas_build_ms": round(info.msTlasBuild, 3), "tlas_refit_ms": round(info.msRefit, 3), "refits_rejected": int(r.bvh_info().refitsRejected),
                "wide_nodes_blas": int(info.blasNodes8), "wide_nodes_tlas": int(info.tlasNodes8), "bytes": int(info.bytesBvh + info.bytesGeometry)},
        "roofline": roofline, "roofline_filters": roofline_filters, "roofline_build": roofline_build, "cpu_baseline": cpu, "parity": parity, "clocks": clocks,
    }
    release()
    return line


def _captured_sample(eng, r, frame_fn):
    """Primary wave + the third extend wave (incoherent bounce rays) of one frame."""
    out = []
    for wave in (0, 2):
        r.set_ray_capture(wave)
        frame_fn(20_000 + wave)
        out.append(r.captured_rays(r.width * r.height).copy())
    r.set_ray_capture(-1)
    return out


def _cpu_rate(eng, rays_list, repeats=3):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    nodes, n = eng.update_nodes()
    use_ref = ol.have_ref()
    scene = ol.OracleScene(eng.scene_desc(), nodes, n, use_ref=use_ref)
    threads = os.cpu_count() or 1
    res, hits = [], []
    for rays in rays_list:
        best = None
        for _ in range(repeats):
            tuv, ids = scene.intersect(rays, threads=threads)
            best = scene.last_seconds if best is None else min(best, scene.last_seconds)
        res.append((len(rays), best))
        hits.append((tuv, ids))
    return use_ref, threads, res, hits


TIE_EPS = 2.0 ** -19  # as tests/test_gpu_parity.py: 16 ulp, GPU distance not the farther one


def hit_parity(gpu_hits, cpu_hits, rays_list=None, brute=None):
    """GPU hit records against the CPU reference's on the same rays (the classification of tests/test_gpu_parity.py):
    ids equal + t/u/v bit-equal; exact-distance ties (different id, same t bits: coincident surfaces); epsilon ties (the
    GPU distance is not farther and within 2^-19 relative); reference misses (the GPU hit is closer by more than that:
    tinybvh's rounded slab test culled the box of the nearest triangle - accepted only when the GPU hit equals the
    exhaustive no-BVH search, checked on a sample); errors (anything else)."""
    tot = dict(rays=0, id_equal=0, tuv_bit_equal=0, exact_t_ties=0, eps_ties=0, reference_misses=0, errors=0, reference_misses_checked=0)
    for k, ((g_tuv, g_ids), (o_tuv, o_ids)) in enumerate(zip(gpu_hits, cpu_hits)):
        differ = (g_ids != o_ids).any(axis=1)
        gb, ob = np.ascontiguousarray(g_tuv).view(np.uint32), np.ascontiguousarray(o_tuv).view(np.uint32)
        tg, to = g_tuv[:, 0].astype(np.float64), o_tuv[:, 0].astype(np.float64)
        exact = differ & (gb[:, 0] == ob[:, 0])
        eps = differ & ~exact & (np.abs(tg - to) <= TIE_EPS * np.maximum(np.abs(tg), np.abs(to))) & (tg <= to)
        closer = differ & ~exact & ~eps & (tg < to)
        errors = differ & ~exact & ~eps & ~closer
        if closer.any() and brute is not None and rays_list is not None:
            idx = np.nonzero(closer)[0][:32]
            b_tuv, _ = brute(rays_list[k][idx])
            ok = np.ascontiguousarray(b_tuv).view(np.uint32)[:, 0] == gb[idx, 0]
            tot["reference_misses_checked"] += int(ok.sum())
            errors[idx[~ok]] = True
            closer[idx[~ok]] = False
        tot["rays"] += len(g_ids); tot["id_equal"] += int((~differ).sum()); tot["tuv_bit_equal"] += int(((gb == ob).all(axis=1) & ~differ).sum())
        tot["exact_t_ties"] += int(exact.sum()); tot["eps_ties"] += int(eps.sum()); tot["reference_misses"] += int(closer.sum()); tot["errors"] += int(errors.sum())
    return tot


def cpu_baseline(eng, r, frame_fn, W, H):
    rays_list = _captured_sample(eng, r, frame_fn)
    use_ref, threads, res, cpu_hits = _cpu_rate(eng, rays_list)
    total = sum(n for n, _ in res)
    secs = sum(s for _, s in res)
    # the GPU traversal on the very same ray buffers (checker only, untimed); tinybvh has no tmin (it accepts t > 0), so the
    # comparison runs with tmin = 0 on both sides
    gpu_hits, rays0_list = [], []
    for rays in rays_list:
        rays0 = rays.copy()
        rays0[:, 3] = 0.0
        rays0_list.append(rays0)
        gpu_hits.append(r.intersect(rays0))

    def brute(sub):  # exhaustive no-BVH search of the oracle, only for rays the reference itself got wrong
        import oracle_lib as ol
        nodes, n = eng.update_nodes()
        return ol.OracleScene(eng.scene_desc(), nodes, n).intersect_bruteforce(sub)

    parity = hit_parity(gpu_hits, cpu_hits, rays0_list, brute)
    parity["against"] = ("real tinybvh (oracle/_ref)" if use_ref else "oracle port") + ", tmin = 0 on both sides"
    return {"value": round(total / secs / 1e6, 3), "unit": "Mrays/s", "cores": threads, "kind": "reference" if use_ref else "port",
            "sample": f"{res[0][0]} primary rays ({res[0][0] / res[0][1] / 1e6:.2f} Mrays/s) + {res[1][0]} third-wave bounce rays ({res[1][0] / max(res[1][1], 1e-9) / 1e6:.2f} Mrays/s) "
                      "captured from the GPU frame, traversal only, best of 3, tinybvh BVH::Intersect over the TLAS" + ("" if use_ref else " (oracle port)")}, parity


def reference_arm(args, scene_name, scene_args, W, H, settings):
    """The reference's CPU implementation of the path (tinybvh ray query, CPUAccelerationStructure.cpp)
    timed on the host cores.  Rays are the camera rays plus cosine-distributed bounce rays derived
    from them on the CPU (no GPU is used by this arm)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gknextrenderer_b200 as gk
    import oracle_lib as ol
    eng = gk.Engine(scene_name, *scene_args)
    eng.set(**settings)
    eng.set(TAA=0)
    nodes, n = eng.update_nodes()
    use_ref = ol.have_ref()
    scene = ol.OracleScene(eng.scene_desc(), nodes, n, use_ref=use_ref)
    threads = os.cpu_count() or 1
    prim = ol.primary_rays(eng.ubo(W, H), W, H, tmax=2000.0)
    stride = max(1, len(prim) // 400000)  # bounded sample: ~0.4M primary + ~0.4M bounce rays per step
    prim = prim[::stride]
    tuv, ids = scene.intersect(prim, threads=threads)
    hit = ids[:, 1] != 0xFFFFFFFF
    rng = np.random.default_rng(1)
    P = prim[hit, 0:3] + prim[hit, 4:7] * tuv[hit, 0:1] * np.float32(0.999)
    d = rng.normal(size=P.shape).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    bounce = np.zeros((len(P), 8), np.float32)
    bounce[:, 0:3], bounce[:, 3], bounce[:, 4:7], bounce[:, 7] = P, 1e-3, d, 1000.0
    rays = np.concatenate([prim, bounce])
    for _ in range(max(1, min(args.warmup, 2))):
        scene.intersect(rays, threads=threads)
    secs = 0.0
    for _ in range(args.steps):
        scene.intersect(rays, threads=threads)
        secs += scene.last_seconds
    value = args.steps * len(rays) / secs / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(secs / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[args.workload], "width": W, "height": H},
            "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": threads, "kind": "reference" if use_ref else "port",
                             "sample": f"{len(prim)} camera rays (every {stride}th pixel) + {len(bounce)} uniformly random bounce rays from their hit points per step; "
                                       "tinybvh BVH::Intersect over the TLAS, traversal only"},
            "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
