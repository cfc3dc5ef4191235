"""Whole-grid probe bake (gk_bake_probes over all 192 x 48 x 192 probes) on a workload: time per pass, probes classified near a
surface, lit faces, and how much the baked grid changes a real-time frame (the ambient-cube terminator of the path tracer).

    python tools/gpu_bake_bench.py [workload] [passes]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gknextrenderer_b200 as gk  # noqa: E402
from bench import WORKLOADS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "room"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene, args, W, H, settings = WORKLOADS[name]
eng = gk.Engine(scene, *args)
eng.set(**settings)
r = gk.Renderer(W, H, device=0)
r.load(eng)
ubo = eng.ubo(W, H)
r.set_ubo(ubo)
r.trace_frame()
before = r.readback("RADIANCE_DIFFUSE_F32").copy()
n = 192 * 48 * 192
for p in range(passes):
    r.synchronize()
    t0 = time.perf_counter()
    r.bake_probes(0, n)
    r.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    cubes, voxels = r.get_probes()
    print(f"{name} bake pass {p}: {ms:.2f} ms for {n} probes ({n / ms / 1e3:.1f} M probes/s), near a surface {int((voxels[:, 1] > 0).sum())}, "
          f"inside geometry {int((voxels[:, 0] > 0).sum())}, lit faces {int((cubes[:, :12] != 0).sum())}", flush=True)
r.set_ubo(ubo)
r.trace_frame()
after = r.readback("RADIANCE_DIFFUSE_F32").copy()
d = np.abs(after[..., :3] - before[..., :3])
print(f"{name}: frame with the baked grid vs un-baked: mean |delta| {float(d.mean()):.5f}, pixels changed {float((d.max(axis=-1) > 0).mean()) * 100:.1f} %, "
      f"mean radiance before {float(before[..., :3].mean()):.4f} after {float(after[..., :3].mean()):.4f}")
