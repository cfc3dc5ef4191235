"""Traversal laboratory: one GPU call compares kernel variants / tuning options on the SAME captured rays.

For each workload it captures the extension-ray buffers of waves 0..2 of one frame, then for every
configuration in CONFIGS (a dict of gk_set_option values) it
  * times gk_intersect_device (closest and any hit) on each captured wave (CUDA events, best of `reps`),
  * checks the hit records bit for bit against the first configuration,
  * times whole frames, and prints the scheduling statistics of one instrumented frame.

    python tools/gpu_trace_lab.py [workload ...] [--reps N] [--configs name,name] [--out file.jsonl]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gknextrenderer_b200 as gk  # noqa: E402
from bench import WORKLOADS  # noqa: E402

def _sched(refill=8, bias=0, keep_n=12, keep_t=10):
    return dict(trace_variant=1, sched_refill_min=refill, sched_bias_node=bias, sched_keep_node=keep_n, sched_keep_tri=keep_t)


CONFIGS = {
    "legacy": dict(trace_variant=0),
    "sched": _sched(),
    "sched_step": _sched(keep_n=33, keep_t=33),  # one step per vote
    "sched_kn8": _sched(keep_n=8), "sched_kn16": _sched(keep_n=16), "sched_kn20": _sched(keep_n=20), "sched_kn24": _sched(keep_n=24),
    "sched_kt4": _sched(keep_t=4), "sched_kt16": _sched(keep_t=16),
    "sched_r4": _sched(refill=4), "sched_r12": _sched(refill=12), "sched_r16": _sched(refill=16),
    "sched_b4": _sched(bias=4),
    "sched_kt4_r4": dict(_sched(refill=4, keep_t=4)),
    "sched_look0": dict(_sched(keep_t=4), wave_lookahead=0), "sched_look1": dict(_sched(keep_t=4), wave_lookahead=1),
    "sched_look3": dict(_sched(keep_t=4), wave_lookahead=3), "sched_look2": dict(_sched(keep_t=4), wave_lookahead=2),
}


def apply(r, cfg):
    for k, v in cfg.items():
        r.set_option(k, v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["room"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--configs", default=",".join(CONFIGS))
    ap.add_argument("--out", default=None)
    ap.add_argument("--frames", type=int, default=4)
    a = ap.parse_args()
    names = [n for n in a.configs.split(",") if n]
    out = open(a.out, "a") if a.out else None
    for wl in a.workloads:
        scene, args, W, H, settings = WORKLOADS[wl]
        eng = gk.Engine(scene, *args)
        eng.set(**settings)
        r = gk.Renderer(W, H, device=0)
        r.load(eng)
        r.set_ubo(eng.ubo(W, H))
        info = r.bvh_info()
        print(f"== {wl}: {info.triangleCount} unique tris, {info.instanceCount} instances, wide nodes blas {info.blasNodes8} tlas {info.tlasNodes8}", flush=True)
        stream = torch.cuda.ExternalStream(r.stream())
        apply(r, CONFIGS["legacy"])
        waves = []
        for wave in (0, 1, 2):
            r.set_ray_capture(wave)
            r.trace_frame()
            waves.append(torch.from_numpy(r.captured_rays(W * H).copy()).cuda())
            r.set_ray_capture(-1)
        base = {}
        for name in names:
            cfg = CONFIGS[name]
            apply(r, cfg)
            rec = {"workload": wl, "config": name, "options": cfg}
            for wi, d in enumerate(waves):
                n = d.shape[0]
                if n == 0:
                    continue
                tuv = torch.empty((n, 3), dtype=torch.float32, device="cuda")
                ids = torch.empty((n, 2), dtype=torch.int32, device="cuda")
                for any_hit in (False, True):
                    best = 1e9
                    for _ in range(a.reps):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        r.intersect_device(d.data_ptr(), n, tuv.data_ptr(), ids.data_ptr(), any_hit)
                        e1.record(stream)
                        r.synchronize()
                        best = min(best, e0.elapsed_time(e1))
                    key = (wi, any_hit)
                    res = (tuv.clone(), ids.clone())
                    if key not in base:
                        base[key] = res
                        same = True
                    else:
                        bt, bi = base[key]
                        if any_hit:
                            same = bool(torch.equal(bi, res[1]))
                        else:
                            same = bool(torch.equal(bi, res[1])) and bool(torch.equal(bt.view(torch.int32), res[0].view(torch.int32)))
                    rec[f"w{wi}_{'any' if any_hit else 'closest'}"] = {"rays": n, "ms": round(best, 4), "grays": round(n / best / 1e6, 4), "same_as_first": same}
                    if not same:
                        bt, bi = base[key]
                        nd = int((bi != res[1]).any(dim=1).sum().item())
                        rec[f"w{wi}_{'any' if any_hit else 'closest'}"]["id_mismatches"] = nd
            # whole frames
            for _ in range(2):
                r.trace_frame()
            ms = []
            for _ in range(a.frames):
                r.trace_frame()
                ms.append(r.stats().msTotal)
            st = r.stats()
            rec["frame_ms"] = round(min(ms), 4)
            rec["frame"] = {"waves": st.waves, "trace_ms": round(st.msTrace, 4), "shade_ms": round(st.msShade, 4), "tail_ms": round(st.msTail, 4),
                            "rays": int(st.primaryRays + st.extensionRays + st.shadowRays)}
            # instrumented frame
            r.set_traversal_stats(True)
            r.trace_frame()
            st = r.stats()
            r.set_traversal_stats(False)
            rays = max(1, st.primaryRays + st.extensionRays + st.shadowRays)
            rec["per_ray"] = {"nodes": round(st.nodeVisits / rays, 2), "tlas_nodes": round(st.tlasVisits / rays, 2), "inst": round(st.instanceEntries / rays, 2),
                              "tris": round(st.triTests / rays, 2), "max_stack": int(st.maxStack)}
            if cfg.get("trace_variant", 0) == 1:
                it, ln = list(st.schedIters), list(st.schedLanes)
                rec["sched"] = {"lanes_per_step": [round(ln[k] / max(1, it[k]), 2) for k in range(3)], "steps": it,
                                "rays_per_refill": round(st.schedRefillLanes / max(1, st.schedRefills), 2),
                                "alive_at_vote": round(st.schedPopLanes / max(1, st.schedPopIters), 2), "votes": int(st.schedPopIters)}
            line = json.dumps(rec)
            print(line, flush=True)
            if out:
                out.write(line + "\n")
                out.flush()
        r.close()


if __name__ == "__main__":
    main()
