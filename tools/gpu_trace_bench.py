"""Traversal micro-benchmark: captures the ray buffers of waves 0..2 of one frame of a workload and
times gk_intersect_device on them (closest hit and any hit), plus the per-ray traversal counters.

    python tools/gpu_trace_bench.py [workload] [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gknextrenderer_b200 as gk  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "room"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    scene, args, W, H, settings = WORKLOADS[name]
    eng = gk.Engine(scene, *args)
    eng.set(**settings)
    r = gk.Renderer(W, H, device=0)
    r.load(eng)
    r.set_ubo(eng.ubo(W, H))
    info = r.bvh_info()
    print(f"[{name}] blas nodes2/8 {info.blasNodes2}/{info.blasNodes8} tlas nodes2/8 {info.tlasNodes2}/{info.tlasNodes8} instances {info.instanceCount} "
          f"tris {info.triangleCount} build ms blas {info.msBlasBuild:.3f} tlas {info.msTlasBuild:.3f}")
    stream = torch.cuda.ExternalStream(r.stream())

    # whole-frame counters
    r.set_traversal_stats(True)
    r.trace_frame()
    st = r.stats()
    rays = st.primaryRays + st.extensionRays + st.shadowRays
    print(f"frame: rays {rays} waves {st.waves} per ray: node visits {st.nodeVisits / rays:.2f} (tlas {st.tlasVisits / rays:.2f}) "
          f"instance entries {st.instanceEntries / rays:.2f} tri tests {st.triTests / rays:.2f}")
    r.set_traversal_stats(False)
    for _ in range(3):
        r.trace_frame()
    st = r.stats()
    print(f"frame ms {st.msTotal:.3f} gen {st.msGenerate:.3f} ext {st.msExtend:.3f} shd {st.msShadow:.3f} shade {st.msShade:.3f} acc {st.msAccumulate:.3f}")

    for wave in (0, 1, 2):
        r.set_ray_capture(wave)
        r.trace_frame()
        buf = r.captured_rays(W * H).copy()
        r.set_ray_capture(-1)
        d = torch.from_numpy(buf).cuda()
        n = len(buf)
        tuv = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        ids = torch.empty((n, 2), dtype=torch.int32, device="cuda")
        for any_hit in (False, True):
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                r.intersect_device(d.data_ptr(), n, tuv.data_ptr(), ids.data_ptr(), any_hit)
                e1.record(stream)
                r.synchronize()
                best = min(best, e0.elapsed_time(e1))
            hits = int((ids[:, 0] != -1).sum().item()) if not any_hit else -1
            print(f"wave {wave} {'any' if any_hit else 'closest'}: {n} rays {best:.3f} ms {n / best / 1e6:.3f} Grays/s hits {hits}")


if __name__ == "__main__":
    main()
