"""Profiling range around instance updates of the brick workload (TLAS refit / rebuild + one frame):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/prof_update.py bricks 2
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gknextrenderer_b200 as gk  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "bricks"
    updates = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    scene, args, W, H, settings = WORKLOADS[name]
    eng = gk.Engine(scene, *args)
    eng.set(**settings)
    r = gk.Renderer(W, H, device=0)
    r.load(eng)
    cudart = ctypes.CDLL(None)
    r.set_ubo(eng.ubo(W, H)); r.render_frame(); r.synchronize()
    cudart.cudaProfilerStart()
    for f in range(updates):
        eng.step_scene(f + 1)
        t0 = time.perf_counter()
        nodes, n = eng.update_nodes()
        t1 = time.perf_counter()
        r.update_instances(nodes, n, refit=True)
        t2 = time.perf_counter()
        r.set_ubo(eng.ubo(W, H)); r.render_frame(); r.synchronize()
        t3 = time.perf_counter()
        info = r.bvh_info()
        st = r.stats()
        print(f"update {f}: host UpdateNodes {1e3 * (t1 - t0):.2f} ms, gk_update_instances wall {1e3 * (t2 - t1):.2f} ms (device: build {info.msTlasBuild:.3f} refit {info.msRefit:.3f}, "
              f"rejected {info.refitsRejected}), frame wall {1e3 * (t3 - t2):.2f} ms device {st.msTotal:.3f} ext {st.msExtend:.3f} shd {st.msShadow:.3f} tail {st.msTail:.3f}")
    cudart.cudaProfilerStop()


if __name__ == "__main__":
    main()
