"""One profiled frame of a workload: warm-up frames run outside the profiling range, then exactly
one frame inside cudaProfilerStart/Stop, so `ncu --profile-from-start off -k regex:<kernel> -c N`
captures the first N launches of that kernel in wave order (primary wave first).

    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:k_extend -c 3 -o gpurun_out/prof python tools/prof_frame.py room
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gknextrenderer_b200 as gk  # noqa: E402

sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "room"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    scene, args, W, H, settings = WORKLOADS[name]
    eng = gk.Engine(scene, *args)
    eng.set(**settings)
    tiles = int(os.environ.get("GK_PROF_TILES", "1"))  # >1: the per-rank share of a tile-partitioned frame
    r = gk.Renderer(W, H, device=0, tile_index=0, tile_count=tiles, tile_rows=16)
    r.load(eng)
    cudart = ctypes.CDLL(None)  # libcudart is already in the process (pulled in by libgknext_cuda.so, RTLD_GLOBAL)
    for f in range(3):
        r.set_ubo(eng.ubo(W, H)); r.render_frame(); eng.advance_frame()
    r.synchronize()
    cudart.cudaProfilerStart()
    for f in range(frames):
        r.set_ubo(eng.ubo(W, H)); r.render_frame(); eng.advance_frame()
    r.synchronize()
    cudart.cudaProfilerStop()
    st = r.stats()
    print("frame ms", st.msTotal, "waves", st.waves, "rays", st.primaryRays + st.extensionRays + st.shadowRays)


if __name__ == "__main__":
    main()
