"""Per-rank path-tracer time of a tile-partitioned frame, measured on ONE GPU (a context with tileCount = T
traces 1/T of the rows): how do launch parameters behave at the wave sizes of 4- and 8-GPU runs?
    python tools/gpu_tile_experiment.py [workload] [ENV_NAME v1 v2 ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import gknextrenderer_b200 as gk
from bench import WORKLOADS

scene, args, W, H, settings = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "room"]
env_name = sys.argv[2] if len(sys.argv) > 2 else "GK_TRACE_BLOCK"
values = sys.argv[3:] if len(sys.argv) > 3 else ["256", "128"]
eng = gk.Engine(scene, *args); eng.set(**settings)
for tiles in (1, 4, 8):
    for v in values:
        os.environ[env_name] = v
        r = gk.Renderer(W, H, device=0, tile_index=0, tile_count=tiles, tile_rows=16)
        r.load(eng)
        ms, parts = [], np.zeros(5)
        for f in range(8):
            r.set_ubo(eng.ubo(W, H)); r.trace_frame(); st = r.stats(); eng.advance_frame()
            if f >= 3:
                ms.append(st.msTotal); parts += [st.msExtend, st.msShadow, st.msShade, st.msTail, st.waves]
        print(f"tiles {tiles} {env_name}={v}: trace {np.mean(ms):.3f} ms  ext/shd/shade/tail/waves {np.round(parts / len(ms), 3)}", flush=True)
        r.close()
