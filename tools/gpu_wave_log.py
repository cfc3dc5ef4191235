"""Per-wave device times of one frame (GK_WAVE_LOG): python tools/gpu_wave_log.py [workload]"""
import os, sys
os.environ["GK_WAVE_LOG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gknextrenderer_b200 as gk
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "room"
scene, args, W, H, settings = WORKLOADS[name]
eng = gk.Engine(scene, *args); eng.set(**settings)
opts = dict(a.split("=") for a in sys.argv[2:])
tiles = int(opts.pop("tiles", 1))  # >1: the share one rank of a tile-partitioned frame traces
rows, index = int(opts.pop("tile_rows", 16)), int(opts.pop("tile_index", 0))
r = gk.Renderer(W, H, device=0, tile_index=index, tile_count=tiles, tile_rows=rows); r.load(eng)
for k, v in opts.items():
    r.set_option(k, float(v))
os.environ.pop("GK_WAVE_LOG")
for f in range(3):
    r.set_ubo(eng.ubo(W, H)); r.trace_frame(); eng.advance_frame()
os.environ["GK_WAVE_LOG"] = "1"
r.set_ubo(eng.ubo(W, H)); r.trace_frame()
st = r.stats()
print("frame ms", st.msTotal, "waves", st.waves, "rays", st.primaryRays + st.extensionRays + st.shadowRays, "trace", st.msTrace, "shade", st.msShade)
