"""Debug: any-hit gk_intersect_device on captured extension rays (room)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gknextrenderer_b200._native as N
if os.environ.get("GK_LIB_DIR"):
    N.CUDA_LIB_PATH = os.path.join(os.environ["GK_LIB_DIR"], "libgknext_cuda.so")
    N.HOST_LIB_PATH = os.path.join(os.environ["GK_LIB_DIR"], "libgknext_host.so")
    # round-1 library: no gk_set_option, shorter GkFrameStats (we do not read stats here)
    N.CUDA_API.pop("gk_set_option", None)
import numpy as np, torch
import gknextrenderer_b200 as gk
from bench import WORKLOADS
scene, args, W, H, settings = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "room"]
eng = gk.Engine(scene, *args); eng.set(**settings)
r = gk.Renderer(W, H, device=0); r.load(eng); r.set_ubo(eng.ubo(W, H))
r.set_ray_capture(int(sys.argv[2]) if len(sys.argv) > 2 else 0); r.trace_frame()
rays = r.captured_rays(W * H).copy(); r.set_ray_capture(-1)
d = torch.from_numpy(rays).cuda(); n = len(rays)
tuv = torch.empty((n, 3), dtype=torch.float32, device="cuda"); ids = torch.empty((n, 2), dtype=torch.int32, device="cuda")
for any_hit in (False, True):
    r.intersect_device(d.data_ptr(), n, tuv.data_ptr(), ids.data_ptr(), any_hit)
    r.synchronize()
    print("ok", any_hit, int((ids[:, 0] != -1).sum()) if not any_hit else int(ids[:, 0].sum()))
