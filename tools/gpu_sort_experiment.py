"""Does sorting the bounce rays (direction octant + origin Morton cell) pay for the extend kernel?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gknextrenderer_b200 as gk
from bench import WORKLOADS

scene, args, W, H, settings = WORKLOADS["room"]
eng = gk.Engine(scene, *args); eng.set(**settings)
r = gk.Renderer(W, H, device=0); r.load(eng)
ubo = eng.ubo(W, H); r.set_ubo(ubo)
stream = torch.cuda.ExternalStream(r.stream())

def morton3(q):
    def spread(v):
        v = v.astype(np.uint32) & 0x3ff
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])

def time_rays(rays, label, reps=5):
    d = torch.from_numpy(rays).cuda()
    tuv = torch.empty((len(rays), 3), dtype=torch.float32, device="cuda"); ids = torch.empty((len(rays), 2), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); r.intersect_device(d.data_ptr(), len(rays), tuv.data_ptr(), ids.data_ptr()); e1.record(stream)
        r.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"{label:40s} {len(rays):8d} rays  {best:7.3f} ms  {len(rays)/best/1e6:7.3f} Grays/s")
    return best

for wave in (1, 2):
    r.set_ray_capture(wave); r.trace_frame()
    rays = r.captured_rays(W * H).copy()
    time_rays(rays, f"wave {wave} as queued")
    o, d = rays[:, 0:3], rays[:, 4:7]
    octant = ((d[:, 0] < 0).astype(np.uint32) << 2) | ((d[:, 1] < 0).astype(np.uint32) << 1) | (d[:, 2] < 0).astype(np.uint32)
    lo, hi = o.min(0), o.max(0)
    for bits in (4, 6, 8):
        q = np.clip(((o - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
        key = (octant.astype(np.uint64) << (3 * bits)) | morton3(q).astype(np.uint64)
        order = np.argsort(key, kind="stable")
        time_rays(np.ascontiguousarray(rays[order]), f"wave {wave} sorted octant+morton {bits}b/axis")
    order = np.argsort(octant, kind="stable")
    time_rays(np.ascontiguousarray(rays[order]), f"wave {wave} sorted by octant only")
    rng = np.random.default_rng(0)
    time_rays(np.ascontiguousarray(rays[rng.permutation(len(rays))]), f"wave {wave} shuffled")
