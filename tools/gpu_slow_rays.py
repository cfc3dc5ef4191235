"""Finds the rays of the late (tail) waves and measures their individual traversal cost."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gknextrenderer_b200 as gk

W, H = 1920, 1080
eng = gk.Engine("room", 1000000, 1234)
eng.set(TAA=0, NumberOfSamples=1, NumberOfBounces=4, Denoiser=1, TemporalFrames=16)
r = gk.Renderer(W, H, device=0)
r.load(eng)
ubo = eng.ubo(W, H); r.set_ubo(ubo)
for wave in (9, 12, 15):
    r.set_ray_capture(wave); r.trace_frame()
    rays = r.captured_rays(4096)
    print("wave", wave, "rays", len(rays))
    r.set_traversal_stats(True)
    res = []
    for i in range(min(len(rays), 64)):
        r.set_traversal_stats(True)
        t0 = time.perf_counter(); tuv, ids = r.intersect(rays[i:i + 1]); dt = time.perf_counter() - t0
        st = r.stats()
        res.append((st.nodeVisits, st.triTests, dt * 1e6, i))
    r.set_traversal_stats(False)
    res.sort(reverse=True)
    for nv, nt, us, i in res[:6]:
        print("   visits", nv, "tris", nt, "host us", round(us), "ray", rays[i])
