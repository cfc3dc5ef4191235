"""How much does reading the final image back cost per frame: none / synchronous / asynchronous (overlapped)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gknextrenderer_b200 as gk
from bench import WORKLOADS

scene, args, W, H, settings = WORKLOADS["room"]
eng = gk.Engine(scene, *args); eng.set(**settings)
r = gk.Renderer(W, H, device=0); r.load(eng)
nbytes = r.plane_bytes("DENOISED")
bufs = [torch.empty((H, W, 4), dtype=torch.float16).pin_memory() for _ in range(2)]
out = np.empty((H, W, 4), np.float16)

def run(mode, n=20):
    for _ in range(3):
        r.set_ubo(eng.ubo(W, H)); r.render_frame(); eng.advance_frame()
    r.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter(); dev = 0.0; parts = np.zeros(4)
    for i in range(n):
        r.set_ubo(eng.ubo(W, H)); r.render_frame(); eng.advance_frame()
        if mode == "sync":
            r._check(r.lib.gk_readback(r.h, gk.PLANES["DENOISED"], bufs[0].data_ptr(), nbytes))
        elif mode == "async":
            r.readback_async("DENOISED", bufs[i & 1].data_ptr(), nbytes)
        st = r.stats(); dev += st.msTotal; parts += [st.msExtend, st.msShadow, st.msShade, st.msTail]
    if mode == "async":
        r.readback_wait()
    r.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / n
    print(f"{mode:6s}: wall {ms:.3f} ms/frame, device frame {dev / n:.3f} ms, ext/shd/shade/tail {np.round(parts / n, 3)}")

for m in ("none", "sync", "async", "none", "async"):
    run(m)
