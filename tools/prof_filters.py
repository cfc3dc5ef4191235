"""A few filter frames at one resolution for ncu: python tools/prof_filters.py [W H]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import gknextrenderer_b200 as gk
from gknextrenderer_b200 import GkUniformBufferObject
from bench_filters import synth
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
r = gk.Renderer(W, H, device=0)
for name, arr in synth(W, H, W + H).items():
    r.upload_plane(name, arr)
u = GkUniformBufferObject()
u.ViewportRect[:] = [0, 0, W, H]
u.TemporalFrames, u.TotalFrames, u.BFSize = 16, 5, 5
u.BFSigma, u.BFSigmaLum, u.PaperWhiteNit = 2.0, 3.0, 600.0
u.SelectedId = 0xFFFFFFFF
r.set_ubo(u)
for i in range(4):
    r.filter_frame()
st = r.stats()
print("reproject", st.msReproject, "jbf", st.msDenoise)
