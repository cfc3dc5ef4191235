"""Dumps the rays whose GPU hit differs from the real tinybvh's (ids or distance) for offline analysis.
    python tools/gpu_dump_mismatches.py bricks out.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gknextrenderer_b200 as gk, oracle_lib as ol
name, out = sys.argv[1], sys.argv[2]
cfg = {"bricks": ("bricks", (200000, 42), 1920, 1080, 5, (-20, 0.0, -20), (20, 3.0, 20)), "city": ("city", (40, 100, 7, 46), 3840, 2160, 7, (-200, 0.5, -200), (200, 60, 200))}[name]
scene, args, W, H, stride, lo, hi = cfg
eng = gk.Engine(scene, *args); eng.set(TAA=0)
r = gk.Renderer(W, H, device=0); r.upload_scene(eng.scene_desc())
eng.update_nodes(); nodes, n = eng.update_nodes(); r.update_instances(nodes, n)
ref = ol.OracleScene(eng.scene_desc(), nodes, n, use_ref=True)
rng = np.random.default_rng(23)
o = rng.uniform(lo, hi, (300000, 3)).astype(np.float32); d = rng.normal(size=(300000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
inc = np.zeros((300000, 8), np.float32); inc[:, 0:3], inc[:, 3], inc[:, 4:7], inc[:, 7] = o, 0.0, d, 1000.0
rays = np.concatenate([np.ascontiguousarray(ol.primary_rays(eng.ubo(W, H), W, H)[::stride]), inc])
rays[:, 3] = 0.0
g_tuv, g_ids = r.intersect(rays)
t_tuv, t_ids = ref.intersect(rays, threads=os.cpu_count() or 1)
differ = (g_ids != t_ids).any(axis=1) | (g_tuv.view(np.uint32) != t_tuv.view(np.uint32)).any(axis=1)
idx = np.nonzero(differ)[0]
print(name, "rays", len(rays), "differ", len(idx))
np.savez_compressed(out, idx=idx, rays=rays[idx], g_tuv=g_tuv[idx], g_ids=g_ids[idx], t_tuv=t_tuv[idx], t_ids=t_ids[idx])
