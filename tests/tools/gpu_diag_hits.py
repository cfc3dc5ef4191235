"""Classifies hit mismatches between the CUDA traversal, the oracle BVH and brute force."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gknextrenderer_b200 as gk
import oracle_lib as ol

def bits(a): return np.ascontiguousarray(a).view(np.uint32)

def run(scene, args, rays, label):
    eng = gk.Engine(scene, *args); eng.set(TAA=0)
    r = gk.Renderer(64, 64, device=0)
    r.upload_scene(eng.scene_desc())
    eng.update_nodes(); nodes, n = eng.update_nodes()
    r.update_instances(nodes, n)
    orc = ol.OracleScene(eng.scene_desc(), nodes, n)
    g_tuv, g_ids = r.intersect(rays)
    o_tuv, o_ids = orc.intersect(rays, threads=8)
    differ = (g_ids != o_ids).any(axis=1)
    hard = differ & (bits(g_tuv[:, 0]) != bits(o_tuv[:, 0]))
    print(label, "rays", len(rays), "differ", int(differ.sum()), "hard", int(hard.sum()))
    idx = np.nonzero(hard)[0][:12]
    if len(idx):
        b_tuv, b_ids = orc.intersect_bruteforce(rays[idx])
        for k, i in enumerate(idx):
            print("  ray", i, rays[i])
            print("     gpu  ", g_ids[i], g_tuv[i])
            print("     orc  ", o_ids[i], o_tuv[i])
            print("     brute", b_ids[k], b_tuv[k], "-> gpu==brute" if bits(b_tuv[k, 0]) == bits(g_tuv[i, 0]) else ("-> orc==brute" if bits(b_tuv[k, 0]) == bits(o_tuv[i, 0]) else "-> neither"))

g = np.load(os.path.join(ROOT, "tests", "golden", "cornell_tinybvh.npz"))
run("cornell", (), g["rays"], "cornell golden")
g = np.load(os.path.join(ROOT, "tests", "golden", "room20k_tinybvh.npz"))
run("room", (20000, 99), g["rays"], "room20k golden")
rng = np.random.default_rng(11)
def rr(n, lo, hi):
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32); d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros((n, 8), np.float32); r[:, 0:3], r[:, 3], r[:, 4:7], r[:, 7] = o, 1e-3, d, 1000.0; return r
run("cornell", (), rr(200000, (-2.7, 0.05, -2.7), (2.7, 5.5, 2.7)), "cornell incoherent")
run("room", (60000, 5), rr(300000, (-9.5, 0.1, -9.5), (9.5, 3.9, 9.5)), "room60k incoherent")
