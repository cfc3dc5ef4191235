"""C5: reprojection + JBF denoise alone on synthetic G-buffers, 720p .. 8K.

Prints one JSON line per resolution: ms of k_reproject / k_denoise_jbf (CUDA events inside the library),
algorithmic GB/s (96 B/px and 48 B/px) and the fraction of the measured HBM peak.

    python tools/bench_filters.py [reps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gknextrenderer_b200 as gk  # noqa: E402
from gknextrenderer_b200 import GkUniformBufferObject  # noqa: E402

RES = [(1280, 720), (1920, 1080), (2560, 1440), (3840, 2160), (7680, 4320)]


def synth(W, H, seed):
    """radiance ~ LogNormal(0,1) RGBA16F, normals from a 64-px checker of planes, ids = checker index, motion U[-2,2] px"""
    rng = np.random.default_rng(seed)
    f16 = np.float16
    cell = 64
    yy, xx = np.mgrid[0:H, 0:W]
    ids = ((yy // cell) * ((W + cell - 1) // cell) + (xx // cell)).astype(np.uint32) % 65000
    tab = rng.normal(size=(4096, 3)).astype(np.float32)
    tab /= np.linalg.norm(tab, axis=1, keepdims=True)
    normal = np.zeros((H, W, 4), f16)
    normal[..., :3] = tab[ids % 4096]
    normal[..., 3] = 0.5

    def logn(scale=1.0):
        return (scale * np.exp(rng.standard_normal((H, W, 4), dtype=np.float32))).astype(f16)

    motion = rng.uniform(-2, 2, (H, W, 2)).astype(np.float32)
    motion[rng.uniform(size=(H, W)) < 0.5] = 0.0
    return {"OUTPUT_DIFFUSE": logn(), "OUTPUT_SPECULAR": logn(0.3), "ALBEDO": rng.uniform(0.05, 1, (H, W, 4)).astype(f16), "NORMAL": normal,
            "HISTORY_DIFFUSE": logn(), "HISTORY_SPECULAR": logn(0.3), "HISTORY_ALBEDO": rng.uniform(0.05, 1, (H, W, 4)).astype(f16),
            "OBJECT_ID0": ids, "OBJECT_ID1": ids.copy(), "MOTION": motion}


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    for W, H in RES:
        r = gk.Renderer(W, H, device=0)
        for name, arr in synth(W, H, W + H).items():
            r.upload_plane(name, arr)
        u = GkUniformBufferObject()
        u.ViewportRect[:] = [0, 0, W, H]
        u.TemporalFrames, u.TotalFrames, u.BFSize = 16, 5, 5
        u.BFSigma, u.BFSigmaLum, u.PaperWhiteNit = 2.0, 3.0, 600.0
        u.SelectedId = 0xFFFFFFFF
        r.set_ubo(u)
        rep, jbf = [], []
        for i in range(reps + 3):
            r.filter_frame()
            st = r.stats()
            if i >= 3:
                rep.append(st.msReproject), jbf.append(st.msDenoise)
        px = W * H
        rep_ms, jbf_ms = float(np.median(rep)), float(np.median(jbf))
        line = {"workload": f"C5 filters {W}x{H}", "reproject_ms": round(rep_ms, 4), "denoise_ms": round(jbf_ms, 4),
                "reproject_GBs": round(96 * px / rep_ms / 1e6, 1), "denoise_GBs": round(48 * px / jbf_ms / 1e6, 1),
                "both_frac_of_hbm_peak": round(144 * px / (rep_ms + jbf_ms) / 1e6 / hbm, 4), "hbm_peak_GBs": hbm,
                "Gpx_per_s": round(px / (rep_ms + jbf_ms) / 1e6, 2)}
        print(json.dumps(line), flush=True)
        r.close()


if __name__ == "__main__":
    main()
