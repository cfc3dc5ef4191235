"""Phase times of the BLAS build (GK_BUILD_LOG): python tools/gpu_build_log.py [workload]; the scene is uploaded three times
(first = with allocations, then warm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gknextrenderer_b200 as gk
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "roomu"
scene, args, W, H, settings = WORKLOADS[name]
eng = gk.Engine(scene, *args); eng.set(**settings)
r = gk.Renderer(W, H, device=0)
for k, v in [a.split("=") for a in sys.argv[2:]]:
    r.set_option(k, float(v))
for rep in range(3):
    os.environ["GK_BUILD_LOG"] = "1"
    r.upload_scene(eng.scene_desc())
    os.environ.pop("GK_BUILD_LOG")
    info = r.bvh_info()
    print(f"upload {rep}: with log {info.msBlasBuild:.3f} ms", flush=True)
    r.upload_scene(eng.scene_desc())
    info = r.bvh_info()
    print(f"upload {rep}: blas build {info.msBlasBuild:.3f} ms, {info.triangleCount} tris, {info.blasNodes8} wide nodes", flush=True)
