"""Developer diagnostic run on the GPU box (not a test): prints parity and timing details."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gknextrenderer_b200 as gk
import oracle_lib as ol

def main():
    W, H = 640, 360
    eng = gk.Engine("cornell")
    eng.set(TAA=0, NumberOfSamples=8, NumberOfBounces=4)
    r = gk.Renderer(W, H, device=0)
    t0 = time.time(); r.load(eng); print("load s", time.time() - t0)
    info = r.bvh_info()
    print("bvh: blas", info.blasCount, "inst", info.instanceCount, "tris", info.triangleCount, "nodes8", info.blasNodes8, info.tlasNodes8,
          "ms blas/tlas", info.msBlasBuild, info.msTlasBuild)
    ubo = eng.ubo(W, H)
    rays = ol.primary_rays(ubo, W, H)
    nodes, n = eng.update_nodes()
    orc = ol.OracleScene(eng.scene_desc(), nodes, n)
    o_tuv, o_ids = orc.intersect(rays, threads=8)
    g_tuv, g_ids = r.intersect(rays)
    bad = (g_ids != o_ids).any(axis=1)
    print("primary rays:", len(rays), "id mismatches:", int(bad.sum()), "t bit-equal:", int((g_tuv[:, 0].view(np.uint32) == o_tuv[:, 0].view(np.uint32)).sum()))
    if bad.any():
        idx = np.nonzero(bad)[0][:10]
        for i in idx: print("  ", i, g_ids[i], o_ids[i], g_tuv[i], o_tuv[i])
    r.set_ubo(ubo)
    r.set_traversal_stats(True)
    r.trace_frame()
    st = r.stats()
    print("trace: ms", st.msTotal, "waves", st.waves, "launches", st.launches, "rays p/e/s", st.primaryRays, st.extensionRays, st.shadowRays,
          "ms gen/ext/shd/shade/acc", st.msGenerate, st.msExtend, st.msShadow, st.msShade, st.msAccumulate, "visits", st.nodeVisits, st.triTests)
    r.set_traversal_stats(False)
    t0 = time.time(); o = orc.render(ubo, W, H, threads=8); print("oracle render s", time.time() - t0)
    ids = r.readback("PRIMARY_IDS")
    print("frame primary id mismatches", int((ids != o["primIds"]).any(axis=2).sum()))
    for name, key in (("RADIANCE_DIFFUSE_F32", "diffuse"), ("RADIANCE_SPECULAR_F32", "spec")):
        g = r.readback(name)[..., :3]; ref = o[key][..., :3]
        rel = np.sqrt(np.mean((g - ref) ** 2)) / (np.sqrt(np.mean(ref ** 2)) + 1e-12)
        print(name, "mean gpu/orc", g.mean(), ref.mean(), "relRMSE", rel, "max abs", np.abs(g - ref).max(), "exact px", int((g == ref).all(axis=2).sum()), "/", W * H)
    print("rayCount equal px", int((r.readback("RAY_COUNT") == o["rayCount"]).sum()), "sum gpu/orc", int(r.readback("RAY_COUNT").sum()), int(o["rayCount"].sum()))
    print("objectId equal", bool((r.readback("OBJECT_ID0") == o["objectId"]).all()))
    alb = r.readback("ALBEDO").astype(np.float32); print("albedo max diff", np.abs(alb - o["albedo"]).max())
    nrm = r.readback("NORMAL").astype(np.float32); print("normal max diff", np.abs(nrm - o["normal"]).max())
    print("motion max diff", np.abs(r.readback("MOTION") - o["motion"]).max(), "depth max diff", np.abs(r.readback("DEPTH") - o["depth"]).max())
    # timing: a few frames
    for f in range(3):
        r.trace_frame(); st = r.stats(); print("frame", f, "ms", st.msTotal, "ext", st.msExtend, "shade", st.msShade)
    total = st.primaryRays + st.extensionRays + st.shadowRays
    print("Mrays/s", total / st.msTotal / 1e3)
    # filters
    eng.set(Denoiser=1); ubo2 = eng.ubo(W, H); r.set_ubo(ubo2); r.render_frame(); st = r.stats(); print("render_frame ms", st.msTotal, "reproject", st.msReproject, "denoise", st.msDenoise)
    fin = r.readback("DENOISED").astype(np.float32); print("denoised mean", fin[..., :3].mean(), "finite", bool(np.isfinite(fin).all()))



def diag2():
    W, H = 640, 360
    eng = gk.Engine("cornell")
    eng.set(TAA=0, NumberOfSamples=2, NumberOfBounces=4, Denoiser=1)
    r = gk.Renderer(W, H, device=0)
    r.load(eng)
    ubo = eng.ubo(W, H)
    nodes, n = eng.update_nodes()
    orc = ol.OracleScene(eng.scene_desc(), nodes, n)
    o = orc.render(ubo, W, H, threads=8)
    r.set_ubo(ubo); r.render_frame()
    m = r.readback("MOTION"); d = np.abs(m - o["motion"])
    print("motion: mismatching px", int((d > 0).any(axis=2).sum()), "max", d.max(), "gpu absmax", np.abs(m).max(), "orc absmax", np.abs(o["motion"]).max(),
          "nan gpu/orc", int(np.isnan(m).sum()), int(np.isnan(o["motion"]).sum()))
    bad = np.argwhere((d > 0).any(axis=2))[:5]
    for y, x in bad: print("   px", x, y, m[y, x], o["motion"][y, x])
    for p in ("OUTPUT_DIFFUSE", "OUTPUT_SPECULAR", "ALBEDO", "NORMAL", "ACCUM_DIFFUSE", "ACCUM_SPECULAR", "ACCUM_ALBEDO", "HISTORY_DIFFUSE", "DENOISED"):
        a = r.readback(p).astype(np.float32)
        print(p, "nan", int(np.isnan(a).sum()), "inf", int(np.isinf(a).sum()), "min", np.nanmin(a), "max", np.nanmax(a))
    a = r.readback("DENOISED").astype(np.float32)
    bad = np.argwhere(np.isnan(a).any(axis=2))[:5]
    for y, x in bad: print("   nan px", x, y, a[y, x], r.readback("ACCUM_DIFFUSE")[y, x], r.readback("ACCUM_SPECULAR")[y, x], r.readback("ACCUM_ALBEDO")[y, x])


def room():
    W, H = 1920, 1080
    t0 = time.time(); eng = gk.Engine("room"); print("room build s", time.time() - t0, "tris", eng.triangles(), eng.triangles(True))
    eng.set(TAA=0, NumberOfSamples=1, NumberOfBounces=4, Denoiser=1, TemporalFrames=16)
    r = gk.Renderer(W, H, device=0)
    t0 = time.time(); r.load(eng); print("load s", time.time() - t0)
    info = r.bvh_info()
    print("bvh: blas", info.blasCount, "inst", info.instanceCount, "tris", info.triangleCount, "instanced", info.instancedTriangles, "nodes8", info.blasNodes8, info.tlasNodes8,
          "ms blas/tlas", info.msBlasBuild, info.msTlasBuild, "bytes", info.bytesGeometry, info.bytesBvh)
    for f in range(4):
        ubo = eng.ubo(W, H); r.set_ubo(ubo)
        if f == 3: r.set_traversal_stats(True)
        r.render_frame(); st = r.stats(); eng.advance_frame()
        total = st.primaryRays + st.extensionRays + st.shadowRays
        print("frame", f, "ms", st.msTotal, "waves", st.waves, "rays", st.primaryRays, st.extensionRays, st.shadowRays, "ext", st.msExtend, "shd", st.msShadow, "shade", st.msShade,
              "repro", st.msReproject, "jbf", st.msDenoise, "Mrays/s", total / st.msTotal / 1e3, "visits/ray", st.nodeVisits / max(total, 1), st.triTests / max(total, 1))
    # parity of primary ids on the big scene against the oracle
    nodes, n = eng.update_nodes()
    t0 = time.time(); orc = ol.OracleScene(eng.scene_desc(), nodes, n); print("oracle build s", time.time() - t0)
    rays = ol.primary_rays(eng.ubo(W, H), W, H)[::7]
    t0 = time.time(); o_tuv, o_ids = orc.intersect(rays, threads=8); print("oracle intersect s", time.time() - t0, "Mrays/s", len(rays) / orc.last_seconds / 1e6)
    g_tuv, g_ids = r.intersect(rays)
    bad = (g_ids != o_ids).any(axis=1)
    print("room primary sample:", len(rays), "id mismatches", int(bad.sum()), "t bit-equal", int((g_tuv[:, 0].view(np.uint32) == o_tuv[:, 0].view(np.uint32)).sum()))
    for i in np.nonzero(bad)[0][:8]: print("   ", i, g_ids[i], o_ids[i], g_tuv[i], o_tuv[i])


if __name__ == "__main__":
    for a in (sys.argv[1:] or ["main"]):
        {"main": main, "diag2": diag2, "room": room}[a]()
