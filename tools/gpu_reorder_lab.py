"""How much does the ORDER of a wave's rays matter to the traversal kernel?  Captures bounce waves, times gk_intersect_device on the
same rays in several orders (as emitted, octant-binned within blocks, globally sorted, shuffled).

    python tools/gpu_reorder_lab.py [workload ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gknextrenderer_b200 as gk  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def octant(d):
    return ((d[:, 4] < 0).long() | ((d[:, 5] < 0).long() << 1) | ((d[:, 6] < 0).long() << 2))


def morton(o, bits=7):
    lo, hi = o.min(0).values, o.max(0).values
    q = ((o - lo) / (hi - lo + 1e-20) * ((1 << bits) - 1)).long()
    code = torch.zeros(o.shape[0], dtype=torch.long, device=o.device)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return code


def block_sort(key, block):
    n = key.shape[0]
    blk = torch.arange(n, device=key.device) // block
    return torch.argsort(blk * (int(key.max().item()) + 1) + key, stable=True)


def orders(d):
    n = d.shape[0]
    oc = octant(d)
    mo = morton(d[:, 0:3])
    yield "as emitted", torch.arange(n, device=d.device)
    for block in (128, 256, 1024, 4096):
        yield f"octant within {block}", block_sort(oc, block)
    yield "octant+major-axis within 256", block_sort(oc * 3 + d[:, 4:7].abs().argmax(1), 256)
    yield "global octant, then origin morton", torch.argsort(oc * (1 << 21) + mo, stable=True)
    yield "global origin morton, then octant", torch.argsort(mo * 8 + oc, stable=True)
    yield "origin morton(4 bits/axis) x octant", torch.argsort(morton(d[:, 0:3], 4) * 8 + oc, stable=True)
    yield "shuffled", torch.randperm(n, device=d.device)


def main():
    for wl in sys.argv[1:] or ["room"]:
        scene, args, W, H, settings = WORKLOADS[wl]
        eng = gk.Engine(scene, *args)
        eng.set(**settings)
        r = gk.Renderer(W, H, device=0)
        r.load(eng)
        r.set_ubo(eng.ubo(W, H))
        stream = torch.cuda.ExternalStream(r.stream())
        waves = []
        for wave in (1, 2):
            r.set_ray_capture(wave)
            r.trace_frame()
            waves.append(torch.from_numpy(r.captured_rays(W * H).copy()).cuda())
            r.set_ray_capture(-1)
        for wi, d in enumerate(waves):
            n = d.shape[0]
            tuv = torch.empty((n, 3), dtype=torch.float32, device="cuda")
            ids = torch.empty((n, 2), dtype=torch.int32, device="cuda")
            for name, perm in orders(d):
                rays = d[perm].contiguous()
                torch.cuda.synchronize()
                best = 1e9
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    r.intersect_device(rays.data_ptr(), n, tuv.data_ptr(), ids.data_ptr(), False)
                    e1.record(stream)
                    r.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                print(f"{wl} wave {wi + 1} {n:8d} rays  {name:38s} {best:7.3f} ms  {n / best / 1e6:6.3f} Grays/s", flush=True)


if __name__ == "__main__":
    main()
