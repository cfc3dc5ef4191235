"""profiles/<round>_bench/n<N>_<workload>.json -> profiles/<round>_scaling.md      python tools/scaling_table.py [r02]"""
import glob
import json
import os
import sys

ROUND = sys.argv[1] if len(sys.argv) > 1 else "r01"

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", ROUND + "_bench", "n*_*.json"))):
    d = json.load(open(f))
    if d.get("impl") == "reference":
        continue
    mode = "frames" if d.get("scaling") == "weak" else "tiles"
    rows.setdefault((d["config"]["workload"].split(":")[0], mode), {})[d["n_gpus"]] = d
out = [f"# Round {int(ROUND[1:])}: throughput and scaling on B200 (bench.py, 20 timed frames after 5 warm-up frames)\n",
       "`value` = rays traced by all ranks / max-over-ranks device time; `e2e` adds the UBO upload and the read-back of the final image on the presenting rank.",
       "Speed-ups are against the 1-GPU line of the same workload; the driver computes its own from the per-N values.\n"]
for (wl, mode), by_n in sorted(rows.items()):
    base = by_n.get(1) or rows.get((wl, "tiles"), {}).get(1)
    title = base['config']['workload'] if base else wl
    if mode == "frames":
        out.append(f"## {title} - FRAME-SHARDED (a step is N consecutive progressive frames, one per rank: weak scaling)\n")
    else:
        out.append(f"## {title} - image tiles of one frame (strong scaling)\n")
    out.append("| GPUs | Mrays/s | ms/frame | speed-up | e2e Mrays/s | trace per rank ms (min..max) | exchange ms | filters ms | tail ms | launches/frame |")
    out.append("|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|")
    for n in sorted(by_n):
        d = by_n[n]
        b = d["breakdown_ms_per_step"]
        pr = d.get("per_rank") or {}
        tr = pr.get("trace_ms")
        trs = f"{min(tr):.2f}..{max(tr):.2f}" if tr else f"{sum(b[k] for k in ('gen', 'ext', 'shade', 'tail', 'acc')):.2f} (sum of kernels)"
        sp = f"{d['value'] / base['value']:.2f}x" if base else "-"
        out.append(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {sp} | {d['e2e']['value']:.0f} | {trs} | {b.get('xchg', 0):.2f} | {b['rep'] + b['jbf']:.2f} | {b['tail']:.2f} | "
                   f"{d['gpu_launches'] / d['steps'] / n:.0f} |")
    if mode == "tiles" and base and base.get("cpu_baseline"):
        c = base["cpu_baseline"]
        out.append(f"\nCPU baseline on the same box: {c['value']:.1f} Mrays/s on {c['cores']} cores ({c['kind']}: {c['sample']}).")
    out.append("")
if ROUND == "r01":
    out.append("Note: the 2- and 4-GPU lines of C2 and the multi-GPU lines of C4 were measured before the last change of the round (tight world boxes of rotated "
               "instances: -7 % frame time on one GPU for C2, neutral for C4 whose instances are axis aligned); the 1-GPU lines and the 8-GPU line of C2 are final.")
else:
    # the C4 results that ride on the C2 lines at N > 1 (bench.py "secondary")
    out.append("## C4 city 4K progressive, measured inside the N-GPU runs of C2 (`secondary` of the bench line)\n")
    out.append("| GPUs | shard | Mrays/s | ms/step | speed-up vs one GPU of the same run | e2e Mrays/s | exchange ms |")
    out.append("|---:|---|---:|---:|---:|---:|---:|")
    for (wl, mode), by_n in sorted(rows.items()):
        for n in sorted(by_n):
            for sec in by_n[n].get("secondary") or []:
                out.append(f"| {sec['n_gpus']} | {sec['shard']} | {sec['value']:.0f} | {sec['ms_per_step']:.2f} | {sec['speedup_vs_one_gpu_same_run']:.2f}x "
                           f"(1 GPU: {sec['one_gpu_same_run']['value']:.0f}) | {sec['e2e']['value']:.0f} | {sec['breakdown_ms_per_step'].get('xchg', 0):.2f} |")
    out.append("")
open(os.path.join(ROOT, "profiles", ROUND + "_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
