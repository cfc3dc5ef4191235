"""Per-source-line view of one kernel of an `ncu --set full --import-source on` report (binary built with -lineinfo):
warp-stall samples and executed warp instructions per CUDA source line, summed over the launches of the kernel.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-name regex> [top N] > profiles/<name>_lines.md
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, pattern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = csv.reader(raw.splitlines())
    per_line = collections.defaultdict(lambda: [0, 0, 0, ""])  # (file, line) -> samples, warp instr, thread instr, text
    fname, func, use, hdr = "", "", False, None
    launches = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            use = re.search(pattern, func) is not None
            continue
        if r[0] == "Kernel Name":
            continue
        if r[0] == "Line No":
            hdr = r
            if use and fname:
                pass
            continue
        if not use or hdr is None or not r[0].isdigit():
            continue
        col = {k: i for i, k in enumerate(hdr)}
        try:
            samples = int(r[col["# Samples"]] or 0)
            winst = int(r[col["Instructions Executed"]] or 0)
            tinst = int(r[col["Thread Instructions Executed"]] or 0)
        except (ValueError, KeyError):
            continue
        e = per_line[(fname, int(r[0]))]
        e[0] += samples
        e[1] += winst
        e[2] += tinst
        e[3] = r[1].strip()
    tot_s = sum(v[0] for v in per_line.values()) or 1
    tot_i = sum(v[1] for v in per_line.values()) or 1
    print(f"# per-source-line profile: kernels matching `{pattern}` in {rep.split('/')[-1]}\n")
    print(f"{tot_s} warp-stall samples, {tot_i} executed warp instructions over all matching launches.\n")
    print("| file:line | stall samples % | warp instr % | threads / instr | source |")
    print("|---|---:|---:|---:|---|")
    for (f, ln), (s, wi, ti, text) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {f}:{ln} | {100 * s / tot_s:.1f} | {100 * wi / tot_i:.1f} | {ti / wi if wi else 0:.1f} | `{text[:110].replace('|', '/')}` |")


if __name__ == "__main__":
    main()
