"""Turn ncu output into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv  > profiles/r01_launches.md
    python tools/ncu_summary.py report   gpurun_out/prof_trace_r01.ncu-rep > profiles/r01_k_trace_full.md

`launches` aggregates a `--metrics gpu__time_duration.sum --csv` launch list per kernel (count, total,
share of the captured window).  `report` prints the metrics we reason with in DESIGN.md for every
launch in a `--set full` report (read with `ncu -i ... --page raw --csv`).
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "global load sector use %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__sass_inst_executed_op_local.sum", "local-memory instr"),
]


def short(name):
    name = re.sub(r"gk::|<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(<[^>]*>)?)", name)
    return m.group(1) if m else name[:60]


def launches(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ns
        a[2] = max(a[2], ns)
        total += ns
    print(f"# launch list summary: {path}\n")
    print(f"{len(rows)} launches captured, {total / 1e6:.3f} ms of kernel time (serialised, cold-cache ncu replay: use the SHARES).\n")
    print("| kernel | launches | total ms | share % | mean us | max us |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t / 1e6:.3f} | {100 * t / total:.1f} | {t / n / 1e3:.1f} | {mx / 1e3:.1f} |")


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full summary: {path}\n")
    names = [short(r[hdr.index("Kernel Name")]) for r in body]
    grids = [r[hdr.index("Grid Size")] for r in body]
    print("| metric | unit | " + " | ".join(f"#{i} `{n}` {g}" for i, (n, g) in enumerate(zip(names, grids))) + " |")
    print("|---|---|" + "---:|" * len(body))
    for key, label in KEEP:
        if key not in hdr:
            continue
        i = hdr.index(key)
        print(f"| {label} (`{key}`) | {units[i]} | " + " | ".join(r[i] for r in body) + " |")


def _num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def _to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def traffic(path, workload, out_json):
    """profiles/r02_traffic.json: per-launch averages over the traversal launches (k_trace_sched / k_trace) of a `--set full` report,
    merged under `workload`: DRAM bytes (roofline.traffic), L2 and L1 bytes, issue-slot use and active threads per warp instruction."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(k) for k in hdr}
    sel = [r for r in body if re.search(r"k_trace(_sched)?\b|k_trace<|k_trace_sched<", r[col["Kernel Name"]]) and "k_trace_coop" not in r[col["Kernel Name"]]]
    if not sel:
        raise SystemExit("no traversal launches in " + path)

    def colbytes(key):
        if key not in col:
            return None
        return [_to_bytes(_num(r[col[key]]), units[col[key]]) for r in sel]

    def colval(key):
        return [_num(r[col[key]]) for r in sel] if key in col else None

    dur = colval("gpu__time_duration.sum")
    dur_unit = units[col["gpu__time_duration.sum"]]
    scale = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3}.get(dur_unit, 1e-9)
    dur_s = [d * scale for d in dur]
    dr, dw = colbytes("dram__bytes_read.sum"), colbytes("dram__bytes_write.sum")
    l2, l1 = colbytes("lts__t_bytes.sum"), colbytes("l1tex__t_bytes.sum")
    if l2 is None and "lts__t_sectors.sum" in col:  # 32-byte sectors
        l2 = [32.0 * v for v in colval("lts__t_sectors.sum")]
    if l1 is None and "l1tex__t_sectors.sum" in col:
        l1 = [32.0 * v for v in colval("l1tex__t_sectors.sum")]
    issue = colval("smsp__issue_active.avg.pct_of_peak_sustained_active")
    lanes = colval("smsp__thread_inst_executed_per_inst_executed.ratio")
    inst = colval("smsp__inst_executed.sum")
    n = len(sel)
    w = dur_s  # duration-weighted means for the percentages
    wsum = sum(w)
    rec = {
        "source": f"{os.path.basename(path)} (ncu --set full --clock-control none, {n} traversal launches of one frame); tools/ncu_summary.py traffic",
        "launches": n,
        "kernels": sorted({short(r[col["Kernel Name"]]) for r in sel}),
        "dram_bytes_per_wave_avg": round(sum(a + b for a, b in zip(dr, dw)) / n, 0),
        "l2_bytes_per_wave_avg": round(sum(l2) / n, 0) if l2 else None,
        "l1_bytes_per_wave_avg": round(sum(l1) / n, 0) if l1 else None,
        "l2_gbs_under_ncu": round(sum(l2) / wsum / 1e9, 1) if l2 else None,
        "l1_gbs_under_ncu": round(sum(l1) / wsum / 1e9, 1) if l1 else None,
        "dram_gbs_under_ncu": round(sum(a + b for a, b in zip(dr, dw)) / wsum / 1e9, 1),
        "issue": {"issue_pct": round(sum(a * b for a, b in zip(issue, w)) / wsum, 1) if issue else None,
                  "lanes_per_inst": round(sum(a * b for a, b in zip(lanes, inst)) / sum(inst), 2) if lanes and inst else None,
                  "lane_efficiency": round(sum(a * b for a, b in zip(lanes, inst)) / sum(inst) / 32.0, 3) if lanes and inst else None,
                  "warp_instructions_per_wave_avg": round(sum(inst) / n, 0) if inst else None},
        "per_launch": [{"kernel": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "us": round(d * 1e6, 1),
                        "dram_bytes": round(a + b, 0), "l2_bytes": round(c, 0) if l2 else None}
                       for r, d, a, b, c in zip(sel, dur_s, dr, dw, l2 or [0] * n)],
    }
    data = {}
    if os.path.exists(out_json):
        data = json.load(open(out_json))
    data[workload] = rec
    json.dump(data, open(out_json, "w"), indent=1)
    print(json.dumps({k: v for k, v in rec.items() if k != "per_launch"}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
