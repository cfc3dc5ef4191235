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


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
