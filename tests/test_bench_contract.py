"""The bench.py contract, checked on the committed bench lines of the round (profiles/r02_bench/*.json): every key the driver and
the judge read must be present with the right type, the reference arm must carry its own keys, and the multi-GPU line must carry
the C4 results.  (The numbers themselves are measurements; this only guards the shape of the line.)"""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench", "*.json")))


def _load(path):
    return json.load(open(path))


def test_there_are_committed_bench_lines():
    names = {os.path.basename(p) for p in LINES}
    assert {"n1_room_tiles.json", "n1_room_reference.json", "n8_room_tiles.json", "n1_bricks_tiles.json", "n1_city_tiles.json"} <= names


@pytest.mark.parametrize("path", [p for p in LINES if "reference" not in p], ids=os.path.basename)
def test_our_arm_line_has_the_contract_keys(path):
    d = _load(path)
    for key, typ in (("metric", str), ("value", (int, float)), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int), ("ms_per_step", (int, float)),
                     ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str), ("config", dict), ("gpu_launches", int), ("e2e", dict),
                     ("roofline", dict), ("clocks", dict)):
        assert key in d and isinstance(d[key], typ), key
    assert d["vs_baseline"] is None and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] in ("l2", "hbm", "tensor") and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["unit"] in ("GB/s", "TFLOP/s")
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
        p = d["parity"]
        assert p["errors"] == 0 and p["rays"] >= 300000 and p["id_equal"] + p["exact_t_ties"] + p["eps_ties"] + p["reference_misses"] == p["rays"]
        assert p["reference_misses_checked"] == min(p["reference_misses"], p["reference_misses_checked"]) or p["reference_misses"] == 0
    else:
        assert d["per_rank"] and len(d["per_rank"]["trace_ms"]) == d["n_gpus"]


def test_room_line_carries_measured_traffic():
    d = _load(os.path.join(ROOT, "profiles", "r02_bench", "n1_room_tiles.json"))
    r = d["roofline"]
    assert r["traffic"] and r["traffic_source"] and r["issue"]["ncu"]["lanes_per_inst"] > 0 and r["hbm"]["frac"] < 0.05
    t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    assert abs(t["room"]["dram_bytes_per_wave_avg"] - r["traffic"]) < 1.0


def test_reference_arm_line():
    d = _load(os.path.join(ROOT, "profiles", "r02_bench", "n1_room_reference.json"))
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_multi_gpu_line_carries_the_c4_results():
    d = _load(os.path.join(ROOT, "profiles", "r02_bench", "n8_room_tiles.json"))
    sec = {s["shard"]: s for s in d["secondary"]}
    assert set(sec) == {"tiles", "frames"} and sec["frames"]["scaling"] == "weak" and sec["frames"]["speedup_vs_one_gpu_same_run"] >= 7.0
