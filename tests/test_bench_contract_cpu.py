"""The bench line's contract (driver side): the last committed line of this round must carry every key the driver and the
judge read, with consistent values.  Runs on the CPU: it checks the committed artefact, not a new measurement."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last_line(path):
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_committed_bench_line_has_the_contract_keys():
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_final.json")))
    assert paths, "no committed bench line under profiles/"
    d = last_line(paths[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "gpu_launches", "clocks", "roofline", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "DFSPH particle-steps/s" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "workload" in d["config"] and d["config"]["particles"] == 1000000
    # value is the whole-job throughput of the timed region
    assert abs(d["value"] - d["config"]["particles"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] >= 0.5 * r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] <= d["value"] * 1.05
    k = d["clocks"]
    assert "sm_mhz" in k and "sm_max_mhz" in k and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_ncu_traffic_file_matches_the_bench_workload():
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert t["particles"] == 1000000 and t["dram_bytes_per_launch"]["visc_matvec"] > 1e8
