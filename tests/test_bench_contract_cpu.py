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
    if os.path.basename(paths[-1]) >= "r02":
        # round 2 on: the 1M parity block (the cpu_baseline leg's first reference step checks the GPU's step) and the reference's
        # own CUDA solver on the same GPU and state
        p = d["parity"]
        assert p["particles"] == 1000000 and p["ok"] is True and p["value"] <= p["tol"]
        assert d["steps"] >= 100
        assert d["reference_cuda"]["ieee"]["ms_per_step"] > d["ms_per_step"]
        assert d["e2e"]["default_frame_length"]["frames"] > 0


def test_scene_of_n_gpus_repeats_the_one_gpu_scene_along_x():
    """bench.scene(side, world): the block and the box grow along x only (slabs side by side), y and z — the direction the block
    collapses in — are the 1-GPU scene's; bench_multi builds the same box and its ranks' lattice shares tile the block exactly."""
    import sys
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    import bench_multi
    bench.CONFIG_NAME = bench.CONFIGS[3]["name"]
    p1, b1, r1 = bench.scene(6, 1)
    p3, b3, r3 = bench.scene(6, 3)
    assert len(p3) == 3 * len(p1)
    assert b3[1][1] == b1[1][1] and b3[1][2] == b1[1][2] and b3[1][0] > b1[1][0]
    assert b1[1][2] > b1[1][0]                         # twice the block's depth along z: that is where it flows
    nx, box, res = bench_multi.dist_scene(6, 3, "dam")
    assert nx == 18 and np.allclose(box[1], b3[1]) and tuple(res) == tuple(r3)
    # weak scaling: the ranks' shares are the block, each particle once, ids = index in bench.scene's order along x-slabs
    parts = [bench_multi.rank_positions(6, 3, r) for r in range(3)]
    allp = np.concatenate([p for p, _ in parts])
    ids = np.concatenate([i for _, i in parts])
    assert len(np.unique(ids)) == len(p3) == len(allp)
    assert {tuple(np.round(x, 5)) for x in allp} == {tuple(np.round(x, 5)) for x in p3}
    # strong scaling: ONE 6^3 block cut into three x-ranges
    parts = [bench_multi.rank_positions(6, 3, r, strong=True) for r in range(3)]
    allp = np.concatenate([p for p, _ in parts])
    ids = np.concatenate([i for _, i in parts])
    assert len(allp) == 216 and sorted(ids.tolist()) == list(range(216))
    assert {tuple(np.round(x, 5)) for x in allp} == {tuple(np.round(x, 5)) for x in p1}
    assert "8 x 100^3 = 8000000" in bench.workload(100, 8, 200) and "100^3 = 1000000" in bench.workload(100, 1, 200)


def test_ncu_traffic_file_matches_the_bench_workload():
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert t["particles"] == 1000000 and t["dram_bytes_per_launch"]["visc_matvec"] > 1e8
