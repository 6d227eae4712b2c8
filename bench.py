#!/usr/bin/env python3
"""bench.py — DFSPH particle-steps/s on a synthetic dam break (BASELINE.json config 3: 1M particles,
DFSPH + implicit viscosity + surface tension, one B200), with the HBM roofline of the dominant kernel and
the reference's own solver sources timed on the host cores beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--side 100]

A "step" is one DFSPHImplementation::OnUpdate of the whole scene.  `value` is measured with the state
resident in HBM (CUDA events on the solver's stream); `e2e` goes through the C ABI with host buffers:
vfd_dfsph_set_particles (H2D) + vfd_dfsph_simulate with every step captured as a frame (D2H, 36 B/particle).
Nothing here reads /root/reference; oracle/_ref/*.so (the reference's sources compiled by
oracle/build_ref.py) is used only for the cpu_baseline leg and for --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R = 0.025
D = 2 * R
H = 4 * R

# algorithmic HBM bytes per particle and launch of the neighbour-sum kernels:
# own fields read/written once, neighbour ids 4 B each, neighbour fields served from L2
ALGO_BYTES = {
    # kernel class: (fixed bytes per particle, bytes per neighbour)
    "search_bounds": (16, 0), "search_hist": (24, 0), "search_scatter": (12, 0), "search_reorder": (168, 0), "search_build_list": (20, 4),
    "boundary": (32, 0), "density_factor": (76, 4),
    "div_source": (68, 4), "div_accel": (56, 4), "div_solve": (72, 4), "div_finish": (96, 4),
    "press_source": (68, 4), "press_accel": (56, 4), "press_solve": (72, 4), "press_finish": (88, 4),
    "st_classify": (48, 4), "st_smooth": (56, 4), "st_apply": (76, 0),
    "visc_setup": (120, 4), "visc_matvec0": (136, 4), "visc_matvec": (56, 4), "visc_update": (156, 0), "visc_direction": (64, 0), "visc_step": (164, 0), "visc_apply": (80, 0),
    "cfl": (32, 0), "velocity": (48, 0), "position": (48, 0),
}


def block_positions(nx, ny, nz, origin):
    """Lattice at spacing 2r, (i + 1/2) d + origin: what the reference's SampleMode::MinDensity yields (ParticleSampler.cpp:39-43)."""
    d = np.float32(D)
    ax = [(np.arange(n, dtype=np.float32) + np.float32(0.5)) * d + np.float32(o) for n, o in zip((nx, ny, nz), origin)]
    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)


# BASELINE.json configs that fit one GPU: block edge, solver switches, settle steps
CONFIGS = {
    1: dict(side=30, settle=200, name="dam break %d^3 = %d particles, DFSPH (2 divergence + 2 pressure Jacobi iterations), viscosity and surface tension off",
            desc=dict(EnableViscositySolver=0, EnableSurfaceTensionSolver=0)),
    2: dict(side=46, settle=200, name="viscous block drop %d^3 = %d particles, DFSPH (2+2 Jacobi iterations) + implicit viscosity PCG (nu 10, honey-like), surface tension off",
            desc=dict(EnableSurfaceTensionSolver=0)),
    3: dict(side=100, settle=200, name="dam break %d^3 = %d particles, DFSPH (2 divergence + 2 pressure Jacobi iterations) + implicit viscosity PCG (nu 10) + surface tension",
            desc=dict()),
}


def scene(side, world=1):
    """Dam break for `world` GPUs: a lattice block of world*side x side x side particles standing against one long wall of an
    inverted box of the same length and twice the block's depth, so that it collapses along z (config 5:
    "slabs side by side along x").  Every x-slice of the scene is the 1-GPU scene (world = 1): the flow runs across the slabs'
    normal, per-GPU physics, neighbour counts and PCG iteration counts do not depend on the number of GPUs."""
    nx = side * world
    pos = block_positions(nx, side, side, (2 * D, 2 * D, 2 * D))
    L = side * D
    box = ((0.0, 0.0, 0.0), (nx * D + 4 * D, 1.4 * L + 4 * D, 2 * L + 4 * D))
    ext = [b - a + 2 * (8 * H - R) for a, b in zip(*box)]
    res = tuple(min(256, max(8, int(np.ceil(e / (4 * H))))) for e in ext)
    return pos, box, res


def workload(side, world, settle):
    """The workload's name: the same string in both arms' `config.workload` (what differs between the arms goes to `config.notes`)."""
    n = world * side ** 3
    shape = "%d x %d^3" % (world, side) if world > 1 else "%d^3" % side
    return (CONFIG_NAME % (side, side ** 3)).replace("%d^3 = %d" % (side, side ** 3), "%s = %d" % (shape, n)) + "; %d settle steps" % settle


CONFIG_DESC = {}          # solver switches of the selected --config (set by main)
CONFIG_NAME = CONFIGS[3]["name"]


def description(api_mod, frames=0, **kw):
    """Defaults of the reference (viscosity 10, boundary viscosity 10, surface tension 1) with the Jacobi iteration
    counts pinned to 2 (the reference's loops run exactly `Min` iterations: SURVEY.md F4/F5), plus the --config's switches."""
    d = dict(FrameCount=frames, FrameLength=0.0,
             MinPressureSolverIterations=2, MaxPressureSolverIterations=2,
             MinDivergenceSolverIterations=2, MaxDivergenceSolverIterations=2)
    d.update(CONFIG_DESC)
    d.update(kw)
    return api_mod(**d)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        """nvidia-smi takes ~0.1 s to deliver its first line: started before the warm-up so that it is streaming (one line per
        50 ms: the driver's 20-step run times ~0.2 s) when the timed region begins."""
        if self.proc is not None:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __enter__(self):
        self.start()
        self.t0 = time.perf_counter()
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)
            self.proc = None

    def summary(self):
        """Samples taken inside the timed region; a region shorter than the sampling period borrows the samples of the half
        second on either side (warm-up before, the event-bracketed pass after: the same kernels under the same load)."""
        self.stop()
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        window = "timed region"
        if not rows:
            rows = [r for t, r in self.rows if t0 - 0.5 <= t <= t1 + 0.5]
            window = "timed region +- 0.5 s (region shorter than the 50 ms sampling period)"
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_settled_state(args, refsim, desc):
    """Scene preparation of the reference arm WITHOUT the product: the workload's settled state (`--settle` steps from the
    lattice) computed by the reference itself — its CUDA build for this GPU (oracle/_ref/libvfd_ref_gpu.so, the IEEE
    variant: SURVEY.md F6) when the box has a GPU and the build travelled, which takes seconds; else on the host cores on a
    smaller scene (`--ref-side`).  Returns (state, dt, (sample count, Monte-Carlo factor), positions, box, map resolution, how)."""
    def settle(kind, side, steps):
        pos, box, res = scene(side, max(1, args.gpus))
        sim = refsim.RefSim(desc, kind=kind, threads=os.cpu_count() or 1)
        sim.set_particles(pos)
        sim.add_box_body(box[0], box[1], inverted=True, padding=0.0, res=res)
        sim.commit_bodies()
        sim.step(steps)
        info = sim.info_bytes().tobytes()             # DFSPHSimulationInfo: SurfaceTensionSampleCount @84, MonteCarloFactor @112
        st = (int(np.frombuffer(info[84:88], np.uint32)[0]), float(np.frombuffer(info[112:116], np.float32)[0]))
        return sim.particles(), float(sim.debug()["dt"]), st, pos, box, res
    gpu = False
    try:
        import torch
        gpu = torch.cuda.is_available()
    except Exception:
        pass
    if gpu and refsim.available("gpu"):
        try:
            out = settle("gpu", args.side, args.settle)
            return out + ("%d settle steps by the reference's own CUDA build on this GPU (untimed)" % args.settle, args.side)
        except Exception as ex:
            print("reference arm: settling on the reference's CUDA build failed (%r); settling %d^3 on the CPU" % (ex, args.ref_side), file=sys.stderr)
    out = settle("cpu", args.ref_side, args.ref_settle)
    return out + ("%d settle steps on the host cores (untimed)" % args.ref_settle, args.ref_side)


def run_reference(args, rank, world):
    """--impl reference: the reference's own solver sources (oracle/_ref/libvfd_ref_cpu.so, built by oracle/build_ref.py
    from /root/reference) on all host cores, timed on the arm's workload: K steps from the workload's settled state.
    Nothing of vfd_b200 is imported, loaded or run by this arm — the settled state comes from the reference too."""
    if rank != 0:
        return 0
    from oracle import refsim
    threads = os.cpu_count() or 1
    desc = description(refsim.Desc)
    steps = min(args.steps, max(3, args.ref_steps // max(1, args.gpus)))     # each step is ~1 s of 16 cores per million particles: a bounded sample
    with refsim.quiet_stdout():
        state, dt0, st0, pos, box, res, how, side = reference_settled_state(args, refsim, desc)
        sim = refsim.RefSim(desc, threads=threads)
        sim.set_particles(pos)
        sim.add_box_body(box[0], box[1], inverted=True, padding=0.0, res=res)
        sim.commit_bodies()
        sim.set_particles_full(state)
        sim.set_time_step(dt0)
        sim.set_st_state(*st0)
        sim.step(min(args.warmup, 3))
        t0 = time.perf_counter()
        its = []
        for _ in range(steps):
            sim.step(1)
            its.append(sim.debug()["visc_it"])
        dt = time.perf_counter() - t0
    n = len(pos)
    v = n * steps / dt
    full = side == args.side
    wl = workload(args.side, max(1, args.gpus), args.settle) if full else CONFIG_NAME % (side, n) + " (the box has no GPU for the reference's CUDA build to prepare the full scene)"
    sample = "%s; %s; %d warm-up + %d timed steps (a bounded sample of the arm's %d); mean PCG it %.1f" % (
        wl, how, min(args.warmup, 3), steps, args.steps, float(np.mean(its)))
    print(json.dumps({
        "impl": "reference", "metric": "DFSPH particle-steps/s", "value": v, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "notes": "the reference's solver sources on %d host threads; %s" % (threads, how), "baseline_config": args.config,
                   "particles": n, "pcg_iterations_mean": float(np.mean(its))},
        "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def ncu_traffic(kernel, n):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel class from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json, written from the round's capture by
    tools_ncu_summary.py); None when the capture does not cover this kernel or particle count."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if int(t.get("particles", 0)) != int(n):
            return None
        v = t.get("dram_bytes_per_launch", {}).get(kernel)
        return None if v is None else float(v)
    except Exception:
        return None


def roofline_pass(sim, api, n, mbar, steps, skip=False):
    """Second pass over the next K steps with every launch bracketed by CUDA events on the solver's stream: per-kernel
    device time -> roofline of the dominant kernel class (the event pairs cost a few us per launch, so this pass is not
    the one `value` is taken from).  Returns (roofline, per-kernel table, ms of the pass)."""
    sim.set_option(api.VFD_OPT_KERNEL_TIMERS, 0 if skip else 1)
    sim.kernel_times(reset=True)
    sim.record_event(2)
    for _ in range(0 if skip else steps):
        sim.OnUpdate()
    sim.record_event(3)
    ms_prof = sim.elapsed_ms(2, 3)
    ktimes = sim.kernel_times()
    sim.set_option(api.VFD_OPT_KERNEL_TIMERS, 0)
    peak, peak_src = peaks()
    roof, table = None, {}
    total_kernel_ms = sum(v[2] for v in ktimes.values())
    for name, (msa, lna, msall, lnall) in sorted(ktimes.items(), key=lambda kv: -kv[1][2]):
        fixed, per = ALGO_BYTES.get(name, (0, 0))
        b = (fixed + per * mbar) * n
        avg = msa / lna if lna else 0.0
        gbs = b / (avg * 1e-3) / 1e9 if avg > 0 and b > 0 else None
        table[name] = {"ms_per_step": round(msall / steps, 4), "launches_per_step": round(lnall / steps, 2),
                       "avg_ms_active": round(avg, 5), "algo_GBps": None if gbs is None else round(gbs, 1),
                       "frac_of_peak": None if gbs is None else round(gbs / peak, 3), "share": round(msall / total_kernel_ms, 4)}
        if roof is None:
            roof = {"bound": "hbm", "kernel": name, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak if gbs else None,
                    "traffic": ncu_traffic(name, n), "peak_source": peak_src, "avg_launch_ms": avg, "active_launches": lna,
                    "share_of_step": msall / total_kernel_ms, "algorithmic_bytes_per_launch": b,
                    "note": "event-bracketed pass over the %d steps following the timed region (%.3f ms/step with brackets); algorithmic bytes per particle = "
                            "%d + %d x mean neighbours (SURVEY.md section 8d); traffic = DRAM bytes per launch from the round's committed ncu capture "
                            "(profiles/ncu_traffic.json), not measured in this run" % (steps, ms_prof / steps, fixed, per)}
    return roof, table, ms_prof


PARITY_FIELDS = ["Position", "Velocity", "Acceleration", "PressureAcceleration", "VelocityDifference", "MonteCarloSurfaceNormal",
                 "MonteCarloSurfaceNormalSmooth", "Density", "DensityAdvection", "PressureRho2", "PressureRho2V", "Factor",
                 "MonteCarloSurfaceCurvature", "MonteCarloSurfaceCurvatureSmooth", "DeltaFinalCurvature"]


def cpu_baseline(state, dt, st, pos0, box, res, steps=10, gpu_step=None):
    """The reference's solver sources on this box's host cores, from the GPU run's settled state; its FIRST step is also the
    checker of the GPU's step from the same state (`gpu_step`: the 120-B particle state after one vfd_b200 step): error of
    every field relative to the field's scale, as the parity tests measure it."""
    from oracle import refsim
    if not refsim.available("cpu"):
        return None, None
    threads = os.cpu_count() or 1
    with refsim.quiet_stdout():
        sim = refsim.RefSim(description(refsim.Desc), threads=threads)
        sim.set_particles(pos0)
        sim.add_box_body(box[0], box[1], inverted=True, padding=0.0, res=res)
        sim.commit_bodies()
        sim.set_particles_full(state)
        sim.set_time_step(dt)
        sim.set_st_state(*st)
        ref_map = sim.volume_map(0)       # the reference's own host precompute of the box's volume map: the same input for both
        sim.step(1)                       # untimed: first-touch of the reference's buffers; the parity reference
        first = sim.particles()
        t0 = time.perf_counter()
        sim.step(steps)
        el = time.perf_counter() - t0
    n = len(pos0)
    par = None
    if gpu_step is not None:
        gpu_step = gpu_step(ref_map)
        # the reference's own run-to-run spread on this very step (its neighbour order and its reductions depend on the thread
        # schedule: SURVEY.md Appendix E): a second, single-threaded run of the same step — affordable up to a few 100k particles
        noise = {}
        if n <= 200000:
            with refsim.quiet_stdout():
                again = refsim.RefSim(description(refsim.Desc), serial=True)
                again.set_particles(pos0)
                again.add_box_body(box[0], box[1], inverted=True, padding=0.0, res=res)
                again.commit_bodies()
                again.set_particles_full(state)
                again.set_time_step(dt)
                again.set_st_state(*st)
                again.step(1)
                second = again.particles()
            for f in PARITY_FIELDS:
                y, z = np.asarray(first[f], np.float64), np.asarray(second[f], np.float64)
                scale = np.abs(y).max()
                noise[f] = 0.0 if scale == 0.0 else float(np.abs(y - z).max() / scale)
        worst = ("", 0.0, 2.0e-5, 0.0)
        ok = True
        for f in PARITY_FIELDS:
            x, y = np.asarray(gpu_step[f], np.float64), np.asarray(first[f], np.float64)
            scale = np.abs(y).max()
            e = 0.0 if scale == 0.0 else float(np.abs(x - y).max() / scale)
            tol = max(2.0e-5, 3.0 * noise.get(f, 0.0))
            ok = ok and e <= tol
            if e / tol > worst[3]:
                worst = (f, e, tol, e / tol)
        par = {"worst_field": worst[0], "value": worst[1], "tol": worst[2], "ok": bool(ok), "particles": n,
               "reference_self_deviation": (noise.get(worst[0]) if noise else None),
               "note": "one step from the timed run's settled state, both sides fed the reference's own volume map (its mesh distance is evaluated in "
                       "fp32 on a 10-m box, which moves wall distances by up to 1e-3 against the analytic box distance the timed run's GPU-built map "
                       "uses: DESIGN.md section 2): vfd_b200 (search without FMA contraction, as the host-compiled oracle) vs the reference's sources "
                       "on the host cores; max |error| / max |field| over the fields of the 120-B particle state; tolerance as in "
                       "tests/test_gpu_scale.py: max(2e-5, 3 x the reference's own deviation between a threaded and a single-threaded run of this "
                       "step — measured here up to 200k particles; it moves by ~1e-5 against itself at 1M)"}
    return {"value": n * steps / el, "unit": "particle-steps/s", "cores": threads, "kind": "reference",
            "sample": "%d steps of the same %d-particle settled state (after 1 untimed step), %.1f s; PCG it of last step %d" % (
                steps, n, el, sim.debug()["visc_it"])}, par


def reference_cuda(state, dt, st, pos0, box, res, steps=5):
    """The reference's own CUDA solver (its sources compiled by oracle/build_ref.py --gpu for sm_100a) on this GPU and this
    state: the only pre-existing GPU implementation of the path.  `ieee`: without fast-math and with the
    preconditioner's accumulator zero-initialised; `fast_math`: as shipped."""
    from oracle import refsim
    out = {}
    for key, kind, k in (("ieee", "gpu", steps), ("fast_math", "gpu_fast", 2)):
        if not refsim.available(kind):
            out[key] = None
            continue
        try:
            with refsim.quiet_stdout():
                sim = refsim.RefSim(description(refsim.Desc), kind=kind)
                sim.set_particles(pos0)
                sim.add_box_body(box[0], box[1], inverted=True, padding=0.0, res=res)
                sim.commit_bodies()
                sim.set_particles_full(state)
                sim.set_time_step(dt)
                sim.set_st_state(*st)
                sim.step(1)
                t0 = time.perf_counter()
                its, tm = [], {}
                for _ in range(k):
                    sim.step(1)
                    dbg = sim.debug()
                    its.append(dbg["visc_it"])
                    for name, us in dbg["timers_us"].items():
                        tm[name] = tm.get(name, 0.0) + us / 1e3 / k
                fin = np.isfinite(sim.particles()["Position"]).all()       # device -> host read: the steps have completed
                el = time.perf_counter() - t0
            out[key] = {"ms_per_step": 1e3 * el / k, "pcg_it_mean": float(np.mean(its)), "steps": k, "phase_ms": {a: round(b, 2) for a, b in tm.items()},
                        "positions_finite": bool(fin)}
        except Exception as ex:
            out[key] = {"error": repr(ex)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3], help="BASELINE.json config: 1 = 27k dam break, DFSPH only; 2 = 100k viscous block; "
                                                                             "3 = 1M dam break, DFSPH + viscosity + surface tension (the metric's config, default)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--side", type=int, default=None, help="block edge in particles (default: the config's; 100 -> 1M)")
    ap.add_argument("--settle", type=int, default=None, help="scene-preparation steps before warm-up (SURVEY.md §8d; default 200)")
    ap.add_argument("--ref-steps", type=int, default=24, help="--impl reference: at most this many timed steps, divided by the GPU count (a step is ~1.2 s of 16 cores per "
                                                              "million particles: the driver's 20 steps run in full at N = 1, a bounded sample of them at N > 1)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference's own CUDA build on this GPU")
    ap.add_argument("--scene", default="dam", choices=["dam", "tank"], help="N > 1: the dam break stretched along x (default; the N = 1 scene at N = 1) or a closed tank")
    ap.add_argument("--strong", action="store_true", help="N > 1: strong scaling — ONE side^3 scene (BASELINE.json config 4: --side 200) cut into N slabs, instead of side^3 per GPU")
    ap.add_argument("--ref-side", type=int, default=50)
    ap.add_argument("--ref-settle", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scene-prep", action="store_true", help="skip the scene-preparation block (tools_scene_prep.py in a child process)")
    ap.add_argument("--ncu-visc-it", type=int, default=0, help="with --ncu: cap the PCG at this many iterations for the profiled steps "
                                                                  "(keeps an `ncu --set full` capture of one step short)")
    ap.add_argument("--ncu", action="store_true", help="bracket the timed steps with cudaProfilerStart/Stop (run under ncu --profile-from-start off); "
                                                        "skips the event-bracketed pass, e2e and cpu_baseline; numbers printed under a profiler are not bench values")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3)
    global CONFIG_NAME
    cfg = CONFIGS[args.config]
    args.side = args.side or cfg["side"]
    args.settle = cfg["settle"] if args.settle is None else args.settle
    CONFIG_DESC.clear(); CONFIG_DESC.update(cfg["desc"])
    CONFIG_NAME = cfg["name"]
    if args.impl == "reference":
        return run_reference(args, rank, world)

    from vfd_b200 import api, build
    build.build()
    if world > 1:
        import bench_multi
        return bench_multi.main(args, rank, local, world)

    pos, box, res = scene(args.side)
    n = len(pos)
    vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=R, device=local)
    sim = api.DFSPHSimulation(description(api.DFSPHSimulationDescription), device=local)
    sim.SetFluidObjects([api.FluidObject(pos)])
    sim.SetRigidBodies([vm])
    sim.steps(args.settle)
    sim.synchronize()
    settled = sim.particles()
    info = sim.GetInfo()
    settled_dt, settled_st = info.TimeStepSize, (info.SurfaceTensionSampleCount, info.MonteCarloFactor)

    clk = ClockSampler(local).start()
    sim.steps(args.warmup)
    sim.synchronize()
    sim.launch_count(reset=True)
    if args.ncu:
        import torch
        args.no_e2e = args.no_cpu_baseline = True
        if args.ncu_visc_it:
            dt_now = sim.GetCurrentTimeStepSize()
            sim.SetDescription(description(api.DFSPHSimulationDescription, MaxViscositySolverIterations=args.ncu_visc_it))
            sim.set_time_step(dt_now)
        torch.cuda.cudart().cudaProfilerStart()
    with clk:
        sim.record_event(0)
        for _ in range(args.steps):
            sim.OnUpdate()
        sim.record_event(1)
        ms = sim.elapsed_ms(0, 1)
    if args.ncu:
        sim.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    launches = sim.launch_count()
    value = n * args.steps / (ms * 1e-3)
    dbg = sim.GetDebugInfo()
    counts, _, _ = sim.neighbors()
    mbar = float(counts.mean())
    tstats = sim.tile_stats()

    roof, table, ms_prof = roofline_pass(sim, api, n, mbar, args.steps, skip=args.ncu)

    out = {
        "metric": "DFSPH particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.side, 1, args.settle),
                   "notes": ("working set > L2 (neighbour list and pair coefficients alone %.0f MB), no flush" % (n * 72 * 6 / 1e6) if n * 72 * 6 > 126e6 else
                             "working set %.0f MB of list and coefficients: L2-resident, flushed between steps by nothing (stated: NOT flushed)" % (n * 72 * 6 / 1e6)),
                   "baseline_config": args.config,
                   "particles": n, "mean_neighbours": mbar, "pcg_iterations_last_step": int(dbg.ViscositySolverIterationCount),
                   "grid_tiles": tstats["tiles"], "tile_passes_on_slow_path": tstats["fallback_tile_passes"],
                   "kernels": table},
        "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof,
    }

    if not args.no_e2e:
        # end to end through the C ABI with host buffers: upload the settled positions/velocities, bake K frames
        hp = np.ascontiguousarray(settled["Position"])
        hv = np.ascontiguousarray(settled["Velocity"])
        e = api.DFSPHSimulation(description(api.DFSPHSimulationDescription, frames=args.steps), device=local)
        e.SetFluidObjects([api.FluidObject(hp, velocities=hv)])
        e.SetRigidBodies([vm])
        e.Simulate()                                                  # untimed first bake: sizes every buffer, incl. the pinned frame store
        e.synchronize()
        t0 = time.perf_counter()
        e.SetFluidObjects([api.FluidObject(hp, velocities=hv)])      # H2D of the inputs
        e.SetRigidBodies([vm])
        e.Simulate()                                                  # K steps, each captured: D2H 36 B/particle
        frame, _, _ = e.GetFrame(args.steps - 1)
        el = time.perf_counter() - t0
        out["e2e"] = {"value": n * args.steps / el, "unit": "particle-steps/s", "h2d_bytes_per_step": int(24 * n / args.steps),
                      "d2h_bytes_per_step": 36 * n, "ms_per_step": 1e3 * el / args.steps,
                      "note": "vfd_dfsph_set_particles + set_rigid_bodies + simulate (FrameLength 0: every step baked to a host frame) + get_frame of the last one, "
                              "wall clock; second bake of the handle (an untimed first bake sized the buffers)"}
        # the same through the reference's DEFAULT frame length (0.0016 s of simulated time per frame): which steps become
        # frames is decided on the device, the host never reads the time step back
        try:
            nf = max(4, args.steps // 2)
            e.SetDescription(description(api.DFSPHSimulationDescription, frames=nf, FrameLength=0.0016))
            e.SetFluidObjects([api.FluidObject(hp, velocities=hv)])
            e.SetRigidBodies([vm])
            e.Simulate()
            e.synchronize()
            t0 = time.perf_counter()
            e.SetFluidObjects([api.FluidObject(hp, velocities=hv)])
            e.SetRigidBodies([vm])
            e.Simulate()
            frame, _, _ = e.GetFrame(nf - 1)
            el2 = time.perf_counter() - t0
            steps2 = int(e.GetDebugInfo().IterationCount)
            out["e2e"]["default_frame_length"] = {"value": n * steps2 / el2, "unit": "particle-steps/s", "ms_per_step": 1e3 * el2 / steps2, "steps": steps2, "frames": nf,
                                                  "d2h_bytes_per_step": int(36 * n), "note": "FrameLength 0.0016 (the reference's default): %d steps made %d frames; the "
                                                  "export and the copy are enqueued for every step, the frame pipe keeps the ones the device marked" % (steps2, nf)}
        except Exception as ex:
            out["e2e"]["default_frame_length"] = {"error": repr(ex)}
        e.close()

    if not args.no_cpu_baseline:
        try:
            # the GPU's step from the settled state, for the parity block (host-order search: the oracle is host-compiled)
            def gpu_step(m):
                g = api.DFSPHSimulation(description(api.DFSPHSimulationDescription), device=local)
                g.set_option(api.VFD_OPT_SEARCH_FMA, 0)
                g.SetFluidObjects([api.FluidObject(pos)])
                g.SetRigidBodies([api.VolumeMap(m["domain_min"], m["domain_max"], m["resolution"], m["cell_size"], m["cell_size_inv"], m["field_count"],
                                                m["node_count"], m["cell_count"], m["cell_map_count"], m["nodes"], m["cells"], m["cell_map"])])
                g.set_particles_full(settled)
                g.set_time_step(settled_dt)
                g.set_surface_tension_state(*settled_st)
                g.OnUpdate()
                got = g.particles()
                g.close()
                return got
            out["cpu_baseline"], out["parity"] = cpu_baseline(settled, settled_dt, settled_st, pos, box, res, gpu_step=gpu_step)
        except Exception as ex:      # the baseline must not take the GPU number down with it
            out["cpu_baseline"] = {"error": repr(ex)}
    if not args.no_ref_cuda and not args.no_cpu_baseline:
        try:
            sim.close()
            out["reference_cuda"] = reference_cuda(settled, settled_dt, settled_st, pos, box, res)
            ie = (out["reference_cuda"] or {}).get("ieee") or {}
            if ie.get("ms_per_step"):
                out["reference_cuda"]["speedup_over_ieee"] = ie["ms_per_step"] / (ms / args.steps)
        except Exception as ex:
            out["reference_cuda"] = {"error": repr(ex)}
    if not args.no_cpu_baseline and not args.no_scene_prep:
        # the steps next to the path: mesh sampling and volume-map precompute on this GPU, checked
        # against and timed beside the reference's host code — part of the CPU-baseline leg (the one place the bench runs
        # oracle/), in a child process so that nothing there can touch the numbers above
        try:
            import subprocess
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools_scene_prep.py"), "--device", str(local)], stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True, timeout=150)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            out["scene_prep"] = json.loads(lines[-1]) if r.returncode == 0 and lines else {"error": "exit code %d: %s" % (r.returncode, r.stderr[-300:])}
        except Exception as ex:
            out["scene_prep"] = {"error": repr(ex)}
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner, the reference's banners)
    # are routed to stderr while the bench runs; print() keeps writing to the real stdout
    sys.stdout.flush()
    _real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real, "w")
    rc = main()
    sys.stdout.flush()
    sys.exit(rc)
