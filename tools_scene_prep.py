#!/usr/bin/env python3
"""tools_scene_prep.py — the scene-preparation steps next to the solver path measured on this
box: particle sampling of a mesh volume and the volume map of a rigid body, on the GPU through the C ABI
(vfd_sample_mesh_volume, vfd_volume_map_build_mesh / _build_box) and — the checker and the host baseline — by the reference's
own host code (oracle/_ref: FluidObject -> ParticleSampler, RigidBody -> MeshDistance / SDF / Gauss quadrature, OpenMP).
Prints ONE JSON object.  bench.py runs it in a child process inside its cpu_baseline leg (`scene_prep` block of the line): a
fault here cannot take the solver's numbers with it.

The scene is the shape of the reference's shipped DFSPH scene (a cone of fluid over a slab); `--scale` enlarges the cone."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
R = 0.025


def cone(segments=32, radius=1.0, height=2.0):
    a = 2.0 * np.pi * np.arange(segments) / segments
    rim = np.stack([radius * np.sin(a), np.full(segments, -0.5 * height), -radius * np.cos(a)], 1)
    v = np.concatenate([rim, [[0.0, 0.5 * height, 0.0]]]).astype(np.float32)
    t = [(i, segments, (i + 1) % segments) for i in range(segments)] + [(0, i, i + 1) for i in range(1, segments - 1)]
    t = np.array(t, np.uint32)
    w = v.astype(np.float64)
    vol = np.einsum("ij,ij->i", w[t[:, 0]], np.cross(w[t[:, 1]], w[t[:, 2]])).sum()
    return v, (t if vol > 0 else np.ascontiguousarray(t[:, ::-1]))


def box_mesh(lo, hi):
    a, b = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    v = np.array([[a[0], a[1], a[2]], [b[0], a[1], a[2]], [b[0], a[1], b[2]], [a[0], a[1], b[2]],
                  [a[0], b[1], a[2]], [b[0], b[1], a[2]], [b[0], b[1], b[2]], [a[0], b[1], b[2]]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3], [4, 7, 6], [4, 6, 5], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7]], np.uint32)
    return v, t


def timed(fn, repeat=1):
    best, out = None, None
    for _ in range(repeat):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out, best


def measure(api, refsim, scale, device=0):
    cv, ct = cone()
    cT = np.array([[scale, 0, 0, 0], [0, -scale, 0, 1.0 + 2.0 * scale], [0, 0, -scale, 0], [0, 0, 0, 1]], np.float32)     # upside down, above the slab
    sv, st = box_mesh((-1, -1, -1), (1, 1, 1))
    sT = np.diag([2.0 * scale, 0.2, 2.0 * scale, 1.0]).astype(np.float32)
    res = (20, 20, 20)
    out = {"scene": "cone of fluid (33 vertices, 62 triangles, scale %g, MediumDensity sampling, distance grid 20^3) over a slab (12 triangles, volume map 20^3); "
                    "particle radius %g" % (scale, R)}
    api.mesh_signed_distance(sv, st, np.zeros((1, 3), np.float32), device=device)        # context and module load
    pos, t_s = timed(lambda: api.sample_mesh_volume(cv, ct, R, res, False, 1, transform=cT, device=device), repeat=3)
    vm, t_m = timed(lambda: api.VolumeMap.build_mesh(sv, st, transform=sT, resolution=res, particle_radius=R, device=device), repeat=3)
    out["gpu"] = {"particles_sampled": int(len(pos)), "sampling_ms": 1e3 * t_s, "volume_map_ms": 1e3 * t_m, "map_nodes": int(vm.node_count),
                  "note": "host wall clock of the C-ABI call: mesh upload, kernels, result back in host memory; best of 3"}
    if refsim is not None and refsim.available("cpu") and hasattr(refsim._load("cpu"), "ref_add_mesh_body"):
        threads = os.cpu_count() or 1
        with refsim.quiet_stdout():
            rpos, rt_s = timed(lambda: refsim.sample_mesh_volume(cv, ct, R, res, False, 1, transform=cT))
            sim = refsim.RefSim(refsim.Desc(ParticleRadius=R), threads=threads)
            sim.set_particles(np.zeros((1, 3), np.float32))
            _, rt_m = timed(lambda: sim.add_mesh_body(sv, st, transform=sT, inverted=False, padding=0.0, res=res))
            m = sim.volume_map(0)
        n = int(m["node_count"])
        f1 = float(np.abs(m["nodes"][n:]).max())
        out["reference"] = {"cores": threads, "sampling_ms": 1e3 * rt_s / 2, "volume_map_ms": 1e3 * rt_m / 2,
                            "note": "the reference's host code (FluidObject / RigidBody constructors); the oracle builds each twice until two builds agree "
                                    "(its randomised sphere tree: DESIGN.md section 2) — half the measured time is reported"}
        out["parity"] = {"samples_equal": bool(rpos.shape == pos.shape and np.array_equal(rpos, pos)), "samples_reference": int(len(rpos)),
                         "map_nodes_equal": bool(n == vm.node_count),
                         "distance_field_bit_exact": bool(n == vm.node_count and np.array_equal(vm.nodes[:n], m["nodes"][:n])),
                         "volume_field_max_err_of_scale": (float(np.abs(vm.nodes[n:] - m["nodes"][n:]).max()) / f1) if n == vm.node_count and f1 > 0 else None}
        out["speedup"] = {"sampling": out["reference"]["sampling_ms"] / out["gpu"]["sampling_ms"], "volume_map": out["reference"]["volume_map_ms"] / out["gpu"]["volume_map_ms"]}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=2.0, help="size of the cone (1: the fixture's 18 849 particles; 2: ~150 000)")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    real = os.dup(1)
    os.dup2(2, 1)                       # the reference prints banners on stdout
    from vfd_b200 import api
    try:
        from oracle import refsim
    except Exception:
        refsim = None
    result = measure(api, refsim, a.scale, a.device)
    os.write(real, (json.dumps(result) + "\n").encode())
