#!/usr/bin/env python3
"""tests/golden/make_golden.py — generates the committed golden fixtures from the REFERENCE'S OWN solver
sources (oracle/_ref/libvfd_ref_cpu.so, built by oracle/build_ref.py from /root/reference).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
Each fixture <scene>.npz holds, for a small scene of tests/scenes.py:
  pos0, vel0            initial state
  map_*                 the rigid body's flattened volume map (reference SDF output)
  state_in, dt_in, st_in  full 120-B particle state after K warm-up steps + the running dt and
                        (SurfaceTensionSampleCount, MonteCarloFactor)
  state_out, dt_out     the state one reference step later (serial emulation => deterministic)
  nbr_counts/offsets/ids  the CSR neighbour list the reference built during that step (sorted per row)
  kernel                the reference's PrecomputedDFSPHCubicKernel bytes (W[10000], gradW[10001], 6 scalars)
The reference has no tests or golden vectors of its own: these outputs of the reference
itself are what pins the oracle and the CUDA path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle.refsim import Desc, RefSim  # noqa: E402
import scenes  # noqa: E402

WARMUP = {"dfsph": 60, "dfsph_default": 60, "viscous": 60, "full": 60}


def sorted_rows(counts, offsets, ids):
    out = ids.copy()
    for i in range(len(counts)):
        o, c = int(offsets[i]), int(counts[i])
        out[o:o + c] = np.sort(ids[o:o + c])
    return out


def make(name):
    sc = scenes.SCENES[name]
    pos = scenes.scene_positions(name)
    desc = Desc(**sc["desc"])
    sim = RefSim(desc, serial=True)
    sim.set_particles(pos)
    sim.add_box_body(sc["box"][0], sc["box"][1], inverted=True, padding=0.0, res=sc["res"])
    sim.commit_bodies()
    vm = sim.volume_map()
    sim.step(WARMUP[name])
    state_in = sim.particles()
    dbg_in = sim.debug()
    info = sim.info_bytes()
    st_in = np.array([info[84:88].view(np.uint32)[0], info[112:116].view(np.float32)[0]], dtype=np.float64)
    sim.step(1)
    state_out = sim.particles()
    dbg_out = sim.debug()
    c, o, ids = sim.neighbors()
    ids = sorted_rows(c, o, ids)
    bxj, bvol = sim.boundary(0)
    out = dict(pos0=pos, vel0=np.zeros_like(pos), state_in=state_in, dt_in=np.float32(dbg_in["dt"]), st_in=st_in,
               state_out=state_out, dt_out=np.float32(dbg_out["dt"]),
               its_out=np.array([dbg_out["div_it"], dbg_out["press_it"], dbg_out["visc_it"]], np.uint32),
               visc_err_out=np.float32(dbg_out["visc_err"]), max_vel2_out=np.float32(dbg_out["max_vel2"]),
               nbr_counts=c, nbr_offsets=o, nbr_ids=ids, boundary_xj=bxj, boundary_vol=bvol,
               kernel=sim.kernel_bytes(), info_out=sim.info_bytes())
    # The reference's own reproducibility floor for this step: its neighbour order (atomic arrival order,
    # ParticleSearchKernels.cu:77) and reduction order change from run to run when it runs in parallel; the
    # same step is repeated 16 times with 8 OpenMP threads and the worst deviation from the serial run is recorded
    # per field (relative to the field's scale).  The parity tests accept max(1e-5, 3 x this): the statistic is
    # a maximum over particles of a PCG-amplified rounding difference, so a single further sample (ours) sits at
    # 1-2 x the recorded maximum about as often as the reference's own next run would.
    import parity
    noise = {f: 0.0 for f in parity.ALL_FIELDS}
    for _ in range(16):
        par = RefSim(desc, serial=False, threads=8)
        par.set_particles(pos)
        par.add_box_body(sc["box"][0], sc["box"][1], inverted=True, padding=0.0, res=sc["res"])
        par.commit_bodies()
        par.set_particles_full(state_in)
        par.set_time_step(float(dbg_in["dt"]))
        par.set_st_state(int(st_in[0]), float(st_in[1]))
        par.step(1)
        for f, v in parity.field_errors(par.particles(), state_out).items():
            noise[f] = max(noise[f], float(v[0]))
    out["noise_fields"] = np.array(parity.ALL_FIELDS)
    out["noise_values"] = np.array([noise[f] for f in parity.ALL_FIELDS], np.float64)
    for k, v in vm.items():
        out["map_" + k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "n =", len(pos), "its", out["its_out"], "mean nbrs %.2f max %d" % (c.mean(), c.max()),
          "boundary particles", int((bvol > 0).sum()), "->", os.path.getsize(path) // 1024, "kB")


def make_halton():
    """sha256 + probes of the reference's 16 384-point Halton sphere table (its data file HaltonVec323.cuh, parsed as
    text and narrowed to fp32 exactly as a C compiler narrows the literals): pins vfd_halton_table_build."""
    import hashlib
    import json
    import re
    ref = os.environ.get("VFD_REFERENCE", "/root/reference")
    txt = open(os.path.join(ref, "VFD", "Source", "Simulation", "DFSPH", "HaltonVec323.cuh")).read()
    body = txt[txt.index("{") + 1:txt.rindex("}")]
    vals = re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", body)
    t = np.array([float(v) for v in vals], dtype=np.float64).astype(np.float32)
    assert t.shape == (49152,)
    probe = [0, 1, 2, 3, 4, 5, 3 * 1000, 3 * 1000 + 1, 3 * 1000 + 2, 49149, 49150, 49151]
    out = dict(count=int(t.size), sha256_f32le=hashlib.sha256(t.astype("<f4").tobytes()).hexdigest(),
               probe_index=probe, probe_value=[float(t[i]) for i in probe])
    with open(os.path.join(HERE, "halton_ref.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("halton_ref.json", out["sha256_f32le"])


if __name__ == "__main__":
    names = sys.argv[1:] or (list(scenes.SCENES) + ["halton"])
    for n in names:
        make_halton() if n == "halton" else make(n)
