"""tests/ref_gpu_timing.py — TEST INFRASTRUCTURE (a script, not collected by pytest): the reference's own CUDA solver (its sources compiled by `oracle/build_ref.py --gpu` for sm_100a, as
shipped with --use_fast_math and in the IEEE variant) timed on the B200 on bench.py's workload: K steps of the settled
1M-particle dam break.  It is the only pre-existing GPU implementation of the path: a reported number,
recorded under profiles/, not part of bench.py.  Usage (GPU box): python tests/ref_gpu_timing.py [steps] [side]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from oracle import refsim

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
side = int(sys.argv[2]) if len(sys.argv) > 2 else 100


class Args:
    pass


a = Args(); a.side = side; a.settle = 200
state, dt0, st0, pos, box, res = bench.settled_state_on_gpu(a, 0)
out = {}
for kind in ("gpu_fast", "gpu"):
    if not refsim.available(kind):
        out[kind] = "not built"
        continue
    with refsim.quiet_stdout():
        sim = refsim.RefSim(bench.description(refsim.Desc), kind=kind)
        sim.set_particles(pos)
        sim.add_box_body(box[0], box[1], inverted=True, padding=0.0, res=res)
        sim.commit_bodies()
        sim.set_particles_full(state)
        sim.set_time_step(dt0)
        sim.set_st_state(*st0)
        sim.step(2)
        t0 = time.perf_counter()
        its = []
        for _ in range(steps):
            sim.step(1)
            its.append(sim.debug()["visc_it"])
        sim.particles()        # device -> host read: the steps have completed
        el = time.perf_counter() - t0
    out[kind] = {"ms_per_step": 1e3 * el / steps, "particle_steps_per_s": len(pos) * steps / el, "pcg_it_mean": float(np.mean(its)), "steps": steps, "particles": len(pos)}
print(json.dumps({"reference_cuda_on_this_gpu": out}))
