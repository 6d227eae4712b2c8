# e2e bake timing, repeated: set_particles + set_rigid_bodies + simulate (every step a frame) + get_frame, wall clock
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from vfd_b200 import api
bench.CONFIG_NAME = bench.CONFIGS[3]["name"]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pos, box, res = bench.scene(100)
n = len(pos)
vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=bench.R)
sim = api.DFSPHSimulation(bench.description(api.DFSPHSimulationDescription))
sim.SetFluidObjects([api.FluidObject(pos)]); sim.SetRigidBodies([vm]); sim.steps(200); sim.synchronize()
st = sim.particles(); hp = np.ascontiguousarray(st["Position"]); hv = np.ascontiguousarray(st["Velocity"]); sim.close()
for fl in (0.0, 0.0016, 0.0, 0.0016):
    e = api.DFSPHSimulation(bench.description(api.DFSPHSimulationDescription, frames=steps, FrameLength=fl))
    out = []
    for rep in range(4):
        t0 = time.perf_counter()
        e.SetFluidObjects([api.FluidObject(hp, velocities=hv)]); t1 = time.perf_counter()
        e.SetRigidBodies([vm]); t2 = time.perf_counter()
        e.Simulate(); t3 = time.perf_counter()
        f, _, _ = e.GetFrame(steps - 1); t4 = time.perf_counter()
        out.append("%.2f (set %.0f+%.0f ms, bake %.0f ms, get %.0f ms; %d steps)" % (1e3 * (t4 - t0) / steps, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), e.GetDebugInfo().IterationCount))
    print("FrameLength", fl, "ms/step per bake:", " | ".join(out))
    e.close()
