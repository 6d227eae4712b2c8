"""tools_ablate.py — ms per launch of the pipelined PCG mat-vec on the settled 1M scene (tuning aid; run on the GPU box
with VFD_LIB pointing at a PIPE_ABLATE build to see what each part of the pass costs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from vfd_b200 import api
side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pos, box, res = bench.scene(side)
vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=bench.R, device=0)
sim = api.DFSPHSimulation(bench.description(api.DFSPHSimulationDescription), device=0)
sim.SetFluidObjects([api.FluidObject(pos)])
sim.SetRigidBodies([vm])
state = os.environ.get("VFD_ABLATE_STATE")
if state and os.path.exists(state):
    import numpy as np
    d = np.load(state, allow_pickle=False)
    sim.set_particles_full(d["state"]); sim.set_time_step(float(d["dt"])); sim.set_surface_tension_state(int(d["st"][0]), float(d["st"][1]))
    sim.steps(1)
else:
    sim.steps(200)
    if state:
        import numpy as np
        info = sim.GetInfo()
        np.savez(state, state=sim.particles(), dt=info.TimeStepSize, st=np.array([info.SurfaceTensionSampleCount, info.MonteCarloFactor]))
sim.synchronize()
print("[%s] matvec %.4f ms/launch" % (os.environ.get("VFD_LIB", "default"), sim.time_matvec(40)))
