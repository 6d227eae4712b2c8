import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import bench, parity
from vfd_b200 import api
from oracle import refsim
side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pos, box, res = bench.scene(side)
vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=bench.R)
sim = api.DFSPHSimulation(bench.description(api.DFSPHSimulationDescription))
sim.set_option(api.VFD_OPT_SEARCH_FMA, 0)
sim.SetFluidObjects([api.FluidObject(pos)]); sim.SetRigidBodies([vm])
sim.steps(200); sim.synchronize()
state = sim.particles(); info = sim.GetInfo()
sim.OnUpdate(); out = sim.particles()
counts, offs, ids = sim.neighbors()
print("neighbours: mean %.2f max %d, at the cap: %d" % (counts.mean(), counts.max(), (counts >= 70).sum()))
with refsim.quiet_stdout():
    ref = refsim.RefSim(bench.description(refsim.Desc), threads=os.cpu_count())
    ref.set_particles(pos); ref.add_box_body(box[0], box[1], inverted=True, padding=0.0, res=res); ref.commit_bodies()
    ref.set_particles_full(state); ref.set_time_step(info.TimeStepSize); ref.set_st_state(int(info.SurfaceTensionSampleCount), float(info.MonteCarloFactor))
    ref.step(1)
    want = ref.particles()
    rc, ro, ri = ref.neighbors()
print("reference neighbours: mean %.2f max %d at cap %d; count mismatches %d" % (rc.mean(), rc.max(), (rc >= 70).sum(), (rc != counts).sum()))
for f in ("PressureAcceleration", "Velocity", "Density", "Position", "Acceleration"):
    x, y = np.asarray(out[f], np.float64), np.asarray(want[f], np.float64)
    d = np.abs(x - y); d = d.max(axis=1) if d.ndim > 1 else d
    sc = np.abs(y).max()
    order = np.argsort(-d)[:5]
    print(f, "scale %.4g worst %.3e  p99.99 %.3e  p99.9 %.3e; worst particles" % (sc, d.max() / sc, np.percentile(d, 99.99) / sc, np.percentile(d, 99.9) / sc), [(int(i), int(counts[i]), int(rc[i])) for i in order])
# the same step with the REFERENCE's volume map fed to our solver (what the parity tests do)
m = ref.volume_map(0)
g = api.DFSPHSimulation(bench.description(api.DFSPHSimulationDescription))
g.set_option(api.VFD_OPT_SEARCH_FMA, 0)
g.SetFluidObjects([api.FluidObject(pos)])
g.SetRigidBodies([api.VolumeMap(m["domain_min"], m["domain_max"], m["resolution"], m["cell_size"], m["cell_size_inv"], m["field_count"],
                                m["node_count"], m["cell_count"], m["cell_map_count"], m["nodes"], m["cells"], m["cell_map"])])
g.set_particles_full(state); g.set_time_step(info.TimeStepSize); g.set_surface_tension_state(int(info.SurfaceTensionSampleCount), float(info.MonteCarloFactor))
g.OnUpdate(); out2 = g.particles()
print("with the reference's own volume map:")
print(parity.format_errors(parity.field_errors(out2, want)))
n = int(m["node_count"])
ours, theirs = np.asarray(vm.nodes, np.float64).reshape(2, -1), np.asarray(m["nodes"], np.float64).reshape(2, n)
print("map: resolution ours %s theirs %s; node counts %d %d" % (list(vm.resolution), list(m["resolution"]), ours.shape[1], n))
if ours.shape == theirs.shape:
    fin = np.abs(theirs[0]) < 1e30
    print("distance field max abs err %.3e; volume field max err %.3e of scale %.3g" % (np.abs(ours[0] - theirs[0])[fin].max(), np.abs(ours[1] - theirs[1]).max() / np.abs(theirs[1]).max(), np.abs(theirs[1]).max()))
    i = np.argmax(np.abs(ours[1] - theirs[1])); print("worst volume node", i, ours[1][i], theirs[1][i], "distance there", ours[0][i], theirs[0][i])
