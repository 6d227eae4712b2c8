# tile occupancy statistics of the bench scene after settling (to size CTAs / staging buffers)
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from vfd_b200 import api
pos, box, res = bench.scene(100)
vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=bench.R)
sim = api.DFSPHSimulation(bench.description(api.DFSPHSimulationDescription))
sim.SetFluidObjects([api.FluidObject(pos)]); sim.SetRigidBodies([vm]); sim.steps(200); sim.synchronize()
x = sim.particles()["Position"].astype(np.float64)
h = 0.1 * (1 + 1 / 1023.0)
c = np.floor((x - x.min(0)) / h).astype(np.int64) + 2
t = c // 4
key = (t[:, 0] * 1000 + t[:, 1]) * 1000 + t[:, 2]
u, cnt = np.unique(key, return_counts=True)
print("particles", len(x), "non-empty tiles", len(u), "mean per tile %.1f" % cnt.mean(), "median", np.median(cnt), "max", cnt.max())
print("histogram of tile population (bins of 64):", np.bincount(cnt // 64))
print("fraction of particles in tiles with >=448:", cnt[cnt >= 448].sum() / len(x), " >512:", cnt[cnt > 512].sum() / len(x), " >544:", cnt[cnt > 544].sum() / len(x))
ck = (c[:, 0] * 10000 + c[:, 1]) * 10000 + c[:, 2]
cu, cc = np.unique(ck, return_counts=True)
print("cells: mean per non-empty cell %.2f max %d" % (cc.mean(), cc.max()))
