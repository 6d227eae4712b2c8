# lane utilisation of the pair passes on the bench scene: one particle per lane, a warp runs max-over-lanes groups of four
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from vfd_b200 import api
side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pos, box, res = bench.scene(side)
vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=bench.R)
sim = api.DFSPHSimulation(bench.description(api.DFSPHSimulationDescription))
sim.SetFluidObjects([api.FluidObject(pos)]); sim.SetRigidBodies([vm]); sim.steps(200); sim.synchronize()
x = sim.particles()["Position"].astype(np.float64)
counts, _, _ = sim.neighbors()
h = 0.1 * (1 + 1 / 1023.0)
c = np.floor((x - x.min(0)) / h).astype(np.int64) + 2
t = c // 4
tile = (t[:, 0] * 1000 + t[:, 1]) * 1000 + t[:, 2]
cell = ((c[:, 2] & 3) << 4) | ((c[:, 1] & 3) << 2) | (c[:, 0] & 3)
order = np.lexsort((np.arange(len(x)), cell, tile))
m = counts[order].astype(np.int64); tl = tile[order]
g = (m + 3) // 4
# batches: 32 consecutive particles of a tile
start = np.r_[0, np.nonzero(np.diff(tl))[0] + 1, len(tl)]
def stats(gs, name):
    tot_slots = 0; nb = 0
    for a, b in zip(start[:-1], start[1:]):
        gg = gs(g[a:b])
        n = len(gg); pad = (-n) % 32
        gg = np.r_[gg, np.zeros(pad, np.int64)].reshape(-1, 32)
        tot_slots += int(gg.max(1).sum()) * 32 * 4; nb += gg.shape[0]
    print("%-28s batches %d  warp slots %d  real pairs %d  lane utilisation %.3f  groups per batch %.2f" % (name, nb, tot_slots, m.sum(), m.sum() / tot_slots, tot_slots / 128 / nb))
stats(lambda v: v, "cell order (now)")
stats(lambda v: np.sort(v)[::-1], "sorted by count in tile")
print("mean m %.2f; tiles %d; mean batch fill %.3f" % (m.mean(), len(start) - 1, len(m) / (32.0 * sum((b - a + 31) // 32 for a, b in zip(start[:-1], start[1:])))))
print("histogram of m (bins of 4):", np.bincount(m // 4))
