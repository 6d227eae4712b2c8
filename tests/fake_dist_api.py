"""TEST INFRASTRUCTURE: a stand-in for vfd_b200.api with just the calls bench_multi's end-to-end children make, and no GPU
behind it — so that the plumbing around those calls (parent ranks -> state files -> one child per rank -> file rendezvous ->
re-homing of strays -> two bakes -> result file -> the parents' JSON) runs in the CPU suite (tests/test_bench_e2e_multi_cpu.py;
selected with VFD_E2E_API=fake_dist_api).  It checks what a decomposed solver would insist on: a 128-byte id that is the same
on every rank, every uploaded particle inside the rank's slab, the same global count everywhere."""
import json
import os
import time

import numpy as np

from vfd_b200 import partition

H = 0.1
PARTICLE_SIMPLE_DTYPE = np.dtype([("Position", "<f4", 3), ("Velocity", "<f4", 3), ("Acceleration", "<f4", 3)])


class DFSPHSimulationDescription:
    def __init__(self, **kw):
        self.kw = kw


class VolumeMap:
    @staticmethod
    def build_box(lo, hi, **kw):
        return ("box", tuple(lo), tuple(hi))


def dist_unique_id():
    return bytes(range(128))


class DFSPHSimulation:
    def __init__(self, desc, device=0):
        self.frames = int(desc.kw["FrameCount"])
        self.bakes = 0

    def init_distributed(self, rank, world, uid, lo, hi):
        assert bytes(uid) == bytes(range(128))
        self.rank, self.world = rank, world

    def grid(self):
        return np.array([-0.2, -0.2, -0.2], np.float32), H, np.array([8, 3, 3], np.uint32)

    def set_slab(self, lo, hi):
        assert 0 <= lo < hi <= 8
        self.lo, self.hi = lo, hi

    def slab(self):
        return {"lo": self.lo, "hi": self.hi, "shifts": self.bakes, "peer_memory": False}

    def set_particles_distributed(self, pos, vel, ids, n_global, capacity):
        origin, _, tiles = self.grid()
        cols = partition.tile_columns(pos[:, 0], origin[0], H, tiles[0])
        assert np.all((cols >= self.lo) & (cols < self.hi)), "a particle outside the rank's slab was uploaded"
        assert pos.dtype == np.float32 and vel.dtype == np.float32 and ids.dtype == np.uint32 and capacity >= len(pos)
        self.n, self.n_global, self.ids = len(pos), int(n_global), ids.copy()

    def SetRigidBodies(self, maps):
        assert maps and maps[0][0] == "box"

    def Simulate(self):
        time.sleep(0.02)
        self.bakes += 1
        if os.environ.get("FAKE_DIST_HANG_RANK") == str(self.rank) and self.bakes == 2:
            time.sleep(3600)                       # a child that never comes back from its second bake
        with open(os.path.join(os.environ["FAKE_DIST_DIR"], "uploaded_%d_%d.json" % (self.rank, self.bakes)), "w") as f:
            json.dump(sorted(int(i) for i in self.ids), f)
        if self.bakes == 1 and self.world == 2:     # the bake re-balanced: the boundary between the two slabs moved one column to the left
            if self.rank == 0:
                self.hi -= 1
            else:
                self.lo -= 1

    def GetFrame(self, index):
        assert index == self.frames - 1 and self.rank == 0
        out = np.zeros(self.n_global, PARTICLE_SIMPLE_DTYPE)
        out["Position"][:, 0] = np.arange(self.n_global)
        return out, 0.0, 0.001

    def synchronize(self):
        pass
