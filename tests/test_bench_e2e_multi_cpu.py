"""bench_multi.py's end-to-end leg at N > 1 without a GPU: two parent ranks over gloo (as under torch.distributed.run), a
stand-in for the decomposed solver (tests/fake_dist_api.py) in the children.  Checks the plumbing: state files, one child per
rank, the file rendezvous, strays re-homed to the rank whose slab holds them, first and second bake, the result on rank 0 only
— and that a child that hangs is killed while the first bake's figure survives."""
import json
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
H, D = 0.1, 0.05


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeSim:
    """What the parent asks of its (real) solver: the owned particles and the slab bounds."""

    def __init__(self, rank):
        from vfd_b200 import api
        # columns 0..3 belong to rank 0, 4..7 to rank 1 (tile column = 4 cells of ~0.1 from x = -0.2); rank 0 still holds two
        # particles that have crossed into rank 1's first column, rank 1 one that went the other way
        xs = {0: [0.05, 0.4, 0.9, 1.35, 1.45, 1.5], 1: [1.45, 1.9, 2.4, 1.38]}[rank]
        self.part = np.zeros(len(xs), api.PARTICLE_DTYPE)
        self.part["Position"][:, 0] = xs
        self.part["Velocity"][:, 1] = -1.0
        self.ids = (np.arange(len(xs)) + 100 * rank).astype(np.uint32)
        self.rank = rank

    def owned(self):
        return self.ids, self.part

    def slab(self):
        return {"lo": 0, "hi": 4, "shifts": 0, "peer_memory": False} if self.rank == 0 else {"lo": 4, "hi": 8, "shifts": 0, "peer_memory": False}


def _parent(rank, world, port, outdir, hang, timeout_s):
    import argparse
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      VFD_E2E_API="fake_dist_api", FAKE_DIST_DIR=outdir, PYTHONPATH=HERE + os.pathsep + ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    if hang is not None:
        os.environ["FAKE_DIST_HANG_RANK"] = str(hang)
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    import bench_multi
    bench_multi.E2E_TIMEOUT_S = timeout_s
    args = argparse.Namespace(steps=4, config=3, side=10, scene="dam", strong=False)
    out = bench_multi.e2e_children(args, rank, rank, world, FakeSim(rank), 10)
    with open(os.path.join(outdir, "parent_%d.json" % rank), "w") as f:
        json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


def _run(tmp_path, hang=None, timeout_s=120):
    import torch.multiprocessing as mp
    mp.spawn(_parent, args=(2, _free_port(), str(tmp_path), hang, timeout_s), nprocs=2, join=True)
    return [json.load(open(os.path.join(str(tmp_path), "parent_%d.json" % r))) for r in range(2)]


def test_children_bake_twice_and_rank_0_reports(tmp_path):
    r0, r1 = _run(tmp_path)
    assert r1 is None
    assert "error" not in r0, r0
    assert r0["rebake_identical"] is True and r0["bake"].startswith("second bake") and r0["n_gpus"] == 2
    assert r0["value"] > 0 and r0["unit"] == "particle-steps/s" and r0["d2h_bytes_per_step"] == 360 and r0["h2d_bytes_per_step"] == 28 * 10 // 4
    assert abs(r0["value"] - 10 * 4 / (r0["ms_per_step"] * 4e-3)) < 1e-6 * r0["value"]
    # every particle uploaded exactly once, by the rank whose slab holds it — also in the second bake, after the stand-in moved
    # the boundary one column to the left (what re-balancing does to a real handle)
    up = [[json.load(open(os.path.join(str(tmp_path), "uploaded_%d_%d.json" % (r, bake)))) for r in range(2)] for bake in (1, 2)]
    assert up[0][0] == [0, 1, 2, 3, 103] and up[0][1] == [4, 5, 100, 101, 102]
    assert up[1][0] == [0, 1, 2] and up[1][1] == [3, 4, 5, 100, 101, 102, 103]


def test_a_hanging_child_is_killed_and_the_first_bake_survives(tmp_path):
    r0, r1 = _run(tmp_path, hang=1, timeout_s=8)
    assert r1 is None
    assert "error" not in r0, r0
    assert r0["bake"].startswith("first bake") and r0["children_failed"] >= 1 and r0["value"] > 0
