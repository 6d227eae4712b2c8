"""N-rank correctness of the slab decomposition on real GPUs: tests/dist_check.py under torch.distributed.run.

The reference has no multi-GPU path: the obligation is that an N-rank run of this library reproduces its
own 1-rank run, which the single-GPU tests pin to the reference.  Skipped on boxes with fewer GPUs than ranks."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,mode,p2p", [(2, "balanced", "1"), (2, "unbalanced", "1"), (2, "unbalanced", "0"), (4, "unbalanced", "1")])
def test_n_ranks_reproduce_one_rank(world, mode, p2p, lib_built):
    """balanced: the planner's slabs; unbalanced: a lopsided start with a boundary move allowed every step, so tile columns
    change hands while the solver runs.  p2p 1: halos and all-reduces through peer memory (the default), 0: NCCL only."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, this box has %d" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check.py"), "10" if mode == "unbalanced" else "6", mode]
    env = dict(os.environ, VFD_DIST_P2P=p2p, VFD_DIST_REBALANCE="1" if mode == "unbalanced" else "4")
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-2000:]
