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


@pytest.mark.parametrize("world", [2, 4])
def test_n_ranks_reproduce_one_rank(world, lib_built):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, this box has %d" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check.py"), "6"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-2000:]
