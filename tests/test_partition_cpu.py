"""Host-side logic of the slab decomposition (vfd_b200/partition.py), single process and 2 ranks over gloo.

The N-GPU path cuts the domain into slabs of tile columns along x; the plan is computed from a histogram of particles per
tile column that every rank builds for its own particles and all-reduces (bench_multi.setup).  These tests run the same
sequence on CPU: histogram -> all_reduce -> plan -> owner -> all-to-all hand-over, and check that both ranks derive the
same plan, that every particle ends up on exactly one rank, and that the device's cell arithmetic is reproduced."""
import os
import socket

import numpy as np
import pytest

from vfd_b200 import partition

R, D, H = 0.025, 0.05, 0.1


def lattice(nx, ny, nz, origin):
    i = (np.arange(nx, dtype=np.float32) + np.float32(0.5)) * np.float32(D) + np.float32(origin[0])
    j = (np.arange(ny, dtype=np.float32) + np.float32(0.5)) * np.float32(D) + np.float32(origin[1])
    k = (np.arange(nz, dtype=np.float32) + np.float32(0.5)) * np.float32(D) + np.float32(origin[2])
    Z, Y, X = np.meshgrid(k, j, i, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)


def test_plan_is_balanced_and_contiguous():
    rng = np.random.RandomState(0)
    hist = rng.randint(0, 5000, size=64)
    for world in (1, 2, 3, 4, 8):
        b = partition.plan_slabs(hist, world)
        assert b[0] == 0 and b[-1] == len(hist) and len(b) == world + 1
        assert np.all(np.diff(b) >= 1)
        # no slab is more than one column's population above the ideal share
        cum = np.concatenate([[0], np.cumsum(hist)])
        pop = np.diff(cum[b])
        assert pop.max() <= hist.sum() / world + hist.max()
        assert partition.imbalance(hist, b) < 1.5


def test_plan_rejects_more_ranks_than_columns():
    with pytest.raises(ValueError):
        partition.plan_slabs(np.ones(3, np.int64), 4)


def test_degenerate_histograms():
    # all particles in one column: slabs stay one column wide at least, the full range is covered
    hist = np.zeros(16, np.int64)
    hist[5] = 1000
    b = partition.plan_slabs(hist, 4)
    assert b[0] == 0 and b[-1] == 16 and np.all(np.diff(b) >= 1)
    # empty scene
    b = partition.plan_slabs(np.zeros(8, np.int64), 2)
    assert b[0] == 0 and b[-1] == 8 and np.all(np.diff(b) >= 1)


def test_tile_columns_follow_the_device_arithmetic():
    # cell = h / (1 - 1/1024) in fp32 (tile.cuh cell_inv); column = clamp(floor((x - origin) * invCell), 1, 4*tiles - 2) >> 2
    origin, tiles = np.float32(-0.2), 12
    x = np.linspace(-1.0, 6.0, 20001).astype(np.float32)
    got = partition.tile_columns(x, origin, H, tiles)
    inv = (np.float32(1.0) / np.float32(H)) * (np.float32(1.0) - np.float32(1.0) / np.float32(1024.0))
    want = np.clip(np.floor((x - origin) * inv).astype(np.int64), 1, 4 * tiles - 2) >> 2
    assert np.array_equal(got, want)
    assert got.min() == 0 and got.max() == tiles - 1
    assert np.all(np.diff(got) >= 0)
    assert partition.column_histogram(x, origin, H, tiles).sum() == len(x)


def test_owner_of_maps_every_column_to_one_rank():
    b = np.asarray([0, 3, 7, 8, 12])
    own = partition.owner_of(np.arange(12), b)
    assert own.tolist() == [0, 0, 0, 1, 1, 1, 1, 2, 3, 3, 3, 3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, side, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        # the sequence of bench_multi.setup / redistribute with CPU tensors
        origin_x, tiles_x = np.float32(-2 * H), (side * world * 2 + 40) // 8 + 2
        pos = lattice(side, 6, 6, (2 * D + rank * side * D, 2 * D, 2 * D))
        if rank == 0:                     # skew the scene: rank 0 generates twice as many particles
            pos = np.concatenate([pos, lattice(side, 6, 6, (2 * D + world * side * D, 2 * D, 2 * D))])
        ids = (np.arange(len(pos), dtype=np.uint32) + np.uint32(rank * 1000000))
        hist = torch.from_numpy(partition.column_histogram(pos[:, 0], origin_x, H, tiles_x))
        dist.all_reduce(hist)
        bounds = partition.plan_slabs(hist.numpy(), world)
        owner = partition.owner_of(partition.tile_columns(pos[:, 0], origin_x, H, tiles_x), bounds)
        rec = np.concatenate([pos, ids.view(np.float32).reshape(-1, 1)], axis=1)
        send = [torch.from_numpy(np.ascontiguousarray(rec[owner == r])) for r in range(world)]
        counts = torch.tensor([len(s) for s in send], dtype=torch.int64)
        allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allc, counts)
        # gloo has no all_to_all on CPU tensors in every build: pairwise send/recv in a fixed order
        recv = [torch.zeros((int(allc[r][rank]), 4), dtype=torch.float32) for r in range(world)]
        recv[rank] = send[rank]
        for a in range(world):
            for b in range(world):
                if a == b:
                    continue
                if rank == a:
                    dist.send(send[b], dst=b)
                elif rank == b:
                    dist.recv(recv[a], src=a)
        got = torch.cat(recv).numpy()
        mine_pos, mine_ids = got[:, :3], np.ascontiguousarray(got[:, 3]).view(np.uint32)
        cols = partition.tile_columns(mine_pos[:, 0], origin_x, H, tiles_x)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), bounds=bounds, ids=mine_ids, cols=cols, hist=hist.numpy(),
                 generated=len(pos))
    finally:
        dist.destroy_process_group()


def test_two_ranks_agree_on_the_plan_and_partition_the_particles(tmp_path):
    import torch.multiprocessing as mp
    world, side = 2, 24
    port = _free_port()
    mp.spawn(_rank_main, args=(world, port, side, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]
    # same plan and same global histogram on both ranks
    assert np.array_equal(r[0]["bounds"], r[1]["bounds"])
    assert np.array_equal(r[0]["hist"], r[1]["hist"])
    bounds = r[0]["bounds"]
    total = int(r[0]["generated"]) + int(r[1]["generated"])
    assert int(r[0]["hist"].sum()) == total
    # every particle on exactly one rank, inside that rank's slab
    all_ids = np.concatenate([r[0]["ids"], r[1]["ids"]])
    assert len(all_ids) == total and len(np.unique(all_ids)) == total
    for k in range(world):
        assert np.all((r[k]["cols"] >= bounds[k]) & (r[k]["cols"] < bounds[k + 1]))
    # the skewed scene (2:1) is rebalanced to within one column's population
    pops = np.asarray([len(r[k]["ids"]) for k in range(world)], np.float64)
    assert pops.max() / pops.mean() < 1.15
