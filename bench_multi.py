"""bench_multi.py — the N-GPU leg of bench.py (one process per GPU, launched by torch.distributed.run).

Weak scaling: every rank contributes side^3 particles; together they form one block N*side x side x side that fills
a closed tank along x (so the slabs stay balanced without re-partitioning), DFSPH + viscosity + surface tension as in
the 1-GPU workload.  The domain is cut into slabs of tile columns (vfd_b200/partition.py plans them from the global
histogram of particles per column); each step exchanges migrants/ghosts, halos and the solver's scalars over NCCL
inside libvfd_dfsph.so (csrc/distributed.cu).  torch.distributed is only the launcher plumbing here: rendezvous,
broadcast of the NCCL id, MAX over ranks of the device-timed region.
"""
import json
import os
import time

import numpy as np

R, D, H = 0.025, 0.05, 0.1


def dist_scene(side, world, kind="dam"):
    """The global block of world*side x side x side particles: "dam" — bench.scene(side, world): the 1-GPU dam break repeated
    along x (it collapses along z: at world = 1 the very same scene, at any world size the same per-GPU physics); "tank" —
    filling a closed tank along x."""
    nx = side * world
    if kind == "dam":
        L = side * D
        box = ((0.0, 0.0, 0.0), (nx * D + 4 * D, 1.4 * L + 4 * D, 2 * L + 4 * D))       # = bench.scene(side, world)'s box
    else:
        box = ((0.0, 0.0, 0.0), (nx * D + 4 * D, 1.4 * side * D + 4 * D, side * D + 4 * D))
    ext = [b - a + 2 * (8 * H - R) for a, b in zip(*box)]
    res = tuple(min(256, max(8, int(np.ceil(e / (4 * H))))) for e in ext)
    return nx, box, res


def rank_positions(side, world, rank, strong=False):
    """This rank's share of the lattice and its global ids: weak scaling — the x-range [rank*side, (rank+1)*side) of a block of
    world*side x side x side; strong scaling (BASELINE.json config 4) — the x-range [rank*side/world, (rank+1)*side/world) of ONE
    side^3 block, whatever the number of GPUs."""
    from vfd_b200 import api
    if strong:
        lo, hi = rank * side // world, (rank + 1) * side // world
        pos = api.block_positions(hi - lo, side, side, R, origin=(2 * D + lo * D, 2 * D, 2 * D))
        ids = (np.arange(len(pos), dtype=np.uint64) + np.uint64(lo) * np.uint64(side * side)).astype(np.uint32)
        return pos, ids
    pos = api.block_positions(side, side, side, R, origin=(2 * D + rank * side * D, 2 * D, 2 * D))
    ids = (np.arange(len(pos), dtype=np.uint64) + np.uint64(rank) * np.uint64(side ** 3)).astype(np.uint32)
    return pos, ids


def setup(args, rank, local, world, description, frames=0):
    """Creates this rank's solver: communicator, grid, slab plan from the global column histogram, particles, tank."""
    import torch
    import torch.distributed as dist
    from vfd_b200 import api, partition
    strong = bool(getattr(args, "strong", False))
    # strong scaling: the ONE side^3 scene (= the 1-GPU scene of that size) cut into `world` slabs
    nx, box, res = dist_scene(args.side, 1 if strong else world, getattr(args, "scene", "dam"))
    sim = api.DFSPHSimulation(description(api.DFSPHSimulationDescription, frames=frames), device=local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(api.dist_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    sim.init_distributed(rank, world, uid.cpu().numpy().tobytes(), box[0], box[1])
    origin, cell, tiles = sim.grid()
    pos, ids = rank_positions(args.side, world, rank, strong)
    hist = torch.from_numpy(partition.column_histogram(pos[:, 0], origin[0], H, tiles[0])).cuda()
    dist.all_reduce(hist)
    bounds = partition.plan_slabs(hist.cpu().numpy(), world)
    # every rank generated the particles of "its" x-range; hand over the few that the plan assigns elsewhere
    owner = partition.owner_of(partition.tile_columns(pos[:, 0], origin[0], H, tiles[0]), bounds)
    pos, ids = redistribute(pos, ids, owner, rank, world)
    sim.set_slab(int(bounds[rank]), int(bounds[rank + 1]))
    n_global = args.side ** 3 if strong else world * args.side ** 3
    ghost = int(tiles[1]) * int(tiles[2]) * 64 * 12 * 2
    sim.set_particles_distributed(pos, None, ids, n_global, int(1.5 * len(pos)) + 2 * ghost)
    vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=R, device=local)
    sim.SetRigidBodies([vm])
    return sim, n_global, bounds, hist.cpu().numpy()


def redistribute(pos, ids, owner, rank, world):
    """All-to-all of the particles whose planned owner is another rank (host side, once at start-up)."""
    import torch
    import torch.distributed as dist
    rec = np.concatenate([pos.astype(np.float32), ids.view(np.float32).reshape(-1, 1)], axis=1)
    send = [torch.from_numpy(np.ascontiguousarray(rec[owner == r])).cuda() for r in range(world)]
    counts = torch.tensor([len(s) for s in send], dtype=torch.int64, device="cuda")
    allc = [torch.zeros(world, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(allc, counts)
    recv = [torch.zeros((int(allc[r][rank]), 4), dtype=torch.float32, device="cuda") for r in range(world)]
    dist.all_to_all(recv, send)
    got = torch.cat(recv).cpu().numpy()
    return np.ascontiguousarray(got[:, :3]), np.ascontiguousarray(got[:, 3]).view(np.uint32)


def main(args, rank, local, world):
    import torch
    import torch.distributed as dist
    import bench
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    sim, n_global, bounds, hist = setup(args, rank, local, world, bench.description)
    sim.steps(args.settle)
    clk = bench.ClockSampler(local).start()
    sim.steps(args.warmup)
    sim.synchronize()
    sim.launch_count(reset=True)
    stats0 = sim.comm_stats()
    dist.barrier()
    torch.cuda.synchronize()
    with clk:
        sim.record_event(0)
        for _ in range(args.steps):
            sim.OnUpdate()
        sim.record_event(1)
        ms = sim.elapsed_ms(0, 1)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    stats1 = sim.comm_stats()
    launches = sim.launch_count()
    dbg = sim.GetDebugInfo()
    ids, _ = sim.owned()
    own = torch.tensor([len(ids)], dtype=torch.int64, device="cuda")
    owns = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(owns, own)
    # roofline of rank 0's dominant kernel class: a second pass with per-launch event brackets (every rank steps along:
    # the halo exchanges and all-reduces are collective)
    from vfd_b200 import api as api_mod
    try:
        counts, _, _ = sim.neighbors()
        mbar = float(counts[counts > 0].mean())
    except Exception:
        mbar = 36.0
    if rank == 0:
        roof, table, _ = bench.roofline_pass(sim, api_mod, len(ids), mbar, args.steps)
    else:
        sim.steps(args.steps)
        sim.synchronize()
        roof, table = None, {}
    dist.barrier()
    if rank == 0:
        value = n_global * args.steps / (ms_max * 1e-3)
        per_step = {k: (stats1[k] - stats0[k]) / args.steps for k in stats1}
        out = {
            "metric": "DFSPH particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if getattr(args, "strong", False) else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": bench.workload(args.side, 1 if getattr(args, "strong", False) else world, args.settle) if args.scene == "dam" else
                                   "closed tank, %d x %d^3 = %d particles, DFSPH (2+2 Jacobi iterations) + implicit viscosity PCG (nu 10) + surface tension; %d settle steps" % (world, args.side, n_global, args.settle),
                       "notes": ("strong scaling: ONE %d^3 scene cut into slabs; " % args.side if getattr(args, "strong", False) else "") + "%d^3 particles per GPU, the 1-GPU scene repeated along x (it collapses along z); slabs of tile columns along x re-balanced while stepping, halos and "
                                "all-reduces through peer memory (%s); working set > L2, no flush" % (args.side, "CUDA IPC over NVLink" if sim.slab()["peer_memory"] else "off: NCCL only"),
                       "baseline_config": args.config,
                       "slab_rank0_now": sim.slab(),
                       "particles": n_global, "slab_bounds_tile_columns": [int(b) for b in bounds], "owned_per_rank": [int(o.item()) for o in owns],
                       "pcg_iterations_last_step": int(dbg.ViscositySolverIterationCount),
                       "per_step_rank0": {"halo_exchanges": per_step["halos"], "all_reduces": per_step["reductions"],
                                          "halo_bytes_sent": per_step["halo_bytes"], "state_bytes_sent": per_step["state_bytes"]},
                       "kernels_rank0": table},
            "gpu_launches": int(launches), "clocks": clk.summary(),
            "roofline": roof, "cpu_baseline": None,     # cpu_baseline: reported at N = 1 only
            "e2e": None,                                   # end to end through host buffers is measured at N = 1 (bench.py)
        }
        print(json.dumps(out))
    clk.stop()
    dist.barrier()
    dist.destroy_process_group()
    return 0
