"""tests/dist_check.py — a script run by tests/test_gpu_dist.py (and by hand) under torch.distributed.run, not collected by pytest.
N-rank vs 1-rank consistency of the slab decomposition (run under torch.distributed.run on a multi-GPU box):
the same scene is stepped K times by the distributed solver and, on rank 0, by a plain single-GPU solver; the
gathered distributed state must agree with the single-GPU state to rounding (summation orders differ)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import parity
from vfd_b200 import api, partition

R, D, H = 0.025, 0.05, 0.1
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
UNBALANCED = len(sys.argv) > 2 and sys.argv[2] == "unbalanced"      # start from a lopsided plan: the slabs re-balance while stepping
dims = (48, 14, 12)
box = ((0.0, 0.0, 0.0), (dims[0] * D + 6 * D, 1.6 * dims[1] * D, dims[2] * D + 4 * D))
rng = np.random.RandomState(7)
pos_all = api.block_positions(*dims, R, origin=(2 * D, 2 * D, 2 * D))
pos_all = (pos_all + rng.uniform(-0.2 * R, 0.2 * R, pos_all.shape)).astype(np.float32)
vel_all = (rng.uniform(-0.3, 0.3, pos_all.shape)).astype(np.float32)          # non-uniform velocities: the PCG iterates
n = len(pos_all)
kw = dict(FrameCount=K, FrameLength=0.0, MinPressureSolverIterations=2, MaxPressureSolverIterations=2, MinDivergenceSolverIterations=2, MaxDivergenceSolverIterations=2, CSDFix=16)
vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=(12, 8, 8), particle_radius=R, device=local)

sim = api.DFSPHSimulation(api.DFSPHSimulationDescription(**kw), device=local)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.frombuffer(bytearray(api.dist_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(uid, 0)
sim.init_distributed(rank, world, uid.cpu().numpy().tobytes(), box[0], box[1])
origin, cell, tiles = sim.grid()
cols = partition.tile_columns(pos_all[:, 0], origin[0], H, tiles[0])
bounds = partition.plan_slabs(np.bincount(cols, minlength=int(tiles[0])), world)
if UNBALANCED:
    bounds = np.asarray([0] + [2 + r for r in range(world - 1)] + [int(tiles[0])], np.int64)     # rank 0..n-2 one or two columns, the last rank the rest
mine = partition.owner_of(cols, bounds) == rank
sim.set_slab(int(bounds[rank]), int(bounds[rank + 1]))
sim.set_particles_distributed(pos_all[mine], vel_all[mine], np.nonzero(mine)[0].astype(np.uint32), n, int(mine.sum()) * 2 + 20000)
sim.SetRigidBodies([vm])
sim.Simulate()                       # K steps, every step baked: whole-scene frames gathered by persistent id on rank 0
sim.synchronize()
ids, part = sim.owned()
dbg = sim.GetDebugInfo()
stats = sim.comm_stats()
# gather on rank 0
cnt = torch.tensor([len(ids)], dtype=torch.int64, device="cuda")
cnts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
dist.all_gather(cnts, cnt)
mx = int(max(c.item() for c in cnts))
buf = np.zeros((mx, 31), np.float32)
buf[:len(ids), :30] = part.view(np.float32).reshape(-1, 30)
buf[:len(ids), 30] = ids.view(np.float32)
t = torch.from_numpy(buf).cuda()
outs = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(outs, t)
ok = True
if rank == 0:
    full = np.zeros(n, api.PARTICLE_DTYPE)
    seen = np.zeros(n, np.int32)
    for r in range(world):
        a = outs[r].cpu().numpy()[:int(cnts[r].item())]
        i = np.ascontiguousarray(a[:, 30]).view(np.uint32)
        full[i] = np.ascontiguousarray(a[:, :30]).view(api.PARTICLE_DTYPE).reshape(-1)
        seen[i] += 1
    ref = api.DFSPHSimulation(api.DFSPHSimulationDescription(**kw), device=local)
    ref.SetFluidObjects([api.FluidObject(pos_all, velocities=vel_all)])
    ref.SetRigidBodies([vm])
    ref.Simulate()
    want = ref.particles()
    frames_ok = sim.GetFrameCount() == K
    for fi in (0, K - 1):
        a, av, adt = sim.GetFrame(fi)
        b, bv, bdt = ref.GetFrame(fi)
        frames_ok = frames_ok and a.tobytes() == b.tobytes() and av == bv and adt == bdt
    print("baked frames gathered on rank 0 equal the single-GPU frames:", frames_ok)
    rdbg = ref.GetDebugInfo()
    errs = parity.field_errors(full, want)
    worst = max(v[0] for v in errs.values())
    print("ranks %d  particles %d  steps %d  every particle owned exactly once: %s" % (world, n, K, bool((seen == 1).all())))
    print("owned per rank", [int(c.item()) for c in cnts], "initial slab bounds", bounds.tolist(), "rank 0 slab now", sim.slab(), "comm", stats)
    print("PCG iterations: distributed %d, single %d; dt %.9g vs %.9g" % (dbg.ViscositySolverIterationCount, rdbg.ViscositySolverIterationCount,
                                                                          sim.GetCurrentTimeStepSize(), ref.GetCurrentTimeStepSize()))
    print(parity.format_errors(errs))
    ok = bool((seen == 1).all()) and worst < 2e-4 and frames_ok
    print("DIST_CHECK", "PASS" if ok else "FAIL", "worst %.3e" % worst)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
