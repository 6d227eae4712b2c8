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
    # end to end at N GPUs: the settled state of every rank goes through host memory into a NEW decomposed solver that bakes K
    # frames (each gathered by persistent id into a host frame on rank 0) — in child processes, one per GPU, so that nothing
    # there (a hang, a crash) can take the numbers above with it
    e2e = None
    if not getattr(args, "no_e2e", False):
        try:
            e2e = e2e_children(args, rank, local, world, sim, n_global)
        except Exception as ex:
            e2e = {"error": repr(ex)}
        try:
            dist.barrier()
        except Exception:                   # the line below needs no collective: print it whatever became of this leg
            pass
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
            "e2e": e2e,
        }
        print(json.dumps(out))
    clk.stop()
    dist.barrier()
    dist.destroy_process_group()
    return 0


# ---- end to end at N GPUs -------------------------------------------------------------------------------------------------
E2E_TIMEOUT_S = 300


def e2e_children(args, rank, local, world, sim, n_global):
    """Parent side: every rank writes its owned particles (positions, velocities, persistent ids: host memory) and the slab
    bounds to a file and starts ONE child process on its GPU (`bench_multi.py --e2e-child`); the children build their own
    decomposed solver (a new communicator; they meet through files in a scratch directory — one node, no second network
    rendezvous), bake, and rank 0's child leaves a JSON file.  Returns that JSON on rank 0 (None elsewhere); a child that fails
    or outlives E2E_TIMEOUT_S is killed and reported as an error.  Every rank runs the same sequence of collectives whatever
    goes wrong locally."""
    import shutil
    import signal
    import subprocess
    import sys
    import tempfile
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"

    def gather(values):
        t = torch.tensor(values, dtype=torch.int64, device=dev)
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [[int(v) for v in o.cpu().tolist()] for o in outs]

    port = int(os.environ.get("MASTER_PORT", "29500"))
    tmp = os.path.join(tempfile.gettempdir(), "vfd_e2e_%d" % port)
    state, result, log = os.path.join(tmp, "state_%d.npz" % rank), os.path.join(tmp, "result.json"), os.path.join(tmp, "child_%d.log" % rank)
    err, lo, hi = None, 0, 0
    try:
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
    except Exception as ex:
        err = repr(ex)
    dist.barrier()
    try:
        os.makedirs(tmp, exist_ok=True)
        ids, part = sim.owned()
        sl = sim.slab()
        lo, hi = int(sl["lo"]), int(sl["hi"])
    except Exception as ex:
        err = repr(ex)
    slabs = gather([lo, hi, 0 if err is None else 1])
    if any(f for _, _, f in slabs):
        return {"error": "preparing the end-to-end leg failed on %d rank(s): %s" % (sum(f for _, _, f in slabs), err)} if rank == 0 else None
    bounds = [l for l, _, _ in slabs] + [slabs[-1][1]]
    proc = None
    try:
        np.savez(state, pos=np.ascontiguousarray(part["Position"], np.float32), vel=np.ascontiguousarray(part["Velocity"], np.float32),
                 ids=np.ascontiguousarray(ids, np.uint32), bounds=np.asarray(bounds, np.int64), n_global=np.int64(n_global))
        env = dict(os.environ)
        env["RANK"], env["LOCAL_RANK"], env["WORLD_SIZE"] = str(rank), str(local), str(world)
        cmd = [sys.executable, os.path.abspath(__file__), "--e2e-child", state, "--out", result, "--meet", tmp, "--steps", str(args.steps),
               "--config", str(args.config), "--side", str(args.side), "--scene", str(getattr(args, "scene", "dam"))]
        if getattr(args, "strong", False):
            cmd.append("--strong")
        with open(log, "w") as lf:
            proc = subprocess.Popen(cmd, env=env, stdout=lf, stderr=subprocess.STDOUT, start_new_session=True)
    except Exception as ex:
        err = repr(ex)
    if proc is not None:
        try:
            rc = proc.wait(timeout=E2E_TIMEOUT_S)
            if rc != 0:
                err = "child of rank %d: exit code %d" % (rank, rc)
        except subprocess.TimeoutExpired:
            try:
                os.killpg(proc.pid, signal.SIGKILL)          # the child's own process group (start_new_session): nothing else
            except OSError:
                pass
            proc.wait()
            err = "child of rank %d: killed after %d s" % (rank, E2E_TIMEOUT_S)
    failed = sum(f for f, in gather([0 if err is None else 1]))
    if rank != 0:
        return None
    out = None
    try:
        if os.path.exists(result):
            with open(result) as f:
                out = json.load(f)
    except Exception as ex:
        err = err or repr(ex)
    if out is None:
        tail = ""
        try:
            tail = open(log).read()[-400:]
        except OSError:
            pass
        return {"error": "no result; %d child(ren) failed; rank 0: %s" % (failed, err), "log_tail": tail}
    if failed:
        out["children_failed"] = failed
    return out


class FileMeet:
    """The children's rendezvous: one node, one scratch directory.  put / get of small pickled objects by (tag, rank), with
    atomic renames; barrier and all-gather on top.  (No second network rendezvous beside the parents' — nothing to collide
    with the launcher's store, nothing that depends on the host name resolving.)"""

    def __init__(self, root, rank, world, timeout_s):
        self.root, self.rank, self.world, self.timeout, self.n = root, rank, world, timeout_s, 0

    def _path(self, tag, r):
        return os.path.join(self.root, "meet_%s_%d.pkl" % (tag, r))

    def put(self, tag, obj):
        import pickle
        p = self._path(tag, self.rank)
        with open(p + ".tmp", "wb") as f:
            pickle.dump(obj, f)
        os.replace(p + ".tmp", p)

    def get(self, tag, r):
        import pickle
        p, t0 = self._path(tag, r), time.perf_counter()
        while not os.path.exists(p):
            if time.perf_counter() - t0 > self.timeout:
                raise TimeoutError("rank %d waited %d s for rank %d at %r" % (self.rank, self.timeout, r, tag))
            time.sleep(0.0005)
        with open(p, "rb") as f:
            return pickle.load(f)

    def all_gather(self, obj):
        self.n += 1
        tag = "g%d" % self.n
        self.put(tag, obj)
        return [self.get(tag, r) for r in range(self.world)]

    def barrier(self):
        self.all_gather(None)

    def broadcast(self, obj, src=0):
        self.n += 1
        tag = "b%d" % self.n
        if self.rank == src:
            self.put(tag, obj)
        return self.get(tag, src)


def e2e_child_main(argv):
    """Child side (one process per GPU): a new decomposed solver is given this rank's particles from HOST arrays
    (set_particles_distributed: H2D), bakes K steps with every step a whole-scene frame gathered on rank 0 and copied to host
    memory there (D2H), and rank 0 reads the last frame.  Wall clock from before the upload to after the read, MAX over
    ranks.  The first bake also allocates (device arrays, the pinned frame store): as in bench.py's N = 1 leg the figure of
    record is a SECOND bake of the same handle, which must reproduce the first one's last frame bit for bit; the first
    bake's figure is written out before the second is attempted."""
    import argparse
    import importlib
    import bench
    from vfd_b200 import partition
    ap = argparse.ArgumentParser()
    ap.add_argument("--e2e-child")
    ap.add_argument("--out")
    ap.add_argument("--meet")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--side", type=int, default=100)
    ap.add_argument("--scene", default="dam")
    ap.add_argument("--strong", action="store_true")
    a = ap.parse_args(argv)
    api = importlib.import_module(os.environ.get("VFD_E2E_API", "vfd_b200.api"))       # (the CPU test of this plumbing substitutes a stand-in)
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    meet = FileMeet(a.meet, rank, world, E2E_TIMEOUT_S)
    cfg = bench.CONFIGS[a.config]
    bench.CONFIG_DESC.clear(); bench.CONFIG_DESC.update(cfg["desc"])
    d = np.load(a.e2e_child)
    pos, vel, ids, bounds, n_global = d["pos"], d["vel"], d["ids"], d["bounds"], int(d["n_global"])
    K = a.steps
    _, box, res = dist_scene(a.side, 1 if a.strong else world, a.scene)
    sim = api.DFSPHSimulation(bench.description(api.DFSPHSimulationDescription, frames=K), device=local)
    uid = meet.broadcast(api.dist_unique_id() if rank == 0 else None)
    sim.init_distributed(rank, world, uid, box[0], box[1])
    origin, cell, tiles = sim.grid()
    sim.set_slab(int(bounds[rank]), int(bounds[rank + 1]))
    ghost = int(tiles[1]) * int(tiles[2]) * 64 * 12 * 2
    vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=R, device=local)
    held = {"pos": pos, "vel": vel, "ids": ids}

    def rehome():
        """Every particle to the rank whose slab holds its tile column NOW (host-side partitioning, like bench_multi.setup's, not
        timed): a particle may have crossed its slab's boundary in the timed run's last step (the running solver migrates it at the
        start of the next one), and the slabs of this handle have moved if its previous bake re-balanced them."""
        sl = sim.slab()
        now = meet.all_gather((int(sl["lo"]), int(sl["hi"])))
        b = np.asarray([lo for lo, _ in now] + [now[-1][1]], np.int64)
        p, v, i = held["pos"], held["vel"], held["ids"]
        owner = partition.owner_of(partition.tile_columns(p[:, 0], origin[0], H, tiles[0]), b)
        away = owner != rank
        sent = meet.all_gather((p[away], v[away], i[away], owner[away]))
        mine = [(sp[so == rank], sv[so == rank], si[so == rank]) for sp, sv, si, so in sent]
        held["pos"] = np.ascontiguousarray(np.concatenate([p[~away]] + [m[0] for m in mine]), np.float32)
        held["vel"] = np.ascontiguousarray(np.concatenate([v[~away]] + [m[1] for m in mine]), np.float32)
        held["ids"] = np.ascontiguousarray(np.concatenate([i[~away]] + [m[2] for m in mine]), np.uint32)

    def bake():
        rehome()
        p, v, i = held["pos"], held["vel"], held["ids"]
        capacity = int(1.5 * len(p)) + 2 * ghost
        meet.barrier()
        t0 = time.perf_counter()
        sim.set_particles_distributed(p, v, i, n_global, capacity)             # H2D of the inputs
        sim.SetRigidBodies([vm])
        sim.Simulate()                                                          # K steps, each a gathered host frame on rank 0
        last = sim.GetFrame(K - 1)[0] if rank == 0 else None                   # the frame pipe's D2H copy, read from the host store
        sim.synchronize()
        return max(meet.all_gather(time.perf_counter() - t0)), last

    def report(el, which, extra):
        if rank != 0:
            return
        out = {"value": n_global * K / el, "unit": "particle-steps/s", "h2d_bytes_per_step": int(28 * n_global / K), "d2h_bytes_per_step": 36 * n_global,
               "ms_per_step": 1e3 * el / K, "n_gpus": world, "bake": which,
               "note": "every rank: vfd_dfsph_set_particles_distributed (positions, velocities, ids from host arrays) + set_rigid_bodies + simulate "
                       "(FrameLength 0: every step baked — owned particles gathered by persistent id on rank 0, whole-scene frame copied to host "
                       "memory) + get_frame of the last one on rank 0; wall clock, max over ranks; a new solver per rank in a child process, from the "
                       "timed run's settled state"}
        out.update(extra)
        with open(a.out + ".tmp", "w") as f:
            json.dump(out, f)
        os.replace(a.out + ".tmp", a.out)

    first = "first bake of the handle (includes allocating device arrays and the pinned frame store)"
    el1, f1 = bake()
    report(el1, first, {})
    el2, f2 = bake()
    same = all(meet.all_gather(bool(rank != 0 or (f1 is not None and f2 is not None and f1.tobytes() == f2.tobytes()))))
    if same:
        report(el2, "second bake of the handle (an untimed first bake sized the buffers)", {"rebake_identical": True, "first_bake_ms_per_step": 1e3 * el1 / K})
    else:
        report(el1, first, {"rebake_identical": False})
    meet.barrier()
    os._exit(0)            # the figure is on disk: the teardown of a decomposed handle must not decide the exit code


if __name__ == "__main__":
    import sys
    if "--e2e-child" in sys.argv:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        e2e_child_main(sys.argv[1:])
