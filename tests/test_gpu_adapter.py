"""The adapter of INTEGRATION.md section 2, compiled and run: the reference's `vfd::DFSPHSimulation` (its header untouched) with its
body replaced by tests/adapter/DFSPHSimulator.cpp over libvfd_dfsph.so, driven like the editor's "Bake" — the reference's own
FluidObject / RigidBody (SDF + volume map on the host) / DFSPHParticleBuffer around it — against the reference's solver stepping
the same scene (oracle/_ref).  The binary is built where the reference's headers are (tests/adapter/build_adapter.py, called by
__graft_entry__.build()) and travels to the GPU box."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "adapter", "_bin", "adapter_bake")
R, D, H = 0.025, 0.05, 0.1


@pytest.mark.parametrize("frame_length", [0.0, 0.0016])
def test_bake_through_the_reference_interface(frame_length, tmp_path, lib_built):
    from oracle import refsim
    from vfd_b200 import api
    if not os.path.exists(BIN):
        pytest.skip("tests/adapter/_bin/adapter_bake not built (needs the reference's headers: tests/adapter/build_adapter.py)")
    if not refsim.available("cpu"):
        pytest.skip("oracle/_ref/libvfd_ref_cpu.so not built")
    side, frames, c = 14, 10, 4
    pos = api.block_positions(side, side, side, R, origin=(c * D, c * D, c * D))
    pos = (pos + np.random.RandomState(5).uniform(-0.2 * R, 0.2 * R, pos.shape)).astype(np.float32)
    box = ((0.0, 0.0, 0.0), ((side + 2 * c + 8) * D, (side + 2 * c + 4) * D, (side + 2 * c) * D))
    pfile, ofile = str(tmp_path / "positions.bin"), str(tmp_path / "frames.bin")
    pos.tofile(pfile)
    r = subprocess.run([BIN, str(side), str(frames), repr(frame_length), ofile, pfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-1500:]
    raw = np.fromfile(ofile, np.uint8)
    nf, n = (int(x) for x in raw[:8].view(np.uint32))
    assert nf == frames and n == len(pos)
    rec = raw[8:].view(np.float32).reshape(frames, 2 + 9 * n)

    cfg = dict(MinPressureSolverIterations=2, MaxPressureSolverIterations=2, MinDivergenceSolverIterations=2, MaxDivergenceSolverIterations=2,
               EnableViscositySolver=0, EnableSurfaceTensionSolver=0)
    with refsim.quiet_stdout():
        ref = refsim.RefSim(refsim.Desc(**cfg))
        ref.set_particles(pos)
        ref.add_box_body(box[0], box[1], inverted=True, padding=0.0, res=(10, 10, 10))
        ref.commit_bodies()
        # the reference's frame rule (DFSPHImplementation.cu:148-167, FrameTime accumulated in :427), driven step by step
        t_frame, fi, steps = np.float32(0.0), 0, 0
        while fi < frames and steps < 300:
            ref.step(1)
            steps += 1
            dbg = ref.debug()
            t_frame = np.float32(t_frame + np.float32(dbg["dt"]))
            if t_frame >= np.float32(frame_length):
                want = ref.particles()
                got = rec[fi]
                assert abs(got[1] - dbg["dt"]) <= 1e-6 * dbg["dt"], (fi, got[1], dbg["dt"])
                assert abs(got[0] - dbg["max_vel2"]) <= 1e-4 * max(dbg["max_vel2"], 0.1), (fi, got[0], dbg["max_vel2"])
                p = got[2:].reshape(n, 9)
                for k, field in enumerate(("Position", "Velocity", "Acceleration")):
                    y = np.asarray(want[field], np.float64)
                    e = np.abs(p[:, 3 * k:3 * k + 3].astype(np.float64) - y).max() / max(np.abs(y).max(), 1e-12)
                    assert e <= 5e-5, (fi, field, e)
                t_frame = np.float32(0.0)
                fi += 1
    assert fi == frames
