"""The frame store's pageable fallback (csrc/frame_pipe.cu).  Baked frames land in pinned host
memory up to a budget (VFD_FRAME_PINNED_MB, default 8 GB — 227 frames of a million particles); beyond it a frame is copied
into a pinned staging slot and moved to pageable storage by the pipe's worker.  The bench and the other tests stay below
the budget: here the budget is squeezed so that a bake mixes both kinds of storage, or uses pageable storage only, and the
frames must be the ones of the all-pinned bake, bit for bit (the solver is bit-reproducible run to run).
(File name: sorts after the solver's parity tests — this path had no hardware run before the round's end.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bake(monkeypatch, pinned_mb, frame_length):
    from vfd_b200 import api
    import test_gpu_scale as big
    if pinned_mb is None:
        monkeypatch.delenv("VFD_FRAME_PINNED_MB", raising=False)
    else:
        monkeypatch.setenv("VFD_FRAME_PINNED_MB", str(pinned_mb))
    pos, box, res = big.scene(30)                    # 27 000 particles: a frame is 0.93 MB
    vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=big.R)
    sim = api.DFSPHSimulation(api.DFSPHSimulationDescription(FrameCount=12, FrameLength=frame_length, **big.CONFIGS["dfsph"]))
    sim.SetFluidObjects([api.FluidObject(pos)])
    sim.SetRigidBodies([vm])
    sim.Simulate()
    assert sim.GetFrameCount() == 12
    frames = [sim.GetFrame(i) for i in range(12)]
    sim.close()
    return frames


@pytest.mark.parametrize("frame_length", [0.0, 0.0016])
def test_frames_in_pageable_storage_equal_the_pinned_ones(lib_built, monkeypatch, frame_length):
    pinned = bake(monkeypatch, None, frame_length)
    for budget in (3, 0):                            # three pinned frames then pageable ones; pageable only
        other = bake(monkeypatch, budget, frame_length)
        for i, ((a, va, da), (b, vb, db)) in enumerate(zip(pinned, other)):
            assert va == vb and da == db, (budget, i)
            for f in ("Position", "Velocity", "Acceleration"):
                assert np.array_equal(np.asarray(a[f]).view(np.uint32), np.asarray(b[f]).view(np.uint32)), "frame %d field %s differs with a pinned budget of %d MB" % (i, f, budget)


def test_frame_views_are_the_frames(lib_built, monkeypatch):
    """vfd_dfsph_get_frame_data hands out a frame where it lies in the store (DFSPHParticleBuffer::GetFrame returns a
    reference as well): same records as the copying getter, the same address when asked twice, in pinned and in pageable storage."""
    from vfd_b200 import api
    import test_gpu_scale as big
    monkeypatch.setenv("VFD_FRAME_PINNED_MB", "3")
    pos, box, res = big.scene(30)
    vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=res, particle_radius=big.R)
    sim = api.DFSPHSimulation(api.DFSPHSimulationDescription(FrameCount=6, FrameLength=0.0, **big.CONFIGS["dfsph"]))
    sim.SetFluidObjects([api.FluidObject(pos)])
    sim.SetRigidBodies([vm])
    sim.Simulate()
    for i in range(6):
        copy, vmax, dt = sim.GetFrame(i)
        view, vmax2, dt2 = sim.GetFrameView(i)
        again, _, _ = sim.GetFrameView(i)
        assert (vmax, dt) == (vmax2, dt2) and len(view) == len(pos)
        assert view.__array_interface__["data"][0] == again.__array_interface__["data"][0]
        assert np.array_equal(view.view(np.uint8), copy.view(np.uint8))
    with pytest.raises(api.VfdError):
        sim.GetFrameView(6)
    sim.close()
