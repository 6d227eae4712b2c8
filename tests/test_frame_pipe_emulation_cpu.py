"""The asynchronous frame store (csrc/frame_pipe.cu) run on the CPU: the product's source compiled
unchanged against an emulated CUDA runtime whose streams are threads, whose events complete when their stream gets there and
whose copies take time (tests/host_check/emu/cuda_runtime.h), driven like Solver::step drives it
(tests/host_check/frame_pipe_host.cpp).  Covered here, before any GPU: frames complete and in order however export, copy and
worker interleave; the device's keep/drop decision; the pageable fallback beyond the pinned budget (all frames, or the ones
after the budget is used up); re-use of device slots and staging buffers; the zero-copy view.  tests/test_gpu_z_frame_store.py
repeats the storage cases on hardware."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "host_check"))


@pytest.fixture(scope="module")
def pipe():
    import build_emu
    L = C.CDLL(build_emu.build_frame_pipe())
    L.hc_frame_pipe_bake.restype = C.c_long
    L.hc_frame_pipe_bake.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long), C.c_void_p]
    return L


def bake(L, n, steps, conditional=0, mask=0xffffffff, export_us=100, copy_us=300):
    allocs, nbytes = C.c_long(), C.c_long()
    which = np.zeros(max(steps, 1), np.uint32)
    r = L.hc_frame_pipe_bake(n, steps, conditional, mask, export_us, copy_us, C.byref(allocs), C.byref(nbytes), which.ctypes.data_as(C.c_void_p))
    return r, allocs.value, nbytes.value, which[:max(r, 0)]


@pytest.mark.parametrize("export_us,copy_us", [(0, 0), (50, 600), (600, 50)])
def test_every_step_a_frame(pipe, monkeypatch, export_us, copy_us):
    """FrameLength 0: 40 frames through 3 device slots, whichever of export and copy is the slow one."""
    monkeypatch.delenv("VFD_FRAME_PINNED_MB", raising=False)
    n = 5000
    r, allocs, nbytes, which = bake(pipe, n, 40, export_us=export_us, copy_us=copy_us)
    assert r == 40 and np.array_equal(which, np.arange(40))
    assert nbytes >= 40 * n * 36                                       # every frame lives in pinned storage (+ the scalars' mirror)


@pytest.mark.parametrize("budget_mb,frames_pinned", [(0, 0), (1, 5)])
def test_pageable_storage_beyond_the_pinned_budget(pipe, monkeypatch, budget_mb, frames_pinned):
    """5 000 particles: a frame is 180 000 B, so a budget of 1 MB holds five; the rest — or all, with no budget — goes through
    the slot's pinned staging buffer into pageable storage, copied there by the worker before the slot is reused."""
    monkeypatch.setenv("VFD_FRAME_PINNED_MB", str(budget_mb))
    n = 5000
    r, allocs, nbytes, which = bake(pipe, n, 30, export_us=20, copy_us=200)
    assert r == 30 and np.array_equal(which, np.arange(30))
    frame = n * 36
    staging = 3 * frame                                                # at most one staging buffer per device slot
    assert frames_pinned * frame <= nbytes <= frames_pinned * frame + staging + 4096, (nbytes, frame)


@pytest.mark.parametrize("mask", [0x55555555, 0x00ff00f1, 0x80000001, 0x0])
def test_the_device_decides_which_steps_are_frames(pipe, monkeypatch, mask):
    """A frame length > 0: every step is exported and copied, the worker keeps the ones the device marked (float[2] of the
    slot's scalars) and returns the others' storage to the pool: the published frames are the marked steps, in order."""
    monkeypatch.setenv("VFD_FRAME_PINNED_MB", "1")
    steps = 64
    r, allocs, nbytes, which = bake(pipe, 3000, steps, conditional=1, mask=mask, export_us=30, copy_us=100)
    want = [s for s in range(steps) if (mask >> (s % 32)) & 1]
    assert r == len(want) and which.tolist() == want


def test_no_particles_and_one_particle(pipe, monkeypatch):
    monkeypatch.delenv("VFD_FRAME_PINNED_MB", raising=False)
    assert bake(pipe, 1, 7)[0] == 7
    assert bake(pipe, 0, 4)[0] == 4
