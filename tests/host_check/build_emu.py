#!/usr/bin/env python3
"""tests/host_check/build_emu.py — TEST INFRASTRUCTURE: builds the scene-preparation translation unit of the product
(vfd_b200/csrc/volume_map.cu with volume_map.cuh, mesh_distance.cuh, map_geometry.cuh, and tables.cpp) for the HOST against
the CUDA-on-CPU emulation in tests/host_check/emu/: the sources are copied into a throw-away directory next to the emulation's
solver.h, the kernel launches `k<<<blocks, threads>>>(...)` rewritten to `emu_launch(blocks, threads, k, ...)`, and compiled with
g++ (-ffp-contract=off, like the device build's -fmad=false) into tests/host_check/_bin/libvfd_sceneprep_emu.so, which exports
the same C entry points (vfd_volume_map_build_box / _build_mesh, vfd_mesh_signed_distance, vfd_sample_mesh_volume, ...).
Kernels and host code are the product's, line for line; only the machine underneath is emulated."""
import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "vfd_b200", "csrc")
OUT = os.path.join(HERE, "_bin", "libvfd_sceneprep_emu.so")
SOURCES = ["volume_map.cu", "tables.cpp"]
HEADERS = ["volume_map.cuh", "mesh_distance.cuh", "map_geometry.cuh"]
LAUNCH = re.compile(r'(\w+)<<<(.+?),\s*(\w+)>>>\(')


def deps():
    return [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "emu", f) for f in os.listdir(os.path.join(HERE, "emu"))] + [__file__]


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(d) for d in deps()):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vfd_emu_")
    try:
        for f in HEADERS:
            shutil.copy(os.path.join(CSRC, f), tmp)
        for f in os.listdir(os.path.join(HERE, "emu")):
            shutil.copy(os.path.join(HERE, "emu", f), tmp)
        srcs = [os.path.join(tmp, "emu.cpp")]
        for f in SOURCES:
            with open(os.path.join(CSRC, f)) as fh:
                s = fh.read()
            s, n = LAUNCH.subn(r'emu_launch(\2, \3, \1, ', s)
            dst = os.path.join(tmp, f.replace(".cu", "_cu.cpp"))
            with open(dst, "w") as fh:
                fh.write(s)
            srcs.append(dst)
        cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-pthread", "-w", "-I", tmp, "-I", os.path.join(ROOT, "include"), "-o", OUT] + srcs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed:\n" + r.stdout[-6000:])
        return OUT
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


FRAME_OUT = os.path.join(HERE, "_bin", "libvfd_framepipe_emu.so")


def build_frame_pipe(force=False):
    """csrc/frame_pipe.cu (pure host code over the CUDA runtime) compiled unchanged against the emulated runtime with
    asynchronous streams and events (emu/cuda_runtime.h), plus the driver tests/host_check/frame_pipe_host.cpp."""
    srcs = [os.path.join(CSRC, "frame_pipe.cu"), os.path.join(CSRC, "frame_pipe.h"), os.path.join(HERE, "frame_pipe_host.cpp")]
    emu = [os.path.join(HERE, "emu", f) for f in os.listdir(os.path.join(HERE, "emu"))]
    if not force and os.path.exists(FRAME_OUT) and os.path.getmtime(FRAME_OUT) > max(os.path.getmtime(d) for d in srcs + emu + [__file__]):
        return FRAME_OUT
    os.makedirs(os.path.dirname(FRAME_OUT), exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vfd_emu_")
    try:
        for f in emu:
            shutil.copy(f, tmp)
        # frame_pipe.h includes "../../include/vfd_dfsph.h" relative to csrc/: mirror that layout
        os.makedirs(os.path.join(tmp, "include"))
        shutil.copy(os.path.join(ROOT, "include", "vfd_dfsph.h"), os.path.join(tmp, "include"))
        csrc = os.path.join(tmp, "vfd_b200", "csrc")
        os.makedirs(csrc)
        shutil.copy(os.path.join(CSRC, "frame_pipe.h"), csrc)
        shutil.copy(os.path.join(CSRC, "frame_pipe.cu"), os.path.join(csrc, "frame_pipe_cu.cpp"))
        shutil.copy(os.path.join(HERE, "frame_pipe_host.cpp"), csrc)
        cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-pthread", "-w", "-I", tmp, "-I", csrc, "-o", FRAME_OUT,
               os.path.join(csrc, "frame_pipe_cu.cpp"), os.path.join(csrc, "frame_pipe_host.cpp"), os.path.join(tmp, "emu.cpp")]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed:\n" + r.stdout[-6000:])
        return FRAME_OUT
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(build(force=True))
    print(build_frame_pipe(force=True))
