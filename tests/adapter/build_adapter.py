#!/usr/bin/env python3
"""tests/adapter/build_adapter.py — TEST INFRASTRUCTURE: compiles the adapter of INTEGRATION.md section 2 against the reference's
headers where they lie under /root/reference (with oracle/shim's headless stand-ins for the renderer and the CUDA runtime), links
it with libvfd_dfsph.so (the product) and oracle/_ref/libvfd_ref_cpu.so (the reference's own RigidBody / SDF / DFSPHParticleBuffer /
kernel-table code, built by oracle/build_ref.py), and leaves tests/adapter/_bin/adapter_bake for tests/test_gpu_adapter.py.
Nothing of the reference is copied; the binary is git-ignored and travels to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref as br

OUT = os.path.join(HERE, "_bin", "adapter_bake")


def main():
    if not os.path.isdir(br.SRC):
        print("reference not present at %s — keeping the prebuilt %s" % (br.REF, OUT))
        return 0
    ref_so = os.path.join(br.OUT, "libvfd_ref_cpu.so")
    lib_so = os.path.join(ROOT, "vfd_b200", "lib", "libvfd_dfsph.so")
    for p in (ref_so, lib_so):
        if not os.path.exists(p):
            raise SystemExit("missing %s (run __graft_entry__.build() first)" % p)
    srcs = [os.path.join(HERE, "DFSPHSimulator.cpp"), os.path.join(HERE, "adapter_bake.cpp")]
    if os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(p) for p in srcs + [ref_so, __file__, os.path.join(ROOT, "include", "vfd_dfsph.h")]):
        print("up to date:", OUT)
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    orig_dirs = sorted({os.path.dirname(os.path.join(br.SRC, f)) for f in br.FILES})
    inc = ["-I", os.path.join(br.HERE, "shim", "common"), "-I", os.path.join(br.HERE, "shim", "cpu"), "-I", br.SRC]
    for d in orig_dirs:
        inc += ["-I", d]
    inc += ["-I", br.GLM, "-I", br.TINYOBJ, "-I", os.path.join(ROOT, "include"), "-isystem", "/usr/local/cuda/include"]
    cmd = ["g++", "-std=c++20", "-O2", "-fopenmp", "-DNDEBUG", "-w", "-DVFD_ADAPTER_HEADLESS", "-DVFD_ADAPTER_OPEN_SDF",
           "-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_OMP", "-DTHRUST_HOST_SYSTEM=THRUST_HOST_SYSTEM_CPP"] + inc + srcs + [
           "-o", OUT, ref_so, lib_so, "-Wl,-rpath,$ORIGIN/../../../oracle/_ref", "-Wl,-rpath,$ORIGIN/../../../vfd_b200/lib"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-6000:] + "\n")
        raise SystemExit("adapter build failed")
    print("built", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
