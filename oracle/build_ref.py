#!/usr/bin/env python3
"""oracle/build_ref.py — TEST INFRASTRUCTURE: builds the *reference's own* DFSPH solver sources.

Compiles the solver files where they lie under /root/reference
into  oracle/_ref/libvfd_ref_cpu.so  (CUDA-on-CPU emulation, runs on host cores; used as the
parity oracle and as the `cpu_baseline` / `--impl reference` arm of bench.py) and, with --gpu,
oracle/_ref/libvfd_ref_gpu.so  (nvcc, sm_100a — the only pre-existing GPU implementation of the
path; a timing reference only).  Nothing of the reference is copied into the repository: patched
temporary copies (kernel-launch syntax rewritten for the CPU build, three MSVC-isms fixed —
SURVEY.md F7) live in a throw-away directory under /tmp; only the .so files land in oracle/_ref/,
which is git-ignored but travels to the GPU box.

The reference's own build system (premake/VS2022, Windows-only) is not used.
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VFD_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "VFD", "Source")
GLM = os.path.join(REF, "VFD", "VFD", "ThirdParty", "glm")
TINYOBJ = os.path.join(REF, "VFD", "ThirdParty", "tinyobjloader")
OUT = os.path.join(HERE, "_ref")

# the solver path; FluidObject is replaced by oracle/shim (raw positions)
FILES = [
    "Simulation/DFSPH/DFSPHImplementation.cu",
    "Simulation/DFSPH/DFSPHKernels.cu",
    "Simulation/DFSPH/ParticleSearch/ParticleSearch.cu",
    "Simulation/DFSPH/ParticleSearch/ParticleSearchKernels.cu",
    "Simulation/DFSPH/RigidBody/RigidBody.cu",
    "Simulation/DFSPH/ParticleBuffer/DFSPHParticleBuffer.cu",
    "Utility/SDF/SDF.cu",
    "Utility/SDF/MeshDistance.cpp",
    "Utility/Sampler/ParticleSampler.cpp",       # scene preparation next to the path: oracle of vfd_b200/scene_io.py
    "Renderer/Mesh/EdgeMesh.cpp",
    "Core/Structures/BoundingSphere.cpp",
    "Core/Structures/AxisAlignedBoundingBox.cpp",
    "Core/Math/GaussQuadrature.cpp",
    "Core/Math/Math.cpp",
    "Compute/ComputeHelper.cpp",
]

LAUNCH = re.compile(r'(\w+)\s*<<\s*<\s*([^,]+?),\s*([^>]+?)\s*>>\s*>\s*\(')


def patched_copy(rel, tmp, cpu, ieee):
    """Write a patched copy of SRC/rel into tmp (flat) and return its path."""
    with open(os.path.join(SRC, rel), "r", encoding="utf-8-sig") as f:
        s = f.read()
    s = s.replace("__forceinline__ inline", "__forceinline__").replace("std::sqrtf", "sqrtf")
    if cpu:
        s = LAUNCH.sub(r'cpu_launch(\2, \3, \1, ', s)
    if ieee:
        # the accumulator is uninitialised under nvcc; the intended value is zero
        s = s.replace("glm::mat3x3 result = glm::mat3x3();", "glm::mat3x3 result = glm::mat3x3(0.0f);")
    dst = os.path.join(tmp, os.path.basename(rel) + (".cpp" if cpu and rel.endswith(".cu") else ""))
    with open(dst, "w") as f:
        f.write(s)
    return dst


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-6000:] + "\n")
        raise SystemExit("oracle/_ref build failed")
    return r.stdout


def build(kind, jobs, fast_math=False):
    cpu = kind == "cpu"
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vfd_ref_%s_" % kind)
    try:
        # include order: shim → reference source root → original directories of the (flattened)
        # patched copies → vendored glm → CUDA toolkit (thrust/CCCL only, as a system dir)
        orig_dirs = sorted({os.path.dirname(os.path.join(SRC, f)) for f in FILES})
        inc = ["-I", os.path.join(HERE, "shim", "common")]
        if cpu:
            inc += ["-I", os.path.join(HERE, "shim", "cpu")]
        inc += ["-I", SRC]
        for d in orig_dirs:
            inc += ["-I", d]
        inc += ["-I", GLM, "-I", TINYOBJ]
        if cpu:
            inc += ["-isystem", "/usr/local/cuda/include"]
            base = ["g++", "-std=c++20", "-O2", "-fopenmp", "-fPIC", "-DNDEBUG", "-w",
                    "-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_OMP", "-DTHRUST_HOST_SYSTEM=THRUST_HOST_SYSTEM_CPP"]
            lang = ["-x", "c++"]
        else:
            base = ["nvcc", "-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-DNDEBUG", "-w",
                    "-DVFD_REF_GPU", "-Xcompiler", "-fopenmp,-fPIC", "--expt-relaxed-constexpr"]
            if fast_math:
                base.append("--use_fast_math")   # as shipped: premake5.lua:237
            lang = ["-x", "cu"]
        ieee = not fast_math
        srcs = [patched_copy(f, tmp, cpu, ieee) for f in FILES]
        srcs.append(os.path.join(HERE, "ref_driver.cpp"))
        if cpu:
            srcs.append(os.path.join(HERE, "shim", "cpu", "emu.cpp"))
        objs = [os.path.join(tmp, "o%02d.o" % i) for i in range(len(srcs))]

        def cc(i):
            run(base + lang + inc + ["-c", srcs[i], "-o", objs[i]])
        with ThreadPoolExecutor(max_workers=jobs) as ex:
            list(ex.map(cc, range(len(srcs))))
        name = "libvfd_ref_cpu.so" if cpu else ("libvfd_ref_gpu_fast.so" if fast_math else "libvfd_ref_gpu.so")
        out = os.path.join(OUT, name)
        if cpu:
            run(["g++", "-shared", "-fopenmp", "-o", out] + objs)
        else:
            run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fopenmp", "-o", out] + objs)
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true", help="also build the nvcc/sm_100a variants")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(SRC):
        print("reference not present at %s — keeping prebuilt oracle/_ref" % REF)
        return 0
    targets = [("cpu", False)]
    if a.gpu:
        targets += [("gpu", False), ("gpu", True)]
    for kind, fm in targets:
        name = "libvfd_ref_cpu.so" if kind == "cpu" else ("libvfd_ref_gpu_fast.so" if fm else "libvfd_ref_gpu.so")
        dst = os.path.join(OUT, name)
        newest = max(os.path.getmtime(p) for p in [os.path.join(HERE, "ref_driver.cpp"), __file__] +
                     [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(HERE, "shim")) for f in fs])
        if not a.force and os.path.exists(dst) and os.path.getmtime(dst) > newest:
            print("up to date:", dst)
            continue
        print("built", build(kind, a.jobs, fm))
    return 0


if __name__ == "__main__":
    sys.exit(main())
