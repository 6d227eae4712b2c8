#!/usr/bin/env python3
"""tests/golden/make_golden_mesh.py — golden fixture of the general-mesh scene preparation,
generated from the REFERENCE'S OWN sources (oracle/_ref/libvfd_ref_cpu.so: MeshDistance.cpp, SDF.cu, ParticleSampler.cpp,
RigidBody.cu compiled where they lie under /root/reference).

Run in the build container:   python tests/golden/make_golden_mesh.py      -> tests/golden/mesh.npz
The scene is the shape of the reference's shipped DFSPH scene (Resources/Scenes/New/DFSPH/default.json): a cone of fluid,
upside down above a flat slab — with procedural meshes of tests/meshes.py, not the reference's .obj files.
  cone_*, slab_*          vertices, triangles, row-major 4x4 transforms
  points, sd_cone, sd_torus   query points and MeshDistance::SignedDistance there (torus: 576 faces under a general transform)
  sample_<mode>           ParticleSampler::SampleMeshVolume of the cone (radius 0.025, distance grid 20^3), modes 0 1 2
  slabmap_*, conemap_*, oddmap_*   RigidBody volume maps (field 0 and 1) of the slab (20^3), of the cone as a body (12^3) and of
                          the torus on a non-cubic grid (5 x 9 x 7: the node numbering's axis order)

Every reference output is the MODAL result of RUNS consecutive calls.  The reference's MeshDistance prunes its search with a
sphere tree whose spheres come from a randomised smallest-enclosing-sphere routine (Core/Structures/BoundingSphere.h:133-172,
rand()-driven permutation + 1e-6 perturbation); on the cone, whose rim vertices are cocircular, some rand() states produce
spheres that do not enclose their triangles and the walk then misses the nearest face (first call of a process: 2 184 of
20 000 distances wrong by up to 1.45, 16 691 instead of 16 670 samples; checked against an fp64 brute force).  Where the tree
is sound the reference returns the minimum over all faces — the value the brute-force GPU kernels compute by construction.
"""
import collections
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import refsim  # noqa: E402
import meshes  # noqa: E402

RADIUS = 0.025
RUNS = 7


def modal(fn):
    """The most frequent output (bit for bit) of RUNS calls."""
    outs = [np.ascontiguousarray(fn()) for _ in range(RUNS)]
    keys = [hashlib.md5(o.tobytes()).hexdigest() + str(o.shape) for o in outs]
    best, n = collections.Counter(keys).most_common(1)[0]
    print("   modal result: %d of %d calls agree" % (n, RUNS))
    return outs[keys.index(best)]


def modal_elementwise(fn):
    """Per element the most frequent value of RUNS calls (a 1-D float array)."""
    outs = np.stack([fn() for _ in range(RUNS)])
    srt = np.sort(outs, axis=0)
    med = srt[RUNS // 2]                         # with a majority of equal values the median is that value
    agree = (outs == med).sum(0)
    print("   element-wise: %d of %d elements unanimous, weakest majority %d of %d" % ((agree == RUNS).sum(), outs.shape[1], agree.min(), RUNS))
    return med


def body_map(verts, tris, T, res, inverted=False, padding=0.0):
    sim = refsim.RefSim(refsim.Desc(ParticleRadius=RADIUS))
    sim.set_particles(np.zeros((1, 3), np.float32))
    sim.add_mesh_body(verts, tris, transform=T, inverted=inverted, padding=padding, res=res)
    return sim.volume_map(0)


def main():
    out = {}
    cv, ct = meshes.cone()
    cT = meshes.transform(rotate_x_deg=180.0, translate=(0.0, 3.0, 0.0))
    sv, st = meshes.box((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
    sT = meshes.transform(scale=(2.0, 0.2, 2.0))
    tv, tt = meshes.torus()
    tT = meshes.transform(scale=(1.5, 1.0, 0.7), rotate_x_deg=30.0, translate=(0.2, -0.1, 0.3))
    out.update(cone_verts=cv, cone_tris=ct, cone_T=cT, slab_verts=sv, slab_tris=st, slab_T=sT, torus_verts=tv, torus_tris=tt, torus_T=tT)

    rng = np.random.default_rng(7)
    pc = (np.array([-1.6, 1.4, -1.6]) + np.array([3.2, 3.2, 3.2]) * rng.random((20000, 3))).astype(np.float32)
    pt = (np.array([-2.4, -1.6, -1.6]) + np.array([4.8, 3.2, 3.2]) * rng.random((20000, 3))).astype(np.float32)
    out.update(points_cone=pc, sd_cone=modal_elementwise(lambda: refsim.mesh_signed_distance(cv, ct, pc, transform=cT)),
               points_torus=pt, sd_torus=modal_elementwise(lambda: refsim.mesh_signed_distance(tv, tt, pt, transform=tT)))

    for mode in (0, 1, 2):
        out["sample_%d" % mode] = modal(lambda: refsim.sample_mesh_volume(cv, ct, RADIUS, (20, 20, 20), False, mode, transform=cT))
        print("mode", mode, len(out["sample_%d" % mode]), "samples")

    for name, (v, t, T, res) in {"slabmap": (sv, st, sT, (20, 20, 20)), "conemap": (cv, ct, cT, (12, 12, 12)), "oddmap": (tv, tt, tT, (5, 9, 7))}.items():
        with refsim.quiet_stdout():
            m = body_map(v, t, T, res)
            m["nodes"] = modal(lambda: body_map(v, t, T, res)["nodes"])
        out.update({name + "_domain_min": m["domain_min"], name + "_domain_max": m["domain_max"], name + "_resolution": m["resolution"],
                    name + "_cell_size": m["cell_size"], name + "_cell_size_inv": m["cell_size_inv"],
                    name + "_nodes": m["nodes"], name + "_node_count": m["node_count"]})
        print(name, m["node_count"], "nodes")
    np.savez_compressed(os.path.join(HERE, "mesh.npz"), **out)
    print("wrote", os.path.join(HERE, "mesh.npz"), os.path.getsize(os.path.join(HERE, "mesh.npz")), "bytes")


if __name__ == "__main__":
    main()
