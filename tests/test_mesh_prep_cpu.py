"""Scene preparation for general triangle meshes — what can be pinned without a GPU.

The arithmetic the CUDA kernels of csrc/volume_map.cu run per point (csrc/mesh_distance.cuh: point-triangle distance, sign
from the closest feature's pseudo-normal; csrc/map_geometry.cuh: grid layout; the sampling lattice) is host+device code.
tests/host_check/mesh_host.cpp compiles it with g++ and these tests compare it, bit for bit, with tests/golden/mesh.npz —
outputs of the reference's own MeshDistance / SDF / ParticleSampler / RigidBody (tests/golden/make_golden_mesh.py) — and, when
oracle/_ref is present, with the reference live.  The kernels' own plumbing is covered by tests/test_gpu_zmesh.py."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import meshes  # noqa: E402
import sdf_numpy  # noqa: E402

RADIUS = 0.025
vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "host_check", "mesh_host.cpp")
    out = os.path.join(HERE, "host_check", "_bin", "libmesh_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src] + [os.path.join(ROOT, "vfd_b200", "csrc", f) for f in ("mesh_distance.cuh", "map_geometry.cuh")]
    if not os.path.exists(out) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(out):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-fopenmp", "-o", out, src], check=True)
    return C.CDLL(out)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "mesh.npz"))


def mesh_args(v, t, T):
    v = np.ascontiguousarray(v, np.float32)
    t = np.ascontiguousarray(t, np.uint32)
    T = None if T is None else np.ascontiguousarray(np.asarray(T, np.float32).T).reshape(16)      # row-major -> glm columns
    return v, t, T


def host_sd(hc, v, t, T, p):
    v, t, T = mesh_args(v, t, T)
    p = np.ascontiguousarray(p, np.float32)
    out = np.zeros(len(p), np.float32)
    assert hc.hc_mesh_signed_distance(vp(v), len(v), vp(t), len(t), vp(T), vp(p), len(p), vp(out), None) == 0
    return out


def test_procedural_meshes_are_the_fixtures(g):
    for name, (v, t) in {"cone": meshes.cone(), "torus": meshes.torus(), "slab": meshes.box((-1, -1, -1), (1, 1, 1))}.items():
        assert np.array_equal(g[name + "_tris"], t)
        np.testing.assert_allclose(g[name + "_verts"], v, rtol=0, atol=1e-6)


@pytest.mark.parametrize("name", ["cone", "torus"])
def test_signed_distance_is_the_references_bit_for_bit(hc, g, name):
    sd = host_sd(hc, g[name + "_verts"], g[name + "_tris"], g[name + "_T"], g["points_" + name])
    ref = g["sd_" + name]
    assert np.array_equal(sd, ref), "differ at %d of %d points, max %g" % ((sd != ref).sum(), len(ref), np.abs(sd - ref).max())
    assert (ref < 0).sum() > 300 and (ref > 0).sum() > 300


def test_signed_distance_against_the_live_reference(hc):
    from oracle import refsim
    if not refsim.available("cpu") or not hasattr(refsim._load("cpu"), "ref_mesh_signed_distance"):
        pytest.skip("oracle/_ref without the mesh hooks")
    rng = np.random.default_rng(11)
    for (v, t), T in [(meshes.cone(24, 0.7, 1.3), meshes.transform(scale=(1.0, 2.0, 0.5), rotate_x_deg=77.0, translate=(5.0, -2.0, 1.0))),
                      (meshes.torus(1.0, 0.3, 40, 20), None), (meshes.box((0, 0, 0), (10, 6, 10)), None)]:
        w = v if T is None else (np.c_[v, np.ones(len(v))] @ np.asarray(T, np.float64).T)[:, :3]
        lo, hi = w.min(0) - 0.5, w.max(0) + 0.5
        p = (lo + (hi - lo) * rng.random((30000, 3))).astype(np.float32)
        # the reference's randomised sphere tree is unsound for some rand() states (tests/golden/make_golden_mesh.py): its answer
        # is the per-point majority of five calls
        ref = np.sort(np.stack([refsim.mesh_signed_distance(v, t, p, transform=T) for _ in range(5)]), axis=0)[2]
        got = host_sd(hc, v, t, T, p)
        assert (got != ref).sum() <= 2 and np.abs(got - ref).max() < 5e-6, ((got != ref).sum(), np.abs(got - ref).max())


def body_field0(hc, v, t, T, res, inverted=False, padding=0.0):
    v, t, T = mesh_args(v, t, T)
    r = np.ascontiguousarray(res, np.uint32)
    geom, counts = np.zeros(15, np.float32), np.zeros(5, np.uint32)
    assert hc.hc_body_map_field0(vp(v), len(v), vp(t), len(t), vp(T), int(inverted), C.c_float(padding), vp(r), C.c_float(RADIUS), vp(geom), vp(counts), None) == 0
    nodes = np.zeros(int(counts[0]), np.float32)
    assert hc.hc_body_map_field0(vp(v), len(v), vp(t), len(t), vp(T), int(inverted), C.c_float(padding), vp(r), C.c_float(RADIUS), vp(geom), vp(counts), vp(nodes)) == 0
    return geom, nodes


@pytest.mark.parametrize("name,mesh", [("slabmap", "slab"), ("conemap", "cone"), ("oddmap", "torus")])
def test_body_map_grid_and_distance_field_are_the_references(hc, g, name, mesh):
    geom, nodes = body_field0(hc, g[mesh + "_verts"], g[mesh + "_tris"], g[mesh + "_T"], g[name + "_resolution"])
    assert np.array_equal(geom[0:3], g[name + "_domain_min"]) and np.array_equal(geom[3:6], g[name + "_domain_max"])
    assert np.array_equal(geom[6:9], g[name + "_cell_size"]) and np.array_equal(geom[9:12], g[name + "_cell_size_inv"])
    n = int(g[name + "_node_count"])
    assert len(nodes) == n
    ref0 = g[name + "_nodes"][:n]
    assert np.array_equal(nodes, ref0), "field 0 differs at %d of %d nodes, max %g" % ((nodes != ref0).sum(), n, np.abs(nodes - ref0).max())


def host_sampler(hc, v, t, T, res, mode, inverted=False):
    v, t, T = mesh_args(v, t, T)
    r = np.ascontiguousarray(res, np.uint32)
    geom, counts = np.zeros(15, np.float32), np.zeros(5, np.uint32)
    cap = 1 << 16
    xs, ys, zs = (np.zeros(cap, np.float32) for _ in range(3))
    args = (vp(v), len(v), vp(t), len(t), vp(T), int(inverted), vp(r), C.c_float(RADIUS), int(mode), vp(geom), vp(counts))
    assert hc.hc_sampler_grid(*args, None, vp(xs), vp(ys), vp(zs), cap) == 0
    nodes = np.zeros(int(counts[0]), np.float32)
    assert hc.hc_sampler_grid(*args, vp(nodes), vp(xs), vp(ys), vp(zs), cap) == 0
    xs, ys, zs = xs[:counts[2]].copy(), ys[:counts[3]].copy(), zs[:counts[4]].copy()
    cand = np.zeros((len(xs) * len(ys) * len(zs), 3), np.float32)
    hc.hc_lattice_positions(int(mode), C.c_float(RADIUS), vp(xs), len(xs), vp(ys), len(ys), vp(zs), len(zs), vp(cand))
    phi = sdf_numpy.interpolate(geom[0:3], geom[3:6], res, geom[6:9], geom[9:12], nodes, cand)
    return cand[phi < 0.0]


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_sampling_grid_lattice_and_inside_test_select_the_references_samples(hc, g, mode):
    got = host_sampler(hc, g["cone_verts"], g["cone_tris"], g["cone_T"], (20, 20, 20), mode)
    ref = g["sample_%d" % mode]
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(got, ref)


def test_the_references_sphere_tree_is_what_disagrees_not_the_distance(hc, g):
    """Why the fixtures hold modal outputs and the oracle builds until two builds agree: single builds of the reference's
    MeshDistance on the cone (cocircular rim vertices) — each with the next rand() state — against the minimum over all faces.
    A build either reproduces it (up to one point in 20 000 on an fp32 tie between faces) or misses nearest faces wholesale;
    agreeing builds always reproduce it.  The count of unsound builds is printed, not asserted (it depends on rand())."""
    from oracle import refsim
    if not refsim.available("cpu") or not hasattr(refsim._load("cpu"), "ref_mesh_signed_distance_once"):
        pytest.skip("oracle/_ref without the single-build hook")
    v, t, T, p = g["cone_verts"], g["cone_tris"], g["cone_T"], g["points_cone"][:5000]
    brute = host_sd(hc, v, t, T, p)
    wrong = [int((refsim.mesh_signed_distance(v, t, p, transform=T, once=True) != brute).sum()) for _ in range(24)]
    print("\nsingle builds: points differing from the minimum over all faces, per build:", wrong)
    for _ in range(6):
        agreed = refsim.mesh_signed_distance(v, t, p, transform=T)
        assert (agreed != brute).sum() <= 2 and np.abs(agreed - brute).max() < 5e-6
