"""The scene-preparation kernels EXECUTED on the CPU.  tests/host_check/build_emu.py compiles
the product's csrc/volume_map.cu — kernels and host code, line for line — against a small CUDA-on-CPU emulation
(tests/host_check/emu/: one OS thread per CUDA thread, pthread barriers, static shared memory) and these tests drive it through
the product's own Python wrappers (vfd_b200/api.py, with `lib()` pointed at the emulated library), against the same fixture
the GPU tests use (tests/golden/mesh.npz: outputs of the reference).  What this adds to tests/test_mesh_prep_cpu.py (the
per-point arithmetic) is the plumbing: ragged last blocks, triangle chunks staged between barriers, the ballot/popc
compaction with per-block offsets, the warp-shuffle reduction of the volume quadrature, the ctypes marshalling.  What it
cannot show is the hardware itself; tests/test_gpu_zmesh.py repeats the comparisons there."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "host_check"))
import meshes  # noqa: E402

R = 0.025


@pytest.fixture(scope="module")
def api():
    import build_emu
    from vfd_b200 import api as real
    L = C.CDLL(build_emu.build())
    for name, res, args in real.SYMBOLS:
        if hasattr(L, name):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
    assert all(hasattr(L, n) for n in ("vfd_volume_map_build_mesh", "vfd_volume_map_build_box", "vfd_mesh_signed_distance", "vfd_sample_mesh_volume", "vfd_free", "vfd_volume_map_free"))
    if not hasattr(L, "vfd_dfsph_last_error"):            # lives in api.cu, which is not part of the emulated translation unit
        L.vfd_dfsph_last_error = lambda h: b"(emulated library: no message)"
    saved = real.lib
    real.lib = lambda: L
    yield real
    real.lib = saved


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "mesh.npz"))


@pytest.mark.parametrize("name,count", [("cone", 3001), ("torus", 1500)])
def test_signed_distance_kernel(api, g, name, count):
    """62 faces: one chunk; 576 faces: two chunks of 512; a point count that leaves the last block ragged."""
    p = g["points_" + name][:count]
    sd = api.mesh_signed_distance(g[name + "_verts"], g[name + "_tris"], p, transform=g[name + "_T"])
    assert np.array_equal(sd, g["sd_" + name][:count])


def test_volume_map_kernels_on_a_non_cubic_grid(api, g):
    """The torus map of the fixture (5 x 9 x 7 cells, 2 984 nodes): field 0 bit for bit, field 1 within the quadrature's
    summation-order noise (one warp per node, lanes strided over the 4 096 points, shuffle reduction)."""
    vm = api.VolumeMap.build_mesh(g["torus_verts"], g["torus_tris"], transform=g["torus_T"], resolution=g["oddmap_resolution"], particle_radius=R)
    n = int(g["oddmap_node_count"])
    assert vm.node_count == n and vm.cell_count == 5 * 9 * 7
    assert np.array_equal(vm.domain_min, g["oddmap_domain_min"]) and np.array_equal(vm.cell_size_inv, g["oddmap_cell_size_inv"])
    ref0, ref1 = g["oddmap_nodes"][:n], g["oddmap_nodes"][n:]
    assert np.array_equal(vm.nodes[:n], ref0)
    scale = float(np.abs(ref1).max())
    err = float(np.abs(vm.nodes[n:] - ref1).max()) / scale
    print("\nemulated torus map: field 0 bit-exact, volume field max err %.2e of scale %.3g" % (err, scale))
    assert scale > 0 and err <= 3e-4
    # the cell tables are what the solver's lookup walks: every cell lists 32 distinct nodes, every node is listed
    cells = vm.cells.reshape(2, -1, 32)[0]
    assert all(len(set(c)) == 32 for c in cells.tolist()) and len(np.unique(cells)) == n


def test_box_map_kernel_agrees_with_the_mesh_map_of_the_same_box(api):
    """The analytic box distance (k_map_sdf, the map the bench builds) against the mesh distance of the box's 12 triangles:
    the reference's fp32 point-triangle quadratic agrees with the exact distance to ~1e-4 at this size."""
    lo, hi = (-0.3, 0.1, -0.2), (0.5, 0.6, 0.4)
    a = api.VolumeMap.build_box(lo, hi, inverted=True, padding=0.0, resolution=(4, 3, 5), particle_radius=R)
    v, t = meshes.box(lo, hi)
    b = api.VolumeMap.build_mesh(v, t, inverted=True, padding=0.0, resolution=(4, 3, 5), particle_radius=R)
    n = a.node_count
    assert n == b.node_count and np.array_equal(a.cells, b.cells) and np.array_equal(a.domain_min, b.domain_min)
    assert np.abs(a.nodes[:n] - b.nodes[:n]).max() < 5e-4
    assert np.abs(a.nodes[n:] - b.nodes[n:]).max() < 2e-3 * np.abs(a.nodes[n:]).max()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_sampling_kernels(api, g, mode):
    """Half the fixture's cone (2 400-3 000 samples out of ~10^5 lattice candidates in ~400 blocks): flag, per-block count,
    ordered scatter — against the reference's sampler run live."""
    T = g["cone_T"].copy(); T[:3, :3] *= 0.5; T[1, 3] = 1.1
    from oracle import refsim
    if not refsim.available("cpu") or not hasattr(refsim._load("cpu"), "ref_sample_mesh_volume"):
        pytest.skip("oracle/_ref without the sampler hook")
    got = api.sample_mesh_volume(g["cone_verts"], g["cone_tris"], R, (20, 20, 20), False, mode, transform=T)
    ref = refsim.sample_mesh_volume(g["cone_verts"], g["cone_tris"], R, (20, 20, 20), False, mode, transform=T)
    assert 2000 < len(ref) < 4000
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_sampling_edge_cases(api):
    v, t = meshes.box((0.0, 0.0, 0.0), (0.5, 0.4, 0.3))
    got = api.sample_mesh_volume(v, t, R, (20, 20, 20), False, 0)
    assert len(got) == 10 * 8 * 6
    cells = np.floor(got / 0.05).astype(np.int64)
    assert len(np.unique(cells, axis=0)) == len(got) and np.abs(got - (cells + 0.5) * 0.05).max() < 1e-5
    assert len(api.sample_mesh_volume(v, t, R, (20, 20, 20), True, 0)) == 0            # inverted: nothing inside the bounds
    bad = t.copy(); bad[3, 1] = 99
    with pytest.raises(api.VfdError):
        api.sample_mesh_volume(v, bad, R)
    # a grid whose 32-bit node / cell counts (the interchange format's) would wrap is refused, not built
    for res in ((1024, 1024, 1024), (900, 900, 900), (0, 4, 4)):
        with pytest.raises(api.VfdError):
            api.VolumeMap.build_mesh(v, t, resolution=res, particle_radius=R)
        with pytest.raises(api.VfdError):
            api.sample_mesh_volume(v, t, R, res, False, 0)


def test_the_gpu_tests_themselves_pass_on_the_emulation(api, g):
    """tests/test_gpu_zmesh.py's own test functions (the scene-preparation ones), called with the emulated library behind the
    wrappers: their expectations and the fixture's full-size cases hold before the hardware run.  (The larger ones — the slab's
    62 181-node map, sampling modes 0 and 2 of the whole cone — pass the same way in ~70 s; they are left to the GPU.)"""
    import test_gpu_zmesh as T
    for name in ("cone", "torus"):
        T.test_mesh_signed_distance_matches_the_reference(None, g, name)
    T.test_mesh_volume_map_matches_the_references(None, g, "conemap", "cone")
    T.test_mesh_volume_sampling_matches_the_references(None, g, 1)
    T.test_sampling_a_box_mesh_is_the_lattice_block(None)


def test_the_headless_scene_of_the_gpu_suite_is_prepared_like_the_reference(api):
    """tests/test_gpu_zmesh.py::test_scene_file_runs_headless bakes a scene file through scene_io.build_simulation, i.e. through
    these kernels: its fluid block (a unit cube scaled to 0.5^3, MinDensity sampling) and its slab (20 x 10 x 20 map) prepared by the
    emulated library are the reference's — the 1 000 particles that test expects, the distance field bit for bit."""
    from oracle import refsim
    from vfd_b200 import scene_io
    if not refsim.available("cpu") or not hasattr(refsim._load("cpu"), "ref_add_mesh_body"):
        pytest.skip("oracle/_ref without the mesh hooks")
    v, t = scene_io.UNIT_CUBE
    block = np.eye(4, dtype=np.float32); block[:3, :3] = np.diag([0.25, 0.25, 0.25]); block[:3, 3] = [0.0, 0.6, 0.0]
    floor = np.eye(4, dtype=np.float32); floor[:3, :3] = np.diag([1.0, 0.1, 1.0])
    pos = api.sample_mesh_volume(v, t, R, (20, 20, 20), False, 0, transform=block)
    vm = api.VolumeMap.build_mesh(v, t, transform=floor, inverted=False, padding=0.0, resolution=(20, 10, 20), particle_radius=R)
    with refsim.quiet_stdout():
        rpos = refsim.sample_mesh_volume(v, t, R, (20, 20, 20), False, 0, transform=block)
        sim = refsim.RefSim(refsim.Desc(ParticleRadius=R))
        sim.set_particles(np.zeros((1, 3), np.float32))
        sim.add_mesh_body(v, t, transform=floor, inverted=False, padding=0.0, res=(20, 10, 20))
        m = sim.volume_map(0)
    assert len(pos) == 1000 and rpos.shape == pos.shape and np.array_equal(rpos, pos)
    n = int(m["node_count"])
    assert vm.node_count == n and np.array_equal(vm.nodes[:n], m["nodes"][:n])
    assert np.abs(vm.nodes[n:] - m["nodes"][n:]).max() <= 3e-4 * np.abs(m["nodes"][n:]).max()


def test_the_box_map_test_of_the_gpu_suite_passes_on_the_emulation(api):
    """The box-map kernels (k_map_sdf, k_map_volume: the map every bench scene and most GPU tests are built with) changed after
    their last hardware run — the node numbering was rewritten (map_geometry.cuh).  The GPU suite's own check of them against the
    reference's host precompute, run here on the emulated library: same numbers as the hardware run printed before the rewrite
    (distance field 1.05e-05, volume field 3.85e-05 of scale)."""
    import test_gpu_scale as T
    T.test_box_volume_map_built_on_the_gpu_matches_the_reference(None)
