"""oracle/refsim.py — TEST INFRASTRUCTURE (never imported by vfd_b200/).

ctypes binding of oracle/_ref/libvfd_ref_{cpu,gpu}.so, i.e. of the reference's own DFSPH solver
sources (see oracle/build_ref.py, oracle/ref_driver.cpp).  Used by tests/ as the parity oracle,
by tests/golden/make_golden.py to generate fixtures, and by bench.py's cpu_baseline /
--impl reference legs.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Desc(C.Structure):
    """Field-for-field DFSPHSimulationDescription
    (reference: VFD/Source/Simulation/DFSPH/Structures/DFSPHSimulationDescription.h:9-50), defaults included."""
    _fields_ = [
        ("TimeStepSize", C.c_float), ("MinTimeStepSize", C.c_float), ("MaxTimeStepSize", C.c_float),
        ("FrameLength", C.c_float), ("FrameCount", C.c_uint32),
        ("MinPressureSolverIterations", C.c_uint32), ("MaxPressureSolverIterations", C.c_uint32),
        ("MaxPressureSolverError", C.c_float),
        ("EnableDivergenceSolverError", C.c_uint32), ("MinDivergenceSolverIterations", C.c_uint32),
        ("MaxDivergenceSolverIterations", C.c_uint32), ("MaxDivergenceSolverError", C.c_float),
        ("EnableViscositySolver", C.c_uint32), ("MinViscositySolverIterations", C.c_uint32),
        ("MaxViscositySolverIterations", C.c_uint32), ("MaxViscositySolverError", C.c_float),
        ("Viscosity", C.c_float), ("BoundaryViscosity", C.c_float), ("TangentialDistanceFactor", C.c_float),
        ("EnableSurfaceTensionSolver", C.c_uint32), ("SurfaceTensionSmoothPassCount", C.c_uint32),
        ("SurfaceTension", C.c_float), ("TemporalSmoothing", C.c_uint32),
        ("CSDFix", C.c_int32), ("CSD", C.c_int32),
        ("ParticleRadius", C.c_float), ("Gravity", C.c_float * 3),
    ]

    def __init__(self, **kw):
        super().__init__()
        d = dict(TimeStepSize=0.001, MinTimeStepSize=0.0001, MaxTimeStepSize=0.005, FrameLength=0.0016, FrameCount=200,
                 MinPressureSolverIterations=0, MaxPressureSolverIterations=100, MaxPressureSolverError=10.0,
                 EnableDivergenceSolverError=1, MinDivergenceSolverIterations=0, MaxDivergenceSolverIterations=100,
                 MaxDivergenceSolverError=10.0,
                 EnableViscositySolver=1, MinViscositySolverIterations=0, MaxViscositySolverIterations=100,
                 MaxViscositySolverError=0.1, Viscosity=10.0, BoundaryViscosity=10.0, TangentialDistanceFactor=0.5,
                 EnableSurfaceTensionSolver=1, SurfaceTensionSmoothPassCount=1, SurfaceTension=1.0, TemporalSmoothing=0,
                 CSDFix=-1, CSD=10000, ParticleRadius=0.025)
        d.update(kw)
        g = d.pop("Gravity", (0.0, -9.81, 0.0))
        for k, v in d.items():
            setattr(self, k, v)
        self.Gravity[:] = g


# DFSPHParticle, 120 B AoS (reference: Structures/DFSPHParticle.h:8-32)
PARTICLE_DTYPE = np.dtype([
    ("Position", "<f4", 3), ("Velocity", "<f4", 3), ("Acceleration", "<f4", 3), ("PressureAcceleration", "<f4", 3),
    ("PressureResiduum", "<f4"), ("Density", "<f4"), ("DensityAdvection", "<f4"), ("PressureRho2", "<f4"),
    ("PressureRho2V", "<f4"), ("Factor", "<f4"), ("VelocityDifference", "<f4", 3),
    ("MonteCarloSurfaceNormal", "<f4", 3), ("MonteCarloSurfaceNormalSmooth", "<f4", 3),
    ("MonteCarloSurfaceCurvature", "<f4"), ("MonteCarloSurfaceCurvatureSmooth", "<f4"), ("DeltaFinalCurvature", "<f4"),
])
assert PARTICLE_DTYPE.itemsize == 120


def lib_path(kind="cpu"):
    name = {"cpu": "libvfd_ref_cpu.so", "gpu": "libvfd_ref_gpu.so", "gpu_fast": "libvfd_ref_gpu_fast.so"}[kind]
    return os.path.join(HERE, "_ref", name)


def available(kind="cpu"):
    return os.path.exists(lib_path(kind))


_libs = {}


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    L = C.CDLL(lib_path(kind))
    vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
    L.ref_create.restype = vp
    L.ref_create.argtypes = [C.POINTER(Desc)]
    L.ref_destroy.argtypes = [vp]
    L.ref_set_description.argtypes = [vp, C.POINTER(Desc)]
    L.ref_set_particles.argtypes = [vp, vp, vp, u32]
    L.ref_add_box_body.argtypes = [vp, vp, vp, C.c_int, f32, vp]
    L.ref_commit_bodies.argtypes = [vp]
    L.ref_get_map_sizes.argtypes = [vp, u32, vp, vp, vp, vp]
    L.ref_get_map_data.argtypes = [vp, u32, vp, vp, vp]
    L.ref_begin.argtypes = [vp]
    L.ref_step.argtypes = [vp]
    L.ref_particle_count.argtypes = [vp]
    L.ref_particle_count.restype = u32
    L.ref_get_particles.argtypes = [vp, vp]
    L.ref_set_particles_full.argtypes = [vp, vp]
    L.ref_set_time_step.argtypes = [vp, f32]
    L.ref_set_st_state.argtypes = [vp, u32, f32]
    L.ref_get_info.argtypes = [vp, vp]
    L.ref_info_size.restype = u32
    L.ref_particle_size.restype = u32
    L.ref_kernel_size.restype = u32
    L.ref_get_kernel.argtypes = [vp, vp]
    L.ref_get_debug.argtypes = [vp, vp, vp, vp]
    L.ref_get_neighbors.argtypes = [vp, vp, vp, vp, u32]
    L.ref_get_neighbors.restype = u32
    L.ref_find_neighbors.argtypes = [vp]
    L.ref_get_boundary.argtypes = [vp, u32, vp, vp]
    if hasattr(L, "ref_sample_mesh_volume"):
        L.ref_sample_mesh_volume.restype = u32
        L.ref_sample_mesh_volume.argtypes = [vp, u32, vp, u32, vp, f32, vp, C.c_int, C.c_int, vp, u32]
    if hasattr(L, "ref_mesh_signed_distance"):
        L.ref_mesh_signed_distance.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp]
        L.ref_add_mesh_body.argtypes = [vp, vp, u32, vp, u32, vp, C.c_int, f32, vp]
    if hasattr(L, "ref_mesh_signed_distance_once"):
        L.ref_mesh_signed_distance_once.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp]
    L.ref_set_serial.argtypes = [C.c_int]
    L.ref_set_threads.argtypes = [C.c_int]
    L.ref_get_max_threads.restype = C.c_int
    assert L.ref_particle_size() == 120 and L.ref_info_size() == 128
    _libs[kind] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


BOX_VERTS = lambda a, b: np.array([[a[0], a[1], a[2]], [b[0], a[1], a[2]], [b[0], a[1], b[2]], [a[0], a[1], b[2]],
                                    [a[0], b[1], a[2]], [b[0], b[1], a[2]], [b[0], b[1], b[2]], [a[0], b[1], b[2]]], np.float32)
BOX_TRIS = np.array([[0, 1, 2], [0, 2, 3], [4, 7, 6], [4, 6, 5], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7]], np.uint32)


def sample_mesh_volume(verts, tris, radius, resolution=(20, 20, 20), inverted=False, mode=0, transform=None, kind="cpu"):
    """The reference's FluidObject sampling (FluidObject.cpp:6-26 + ParticleSampler.cpp:7-91) of a raw triangle mesh."""
    L = _load(kind)
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
    r = np.ascontiguousarray(resolution, np.uint32)
    T = None if transform is None else np.ascontiguousarray(np.asarray(transform, np.float32).T).reshape(16)   # row-major matrix -> glm columns
    n = L.ref_sample_mesh_volume(_p(v), len(v), _p(t), len(t), None if T is None else _p(T), float(radius), _p(r), int(inverted), int(mode), None, 0)
    out = np.zeros((max(n, 1), 3), np.float32)
    L.ref_sample_mesh_volume(_p(v), len(v), _p(t), len(t), None if T is None else _p(T), float(radius), _p(r), int(inverted), int(mode), _p(out), n)
    return out[:n]


def _mesh_args(verts, tris, transform):
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
    T = None if transform is None else np.ascontiguousarray(np.asarray(transform, np.float32).T).reshape(16)   # row-major matrix -> glm columns
    return v, t, T


def mesh_signed_distance(verts, tris, points, transform=None, kind="cpu", once=False):
    """The reference's MeshDistance::SignedDistance (MeshDistance.cpp:187-222) of a raw triangle mesh at `points`: from two
    builds of its randomised sphere tree that agree (oracle/ref_driver.cpp), or — `once` — from one build, as is."""
    L = _load(kind)
    v, t, T = _mesh_args(verts, tris, transform)
    p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.zeros(len(p), np.float32)
    (L.ref_mesh_signed_distance_once if once else L.ref_mesh_signed_distance)(_p(v), len(v), _p(t), len(t), None if T is None else _p(T), _p(p), len(p), _p(out))
    return out


class quiet_stdout:
    """The reference prints banners to std::cout (DFSPHImplementation.cu:174-252 ...); callers that must keep
    stdout clean (bench.py prints one JSON line) route fd 1 to fd 2 while the reference runs."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


class RefSim:
    """The reference solver (its own sources) behind the calls the editor makes."""

    def __init__(self, desc=None, kind="cpu", serial=False, threads=None):
        self.L = _load(kind)
        self.kind = kind
        self.desc = desc or Desc()
        self.L.ref_set_serial(1 if serial else 0)
        if threads:
            self.L.ref_set_threads(int(threads))
        self.h = self.L.ref_create(C.byref(self.desc))
        self.nbodies = 0

    def max_threads(self):
        return self.L.ref_get_max_threads()

    def set_description(self, desc):
        self.desc = desc
        self.L.ref_set_description(self.h, C.byref(desc))

    def set_particles(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        n = pos.shape[0]
        velp = None
        if vel is not None:
            vel = np.ascontiguousarray(vel, dtype=np.float32)
            velp = _p(vel)
        self.L.ref_set_particles(self.h, _p(pos), velp, n)
        self.n = n

    def add_box_body(self, bmin, bmax, inverted=True, padding=0.0, res=(20, 20, 20)):
        a = np.asarray(bmin, dtype=np.float32)
        b = np.asarray(bmax, dtype=np.float32)
        r = np.asarray(res, dtype=np.uint32)
        self.L.ref_add_box_body(self.h, _p(a), _p(b), 1 if inverted else 0, float(padding), _p(r))
        self.nbodies += 1

    def add_mesh_body(self, verts, tris, transform=None, inverted=False, padding=0.0, res=(20, 20, 20)):
        """A rigid body from a raw triangle mesh under a transform (row-major 4x4), through RigidBody::RigidBody."""
        v, t, T = _mesh_args(verts, tris, transform)
        r = np.asarray(res, dtype=np.uint32)
        self.L.ref_add_mesh_body(self.h, _p(v), len(v), _p(t), len(t), None if T is None else _p(T), 1 if inverted else 0, float(padding), _p(r))
        self.nbodies += 1

    def commit_bodies(self):
        self.L.ref_commit_bodies(self.h)

    def volume_map(self, body=0):
        """The flattened SDFDeviceData of a body (SDF.cu:227-306) as a dict of numpy arrays."""
        sizes = np.zeros(7, np.uint32)
        dom = np.zeros(6, np.float32)
        cs = np.zeros(3, np.float32)
        ci = np.zeros(3, np.float32)
        self.L.ref_get_map_sizes(self.h, body, _p(sizes), _p(dom), _p(cs), _p(ci))
        fc, nc, cc, cmc = (int(x) for x in sizes[:4])
        nodes = np.zeros(fc * nc, np.float32)
        cells = np.zeros(fc * cc * 32, np.uint32)
        cmap = np.zeros(fc * cmc, np.uint32)
        self.L.ref_get_map_data(self.h, body, _p(nodes), _p(cells), _p(cmap))
        return dict(domain_min=dom[:3].copy(), domain_max=dom[3:].copy(), resolution=sizes[4:7].copy(),
                    cell_size=cs, cell_size_inv=ci, field_count=fc, node_count=nc, cell_count=cc, cell_map_count=cmc,
                    nodes=nodes, cells=cells, cell_map=cmap)

    def begin(self):
        self.L.ref_begin(self.h)

    def step(self, k=1):
        for _ in range(k):
            self.L.ref_step(self.h)

    def particles(self):
        out = np.zeros(self.n, PARTICLE_DTYPE)
        self.L.ref_get_particles(self.h, _p(out))
        return out

    def set_particles_full(self, arr):
        arr = np.ascontiguousarray(arr, dtype=PARTICLE_DTYPE)
        self.L.ref_set_particles_full(self.h, _p(arr))

    def set_time_step(self, dt):
        self.L.ref_set_time_step(self.h, float(dt))

    def set_st_state(self, sample_count, mc_factor):
        self.L.ref_set_st_state(self.h, int(sample_count), float(mc_factor))

    def info_bytes(self):
        b = np.zeros(128, np.uint8)
        self.L.ref_get_info(self.h, _p(b))
        return b

    def kernel_bytes(self):
        b = np.zeros(self.L.ref_kernel_size(), np.uint8)
        self.L.ref_get_kernel(self.h, _p(b))
        return b

    def debug(self):
        it = np.zeros(4, np.uint32)
        f = np.zeros(6, np.float32)
        t = np.zeros(6, np.float32)
        self.L.ref_get_debug(self.h, _p(it), _p(f), _p(t))
        return dict(steps=int(it[0]), div_it=int(it[1]), press_it=int(it[2]), visc_it=int(it[3]),
                    div_err=float(f[0]), press_err=float(f[1]), visc_err=float(f[2]), frame_time=float(f[3]),
                    dt=float(f[4]), max_vel2=float(f[5]),
                    timers_us=dict(zip(("search", "base", "divergence", "surface_tension", "viscosity", "pressure"),
                                       (float(x) for x in t))))

    def find_neighbors(self):
        self.L.ref_find_neighbors(self.h)

    def neighbors(self):
        counts = np.zeros(self.n, np.uint32)
        offsets = np.zeros(self.n, np.uint32)
        total = self.L.ref_get_neighbors(self.h, _p(counts), _p(offsets), None, 0)
        ids = np.zeros(max(total, 1), np.uint32)
        self.L.ref_get_neighbors(self.h, _p(counts), _p(offsets), _p(ids), total)
        return counts, offsets, ids[:total]

    def boundary(self, body=0):
        xj = np.zeros((self.n, 3), np.float32)
        vol = np.zeros(self.n, np.float32)
        self.L.ref_get_boundary(self.h, body, _p(xj), _p(vol))
        return xj, vol


def block_positions(nx, ny, nz, r=0.025, origin=(0.0, 0.0, 0.0)):
    """Lattice block at spacing 2r, positions (i + 1/2)·2r + origin — what SampleMode::MinDensity
    yields (reference: Utility/Sampler/ParticleSampler.cpp:39-43); fp32 like the reference."""
    d = np.float32(2.0 * r)
    i = (np.arange(nx, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[0])
    j = (np.arange(ny, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[1])
    k = (np.arange(nz, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[2])
    Z, Y, X = np.meshgrid(k, j, i, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
