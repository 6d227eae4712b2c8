"""vfd_b200.api — Python host-side mirror of the reference's solver interface over the C ABI.

`DFSPHSimulation` keeps the method names and argument meaning of the reference's
`vfd::DFSPHSimulation` façade (reference: VFD/Source/Simulation/DFSPH/DFSPHSimulator.h:10-56) so tests
read like calls into the reference; everything goes through libvfd_dfsph.so (include/vfd_dfsph.h)
with plain host buffers.  There is no CPU path: if the CUDA library cannot be loaded, or no device
is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VFD_LIB") or os.path.join(HERE, "lib", "libvfd_dfsph.so")   # VFD_LIB: a tuning variant built by build.py --out


class VfdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libvfd_dfsph error %d: %s" % (code, msg))
        self.code = code


class DFSPHSimulationDescription(C.Structure):
    """VfdDfsphDescription == DFSPHSimulationDescription (reference: Structures/DFSPHSimulationDescription.h:9-50)."""
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
        lib().vfd_dfsph_default_description(C.byref(self))
        g = kw.pop("Gravity", None)
        for k, v in kw.items():
            if k not in dict(self._fields_):
                raise AttributeError(k)
            setattr(self, k, v)
        if g is not None:
            self.Gravity[:] = g

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "Gravity"}
        d["Gravity"] = tuple(self.Gravity)
        return d


class VfdDfsphInfo(C.Structure):
    _fields_ = [
        ("ParticleCount", C.c_uint32), ("RigidBodyCount", C.c_uint32),
        ("SupportRadius", C.c_float), ("SupportRadius2", C.c_float), ("ParticleRadius", C.c_float), ("ParticleDiameter", C.c_float),
        ("TimeStepSize", C.c_float), ("TimeStepSize2", C.c_float), ("TimeStepSizeInverse", C.c_float), ("TimeStepSize2Inverse", C.c_float),
        ("Volume", C.c_float), ("Density0", C.c_float), ("ParticleMass", C.c_float), ("ParticleMassInverse", C.c_float),
        ("Viscosity", C.c_float), ("BoundaryViscosity", C.c_float), ("DynamicViscosity", C.c_float), ("DynamicBoundaryViscosity", C.c_float),
        ("TangentialDistanceFactor", C.c_float), ("TangentialDistance", C.c_float),
        ("SurfaceTension", C.c_float), ("SurfaceTensionSampleCount", C.c_uint32),
        ("ClassifierSlope", C.c_float), ("ClassifierConstant", C.c_float),
        ("TemporalSmoothing", C.c_uint8), ("_pad", C.c_uint8 * 3),
        ("SmoothingFactor", C.c_float), ("Factor", C.c_float), ("NeighborParticleRadius", C.c_float), ("MonteCarloFactor", C.c_float),
        ("Gravity", C.c_float * 3),
    ]


assert C.sizeof(VfdDfsphInfo) == 128


class VfdVolumeMap(C.Structure):
    _fields_ = [
        ("domainMin", C.c_float * 3), ("domainMax", C.c_float * 3), ("resolution", C.c_uint32 * 3),
        ("cellSize", C.c_float * 3), ("cellSizeInverse", C.c_float * 3),
        ("fieldCount", C.c_uint32), ("nodeCount", C.c_uint32), ("cellCount", C.c_uint32), ("cellMapCount", C.c_uint32),
        ("nodes", C.c_void_p), ("cells", C.c_void_p), ("cellMap", C.c_void_p),
    ]


class VfdDfsphDebugInfo(C.Structure):
    _fields_ = [
        ("NeighborhoodSearchUs", C.c_float), ("BaseSolverUs", C.c_float), ("DivergenceSolverUs", C.c_float),
        ("SurfaceTensionSolverUs", C.c_float), ("ViscositySolverUs", C.c_float), ("PressureSolverUs", C.c_float),
        ("IterationCount", C.c_uint32), ("DivergenceSolverIterationCount", C.c_uint32),
        ("PressureSolverIterationCount", C.c_uint32), ("ViscositySolverIterationCount", C.c_uint32),
        ("DivergenceSolverError", C.c_float), ("PressureSolverError", C.c_float), ("ViscositySolverError", C.c_float),
        ("FrameTime", C.c_float), ("FrameIndex", C.c_uint32),
    ]


# DFSPHParticle (120 B) and DFSPHParticleSimple (36 B)
PARTICLE_DTYPE = np.dtype([
    ("Position", "<f4", 3), ("Velocity", "<f4", 3), ("Acceleration", "<f4", 3), ("PressureAcceleration", "<f4", 3),
    ("PressureResiduum", "<f4"), ("Density", "<f4"), ("DensityAdvection", "<f4"), ("PressureRho2", "<f4"),
    ("PressureRho2V", "<f4"), ("Factor", "<f4"), ("VelocityDifference", "<f4", 3),
    ("MonteCarloSurfaceNormal", "<f4", 3), ("MonteCarloSurfaceNormalSmooth", "<f4", 3),
    ("MonteCarloSurfaceCurvature", "<f4"), ("MonteCarloSurfaceCurvatureSmooth", "<f4"), ("DeltaFinalCurvature", "<f4"),
])
PARTICLE_SIMPLE_DTYPE = np.dtype([("Position", "<f4", 3), ("Velocity", "<f4", 3), ("Acceleration", "<f4", 3)])
assert PARTICLE_DTYPE.itemsize == 120 and PARTICLE_SIMPLE_DTYPE.itemsize == 36

VFD_OPT_SEARCH_FMA, VFD_OPT_TIMERS, VFD_OPT_MAX_CELLS, VFD_OPT_KERNEL_TIMERS = 1, 2, 3, 4
STATE_NONE, STATE_SIMULATING, STATE_READY = 0, 1, 2

# every symbol include/vfd_dfsph.h declares: (name, restype, argtypes)
_vp, _u32, _u64, _f32, _i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int
SYMBOLS = [
    ("vfd_dfsph_default_description", None, [C.POINTER(DFSPHSimulationDescription)]),
    ("vfd_dfsph_create", _i, [C.POINTER(DFSPHSimulationDescription), _i, C.POINTER(_vp)]),
    ("vfd_dfsph_destroy", None, [_vp]),
    ("vfd_dfsph_last_error", C.c_char_p, [_vp]),
    ("vfd_dfsph_set_description", _i, [_vp, C.POINTER(DFSPHSimulationDescription)]),
    ("vfd_dfsph_get_description", _i, [_vp, C.POINTER(DFSPHSimulationDescription)]),
    ("vfd_dfsph_get_info", _i, [_vp, C.POINTER(VfdDfsphInfo)]),
    ("vfd_dfsph_set_particles", _i, [_vp, _vp, _vp, _u32]),
    ("vfd_dfsph_set_particles_device", _i, [_vp, _vp, _vp, _u32]),
    ("vfd_dfsph_set_rigid_bodies", _i, [_vp, _u32, C.POINTER(VfdVolumeMap)]),
    ("vfd_dfsph_simulate", _i, [_vp]),
    ("vfd_dfsph_begin", _i, [_vp]),
    ("vfd_dfsph_step", _i, [_vp]),
    ("vfd_dfsph_steps", _i, [_vp, _u32]),
    ("vfd_dfsph_synchronize", _i, [_vp]),
    ("vfd_dfsph_get_state", _i, [_vp]),
    ("vfd_dfsph_get_debug_info", _i, [_vp, C.POINTER(VfdDfsphDebugInfo)]),
    ("vfd_dfsph_get_max_velocity_magnitude", _f32, [_vp]),
    ("vfd_dfsph_get_current_time_step_size", _f32, [_vp]),
    ("vfd_dfsph_get_particle_count", _u32, [_vp]),
    ("vfd_dfsph_get_particle_radius", _f32, [_vp]),
    ("vfd_dfsph_get_rigid_body_count", _u32, [_vp]),
    ("vfd_dfsph_get_frame_count", _i, [_vp, C.POINTER(_u32)]),
    ("vfd_dfsph_get_frame", _i, [_vp, _u32, _vp, C.POINTER(_f32), C.POINTER(_f32)]),
    ("vfd_dfsph_get_frame_data", _i, [_vp, _u32, C.POINTER(_vp), C.POINTER(_u32), C.POINTER(_f32), C.POINTER(_f32)]),
    ("vfd_dfsph_get_current_frame", _i, [_vp, _vp]),
    ("vfd_dfsph_get_search_bytes", _i, [_vp, C.POINTER(_u64)]),
    ("vfd_dfsph_get_bounds", _i, [_vp, _vp, _vp]),
    ("vfd_dfsph_get_particles", _i, [_vp, _vp]),
    ("vfd_dfsph_set_particles_full", _i, [_vp, _vp]),
    ("vfd_dfsph_set_time_step", _i, [_vp, _f32]),
    ("vfd_dfsph_set_surface_tension_state", _i, [_vp, _u32, _f32]),
    ("vfd_dfsph_find_neighbors", _i, [_vp]),
    ("vfd_dfsph_get_neighbors", _i, [_vp, _vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    ("vfd_dfsph_get_boundary", _i, [_vp, _u32, _vp, _vp]),
    ("vfd_dfsph_get_kernel_tables", _i, [_vp, _vp, _vp, _vp]),
    ("vfd_dfsph_get_halton_table", _i, [_vp, _vp]),
    ("vfd_kernel_tables_build", _i, [_f32, _vp, _vp, _vp]),
    ("vfd_halton_table_build", _i, [_vp]),
    ("vfd_dfsph_set_option", _i, [_vp, _i, C.c_int64]),
    ("vfd_dfsph_get_launch_count", _i, [_vp, C.POINTER(_u64), _i]),
    ("vfd_dfsph_get_tile_stats", _i, [_vp, _vp]),
    ("vfd_dfsph_time_matvec", _i, [_vp, C.c_uint32, C.POINTER(C.c_float)]),
    ("vfd_dist_unique_id", _i, [_vp]),
    ("vfd_dfsph_init_distributed", _i, [_vp, _i, _i, _vp, _vp, _vp]),
    ("vfd_dfsph_get_grid", _i, [_vp, _vp, C.POINTER(_f32), _vp]),
    ("vfd_dfsph_set_slab", _i, [_vp, _u32, _u32]),
    ("vfd_dfsph_set_particles_distributed", _i, [_vp, _vp, _vp, _vp, _u32, _u32, _u32]),
    ("vfd_dfsph_get_owned", _i, [_vp, _u32, C.POINTER(_u32), _vp, _vp]),
    ("vfd_dfsph_get_comm_stats", _i, [_vp, _vp]),
    ("vfd_dfsph_get_slab", _i, [_vp, _vp]),
    ("vfd_dfsph_get_kernel_times", _i, [_vp, _u32, C.POINTER(_u32), _vp, _vp, _vp, _vp, _vp, _i]),
    ("vfd_dfsph_record_event", _i, [_vp, _u32]),
    ("vfd_dfsph_elapsed_ms", _i, [_vp, _u32, _u32, C.POINTER(_f32)]),
    ("vfd_volume_map_build_box", _i, [_vp, _vp, _i, _f32, _vp, _f32, _i, C.POINTER(VfdVolumeMap)]),
    ("vfd_volume_map_free", None, [C.POINTER(VfdVolumeMap)]),
    ("vfd_volume_map_build_mesh", _i, [_vp, _u32, _vp, _u32, _vp, _i, _f32, _vp, _f32, _i, C.POINTER(VfdVolumeMap)]),
    ("vfd_mesh_signed_distance", _i, [_vp, _u32, _vp, _u32, _vp, _vp, _u32, _i, _vp]),
    ("vfd_sample_mesh_volume", _i, [_vp, _u32, _vp, _u32, _vp, _f32, _vp, _i, _i, _i, C.POINTER(_vp), C.POINTER(_u32)]),
    ("vfd_free", None, [_vp]),
]

_lib = None


def lib():
    """Loads libvfd_dfsph.so (building it first when the sources are newer). Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libvfd_dfsph.so is missing (%s): build it with `python -m vfd_b200.build`; "
                          "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class VolumeMap:
    """Host-side volume map (the flattened SDF the reference uploads: SDF.cu:227-306)."""

    def __init__(self, domain_min, domain_max, resolution, cell_size, cell_size_inv, field_count, node_count,
                 cell_count, cell_map_count, nodes, cells, cell_map):
        self.domain_min = np.ascontiguousarray(domain_min, np.float32)
        self.domain_max = np.ascontiguousarray(domain_max, np.float32)
        self.resolution = np.ascontiguousarray(resolution, np.uint32)
        self.cell_size = np.ascontiguousarray(cell_size, np.float32)
        self.cell_size_inv = np.ascontiguousarray(cell_size_inv, np.float32)
        self.field_count, self.node_count = int(field_count), int(node_count)
        self.cell_count, self.cell_map_count = int(cell_count), int(cell_map_count)
        self.nodes = np.ascontiguousarray(nodes, np.float32)
        self.cells = np.ascontiguousarray(cells, np.uint32)
        self.cell_map = np.ascontiguousarray(cell_map, np.uint32)

    def c_struct(self):
        m = VfdVolumeMap()
        m.domainMin[:] = self.domain_min
        m.domainMax[:] = self.domain_max
        m.resolution[:] = [int(x) for x in self.resolution]
        m.cellSize[:] = self.cell_size
        m.cellSizeInverse[:] = self.cell_size_inv
        m.fieldCount, m.nodeCount, m.cellCount, m.cellMapCount = self.field_count, self.node_count, self.cell_count, self.cell_map_count
        m.nodes, m.cells, m.cellMap = _p(self.nodes).value, _p(self.cells).value, _p(self.cell_map).value
        return m

    @staticmethod
    def build_box(bmin, bmax, inverted=True, padding=0.0, resolution=(20, 20, 20), particle_radius=0.025, device=0):
        """GPU volume-map precompute for an axis-aligned box (reference: RigidBody.cu:32-72)."""
        a = np.ascontiguousarray(bmin, np.float32)
        b = np.ascontiguousarray(bmax, np.float32)
        r = np.ascontiguousarray(resolution, np.uint32)
        m = VfdVolumeMap()
        rc = lib().vfd_volume_map_build_box(_p(a), _p(b), 1 if inverted else 0, float(padding), _p(r), float(particle_radius), int(device), C.byref(m))
        return VolumeMap._take(rc, m)

    @staticmethod
    def build_mesh(vertices, triangles, transform=None, inverted=False, padding=0.0, resolution=(20, 20, 20), particle_radius=0.025, device=0):
        """GPU volume-map precompute for any closed triangle mesh under a (row-major 4x4) transform — RigidBody::RigidBody
        (RigidBody.cu:10-73) with the reference's mesh distance (MeshDistance.cpp:187-222)."""
        v, t, T = _mesh_args(vertices, triangles, transform)
        r = np.ascontiguousarray(resolution, np.uint32)
        m = VfdVolumeMap()
        rc = lib().vfd_volume_map_build_mesh(_p(v), len(v), _p(t), len(t), None if T is None else _p(T), 1 if inverted else 0, float(padding),
                                             _p(r), float(particle_radius), int(device), C.byref(m))
        return VolumeMap._take(rc, m)

    @staticmethod
    def _take(rc, m):
        if rc:
            raise VfdError(rc, (lib().vfd_dfsph_last_error(None) or b"").decode())
        try:
            nn, nc, nm = m.fieldCount * m.nodeCount, m.fieldCount * m.cellCount * 32, m.fieldCount * m.cellMapCount
            nodes = np.ctypeslib.as_array(C.cast(m.nodes, C.POINTER(C.c_float)), (nn,)).copy()
            cells = np.ctypeslib.as_array(C.cast(m.cells, C.POINTER(C.c_uint32)), (nc,)).copy()
            cmap = np.ctypeslib.as_array(C.cast(m.cellMap, C.POINTER(C.c_uint32)), (nm,)).copy()
            return VolumeMap(list(m.domainMin), list(m.domainMax), list(m.resolution), list(m.cellSize), list(m.cellSizeInverse),
                             m.fieldCount, m.nodeCount, m.cellCount, m.cellMapCount, nodes, cells, cmap)
        finally:
            lib().vfd_volume_map_free(C.byref(m))


def _mesh_args(vertices, triangles, transform):
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(triangles, np.uint32).reshape(-1, 3)
    T = None if transform is None else np.ascontiguousarray(np.asarray(transform, np.float32).reshape(4, 4).T).reshape(16)   # row-major -> glm columns
    return v, t, T


def mesh_signed_distance(vertices, triangles, points, transform=None, device=0):
    """MeshDistance::SignedDistance (MeshDistance.cpp:187-222) of a triangle mesh at `points`, on the GPU."""
    v, t, T = _mesh_args(vertices, triangles, transform)
    p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.zeros(len(p), np.float32)
    rc = lib().vfd_mesh_signed_distance(_p(v), len(v), _p(t), len(t), None if T is None else _p(T), _p(p), len(p), int(device), _p(out))
    if rc:
        raise VfdError(rc, "vfd_mesh_signed_distance")
    return out


def sample_mesh_volume(vertices, triangles, particle_radius=0.025, resolution=(20, 20, 20), inverted=False, sample_mode=0, transform=None, device=0):
    """FluidObject::FluidObject -> ParticleSampler::SampleMeshVolume (FluidObject.cpp:6-26, ParticleSampler.cpp:7-91) on the GPU:
    the particle positions inside the mesh, in the reference's order."""
    v, t, T = _mesh_args(vertices, triangles, transform)
    r = np.ascontiguousarray(resolution, np.uint32)
    ptr, n = C.c_void_p(), C.c_uint32(0)
    rc = lib().vfd_sample_mesh_volume(_p(v), len(v), _p(t), len(t), None if T is None else _p(T), float(particle_radius), _p(r),
                                      1 if inverted else 0, int(sample_mode), int(device), C.byref(ptr), C.byref(n))
    if rc:
        raise VfdError(rc, "vfd_sample_mesh_volume")
    try:
        if n.value == 0:
            return np.zeros((0, 3), np.float32)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), (n.value * 3,)).copy().reshape(-1, 3)
    finally:
        lib().vfd_free(ptr)


class FluidObject:
    """What the solver consumes of a reference FluidObject (FluidObject.h:22-32): sampled positions + one velocity.
    `velocities` (per particle) is an extension used by tests."""

    def __init__(self, positions, velocity=(0.0, 0.0, 0.0), velocities=None):
        self.positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        self.velocity = tuple(float(v) for v in velocity)
        self.velocities = None if velocities is None else np.ascontiguousarray(velocities, np.float32).reshape(-1, 3)

    def GetPositions(self):
        return self.positions

    def GetPositionCount(self):
        return self.positions.shape[0]

    def GetVelocity(self):
        return self.velocity


class DFSPHSimulation:
    """Drop-in for vfd::DFSPHSimulation / DFSPHImplementation on one B200."""

    def __init__(self, desc=None, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        self._desc = desc or DFSPHSimulationDescription()
        rc = self.L.vfd_dfsph_create(C.byref(self._desc), int(device), C.byref(self.h))
        if rc:
            raise VfdError(rc, (self.L.vfd_dfsph_last_error(None) or b"").decode())
        self.n = 0
        self._keep = []

    def close(self):
        if self.h:
            self.L.vfd_dfsph_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise VfdError(rc, (self.L.vfd_dfsph_last_error(self.h) or b"").decode())

    # ---- the reference's method set --------------------------------------------------------
    def SetDescription(self, desc):
        self._desc = desc
        self._ck(self.L.vfd_dfsph_set_description(self.h, C.byref(desc)))

    def GetDescription(self):
        d = DFSPHSimulationDescription()
        self._ck(self.L.vfd_dfsph_get_description(self.h, C.byref(d)))
        return d

    def GetInfo(self):
        i = VfdDfsphInfo()
        self._ck(self.L.vfd_dfsph_get_info(self.h, C.byref(i)))
        return i

    def SetFluidObjects(self, fluid_objects):
        """Concatenates the objects' sampled positions (DFSPHImplementation.cu:172-253)."""
        pos = [f.GetPositions() for f in fluid_objects]
        vel = [f.velocities if f.velocities is not None else np.tile(np.asarray(f.GetVelocity(), np.float32), (f.GetPositionCount(), 1))
               for f in fluid_objects]
        pos = np.ascontiguousarray(np.concatenate(pos) if pos else np.zeros((0, 3)), np.float32)
        vel = np.ascontiguousarray(np.concatenate(vel) if vel else np.zeros((0, 3)), np.float32)
        self.n = pos.shape[0]
        self._ck(self.L.vfd_dfsph_set_particles(self.h, _p(pos), _p(vel), self.n))

    def SetRigidBodies(self, volume_maps):
        arr = (VfdVolumeMap * max(len(volume_maps), 1))()
        for k, m in enumerate(volume_maps):
            arr[k] = m.c_struct()
        self._keep = list(volume_maps)
        self._ck(self.L.vfd_dfsph_set_rigid_bodies(self.h, len(volume_maps), arr))

    def Simulate(self):
        self._ck(self.L.vfd_dfsph_simulate(self.h))

    def OnUpdate(self):
        self._ck(self.L.vfd_dfsph_step(self.h))

    def GetSimulationState(self):
        return self.L.vfd_dfsph_get_state(self.h)

    def GetDebugInfo(self):
        d = VfdDfsphDebugInfo()
        self._ck(self.L.vfd_dfsph_get_debug_info(self.h, C.byref(d)))
        return d

    def GetMaxVelocityMagnitude(self):
        return self.L.vfd_dfsph_get_max_velocity_magnitude(self.h)

    def GetCurrentTimeStepSize(self):
        return self.L.vfd_dfsph_get_current_time_step_size(self.h)

    def GetParticleCount(self):
        return self.L.vfd_dfsph_get_particle_count(self.h)

    def GetParticleRadius(self):
        return self.L.vfd_dfsph_get_particle_radius(self.h)

    def GetRigidBodyCount(self):
        return self.L.vfd_dfsph_get_rigid_body_count(self.h)

    def GetFrameCount(self):
        k = C.c_uint32()
        self._ck(self.L.vfd_dfsph_get_frame_count(self.h, C.byref(k)))
        return k.value

    def GetFrame(self, index):
        """One baked DFSPHParticleFrame: (ParticleData[n], MaxVelocityMagnitude, CurrentTimeStep)."""
        out = np.zeros(getattr(self, "n_global", None) or self.n, PARTICLE_SIMPLE_DTYPE)      # a decomposed run bakes whole-scene frames on rank 0
        mv, dt = C.c_float(), C.c_float()
        self._ck(self.L.vfd_dfsph_get_frame(self.h, index, _p(out), C.byref(mv), C.byref(dt)))
        return out, mv.value, dt.value

    def GetFrameView(self, index):
        """The same frame where it lies in the handle's frame store — no copy (DFSPHParticleBuffer::GetFrame returns a reference
        too); a read-only array valid until the next bake, SetFluidObjects or close()."""
        ptr, cnt, mv, dt = C.c_void_p(), C.c_uint32(0), C.c_float(), C.c_float()
        self._ck(self.L.vfd_dfsph_get_frame_data(self.h, index, C.byref(ptr), C.byref(cnt), C.byref(mv), C.byref(dt)))
        if cnt.value == 0:
            return np.zeros(0, PARTICLE_SIMPLE_DTYPE), mv.value, dt.value
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (cnt.value * PARTICLE_SIMPLE_DTYPE.itemsize,))
        view = raw.view(PARTICLE_SIMPLE_DTYPE)
        view.flags.writeable = False
        return view, mv.value, dt.value

    # ---- extensions (no reference equivalent) ----------------------------------------------
    def begin(self):
        self._ck(self.L.vfd_dfsph_begin(self.h))

    def steps(self, k):
        self._ck(self.L.vfd_dfsph_steps(self.h, int(k)))

    def synchronize(self):
        self._ck(self.L.vfd_dfsph_synchronize(self.h))

    def set_option(self, opt, value):
        self._ck(self.L.vfd_dfsph_set_option(self.h, int(opt), int(value)))

    def particles(self):
        out = np.zeros(self.n, PARTICLE_DTYPE)
        self._ck(self.L.vfd_dfsph_get_particles(self.h, _p(out)))
        return out

    def set_particles_full(self, arr):
        arr = np.ascontiguousarray(arr, PARTICLE_DTYPE)
        assert arr.shape[0] == self.n
        self._ck(self.L.vfd_dfsph_set_particles_full(self.h, _p(arr)))

    def set_time_step(self, dt):
        self._ck(self.L.vfd_dfsph_set_time_step(self.h, float(dt)))

    def set_surface_tension_state(self, sample_count, mc_factor):
        self._ck(self.L.vfd_dfsph_set_surface_tension_state(self.h, int(sample_count), float(mc_factor)))

    def current_frame(self, out=None):
        if out is None:
            out = np.zeros(self.n, PARTICLE_SIMPLE_DTYPE)
        self._ck(self.L.vfd_dfsph_get_current_frame(self.h, _p(out)))
        return out

    def find_neighbors(self):
        self._ck(self.L.vfd_dfsph_find_neighbors(self.h))

    def neighbors(self):
        counts = np.zeros(self.n, np.uint32)
        offsets = np.zeros(self.n, np.uint32)
        total = C.c_uint64()
        self._ck(self.L.vfd_dfsph_get_neighbors(self.h, _p(counts), _p(offsets), None, 0, C.byref(total)))
        ids = np.zeros(max(total.value, 1), np.uint32)
        self._ck(self.L.vfd_dfsph_get_neighbors(self.h, _p(counts), _p(offsets), _p(ids), total.value, C.byref(total)))
        return counts, offsets, ids[:total.value]

    def boundary(self, body=0):
        xj = np.zeros((self.n, 3), np.float32)
        vol = np.zeros(self.n, np.float32)
        self._ck(self.L.vfd_dfsph_get_boundary(self.h, body, _p(xj), _p(vol)))
        return xj, vol

    def kernel_tables(self):
        W = np.zeros(10000, np.float32)
        G = np.zeros(10001, np.float32)
        sc = np.zeros(6, np.float32)
        self._ck(self.L.vfd_dfsph_get_kernel_tables(self.h, _p(W), _p(G), _p(sc)))
        return W, G, sc

    def halton_table(self):
        out = np.zeros(49152, np.float32)
        self._ck(self.L.vfd_dfsph_get_halton_table(self.h, _p(out)))
        return out

    def bounds(self):
        a = np.zeros(3, np.float32)
        b = np.zeros(3, np.float32)
        self._ck(self.L.vfd_dfsph_get_bounds(self.h, _p(a), _p(b)))
        return a, b

    def search_bytes(self):
        v = C.c_uint64()
        self._ck(self.L.vfd_dfsph_get_search_bytes(self.h, C.byref(v)))
        return v.value

    def launch_count(self, reset=False):
        v = C.c_uint64()
        self._ck(self.L.vfd_dfsph_get_launch_count(self.h, C.byref(v), 1 if reset else 0))
        return v.value

    def tile_stats(self):
        st = np.zeros(4, np.uint64)
        self._ck(self.L.vfd_dfsph_get_tile_stats(self.h, _p(st)))
        return dict(tiles=int(st[0]), cells=int(st[1]), fallback_tile_passes=int(st[2]), frame_bytes=int(st[3]))

    def time_matvec(self, reps=20):
        """ms per launch of the initial PCG mat-vec on the current state (tuning aid)"""
        ms = C.c_float()
        self._ck(self.L.vfd_dfsph_time_matvec(self.h, int(reps), C.byref(ms)))
        return ms.value / reps

    def kernel_times(self, reset=False):
        """{kernel class: (ms of launches that did work, number of such launches, ms of all launches, all launches)}"""
        cnt = C.c_uint32()
        self._ck(self.L.vfd_dfsph_get_kernel_times(self.h, 0, C.byref(cnt), None, None, None, None, None, 0))
        k = cnt.value
        names = (C.c_char_p * k)()
        ms, msa = np.zeros(k, np.float64), np.zeros(k, np.float64)
        ln, lna = np.zeros(k, np.uint64), np.zeros(k, np.uint64)
        self._ck(self.L.vfd_dfsph_get_kernel_times(self.h, k, C.byref(cnt), C.cast(names, C.c_void_p), _p(ms), _p(ln), _p(msa), _p(lna), 1 if reset else 0))
        return {names[i].decode(): (float(msa[i]), int(lna[i]), float(ms[i]), int(ln[i])) for i in range(k) if ln[i]}

    def record_event(self, slot):
        self._ck(self.L.vfd_dfsph_record_event(self.h, int(slot)))

    def elapsed_ms(self, a, b):
        ms = C.c_float()
        self._ck(self.L.vfd_dfsph_elapsed_ms(self.h, int(a), int(b), C.byref(ms)))
        return ms.value

    # ---- several GPUs (one process per GPU): slabs of tile columns along x -------------------------------
    def init_distributed(self, rank, world, unique_id, domain_min, domain_max):
        a = np.ascontiguousarray(domain_min, np.float32)
        b = np.ascontiguousarray(domain_max, np.float32)
        uid = np.frombuffer(bytes(unique_id), np.uint8).copy()
        assert uid.size == 128
        self._ck(self.L.vfd_dfsph_init_distributed(self.h, int(rank), int(world), _p(uid), _p(a), _p(b)))

    def grid(self):
        origin = np.zeros(3, np.float32)
        tiles = np.zeros(3, np.uint32)
        cell = C.c_float()
        self._ck(self.L.vfd_dfsph_get_grid(self.h, _p(origin), C.byref(cell), _p(tiles)))
        return origin, cell.value, tiles

    def set_slab(self, lo, hi):
        self._ck(self.L.vfd_dfsph_set_slab(self.h, int(lo), int(hi)))

    def set_particles_distributed(self, pos, vel, ids, n_global, capacity):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        ids = np.ascontiguousarray(ids, np.uint32)
        velp = None
        if vel is not None:
            vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 3)
            velp = _p(vel)
        self.n = pos.shape[0]
        self.n_global = int(n_global)
        self._ck(self.L.vfd_dfsph_set_particles_distributed(self.h, _p(pos), velp, _p(ids), self.n, int(n_global), int(capacity)))

    def owned(self):
        """(ids, particles) this rank owns now."""
        cnt = C.c_uint32()
        self._ck(self.L.vfd_dfsph_get_owned(self.h, 0, C.byref(cnt), None, None))
        ids = np.zeros(max(cnt.value, 1), np.uint32)
        out = np.zeros(max(cnt.value, 1), PARTICLE_DTYPE)
        self._ck(self.L.vfd_dfsph_get_owned(self.h, cnt.value, C.byref(cnt), _p(ids), _p(out)))
        return ids[:cnt.value], out[:cnt.value]

    def slab(self):
        """Owned tile columns [lo, hi) now, boundary moves so far, whether the ranks talk through peer memory."""
        st = np.zeros(4, np.uint64)
        self._ck(self.L.vfd_dfsph_get_slab(self.h, _p(st)))
        return {"lo": int(st[0]), "hi": int(st[1]), "shifts": int(st[2]), "peer_memory": bool(st[3])}

    def comm_stats(self):
        st = np.zeros(4, np.uint64)
        self._ck(self.L.vfd_dfsph_get_comm_stats(self.h, _p(st)))
        return dict(halos=int(st[0]), reductions=int(st[1]), halo_bytes=int(st[2]), state_bytes=int(st[3]))

    def set_particles_device(self, d_pos_ptr, d_vel_ptr, n):
        self.n = int(n)
        self._ck(self.L.vfd_dfsph_set_particles_device(self.h, C.c_void_p(d_pos_ptr), C.c_void_p(d_vel_ptr) if d_vel_ptr else None, self.n))


def dist_unique_id():
    """128 bytes identifying a new communicator (rank 0 calls it and hands the bytes to all ranks)."""
    buf = np.zeros(128, np.uint8)
    rc = lib().vfd_dist_unique_id(_p(buf))
    if rc:
        raise VfdError(rc, (lib().vfd_dfsph_last_error(None) or b"").decode())
    return buf.tobytes()


def kernel_tables(support_radius):
    """The reference's 10 000-entry W / gradW tables for a support radius (host arithmetic, no device needed)."""
    W = np.zeros(10000, np.float32)
    G = np.zeros(10001, np.float32)
    sc = np.zeros(6, np.float32)
    rc = lib().vfd_kernel_tables_build(float(support_radius), _p(W), _p(G), _p(sc))
    if rc:
        raise VfdError(rc, "vfd_kernel_tables_build")
    return W, G, sc


def halton_table():
    out = np.zeros(49152, np.float32)
    rc = lib().vfd_halton_table_build(_p(out))
    if rc:
        raise VfdError(rc, "vfd_halton_table_build")
    return out


def block_positions(nx, ny, nz, r=0.025, origin=(0.0, 0.0, 0.0)):
    """Synthetic lattice block at spacing 2r, positions (i + 1/2) 2r + origin, fp32
    (what SampleMode::MinDensity yields, reference: Utility/Sampler/ParticleSampler.cpp:39-43)."""
    d = np.float32(2.0 * r)
    i = (np.arange(nx, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[0])
    j = (np.arange(ny, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[1])
    k = (np.arange(nz, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[2])
    Z, Y, X = np.meshgrid(k, j, i, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
