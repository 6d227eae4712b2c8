/* vfd_dfsph.h — C ABI of libvfd_dfsph.so: a B200-native (sm_100a) implementation of VFD's DFSPH
 * solver step, a drop-in for the reference's  class DFSPHSimulation
 * (reference: VFD/Source/Simulation/DFSPH/DFSPHSimulator.h:10-56, which forwards 1:1 to
 *  DFSPHImplementation, VFD/Source/Simulation/DFSPH/DFSPHImplementation.h:23-137).
 *
 * Plain pointers and sizes only; no C++/torch types.  Ownership: the handle owns all device
 * memory; every pointer argument is caller-owned and only read/written during the call.
 * Errors: every function returns 0 on success or a VFD_E_* code; vfd_dfsph_last_error() gives
 * the message.  The library never calls exit() (the reference prints and exits on any CUDA
 * error: VFD/Source/Compute/Utility/CUDA/cutil.h:772-778).  Threading: one mutating thread plus
 * concurrent read-only getters (the reference polls state/debug info from its UI thread while a
 * worker runs Simulate(): VFD/Source/Simulation/DFSPH/DFSPHSimulator.cpp:156-166).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * VFD_E_CUDA. */
#ifndef VFD_DFSPH_H
#define VFD_DFSPH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFD_OK            0
#define VFD_E_INVALID     1   /* bad argument / call order */
#define VFD_E_CUDA        2   /* CUDA runtime error (message in last_error) */
#define VFD_E_CAPACITY    3   /* search grid or buffer larger than the configured limits */
#define VFD_E_NCCL        4

typedef struct VfdDfsph VfdDfsph;

/* Field-for-field DFSPHSimulationDescription with fixed-width types; the names are the
 * reference's (and its cereal NVPs: VFD/Source/Scene/Components/DFSPHSimulationComponent.h:11-37).
 * Reference: VFD/Source/Simulation/DFSPH/Structures/DFSPHSimulationDescription.h:9-50. */
typedef struct VfdDfsphDescription {
    float    TimeStepSize;                    /* 0.001  */
    float    MinTimeStepSize;                 /* 0.0001 */
    float    MaxTimeStepSize;                 /* 0.005  */
    float    FrameLength;                     /* 0.0016 */
    uint32_t FrameCount;                      /* 200    */
    uint32_t MinPressureSolverIterations;     /* 0   */
    uint32_t MaxPressureSolverIterations;     /* 100 */
    float    MaxPressureSolverError;          /* 10 [%] */
    uint32_t EnableDivergenceSolverError;     /* bool, 1 */
    uint32_t MinDivergenceSolverIterations;   /* 0   */
    uint32_t MaxDivergenceSolverIterations;   /* 100 */
    float    MaxDivergenceSolverError;        /* 10 [%] */
    uint32_t EnableViscositySolver;           /* bool, 1 */
    uint32_t MinViscositySolverIterations;    /* 0   */
    uint32_t MaxViscositySolverIterations;    /* 100 */
    float    MaxViscositySolverError;         /* 0.1 [%] */
    float    Viscosity;                       /* 10  */
    float    BoundaryViscosity;               /* 10  */
    float    TangentialDistanceFactor;        /* 0.5 */
    uint32_t EnableSurfaceTensionSolver;      /* bool, 1 */
    uint32_t SurfaceTensionSmoothPassCount;   /* 1 */
    float    SurfaceTension;                  /* 1 */
    uint32_t TemporalSmoothing;               /* bool, 0 */
    int32_t  CSDFix;                          /* -1 */
    int32_t  CSD;                             /* 10000 */
    float    ParticleRadius;                  /* 0.025 */
    float    Gravity[3];                      /* 0, -9.81, 0 */
} VfdDfsphDescription;

/* Fills *d with the reference's defaults (values above). */
void vfd_dfsph_default_description(VfdDfsphDescription* d);

/* Byte-for-byte DFSPHSimulationInfo (128 B; `bool TemporalSmoothing` at offset 96).
 * Reference: VFD/Source/Simulation/DFSPH/Structures/DFSPHSimulationInfo.h:5-44. */
typedef struct VfdDfsphInfo {
    uint32_t ParticleCount, RigidBodyCount;
    float SupportRadius, SupportRadius2, ParticleRadius, ParticleDiameter;
    float TimeStepSize, TimeStepSize2, TimeStepSizeInverse, TimeStepSize2Inverse;
    float Volume, Density0, ParticleMass, ParticleMassInverse;
    float Viscosity, BoundaryViscosity, DynamicViscosity, DynamicBoundaryViscosity;
    float TangentialDistanceFactor, TangentialDistance;
    float SurfaceTension;
    uint32_t SurfaceTensionSampleCount;
    float ClassifierSlope, ClassifierConstant;
    uint8_t TemporalSmoothing; uint8_t _pad[3];
    float SmoothingFactor, Factor, NeighborParticleRadius, MonteCarloFactor;
    float Gravity[3];
} VfdDfsphInfo;

/* Byte-for-byte DFSPHParticle (120 B).  Reference: .../Structures/DFSPHParticle.h:8-32. */
typedef struct VfdParticle {
    float Position[3], Velocity[3], Acceleration[3], PressureAcceleration[3];
    float PressureResiduum, Density, DensityAdvection, PressureRho2, PressureRho2V, Factor;
    float VelocityDifference[3];
    float MonteCarloSurfaceNormal[3], MonteCarloSurfaceNormalSmooth[3];
    float MonteCarloSurfaceCurvature, MonteCarloSurfaceCurvatureSmooth, DeltaFinalCurvature;
} VfdParticle;

/* Byte-for-byte DFSPHParticleSimple (36 B), the baked-frame / renderer format.
 * Reference: .../Structures/DFSPHParticleSimple.h:8-13. */
typedef struct VfdParticleSimple { float Position[3], Velocity[3], Acceleration[3]; } VfdParticleSimple;

/* One rigid body = the flattened SDFDeviceData the reference uploads
 * (VFD/Source/Utility/SDF/SDFDeviceData.cuh:552-566, flattening in SDF.cu:234-283).
 * Field 0 = signed distance, field 1 = boundary volume (Bender 2019 volume maps).
 * nodes: fieldCount x nodeCount floats; cells: fieldCount x cellCount x 32 node ids;
 * cellMap: fieldCount x cellMapCount.  Host pointers; copied by set_rigid_bodies. */
typedef struct VfdVolumeMap {
    float    domainMin[3], domainMax[3];
    uint32_t resolution[3];
    float    cellSize[3], cellSizeInverse[3];
    uint32_t fieldCount, nodeCount, cellCount, cellMapCount;
    const float*    nodes;
    const uint32_t* cells;
    const uint32_t* cellMap;
} VfdVolumeMap;

/* DFSPHDebugInfo with the six phase timers in microseconds (measured with CUDA events).
 * Reference: VFD/Source/Simulation/DFSPH/Structures/DFSPHDebugInfo.h:8-28. */
typedef struct VfdDfsphDebugInfo {
    float NeighborhoodSearchUs, BaseSolverUs, DivergenceSolverUs, SurfaceTensionSolverUs, ViscositySolverUs, PressureSolverUs;
    uint32_t IterationCount, DivergenceSolverIterationCount, PressureSolverIterationCount, ViscositySolverIterationCount;
    float DivergenceSolverError, PressureSolverError, ViscositySolverError;
    float FrameTime;
    uint32_t FrameIndex;
} VfdDfsphDebugInfo;

enum { VFD_STATE_NONE = 0, VFD_STATE_SIMULATING = 1, VFD_STATE_READY = 2 };  /* DFSPHImplementation.h:27-32 */

/* ---- lifetime: ctor DFSPHImplementation.cu:14-21, dtor :23-31 (reference hard-codes device 0) ---- */
int  vfd_dfsph_create(const VfdDfsphDescription* desc, int device, VfdDfsph** out);
void vfd_dfsph_destroy(VfdDfsph* h);
const char* vfd_dfsph_last_error(const VfdDfsph* h);   /* h may be NULL: last create() error */

/* ---- configuration: SetDescription :318-368, GetDescription :313, GetInfo :370 ---- */
int vfd_dfsph_set_description(VfdDfsph* h, const VfdDfsphDescription* desc);
int vfd_dfsph_get_description(const VfdDfsph* h, VfdDfsphDescription* out);
int vfd_dfsph_get_info(VfdDfsph* h, VfdDfsphInfo* out);

/* ---- scene: SetFluidObjects :172-253 (after host-side sampling), SetRigidBodies :255-271 ----
 * pos/vel: n x 3 floats (vel may be NULL = 0).  All other per-particle fields start at 0.
 * As in the reference, particles must be set before rigid bodies. */
int vfd_dfsph_set_particles(VfdDfsph* h, const float* pos_xyz, const float* vel_xyz, uint32_t n);
/* same, from device pointers (inputs already resident in HBM) */
int vfd_dfsph_set_particles_device(VfdDfsph* h, const float* d_pos_xyz, const float* d_vel_xyz, uint32_t n);
int vfd_dfsph_set_rigid_bodies(VfdDfsph* h, uint32_t count, const VfdVolumeMap* maps);

/* ---- run: Simulate :33-61 (bakes FrameCount frames from the stored initial state), OnUpdate :63-170 ---- */
int vfd_dfsph_simulate(VfdDfsph* h);
int vfd_dfsph_begin(VfdDfsph* h);                 /* what Simulate() does before its loop (:36-51): reset to the initial state */
int vfd_dfsph_step(VfdDfsph* h);                  /* one OnUpdate(); asynchronous w.r.t. the host where the solver settings allow */
int vfd_dfsph_steps(VfdDfsph* h, uint32_t count); /* count x OnUpdate() */
int vfd_dfsph_synchronize(VfdDfsph* h);

/* ---- getters: GetSimulationState :273, GetDebugInfo :390, GetMaxVelocityMagnitude :298,
 *      GetCurrentTimeStepSize :303, GetParticleCount :288, GetParticleRadius :293, GetRigidBodyCount :385 ---- */
int      vfd_dfsph_get_state(const VfdDfsph* h);
int      vfd_dfsph_get_debug_info(VfdDfsph* h, VfdDfsphDebugInfo* out);
float    vfd_dfsph_get_max_velocity_magnitude(VfdDfsph* h);
float    vfd_dfsph_get_current_time_step_size(VfdDfsph* h);
uint32_t vfd_dfsph_get_particle_count(const VfdDfsph* h);
float    vfd_dfsph_get_particle_radius(const VfdDfsph* h);
uint32_t vfd_dfsph_get_rigid_body_count(const VfdDfsph* h);

/* ---- baked frames: GetParticleFrameBuffer :278 -> DFSPHParticleFrame (ParticleBuffer/DFSPHParticleBuffer.h:13-19).
 * out: n x 36 B in ORIGINAL particle order (the renderer's flow lines index particles across
 * frames: VFD/Source/Scene/Scene.cpp:361-368). */
int vfd_dfsph_get_frame_count(const VfdDfsph* h, uint32_t* baked);
int vfd_dfsph_get_frame(VfdDfsph* h, uint32_t index, VfdParticleSimple* out, float* maxVelocityMagnitude, float* currentTimeStep);
/* The same frame where it lies in the handle's frame store, without a copy — what DFSPHParticleBuffer::GetFrame
 * (ParticleBuffer/DFSPHParticleBuffer.cu:54-57) hands out by reference: *data points at *count records in host memory (pinned
 * up to the store's budget, so DFSPHParticleBuffer::SetActiveFrame's upload :38-49 can read it directly); valid until the next
 * bake, the next vfd_dfsph_set_particles or vfd_dfsph_destroy. */
int vfd_dfsph_get_frame_data(VfdDfsph* h, uint32_t index, const VfdParticleSimple** data, uint32_t* count,
                             float* maxVelocityMagnitude, float* currentTimeStep);
/* the current state in frame format (what the next captured frame would hold) */
int vfd_dfsph_get_current_frame(VfdDfsph* h, VfdParticleSimple* out);

/* ---- inspector only: ParticleSearch::GetByteSize ParticleSearch.cu:11, GetBounds ParticleSearch.h:63 ---- */
int vfd_dfsph_get_search_bytes(const VfdDfsph* h, uint64_t* bytes);
int vfd_dfsph_get_bounds(VfdDfsph* h, float bmin[3], float bmax[3]);

/* ---- no reference equivalent: state dump / restore (checkpoint-resume, parity tests) ---- */
int vfd_dfsph_get_particles(VfdDfsph* h, VfdParticle* out);              /* n x 120 B, original order */
int vfd_dfsph_set_particles_full(VfdDfsph* h, const VfdParticle* in);    /* n x 120 B, original order */
int vfd_dfsph_set_time_step(VfdDfsph* h, float dt);                      /* override the running dt (restart from a dump) */
int vfd_dfsph_set_surface_tension_state(VfdDfsph* h, uint32_t sampleCount, float monteCarloFactor);
/* neighbour search only, on the current positions (ParticleSearch::FindNeighbors, ParticleSearch.h:48-61) */
int vfd_dfsph_find_neighbors(VfdDfsph* h);
/* CSR neighbour list of the last search in ORIGINAL particle ids, each list sorted ascending.
 * counts/offsets: n entries; ids: capacity entries; *total receives the list length (pass ids=NULL to size). */
int vfd_dfsph_get_neighbors(VfdDfsph* h, uint32_t* counts, uint32_t* offsets, uint32_t* ids, uint64_t capacity, uint64_t* total);
/* boundary samples of body b after the last step (RigidBody.cuh:39-40): xj n x 3, volume n; original order */
int vfd_dfsph_get_boundary(VfdDfsph* h, uint32_t body, float* xj, float* volume);
/* the 10 000-entry cubic-spline lookup tables (Kernel/DFSPHKernels.h:13-41): W[10000], gradW[10001] */
int vfd_dfsph_get_kernel_tables(VfdDfsph* h, float* W, float* gradW, float* scalars6 /* radius, radius2, invStep, WZero, K, L */);
/* the regenerated 16384-point Halton sphere table (reference data file: HaltonVec323.cuh:4-1643): 49152 floats */
int vfd_dfsph_get_halton_table(VfdDfsph* h, float* out49152);

/* the same two tables without a handle or a device (pure host arithmetic; used by CPU-side tests) */
int vfd_kernel_tables_build(float supportRadius, float* W, float* gradW, float* scalars6);
int vfd_halton_table_build(float* out49152);

/* options (no reference equivalent) */
enum {
    VFD_OPT_SEARCH_FMA = 1,       /* 1 (default): d2 = fma(dz,dz,fma(dx,dx,dy*dy)) as nvcc compiles the reference's test
                                     (ParticleSearchKernels.cu:123-126); 0: (dx*dx+dy*dy)+dz*dz as a host compiler does */
    VFD_OPT_TIMERS = 2,           /* 1: record the six phase timers with CUDA events (default 0) */
    VFD_OPT_MAX_CELLS = 3,        /* upper bound on search-grid cells (default 1<<26) */
    VFD_OPT_KERNEL_TIMERS = 4,    /* 1: bracket every kernel launch with CUDA events and accumulate per-kernel device time
                                     (synchronises at the end of each step; for bench.py's roofline, default 0) */
};
int vfd_dfsph_set_option(VfdDfsph* h, int option, int64_t value);

/* per-kernel accounting for bench.py: number of kernels launched since the last reset */
int vfd_dfsph_get_launch_count(VfdDfsph* h, uint64_t* launches, int reset);
/* search-grid statistics of the last step (inspector only, like ParticleSearch::GetByteSize, ParticleSearch.cu:11):
 * stats[0] tiles (4x4x4 cells) of the grid, [1] cells, [2] tile passes since the last begin() whose 6x6x6-cell
 * neighbourhood did not fit the shared-memory stage and took the slow global-memory path, [3] frame bytes copied D2H */
int vfd_dfsph_get_tile_stats(VfdDfsph* h, uint64_t stats[4]);
/* tuning aid (no reference equivalent): milliseconds of `reps` back-to-back launches of the initial PCG mat-vec on the
 * state of the last step; the PCG work arrays it overwrites are rebuilt by the next step */
int vfd_dfsph_time_matvec(VfdDfsph* h, uint32_t reps, float* ms);

/* ---- several GPUs: one process (rank) per GPU, the domain cut into slabs of tile columns along x ----------------
 * No reference equivalent (the reference drives device 0 only: VFD/Source/Debug/SystemInfo.cpp:34-35).
 * Call order on every rank:  create -> init_distributed -> get_grid / set_slab -> set_particles_distributed ->
 * set_rigid_bodies -> step ...  Ranks own consecutive slabs in rank order; rank r exchanges with r-1 and r+1 only.
 * NCCL (libnccl.so.2) is loaded at run time by init_distributed; single-GPU use never needs it.
 * Baked frames (simulate / step with FrameCount > 0) are whole-scene frames in original particle order — nGlobal entries of
 * VfdParticleSimple, gathered by persistent id — and live on rank 0: vfd_dfsph_get_frame there, nothing on the other ranks.
 * The original-order dumps (get_particles, get_current_frame, get_neighbors, get_boundary) are single-GPU calls; a rank's own
 * particles are read with vfd_dfsph_get_owned. */
int vfd_dist_unique_id(char out[128]);     /* on rank 0; hand the 128 bytes to every rank (ncclGetUniqueId) */
int vfd_dfsph_init_distributed(VfdDfsph* h, int rank, int nranks, const char id[128], const float domainMin[3], const float domainMax[3]);
/* the global search grid over the domain: origin, cell size, tiles (4x4x4 cells) per axis; identical on all ranks */
int vfd_dfsph_get_grid(VfdDfsph* h, float origin[3], float* cellSize, uint32_t tiles[3]);
/* this rank owns the tile columns [lo, hi) along x (lo of rank r+1 == hi of rank r) */
int vfd_dfsph_set_slab(VfdDfsph* h, uint32_t tileColumnLo, uint32_t tileColumnHi);
/* this rank's particles with their global ids (unique over all ranks), the global particle count, and the number
 * of particle slots to allocate (owned + two ghost columns + head room for migration) */
int vfd_dfsph_set_particles_distributed(VfdDfsph* h, const float* pos_xyz, const float* vel_xyz, const uint32_t* ids,
                                        uint32_t n, uint32_t nGlobal, uint32_t capacity);
/* the particles this rank owns now, in local order, with their ids (pass ids = out = NULL to get the count) */
int vfd_dfsph_get_owned(VfdDfsph* h, uint32_t capacity, uint32_t* count, uint32_t* ids, VfdParticle* out);
/* halo exchanges, all-reduces, halo bytes sent, state-exchange bytes sent since creation */
int vfd_dfsph_get_comm_stats(VfdDfsph* h, uint64_t stats[4]);
/* the slab now: set_slab gives the start; every VFD_DIST_REBALANCE steps (environment, default 4, 0 = never) a boundary moves by
 * one tile column towards the heavier of the two ranks it separates.  info[0] lo, [1] hi, [2] boundary moves of
 * this rank so far, [3] 1 when the ranks exchange through peer memory (CUDA IPC over NVLink), 0 when through NCCL only */
int vfd_dfsph_get_slab(VfdDfsph* h, uint64_t info[4]);

/* per-kernel device time accumulated while VFD_OPT_KERNEL_TIMERS is on.  *count receives the number of kernel
 * classes; names/ms/launches (each may be NULL) receive up to `capacity` entries.  msActive/launchesActive count only
 * launches that did work (an iteration kernel whose solver has converged returns immediately). */
int vfd_dfsph_get_kernel_times(VfdDfsph* h, uint32_t capacity, uint32_t* count, const char** names, double* ms, uint64_t* launches,
                               double* msActive, uint64_t* launchesActive, int reset);

/* device-side timing on the solver's own stream (torch.cuda.Event only sees torch's streams):
 * record marks slot (0..15) on the stream; elapsed_ms synchronises on `to` and returns to - from. */
int vfd_dfsph_record_event(VfdDfsph* h, uint32_t slot);
int vfd_dfsph_elapsed_ms(VfdDfsph* h, uint32_t from, uint32_t to, float* ms);

/* ---- scene preparation helper (reference: RigidBody.cu:32-72 over SDF.cu:45-139) ----
 * Builds the two-field volume map of an axis-aligned box on the GPU: field 0 = sign*(d_box - (padding - r)),
 * field 1 = 0.8 * integral over |xi|<h of gamma(phi(x+xi)) by 30^3-point Gauss-Legendre quadrature.
 * Arrays are returned in malloc'ed host memory owned by the map; release with vfd_volume_map_free. */
int  vfd_volume_map_build_box(const float bmin[3], const float bmax[3], int inverted, float padding,
                              const uint32_t resolution[3], float particleRadius, int device, VfdVolumeMap* out);
void vfd_volume_map_free(VfdVolumeMap* map);

/* ---- scene preparation for general triangle meshes ----
 * A mesh is `vertexCount` vertices (3 floats each) and `triangleCount` triangles (3 vertex ids each), closed and
 * outward-facing, under an optional column-major 4x4 transform (glm::mat4 memory order; NULL = identity) — what the editor
 * hands to RigidBody / FluidObject after loading an .obj (TriangleMesh::GetVertices / GetTriangles).
 *
 * vfd_volume_map_build_mesh: RigidBody::RigidBody (RigidBody.cu:10-73) — the two-field volume map of the body, field 0 from
 * the reference's mesh distance (MeshDistance::SignedDistance, MeshDistance.cpp:187-222; every node against every triangle on
 * the GPU instead of a sphere-tree walk per OpenMP thread), field 1 integrated from it as for a box.
 * vfd_mesh_signed_distance: that signed distance at `count` points (3 floats each) -> out[count].
 * vfd_sample_mesh_volume: FluidObject::FluidObject (FluidObject.cpp:6-26) -> ParticleSampler::SampleMeshVolume
 * (ParticleSampler.cpp:7-91): particle positions inside the mesh on the lattice of sampleMode (0 MinDensity, 1 MediumDensity,
 * 2 MaxDensity), in the reference's order; `resolution` is the resolution of the intermediate distance grid
 * (FluidObjectDescription::Resolution).  *positions is malloc'ed (3 floats per sample; NULL when there is none): release it
 * with vfd_free. */
int  vfd_volume_map_build_mesh(const float* vertices, uint32_t vertexCount, const uint32_t* triangles, uint32_t triangleCount,
                               const float* transform16, int inverted, float padding, const uint32_t resolution[3],
                               float particleRadius, int device, VfdVolumeMap* out);
int  vfd_mesh_signed_distance(const float* vertices, uint32_t vertexCount, const uint32_t* triangles, uint32_t triangleCount,
                              const float* transform16, const float* points, uint32_t count, int device, float* out);
int  vfd_sample_mesh_volume(const float* vertices, uint32_t vertexCount, const uint32_t* triangles, uint32_t triangleCount,
                            const float* transform16, float particleRadius, const uint32_t resolution[3], int inverted, int sampleMode,
                            int device, float** positions, uint32_t* count);
void vfd_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* VFD_DFSPH_H */
