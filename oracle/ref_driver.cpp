// oracle/ref_driver.cpp — TEST INFRASTRUCTURE.  Not part of the product; nothing under
// vfd_b200/ may link or call this.
//
// A plain-C façade over the *reference's own* DFSPH solver classes, compiled from the
// sources where they lie under /root/reference (recipe: oracle/build_ref.py) either
//   * for host cores through the CUDA-on-CPU emulation in oracle/shim (libvfd_ref_cpu.so), or
//   * with nvcc for sm_100a (libvfd_ref_gpu.so) — the only pre-existing GPU implementation.
// It drives  DFSPHImplementation  (VFD/Source/Simulation/DFSPH/DFSPHImplementation.h:23-137)
// exactly as the editor does (VFD/Source/Editor/Panels/ComponentPanel.cpp:608-666):
// SetDescription → SetFluidObjects → RigidBody(...) → SetRigidBodies → Simulate()/OnUpdate().
//
// The reference keeps everything private and offers no dump hooks, so this TU (and only
// this TU) is compiled with private/protected opened up.  Class layout is unaffected.
#include "pch.h"
#include <thrust/device_vector.h>
#include <thrust/extrema.h>
#include <thrust/transform_reduce.h>
#include <thrust/iterator/constant_iterator.h>
#include <thrust/functional.h>
#include <thrust/sequence.h>
#include <thrust/gather.h>
#include <omp.h>

#define private public
#define protected public
#include "Simulation/DFSPH/DFSPHImplementation.h"
#include "Utility/SDF/SDF.cuh"
#include "Utility/Sampler/ParticleSampler.h"
#include "Utility/SDF/MeshDistance.h"
#undef private
#undef protected

#include <new>

#ifndef VFD_REF_GPU
extern int g_emu_serial;
#endif

using namespace vfd;

struct RefSim {
    DFSPHImplementation* impl = nullptr;
    void* mem = nullptr;
    std::vector<Ref<FluidObject>> fluids;
    std::vector<Ref<RigidBody>> bodies;
    std::vector<glm::vec3> initVel;   // per-particle initial velocities (reference has one per object)
    bool begun = false;
};

// Mirror of include/vfd_dfsph.h VfdDfsphDescription (field-for-field DFSPHSimulationDescription,
// VFD/Source/Simulation/DFSPH/Structures/DFSPHSimulationDescription.h:9-50) with fixed-width types.
struct CDesc {
    float TimeStepSize, MinTimeStepSize, MaxTimeStepSize, FrameLength;
    uint32_t FrameCount;
    uint32_t MinPressureSolverIterations, MaxPressureSolverIterations; float MaxPressureSolverError;
    uint32_t EnableDivergenceSolverError, MinDivergenceSolverIterations, MaxDivergenceSolverIterations; float MaxDivergenceSolverError;
    uint32_t EnableViscositySolver, MinViscositySolverIterations, MaxViscositySolverIterations; float MaxViscositySolverError;
    float Viscosity, BoundaryViscosity, TangentialDistanceFactor;
    uint32_t EnableSurfaceTensionSolver, SurfaceTensionSmoothPassCount; float SurfaceTension; uint32_t TemporalSmoothing;
    int32_t CSDFix, CSD;
    float ParticleRadius; float Gravity[3];
};

static DFSPHSimulationDescription ToDesc(const CDesc* c) {
    DFSPHSimulationDescription d;
    d.TimeStepSize = c->TimeStepSize; d.MinTimeStepSize = c->MinTimeStepSize; d.MaxTimeStepSize = c->MaxTimeStepSize;
    d.FrameLength = c->FrameLength; d.FrameCount = c->FrameCount;
    d.MinPressureSolverIterations = c->MinPressureSolverIterations; d.MaxPressureSolverIterations = c->MaxPressureSolverIterations;
    d.MaxPressureSolverError = c->MaxPressureSolverError;
    d.EnableDivergenceSolverError = c->EnableDivergenceSolverError != 0;
    d.MinDivergenceSolverIterations = c->MinDivergenceSolverIterations; d.MaxDivergenceSolverIterations = c->MaxDivergenceSolverIterations;
    d.MaxDivergenceSolverError = c->MaxDivergenceSolverError;
    d.EnableViscositySolver = c->EnableViscositySolver != 0;
    d.MinViscositySolverIterations = c->MinViscositySolverIterations; d.MaxViscositySolverIterations = c->MaxViscositySolverIterations;
    d.MaxViscositySolverError = c->MaxViscositySolverError;
    d.Viscosity = c->Viscosity; d.BoundaryViscosity = c->BoundaryViscosity; d.TangentialDistanceFactor = c->TangentialDistanceFactor;
    d.EnableSurfaceTensionSolver = c->EnableSurfaceTensionSolver != 0; d.SurfaceTensionSmoothPassCount = c->SurfaceTensionSmoothPassCount;
    d.SurfaceTension = c->SurfaceTension; d.TemporalSmoothing = c->TemporalSmoothing != 0;
    d.CSDFix = c->CSDFix; d.CSD = c->CSD;
    d.ParticleRadius = c->ParticleRadius; d.Gravity = { c->Gravity[0], c->Gravity[1], c->Gravity[2] };
    return d;
}

template<typename T> static void D2H(const T* dev, T* host, size_t n) {
    COMPUTE_SAFE(cudaMemcpy((void*)host, (const void*)dev, n * sizeof(T), cudaMemcpyDeviceToHost));
}
template<typename T> static void H2D(const T* host, T* dev, size_t n) {
    COMPUTE_SAFE(cudaMemcpy((void*)dev, (const void*)host, n * sizeof(T), cudaMemcpyHostToDevice));
}

// The reference's MeshDistance prunes its search with a sphere tree whose spheres come from a RANDOMISED smallest-enclosing-sphere
// routine (Core/Structures/BoundingSphere.h:133-172: rand()-driven permutation + 1e-6 perturbation).  For some rand() states the
// spheres do not enclose their triangles and the walk misses the nearest face (cocircular vertices — a cone's rim — often, the
// corners of a box about once in a thousand builds).  Where the tree is sound the answer is the minimum
// over all faces and does not depend on the state.  Everything below that builds a MeshDistance therefore builds it until two
// builds AGREE bit for bit (at most three builds; every build advances rand()), so that one unlucky state does not decide a
// parity check.
template<typename Build, typename Same>
static auto BuildUntilTwoAgree(Build build, Same same) -> decltype(build()) {
    auto a = build();
    auto b = build();
    if (same(a, b)) return b;
    auto c = build();
    if (same(a, c)) return a;
    return c;                       // b == c, or no two agree: the last one
}

static Ref<RigidBody> MakeBody(RefSim* s, const RigidBodyDescription& rd) {
    return BuildUntilTwoAgree(
        [&]() { return Ref<RigidBody>::Create(rd, s->impl->GetInfo(), s->impl->GetKernel()); },
        [](Ref<RigidBody>& a, Ref<RigidBody>& b) { return a->m_DensityMap->m_Nodes[0] == b->m_DensityMap->m_Nodes[0]; });
}

extern "C" {

int ref_is_gpu() {
#ifdef VFD_REF_GPU
    return 1;
#else
    return 0;
#endif
}

void ref_set_serial(int s) {
#ifndef VFD_REF_GPU
    g_emu_serial = s;
#else
    (void)s;
#endif
}

void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_get_max_threads() { return omp_get_max_threads(); }

RefSim* ref_create(const CDesc* cd) {
    RefSim* s = new RefSim();
    // zero-filled storage: the reference reads m_Info.ParticleRadius / SurfaceTensionSampleCount /
    // MonteCarloFactor before it ever writes them — pin them to 0.
    s->mem = calloc(1, sizeof(DFSPHImplementation));
    s->impl = new (s->mem) DFSPHImplementation(ToDesc(cd));
    return s;
}

void ref_destroy(RefSim* s) {
    // the reference leaks by design (Q14); keep teardown minimal and safe
    if (!s) return;
    delete s;
}

void ref_set_description(RefSim* s, const CDesc* cd) { s->impl->SetDescription(ToDesc(cd)); }

void ref_set_particles(RefSim* s, const float* pos, const float* vel, uint32_t n) {
    std::vector<glm::vec3> p(n);
    s->initVel.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        p[i] = { pos[3 * i], pos[3 * i + 1], pos[3 * i + 2] };
        s->initVel[i] = vel ? glm::vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]) : glm::vec3(0.0f);
    }
    s->fluids.clear();
    s->fluids.push_back(Ref<FluidObject>::Create(p, glm::vec3(0.0f)));
    s->impl->SetFluidObjects(s->fluids);
    for (uint32_t i = 0; i < n; i++) s->impl->m_Particles[i].Velocity = s->initVel[i];
    s->begun = false;
}

// Rigid body from an axis-aligned box mesh (TriangleMesh(AABB), reference TriangleMesh.cpp:18-37),
// through the reference's own RigidBody constructor (RigidBody.cu:10-73).
void ref_add_box_body(RefSim* s, const float* bmin, const float* bmax, int inverted, float padding, const uint32_t* res) {
    RigidBodyDescription rd;
    rd.Inverted = inverted != 0;
    rd.Padding = padding;
    rd.CollisionMapResolution = { res[0], res[1], res[2] };
    rd.Transform = glm::mat4(1.0f);
    AABB box(glm::vec3(bmin[0], bmin[1], bmin[2]), bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]);
    rd.Mesh = Ref<TriangleMesh>::Create(box);
    s->bodies.push_back(MakeBody(s, rd));
}

void ref_commit_bodies(RefSim* s) { s->impl->SetRigidBodies(s->bodies); }

// The reference's own fluid sampling (FluidObject::FluidObject, FluidObject.cpp:6-26: transform the mesh, then
// ParticleSampler::SampleMeshVolume, ParticleSampler.cpp:7-91) on a raw triangle mesh.  Returns the number of samples;
// writes at most `capacity` positions.
uint32_t ref_sample_mesh_volume(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* transform16,
                                float radius, const uint32_t* res, int inverted, int mode, float* out, uint32_t capacity) {
    std::vector<glm::vec3> v(nv);
    std::vector<glm::uvec3> t(nt);
    glm::mat4 T(1.0f);
    if (transform16) memcpy(&T[0][0], transform16, 16 * sizeof(float));      // column-major like glm
    for (uint32_t i = 0; i < nv; i++) v[i] = T * glm::vec4(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], 1.0f);
    for (uint32_t i = 0; i < nt; i++) t[i] = { tris[3 * i], tris[3 * i + 1], tris[3 * i + 2] };
    const Ref<EdgeMesh> mesh = Ref<EdgeMesh>::Create(v, t);
    const std::vector<glm::vec3> p = BuildUntilTwoAgree(
        [&]() { return ParticleSampler::SampleMeshVolume(mesh, radius, glm::uvec3(res[0], res[1], res[2]), inverted != 0, (SampleMode)mode); },
        [](std::vector<glm::vec3>& a, std::vector<glm::vec3>& b) { return a == b; });
    const uint32_t n = (uint32_t)p.size();
    for (uint32_t i = 0; i < n && i < capacity; i++) { out[3 * i] = p[i].x; out[3 * i + 1] = p[i].y; out[3 * i + 2] = p[i].z; }
    return n;
}

// Rigid body from a raw triangle mesh under a transform (what the editor does with an .obj: RigidBody.cu:10-73).
void ref_add_mesh_body(RefSim* s, const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* transform16,
                       int inverted, float padding, const uint32_t* res) {
    std::vector<glm::vec3> v(nv);
    std::vector<glm::uvec3> t(nt);
    for (uint32_t i = 0; i < nv; i++) v[i] = { verts[3 * i], verts[3 * i + 1], verts[3 * i + 2] };
    for (uint32_t i = 0; i < nt; i++) t[i] = { tris[3 * i], tris[3 * i + 1], tris[3 * i + 2] };
    RigidBodyDescription rd;
    rd.Inverted = inverted != 0;
    rd.Padding = padding;
    rd.CollisionMapResolution = { res[0], res[1], res[2] };
    rd.Transform = glm::mat4(1.0f);
    if (transform16) memcpy(&rd.Transform[0][0], transform16, 16 * sizeof(float));
    rd.Mesh = Ref<TriangleMesh>::Create(v, t);
    s->bodies.push_back(MakeBody(s, rd));
}

// MeshDistance::SignedDistance (MeshDistance.cpp:187-222) at `n` points of a raw triangle mesh under a transform, one
// thread (the per-thread "closest face of the previous query" then evolves in point order).
void ref_mesh_signed_distance(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* transform16,
                              const float* points, uint32_t n, float* out) {
    std::vector<glm::vec3> v(nv);
    std::vector<glm::uvec3> t(nt);
    glm::mat4 T(1.0f);
    if (transform16) memcpy(&T[0][0], transform16, 16 * sizeof(float));
    for (uint32_t i = 0; i < nv; i++) v[i] = T * glm::vec4(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], 1.0f);
    for (uint32_t i = 0; i < nt; i++) t[i] = { tris[3 * i], tris[3 * i + 1], tris[3 * i + 2] };
    const Ref<EdgeMesh> mesh = Ref<EdgeMesh>::Create(v, t);
    const std::vector<float> d = BuildUntilTwoAgree(
        [&]() {
            MeshDistance md(mesh);
            std::vector<float> r(n);
            for (uint32_t i = 0; i < n; i++) r[i] = md.SignedDistance(glm::vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]));
            return r;
        },
        [](std::vector<float>& a, std::vector<float>& b) { return a == b; });
    memcpy(out, d.data(), n * sizeof(float));
}

// The same with ONE build of the reference's MeshDistance, whatever state rand() is in: the reference's raw behaviour
// (tests/test_mesh_prep_cpu.py shows the unsound trees with it; nothing else uses it).
void ref_mesh_signed_distance_once(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* transform16,
                                   const float* points, uint32_t n, float* out) {
    std::vector<glm::vec3> v(nv);
    std::vector<glm::uvec3> t(nt);
    glm::mat4 T(1.0f);
    if (transform16) memcpy(&T[0][0], transform16, 16 * sizeof(float));
    for (uint32_t i = 0; i < nv; i++) v[i] = T * glm::vec4(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], 1.0f);
    for (uint32_t i = 0; i < nt; i++) t[i] = { tris[3 * i], tris[3 * i + 1], tris[3 * i + 2] };
    const Ref<EdgeMesh> mesh = Ref<EdgeMesh>::Create(v, t);
    MeshDistance md(mesh);
    for (uint32_t i = 0; i < n; i++) out[i] = md.SignedDistance(glm::vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]));
}

// Volume-map extraction = exactly what SDF::GetDeviceData flattens (SDF.cu:227-306).
// sizes: [fieldCount, nodeCount, cellCount, cellMapCount, res.x, res.y, res.z]
void ref_get_map_sizes(RefSim* s, uint32_t body, uint32_t* sizes, float* domain6, float* cell3, float* cellInv3) {
    SDF* m = s->bodies[body]->m_DensityMap.Raw();
    sizes[0] = m->m_FieldCount; sizes[1] = (uint32_t)m->m_Nodes[0].size(); sizes[2] = m->m_CellCount;
    sizes[3] = (uint32_t)m->m_CellMap[0].size();
    sizes[4] = m->m_Resolution.x; sizes[5] = m->m_Resolution.y; sizes[6] = m->m_Resolution.z;
    for (int k = 0; k < 3; k++) { domain6[k] = m->m_Domain.min[k]; domain6[3 + k] = m->m_Domain.max[k];
        cell3[k] = m->m_CellSize[k]; cellInv3[k] = m->m_CellSizeInverse[k]; }
}

void ref_get_map_data(RefSim* s, uint32_t body, float* nodes, uint32_t* cells, uint32_t* cellMap) {
    SDF* m = s->bodies[body]->m_DensityMap.Raw();
    size_t o = 0;
    for (auto& f : m->m_Nodes) { memcpy(nodes + o, f.data(), f.size() * sizeof(float)); o += f.size(); }
    o = 0;
    for (auto& f : m->m_Cells) for (auto& c : f) { memcpy(cells + o, c.data(), 32 * sizeof(uint32_t)); o += 32; }
    o = 0;
    for (auto& f : m->m_CellMap) { memcpy(cellMap + o, f.data(), f.size() * sizeof(uint32_t)); o += f.size(); }
}

// What Simulate() does before its loop (DFSPHImplementation.cu:36-51).
void ref_begin(RefSim* s) {
    DFSPHImplementation* I = s->impl;
    H2D(I->m_Particles, I->d_Particles, I->m_Info.ParticleCount);
    I->m_DebugInfo.IterationCount = 0u;
    I->m_Info.TimeStepSize = I->m_Description.TimeStepSize;
    I->m_Info.TimeStepSize2 = I->m_Description.TimeStepSize * I->m_Description.TimeStepSize;
    I->m_Info.TimeStepSizeInverse = 1.0f / I->m_Info.TimeStepSize;
    I->m_Info.TimeStepSize2Inverse = 1.0f / I->m_Info.TimeStepSize2;
    H2D(&I->m_Info, I->d_Info, 1);
    I->m_State = DFSPHImplementation::SimulationState::Simulating;
    I->m_DebugInfo.FrameIndex = 0u;
    I->m_DebugInfo.FrameTime = 0.0f;
    s->begun = true;
}

void ref_step(RefSim* s) { if (!s->begun) ref_begin(s); s->impl->OnUpdate(); }

uint32_t ref_particle_count(RefSim* s) { return s->impl->m_Info.ParticleCount; }

void ref_get_particles(RefSim* s, void* out120) {
    D2H(s->impl->d_Particles, (DFSPHParticle*)out120, s->impl->m_Info.ParticleCount);
}
void ref_set_particles_full(RefSim* s, const void* in120) {
    if (!s->begun) ref_begin(s);
    H2D((const DFSPHParticle*)in120, s->impl->d_Particles, s->impl->m_Info.ParticleCount);
}
// overrides the running time step (host Info + device copy), for restart-from-state tests
void ref_set_time_step(RefSim* s, float dt) {
    DFSPHImplementation* I = s->impl;
    I->m_Info.TimeStepSize = dt; I->m_Info.TimeStepSize2 = dt * dt;
    I->m_Info.TimeStepSizeInverse = 1.0f / dt; I->m_Info.TimeStepSize2Inverse = 1.0f / (dt * dt);
    H2D(&I->m_Info, I->d_Info, 1);
}
void ref_set_st_state(RefSim* s, uint32_t sampleCount, float mcFactor) {
    DFSPHImplementation* I = s->impl;
    I->m_Info.SurfaceTensionSampleCount = sampleCount; I->m_Info.MonteCarloFactor = mcFactor;
    H2D(&I->m_Info, I->d_Info, 1);
}
void ref_get_info(RefSim* s, void* out128) { memcpy(out128, &s->impl->m_Info, sizeof(DFSPHSimulationInfo)); }
uint32_t ref_info_size() { return (uint32_t)sizeof(DFSPHSimulationInfo); }
uint32_t ref_particle_size() { return (uint32_t)sizeof(DFSPHParticle); }
uint32_t ref_kernel_size() { return (uint32_t)sizeof(PrecomputedDFSPHCubicKernel); }
void ref_get_kernel(RefSim* s, void* out) { memcpy(out, &s->impl->m_PrecomputedSmoothingKernel, sizeof(PrecomputedDFSPHCubicKernel)); }

// debug: [iterationCount, divIt, pressIt, viscIt], [divErr, pressErr, viscErr, frameTime, dt, maxVel2], 6 phase timers (us)
void ref_get_debug(RefSim* s, uint32_t* it4, float* f6, float* timers6) {
    const DFSPHDebugInfo& d = s->impl->m_DebugInfo;
    it4[0] = d.IterationCount; it4[1] = d.DivergenceSolverIterationCount; it4[2] = d.PressureSolverIterationCount; it4[3] = d.ViscositySolverIterationCount;
    f6[0] = d.DivergenceSolverError; f6[1] = d.PressureSolverError; f6[2] = d.ViscositySolverError; f6[3] = d.FrameTime;
    f6[4] = s->impl->m_Info.TimeStepSize; f6[5] = s->impl->m_MaxVelocityMagnitude;
    timers6[0] = d.NeighborhoodSearchTimer.GetElapsed<std::chrono::microseconds>();
    timers6[1] = d.BaseSolverTimer.GetElapsed<std::chrono::microseconds>();
    timers6[2] = d.DivergenceSolverTimer.GetElapsed<std::chrono::microseconds>();
    timers6[3] = d.SurfaceTensionSolverTimer.GetElapsed<std::chrono::microseconds>();
    timers6[4] = d.ViscositySolverTimer.GetElapsed<std::chrono::microseconds>();
    timers6[5] = d.PressureSolverTimer.GetElapsed<std::chrono::microseconds>();
}

// Neighbour CSR of the last search (ParticleSearch.cu:141-194). Returns total; pass null ids to size.
uint32_t ref_get_neighbors(RefSim* s, uint32_t* counts, uint32_t* offsets, uint32_t* ids, uint32_t idsCapacity) {
    ParticleSearch& ps = s->impl->m_ParticleSearch;
    const uint32_t n = s->impl->m_Info.ParticleCount;
    std::vector<unsigned int> c(n), o(n);
    D2H(thrust::raw_pointer_cast(ps.d_NeighborCounts.data()), c.data(), n);
    D2H(thrust::raw_pointer_cast(ps.d_NeighborWriteOffsets.data()), o.data(), n);
    const uint32_t total = n ? o[n - 1] + c[n - 1] : 0;
    if (counts) memcpy(counts, c.data(), n * 4);
    if (offsets) memcpy(offsets, o.data(), n * 4);
    if (ids && idsCapacity >= total && total) D2H(thrust::raw_pointer_cast(ps.d_Neighbors.data()), ids, total);
    return total;
}

// only the neighbour search on the current device state (ParticleSearch.h:48-61)
void ref_find_neighbors(RefSim* s) {
    if (!s->begun) ref_begin(s);
    s->impl->m_ParticleSearch.FindNeighbors(s->impl->d_Particles);
}

// boundary samples of body b after the last step (RigidBody.cuh:39-40): xj[3n], vol[n]
void ref_get_boundary(RefSim* s, uint32_t body, float* xj, float* vol) {
    RigidBody* rb = s->bodies[body].Raw();
    const uint32_t n = s->impl->m_Info.ParticleCount;
    D2H((const float*)thrust::raw_pointer_cast(rb->m_BoundaryXJ.data()), xj, 3 * (size_t)n);
    D2H(thrust::raw_pointer_cast(rb->m_BoundaryVolume.data()), vol, n);
}

} // extern "C"
