// api.cu — the C ABI of libvfd_dfsph.so (include/vfd_dfsph.h).  Each entry point names the reference
// method it replaces in the header; this file only adapts calls to vfd::Solver and turns every
// failure into a status code + message (no exceptions, no exit() across the boundary).
#include "solver_impl.h"
#include <cstring>
#include <new>

using vfd::Solver;
// Threading contract (include/vfd_dfsph.h): one mutating thread plus concurrent read-only getters, as the reference's editor
// uses DFSPHSimulation (worker thread bakes, UI thread polls).  `call` is held by every mutating entry point for its whole
// duration; a getter refreshes its snapshot from the device only if it can take the lock (no mutating call in flight),
// otherwise it returns the last snapshot (kept under Solver::dbgMutex) and never touches the solver's stream.
struct VfdDfsph { Solver s; std::mutex call; };
#define LOCKED(h) std::lock_guard<std::mutex> callLock_((h)->call)
static void refresh_if_idle(VfdDfsph* h) {
    std::unique_lock<std::mutex> lk(h->call, std::try_to_lock);
    if (lk.owns_lock() && h->s.state != VFD_STATE_SIMULATING) h->s.sync_debug();
}

static std::string g_createError;
static std::mutex g_createMutex;

#define GUARD(h) if (!(h)) return VFD_E_INVALID
#define TRY(expr) try { return (expr); } catch (const std::bad_alloc&) { return h->s.fail(VFD_E_INVALID, "out of host memory"); } catch (const std::exception& e) { return h->s.fail(VFD_E_INVALID, e.what()); }

extern "C" {

void vfd_dfsph_default_description(VfdDfsphDescription* d) {
    if (!d) return;
    memset(d, 0, sizeof *d);
    d->TimeStepSize = 0.001f; d->MinTimeStepSize = 0.0001f; d->MaxTimeStepSize = 0.005f;
    d->FrameLength = 0.0016f; d->FrameCount = 200u;
    d->MinPressureSolverIterations = 0u; d->MaxPressureSolverIterations = 100u; d->MaxPressureSolverError = 10.0f;
    d->EnableDivergenceSolverError = 1u; d->MinDivergenceSolverIterations = 0u; d->MaxDivergenceSolverIterations = 100u; d->MaxDivergenceSolverError = 10.0f;
    d->EnableViscositySolver = 1u; d->MinViscositySolverIterations = 0u; d->MaxViscositySolverIterations = 100u; d->MaxViscositySolverError = 0.1f;
    d->Viscosity = 10.0f; d->BoundaryViscosity = 10.0f; d->TangentialDistanceFactor = 0.5f;
    d->EnableSurfaceTensionSolver = 1u; d->SurfaceTensionSmoothPassCount = 1u; d->SurfaceTension = 1.0f; d->TemporalSmoothing = 0u;
    d->CSDFix = -1; d->CSD = 10000;
    d->ParticleRadius = 0.025f; d->Gravity[0] = 0.0f; d->Gravity[1] = -9.81f; d->Gravity[2] = 0.0f;
}

int vfd_dfsph_create(const VfdDfsphDescription* desc, int device, VfdDfsph** out) {
    if (!desc || !out) { std::lock_guard<std::mutex> g(g_createMutex); g_createError = "vfd_dfsph_create: null argument"; return VFD_E_INVALID; }
    *out = nullptr;
    VfdDfsph* h = new (std::nothrow) VfdDfsph();
    if (!h) { std::lock_guard<std::mutex> g(g_createMutex); g_createError = "out of host memory"; return VFD_E_INVALID; }
    int rc;
    try { rc = h->s.init(*desc, device); } catch (const std::exception& e) { rc = h->s.fail(VFD_E_INVALID, e.what()); }
    if (rc != VFD_OK) {
        { std::lock_guard<std::mutex> g(g_createMutex); g_createError = h->s.lastError; }
        delete h;
        return rc;
    }
    *out = h;
    return VFD_OK;
}

void vfd_dfsph_destroy(VfdDfsph* h) { delete h; }

const char* vfd_dfsph_last_error(const VfdDfsph* h) {
    // a copy taken under the error mutex: the pointer stays valid (per calling thread) while another thread fails again
    static thread_local std::string copy;
    if (!h) { std::lock_guard<std::mutex> g(g_createMutex); copy = g_createError; return copy.c_str(); }
    Solver& s = const_cast<VfdDfsph*>(h)->s;
    std::lock_guard<std::mutex> g(s.errMutex);
    copy = s.lastError;
    return copy.c_str();
}

int vfd_dfsph_set_description(VfdDfsph* h, const VfdDfsphDescription* d) { GUARD(h); LOCKED(h); if (!d) return h->s.fail(VFD_E_INVALID, "null description"); TRY(h->s.set_description(*d)); }
int vfd_dfsph_get_description(const VfdDfsph* h, VfdDfsphDescription* out) { GUARD(h); if (!out) return VFD_E_INVALID; *out = h->s.desc; return VFD_OK; }
int vfd_dfsph_get_info(VfdDfsph* h, VfdDfsphInfo* out) {
    GUARD(h); if (!out) return h->s.fail(VFD_E_INVALID, "null output");
    refresh_if_idle(h);
    std::lock_guard<std::mutex> g(h->s.dbgMutex);
    *out = h->s.info; return VFD_OK;
}

int vfd_dfsph_set_particles(VfdDfsph* h, const float* pos, const float* vel, uint32_t n) { GUARD(h); LOCKED(h); TRY(h->s.set_particles(pos, vel, n, false)); }
int vfd_dfsph_set_particles_device(VfdDfsph* h, const float* pos, const float* vel, uint32_t n) { GUARD(h); LOCKED(h); TRY(h->s.set_particles(pos, vel, n, true)); }
int vfd_dfsph_set_rigid_bodies(VfdDfsph* h, uint32_t count, const VfdVolumeMap* maps) { GUARD(h); LOCKED(h); TRY(h->s.set_rigid_bodies(count, maps)); }

int vfd_dfsph_simulate(VfdDfsph* h) { GUARD(h); LOCKED(h); TRY(h->s.simulate()); }
int vfd_dfsph_begin(VfdDfsph* h) { GUARD(h); LOCKED(h); TRY(h->s.begin()); }
int vfd_dfsph_step(VfdDfsph* h) { GUARD(h); LOCKED(h); TRY(h->s.step()); }
int vfd_dfsph_steps(VfdDfsph* h, uint32_t count) {
    GUARD(h); LOCKED(h);
    try { for (uint32_t i = 0; i < count; i++) { int rc = h->s.step(); if (rc) return rc; } return VFD_OK; }
    catch (const std::exception& e) { return h->s.fail(VFD_E_INVALID, e.what()); }
}
int vfd_dfsph_synchronize(VfdDfsph* h) { GUARD(h); LOCKED(h); TRY(h->s.synchronize()); }

int vfd_dfsph_get_state(const VfdDfsph* h) { return h ? h->s.state : VFD_STATE_NONE; }
int vfd_dfsph_get_debug_info(VfdDfsph* h, VfdDfsphDebugInfo* out) {
    GUARD(h); if (!out) return h->s.fail(VFD_E_INVALID, "null output");
    refresh_if_idle(h);                            // while a mutating call runs: the last snapshot, as the reference's UI thread sees it
    std::lock_guard<std::mutex> g(h->s.dbgMutex);
    *out = h->s.debug; return VFD_OK;
}
float vfd_dfsph_get_max_velocity_magnitude(VfdDfsph* h) { if (!h) return 0.0f; refresh_if_idle(h); std::lock_guard<std::mutex> g(h->s.dbgMutex); return h->s.maxVel2; }
float vfd_dfsph_get_current_time_step_size(VfdDfsph* h) { if (!h) return 0.0f; refresh_if_idle(h); std::lock_guard<std::mutex> g(h->s.dbgMutex); return h->s.info.TimeStepSize; }
uint32_t vfd_dfsph_get_particle_count(const VfdDfsph* h) { return h ? h->s.info.ParticleCount : 0u; }
float vfd_dfsph_get_particle_radius(const VfdDfsph* h) { return h ? h->s.info.ParticleRadius : 0.0f; }
uint32_t vfd_dfsph_get_rigid_body_count(const VfdDfsph* h) { return h ? h->s.info.RigidBodyCount : 0u; }

int vfd_dfsph_get_frame_count(const VfdDfsph* h, uint32_t* baked) { GUARD(h); if (!baked) return VFD_E_INVALID; *baked = (uint32_t)const_cast<VfdDfsph*>(h)->s.pipe.published(); return VFD_OK; }
int vfd_dfsph_get_frame(VfdDfsph* h, uint32_t index, VfdParticleSimple* out, float* maxVel, float* dt) {
    GUARD(h);
    if (h->s.pipe.read(index, out, maxVel, dt)) return VFD_OK;
    // captured but still in flight (only possible outside Simulate(), which drains before it returns)
    if (index < h->s.frameIndexHost && h->s.state != VFD_STATE_SIMULATING) {
        if (h->s.pipe.drain() != cudaSuccess) return h->s.fail(VFD_E_CUDA, "asynchronous frame copy failed");
        if (h->s.pipe.read(index, out, maxVel, dt)) return VFD_OK;
    }
    return h->s.fail(VFD_E_INVALID, "frame index out of range");
}
int vfd_dfsph_get_frame_data(VfdDfsph* h, uint32_t index, const VfdParticleSimple** data, uint32_t* count, float* maxVel, float* dt) {
    GUARD(h);
    if (!data) return VFD_E_INVALID;
    if (h->s.pipe.view(index, data, count, maxVel, dt)) return VFD_OK;
    if (index < h->s.frameIndexHost && h->s.state != VFD_STATE_SIMULATING) {
        if (h->s.pipe.drain() != cudaSuccess) return h->s.fail(VFD_E_CUDA, "asynchronous frame copy failed");
        if (h->s.pipe.view(index, data, count, maxVel, dt)) return VFD_OK;
    }
    return h->s.fail(VFD_E_INVALID, "frame index out of range");
}
int vfd_dfsph_get_current_frame(VfdDfsph* h, VfdParticleSimple* out) { GUARD(h); LOCKED(h); TRY(h->s.get_current_frame(out)); }

int vfd_dfsph_get_search_bytes(const VfdDfsph* h, uint64_t* bytes) { GUARD(h); if (!bytes) return VFD_E_INVALID; *bytes = h->s.searchBytes; return VFD_OK; }
int vfd_dfsph_get_bounds(VfdDfsph* h, float bmin[3], float bmax[3]) { GUARD(h); LOCKED(h); TRY(h->s.get_bounds(bmin, bmax)); }

int vfd_dfsph_get_particles(VfdDfsph* h, VfdParticle* out) { GUARD(h); LOCKED(h); TRY(h->s.get_particles(out)); }
int vfd_dfsph_set_particles_full(VfdDfsph* h, const VfdParticle* in) { GUARD(h); LOCKED(h); TRY(h->s.set_particles_full(in)); }
int vfd_dfsph_set_time_step(VfdDfsph* h, float dt) { GUARD(h); LOCKED(h); if (!(dt > 0.0f)) return h->s.fail(VFD_E_INVALID, "time step must be positive"); TRY(h->s.set_time_step(dt)); }
int vfd_dfsph_set_surface_tension_state(VfdDfsph* h, uint32_t sc, float mc) { GUARD(h); LOCKED(h); TRY(h->s.set_st_state(sc, mc)); }
int vfd_dfsph_find_neighbors(VfdDfsph* h) { GUARD(h); LOCKED(h); TRY(h->s.search_only()); }
int vfd_dfsph_get_neighbors(VfdDfsph* h, uint32_t* counts, uint32_t* offsets, uint32_t* ids, uint64_t capacity, uint64_t* total) {
    GUARD(h); LOCKED(h); TRY(h->s.get_neighbors(counts, offsets, ids, capacity, total));
}
int vfd_dfsph_get_boundary(VfdDfsph* h, uint32_t body, float* xj, float* vol) { GUARD(h); LOCKED(h); if (!xj || !vol) return h->s.fail(VFD_E_INVALID, "null output"); TRY(h->s.get_boundary(body, xj, vol)); }
int vfd_dfsph_get_kernel_tables(VfdDfsph* h, float* W, float* gradW, float* sc) {
    GUARD(h);
    const vfd::KernelTables& t = h->s.tables;
    if (W) memcpy(W, t.W.data(), t.W.size() * 4);
    if (gradW) memcpy(gradW, t.gradW.data(), t.gradW.size() * 4);
    if (sc) { sc[0] = t.radius; sc[1] = t.radius2; sc[2] = t.invStep; sc[3] = t.wZero; sc[4] = t.k; sc[5] = t.l; }
    return VFD_OK;
}
int vfd_dfsph_get_halton_table(VfdDfsph* h, float* out) { GUARD(h); if (!out) return VFD_E_INVALID; memcpy(out, h->s.halton.data(), h->s.halton.size() * 4); return VFD_OK; }

int vfd_kernel_tables_build(float radius, float* W, float* gradW, float* sc) {
    if (!(radius > 0.0f)) return VFD_E_INVALID;
    vfd::KernelTables t;
    t.build(radius);
    if (W) memcpy(W, t.W.data(), t.W.size() * 4);
    if (gradW) memcpy(gradW, t.gradW.data(), t.gradW.size() * 4);
    if (sc) { sc[0] = t.radius; sc[1] = t.radius2; sc[2] = t.invStep; sc[3] = t.wZero; sc[4] = t.k; sc[5] = t.l; }
    return VFD_OK;
}
int vfd_halton_table_build(float* out) {
    if (!out) return VFD_E_INVALID;
    std::vector<float> t;
    vfd::build_halton_table(t);
    memcpy(out, t.data(), t.size() * 4);
    return VFD_OK;
}

int vfd_dfsph_set_option(VfdDfsph* h, int option, int64_t value) {
    GUARD(h); LOCKED(h);
    switch (option) {
    case VFD_OPT_SEARCH_FMA: h->s.optSearchFma = value ? 1 : 0; return VFD_OK;
    case VFD_OPT_TIMERS: h->s.optTimers = value ? 1 : 0; return VFD_OK;
    case VFD_OPT_KERNEL_TIMERS: h->s.prof.enabled = value != 0; return VFD_OK;
    case VFD_OPT_MAX_CELLS: if (value < (1 << 16) || value > (1ll << 30)) return h->s.fail(VFD_E_INVALID, "max cells out of range"); h->s.optMaxCells = (uint64_t)value; return VFD_OK;
    default: return h->s.fail(VFD_E_INVALID, "unknown option");
    }
}
int vfd_dfsph_get_launch_count(VfdDfsph* h, uint64_t* launches, int reset) { GUARD(h); if (launches) *launches = h->s.launches; if (reset) h->s.launches = 0; return VFD_OK; }

int vfd_dfsph_get_tile_stats(VfdDfsph* h, uint64_t stats[4]) {
    GUARD(h); LOCKED(h); if (!stats) return h->s.fail(VFD_E_INVALID, "null output");
    return h->s.tile_stats(stats);
}

int vfd_dfsph_time_matvec(VfdDfsph* h, uint32_t reps, float* ms) {
    GUARD(h); LOCKED(h); if (!ms) return h->s.fail(VFD_E_INVALID, "null output");
    return h->s.time_matvec(reps, ms);
}

int vfd_dist_unique_id(char out[128]) {
    if (!out) return VFD_E_INVALID;
    std::string err;
    int rc = vfd::dist_unique_id(out, err);
    if (rc) { std::lock_guard<std::mutex> g(g_createMutex); g_createError = err; }
    return rc;
}
int vfd_dfsph_init_distributed(VfdDfsph* h, int rank, int nranks, const char id[128], const float dmin[3], const float dmax[3]) {
    GUARD(h); LOCKED(h); if (!id || !dmin || !dmax) return h->s.fail(VFD_E_INVALID, "null argument"); TRY(h->s.dist_init(rank, nranks, id, dmin, dmax));
}
int vfd_dfsph_get_grid(VfdDfsph* h, float origin[3], float* cellSize, uint32_t tiles[3]) { GUARD(h); if (!origin || !cellSize || !tiles) return h->s.fail(VFD_E_INVALID, "null argument"); TRY(h->s.dist_get_grid(origin, cellSize, tiles)); }
int vfd_dfsph_set_slab(VfdDfsph* h, uint32_t lo, uint32_t hi) { GUARD(h); LOCKED(h); TRY(h->s.dist_set_slab(lo, hi)); }
int vfd_dfsph_set_particles_distributed(VfdDfsph* h, const float* pos, const float* vel, const uint32_t* ids, uint32_t n, uint32_t nGlobal, uint32_t capacity) {
    GUARD(h); LOCKED(h); TRY(h->s.dist_set_particles(pos, vel, ids, n, nGlobal, capacity));
}
int vfd_dfsph_get_owned(VfdDfsph* h, uint32_t capacity, uint32_t* count, uint32_t* ids, VfdParticle* out) { GUARD(h); LOCKED(h); TRY(h->s.dist_get_owned(capacity, count, ids, out)); }
int vfd_dfsph_get_comm_stats(VfdDfsph* h, uint64_t stats[4]) {
    GUARD(h); if (!stats) return VFD_E_INVALID;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (h->s.dist) { stats[0] = h->s.dist->halos; stats[1] = h->s.dist->reductions; stats[2] = h->s.dist->bytesHalo; stats[3] = h->s.dist->bytesState; }
    return VFD_OK;
}

int vfd_dfsph_get_slab(VfdDfsph* h, uint64_t info[4]) {
    GUARD(h); if (!info) return VFD_E_INVALID;
    if (!h->s.dist) return h->s.fail(VFD_E_INVALID, "not distributed");
    info[0] = h->s.dist->colLo; info[1] = h->s.dist->colHi; info[2] = h->s.dist->shifts; info[3] = h->s.dist->p2p ? 1u : 0u;
    return VFD_OK;
}

int vfd_dfsph_get_kernel_times(VfdDfsph* h, uint32_t capacity, uint32_t* count, const char** names, double* ms, uint64_t* launches,
                               double* msActive, uint64_t* launchesActive, int reset) {
    GUARD(h);
    vfd::KernelProf& p = h->s.prof;
    if (count) *count = vfd::KID_COUNT;
    for (uint32_t k = 0; k < capacity && k < (uint32_t)vfd::KID_COUNT; k++) {
        if (names) names[k] = vfd::kKernelNames[k];
        if (ms) ms[k] = p.ms[k];
        if (launches) launches[k] = p.launches[k];
        if (msActive) msActive[k] = p.msActive[k];
        if (launchesActive) launchesActive[k] = p.launchesActive[k];
    }
    if (reset) p.reset();
    return VFD_OK;
}
int vfd_dfsph_record_event(VfdDfsph* h, uint32_t slot) { GUARD(h); LOCKED(h); TRY(h->s.record_event(slot)); }
int vfd_dfsph_elapsed_ms(VfdDfsph* h, uint32_t from, uint32_t to, float* ms) { GUARD(h); LOCKED(h); if (!ms) return h->s.fail(VFD_E_INVALID, "null output"); TRY(h->s.elapsed_ms(from, to, ms)); }

} // extern "C"
