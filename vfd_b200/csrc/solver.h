// solver.h — host-side solver object and kernel launchers of libvfd_dfsph.so.
#pragma once
#include "common.cuh"
#include "../../include/vfd_dfsph.h"
#include <string>
#include <vector>
#include <mutex>

namespace vfd {

// Device pointers of one simulation, in sorted particle order (see common.cuh).
struct Arrays {
    // persistent state: follows the particle through the per-step reorder
    float4 *pos, *vel, *dv, *nbar;          // Position, Velocity, VelocityDifference, MonteCarloSurfaceNormalSmooth
    float  *curv, *curvS, *curvD;           // MonteCarloSurfaceCurvature, ...Smooth, DeltaFinalCurvature
    uint32_t* id;                           // original index
    // reorder targets (swapped with the above after each sort)
    float4 *pos2, *vel2, *dv2, *nbar2;
    float  *curv2, *curvS2, *curvD2;
    uint32_t* id2;
    // per-step state
    float4 *posRho;                         // (x, y, z, density): the position array of everything after the density pass
    float4 *acc, *pacc, *nrm;               // Acceleration, PressureAcceleration, (MonteCarloSurfaceNormal, MonteCarloSurfaceCurvature)
    float  *res, *rho, *rhoAdv, *kappa, *kappaV, *alpha;   // PressureResiduum, Density, DensityAdvection, PressureRho2, PressureRho2V, Factor
    // implicit viscosity (matrix-free PCG)
    float4 *cgG, *cgR, *cgQ, *cgZ;
    float4 *cgXG, *cgXP;                    // (x, y, z, v.x) with v = g at start-up / the search direction p: what the mat-vec gathers (viscosity.cu)
    float2 *cgGyz, *cgPyz;                  // (v.y, v.z)
    float  *minv;                           // 9 x n, SoA: minv[k*n + p], column-major 3x3 like glm
    // boundary samples per rigid body: (x_b.xyz, V_b)
    float4* bx[VFD_MAX_BODIES];
    // neighbour search
    uint32_t *cnt;
    uint16_t *list16;                       // neighbour lists: tile-local indices, warp-blocked ELL (tile.cuh)
    float    *coef;                         // per-pair viscosity coefficient, same ELL layout (frozen during the PCG)
    float4   *bcoef[VFD_MAX_BODIES];        // per-particle boundary-friction coefficients of the 4 tangential samples
    float    *gcoef;                        // per-pair kernel-gradient factor g_ij (gradW = g x_ij), same ELL layout; written by the density pass
    float4   *bgrad[VFD_MAX_BODIES];        // per particle and body: (gradW(x_i - x_b), V_b), zero when out of range; written by the density pass
    uint32_t *key, *rank, *tmpIdx, *cellCount, *cellBegin, *tileSums;
    uint32_t *tileList;                     // [0] = number of non-empty owned tiles, [1..] their indices (search.cu: k_compact_tiles)
    // reductions
    double* partials;
    double* slotSums;                       // per-batch partial sums of the pipelined passes: slotSums[q * slotStride + slot] (common.cuh: fold_slots)
    uint32_t slotStride;
};

// Flattened volume map on the device (reference: SDFDeviceData, Utility/SDF/SDFDeviceData.cuh:552-566)
struct DevVolumeMap {
    float dmin[3], dmax[3];
    uint32_t res[3];
    float cell[3], cellInv[3];
    uint32_t fieldCount, nodeCount, cellCount, cellMapCount;
    const float* nodes; const uint32_t* cells; const uint32_t* cellMap;
};
struct BodySet { DevVolumeMap map[VFD_MAX_BODIES]; };

// Kernel classes for the launch counter and the optional per-kernel device timers (VFD_OPT_KERNEL_TIMERS).
enum KernelId {
    KID_BOUNDS = 0, KID_HIST, KID_SCAN, KID_SCATTER, KID_REORDER, KID_BUILD_LIST,
    KID_BOUNDARY, KID_DENSITY_FACTOR,
    KID_DIV_SOURCE, KID_DIV_ACCEL, KID_DIV_SOLVE, KID_DIV_FINISH,
    KID_ST_CLASSIFY, KID_ST_SMOOTH, KID_ST_APPLY,
    KID_VISC_SETUP, KID_VISC_MATVEC0, KID_VISC_MATVEC, KID_VISC_UPDATE, KID_VISC_DIRECTION, KID_VISC_STEP, KID_VISC_APPLY,
    KID_CFL, KID_VELOCITY,
    KID_PRESS_SOURCE, KID_PRESS_ACCEL, KID_PRESS_SOLVE, KID_PRESS_FINISH,
    KID_POSITION, KID_CLEAR_ACC, KID_IO, KID_COUNT
};
extern const char* const kKernelNames[KID_COUNT];

// Brackets every launch with a pair of CUDA events on the solver's stream; drained (after a stream
// synchronise) into per-class totals.  "active" launches are those that did work: an iteration kernel
// that finds its solver converged returns at once, and is told apart by its duration.
struct KernelProf {
    bool enabled = false;
    struct Rec { int kid; cudaEvent_t a, b; };
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    std::vector<Rec> pending;
    double ms[KID_COUNT] = {}, msActive[KID_COUNT] = {};
    uint64_t launches[KID_COUNT] = {}, launchesActive[KID_COUNT] = {};
    cudaEvent_t take();
    void begin(int kid, cudaStream_t s);
    void end(cudaStream_t s);
    void drain();          // caller has synchronised the stream
    void reset();
    ~KernelProf();
};

struct LaunchCfg {
    cudaStream_t stream;
    int numSMs;
    uint64_t* launchCounter;
    KernelProf* prof;
};

// One per kernel launch: counts it and, when the timers are on, brackets it with events.
struct LaunchScope {
    const LaunchCfg& L;
    bool on;
    LaunchScope(const LaunchCfg& cfg, int kid) : L(cfg), on(cfg.prof && cfg.prof->enabled) {
        *L.launchCounter += 1;
        if (on) L.prof->begin(kid, L.stream);
    }
    ~LaunchScope() { if (on) L.prof->end(L.stream); }
};

// ---- kernel launchers (one per reference kernel group; defined in the .cu files) ----
void launch_search(const LaunchCfg& L, const Params& P, Arrays& A, DevState* S, uint32_t cellCapacity, uint32_t cellEstimate);
void launch_boundary(const LaunchCfg& L, const Params& P, const Arrays& A, const BodySet& B);
void launch_density_factor(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutW, const float* lutG);
void launch_divergence_source(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_divergence_accel(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_divergence_solve(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_divergence_finish(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_pressure_source(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_pressure_accel(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_pressure_solve(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_pressure_finish(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_clear_acceleration(const LaunchCfg& L, const Params& P, const Arrays& A);
void launch_cfl(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S);
void launch_velocity(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S);
void launch_positions(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S);
void launch_st_classify(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* halton);
void launch_st_smooth(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S);
void launch_st_apply(const LaunchCfg& L, const Params& P, const Arrays& A);
void launch_viscosity_setup(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG);
void launch_viscosity_matvec(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, bool init, bool waitHalo = false);
void launch_viscosity_update(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S);
void launch_viscosity_direction(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S);
bool viscosity_step_fits(const LaunchCfg& L, const Params& P);
int launch_viscosity_step(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S);   // update + direction fused (cudaError_t as int)
void launch_viscosity_apply(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S);

// dump / restore helpers (original particle order <-> sorted SoA)
void launch_export_aos(const LaunchCfg& L, const Params& P, const Arrays& A, VfdParticle* dOut);
void launch_import_aos(const LaunchCfg& L, const Params& P, const Arrays& A, const VfdParticle* dIn);
void launch_export_frame(const LaunchCfg& L, const Params& P, const Arrays& A, VfdParticleSimple* dOut);
void launch_import_posvel(const LaunchCfg& L, const Params& P, const Arrays& A, const float* dPos, const float* dVel);
void launch_export_boundary(const LaunchCfg& L, const Params& P, const Arrays& A, uint32_t body, float* dXj, float* dVol);

// host-side construction of the kernel lookup tables and the Halton sphere table
struct KernelTables {
    std::vector<float> W, gradW;      // the reference's raw tables: 10000 and 10001 entries
    std::vector<float> Wc, Gc;        // combined midpoint tables, 10000 entries each (last one padding)
    float radius, radius2, invStep, wZero, k, l;
    void build(float radius);
};
void build_halton_table(std::vector<float>& out);   // 49152 floats

} // namespace vfd
