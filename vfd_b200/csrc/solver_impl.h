// solver_impl.h — the object behind a VfdDfsph handle.
#pragma once
#include "solver.h"
#include "frame_pipe.h"
#include <functional>
#include <memory>

#include <nccl.h>     // types only: the entry points are resolved at run time (no link-time dependency)

namespace vfd {

// NCCL entry points resolved with dlopen/dlsym (distributed.cu)
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    bool load(std::string& err);
};

// state of the slab decomposition (distributed.cu)
struct Dist {
    int rank = 0, nranks = 1;
    NcclApi api;
    ncclComm_t comm = nullptr;
    float origin[3] = {};                 // global grid
    uint32_t gdim[3] = {}, gtiles[3] = {};
    uint32_t colLo = 0, colHi = 0;        // owned tile columns (global)
    uint32_t capacity = 0, haloCap = 0;   // particle slots, records per exchange buffer
    uint32_t nGlobal = 0, nLocal = 0;
    uint32_t ownB = 0, edgeLEnd = 0, edgeRBegin = 0, ownE = 0;   // [ghost L | first col | ... | last col | ghost R] boundaries
    void *sendL = nullptr, *sendR = nullptr, *recvL = nullptr, *recvR = nullptr;
    uint32_t* dCounters = nullptr; uint32_t* hCounters = nullptr;
    uint64_t halos = 0, reductions = 0, bytesHalo = 0, bytesState = 0;
    // peer memory (CUDA IPC over NVLink): one slab per rank = its control block + every array a halo exchange touches, mapped
    // into every other rank of the box; halos are written straight into the neighbour's ghost ranges, solver decisions are
    // all-reduced through the control blocks (control.cuh: PeerCtl) — no NCCL call per solver iteration
    bool p2p = false;
    bool fusedHalo = false;                   // VFD_DIST_FUSED_HALO=1: the PCG direction's halo is written by the fused vector kernel (see solver.cu);
                                              // off by default: validated on 2 GPUs only
    bool fusedNow = false;                    // ... and it carries this step (the same on every rank)
    uint32_t maxOwned = 0;                    // largest owned-particle count over all ranks this step (only maintained with fusedHalo)
    unsigned char* slab = nullptr; size_t slabBytes = 0;
    uint64_t slabNp[8] = {};                  // particle slots of every rank's arrays (region offsets follow from it)
    unsigned char* peerSlab[8] = {};          // every rank's slab in this process' address space (own: slab)
    uint32_t leftOwnE = 0;                    // where the left neighbour's ghost-R range begins (its ownE), renewed every step
    // slab re-balancing: every `rebalanceEvery` steps a boundary moves by one tile column towards the heavier neighbour; both
    // sides of a boundary take the decision from the same six numbers, the column changes hands in the next state exchange
    int rebalanceEvery = 4;
    int pendL = 0, pendR = 0;                 // pending move of my left / right boundary (-1: one column to the left, +1: to the right)
    uint64_t stepsDone = 0, shifts = 0;
    // baked frames of a decomposed run: owned particles as (id, position, velocity, acceleration) records, gathered on rank 0
    void *frameSend = nullptr, *frameRecv = nullptr;
    uint32_t frameRecvCap = 0;
    ~Dist();
};

int dist_unique_id(char* out128, std::string& err);

class Solver {
public:
    ~Solver();
    int init(const VfdDfsphDescription& d, int device);
    int set_description(const VfdDfsphDescription& d);
    int set_particles(const float* pos, const float* vel, uint32_t n, bool onDevice);
    int set_rigid_bodies(uint32_t count, const VfdVolumeMap* maps);
    int begin();
    int step();
    int simulate();
    int synchronize();
    int sync_debug();
    int search_only();
    int get_particles(VfdParticle* out);
    int set_particles_full(const VfdParticle* in);
    int set_time_step(float dt);
    int set_st_state(uint32_t sampleCount, float mcFactor);
    int get_current_frame(VfdParticleSimple* out);
    int get_neighbors(uint32_t* counts, uint32_t* offsets, uint32_t* ids, uint64_t capacity, uint64_t* total);
    int get_boundary(uint32_t body, float* xj, float* vol);
    int get_bounds(float* bmin, float* bmax);
    int tile_stats(uint64_t* stats4);
    int time_matvec(uint32_t reps, float* ms);
    // slab decomposition over several ranks (distributed.cu)
    int dist_init(int rank, int nranks, const char* id128, const float* dmin, const float* dmax);
    int dist_get_grid(float* origin3, float* cellSize, uint32_t* tiles3);
    int dist_set_slab(uint32_t colLo, uint32_t colHi);
    int dist_set_particles(const float* pos, const float* vel, const uint32_t* ids, uint32_t n, uint32_t nGlobal, uint32_t capacity);
    int dist_get_owned(uint32_t capacity, uint32_t* count, uint32_t* ids, VfdParticle* out);
    std::unique_ptr<Dist> dist;
    int record_event(uint32_t slot);
    int elapsed_ms(uint32_t from, uint32_t to, float* ms);
    int fail(int code, const std::string& msg);

    // configuration / host mirrors
    VfdDfsphDescription desc;
    VfdDfsphInfo info;
    Params params;
    KernelTables tables;
    std::vector<float> halton;
    int optSearchFma = 1, optTimers = 0;
    KernelProf prof;
    uint64_t optMaxCells = 1ull << 26;
    uint64_t launches = 0;
    uint64_t allocBytes = 0, searchBytes = 0;
    int state = VFD_STATE_NONE;
    VfdDfsphDebugInfo debug{};
    float maxVel2 = 0.0f;
    FramePipe pipe;                   // baked frames: async D2H ring + host cache (frame_pipe.h)
    std::mutex dbgMutex, errMutex;
    std::string lastError;
    uint32_t frameIndexHost = 0;
    int device = -1;

private:
    int fail_cuda(cudaError_t e, const char* what, int line);
    void refresh_params();
    void free_particles();
    void free_bodies();
    int alloc_particles(uint32_t n, const float* bboxMin, const float* bboxMax);
    int read_state(DevState& out);
    void update_debug(const DevState& s, bool timers);
    int capture_frame(const DevState& s);
    void dist_apply_grid(DevState& s);
    void dist_params(Params& P) const;
    int dist_exchange_state();
    int dist_read_ranges();
    int dist_halo(void* base, uint32_t elemFloats, void* base2 = nullptr, uint32_t elemFloats2 = 0);
    int dist_reduce(int site, bool isMax);
    int dist_frame_step();
    int dist_capture_frame(bool metaFromDevice, float vmax2, float dt);
    int dist_alloc_slab(size_t np);
    void dist_free_slab();
    int dist_upload_grid();
    void l2_window(void* base, size_t bytes);
    void dist_halo_targets(Params& P) const;
    void dist_count_fused_halo(uint64_t bytesPerItem);
    int halo4(float4* a) { return dist ? dist_halo(a, 4) : VFD_OK; }
    int halo42(float4* a, float2* b) { return dist ? dist_halo(a, 4, b, 2) : VFD_OK; }
    int halo1(float* a) { return dist ? dist_halo(a, 1) : VFD_OK; }
    int reduce(int site, bool isMax = false) { return dist ? dist_reduce(site, isMax) : VFD_OK; }
    int run_polled_loop(uint32_t maxIt, uint32_t already, int batch, uint32_t* dFlag, const std::function<int()>& enqueueIteration, uint32_t continueValue);

    int numSMs = 148;
    cudaStream_t stream = nullptr;
    DevState* dState = nullptr;
    DevState* hState = nullptr;       // pinned
    uint32_t* hFlags = nullptr;       // pinned
    cudaEvent_t pollEvent[4] = {};
    cudaEvent_t phaseEvent[7] = {};
    cudaEvent_t userEvent[16] = {};
    float *dLutW = nullptr, *dLutG = nullptr, *dHalton = nullptr;
    Arrays arrays{};
    BodySet bodies{};
    std::vector<void*> bodyAllocs;
    float4 *dPos0 = nullptr, *dVel0 = nullptr;
    uint32_t* dIds0 = nullptr;
    VfdParticleSimple* dFrame = nullptr;
    uint32_t dFrameCapacity = 0;
    uint32_t cellEstimate = 27, cellCapacity = 0;
    size_t allocParticles = 0;        // particle slots the arrays are currently sized for
    bool began = false, searched = false;
    void* pcgArena = nullptr; size_t pcgArenaBytes = 0;      // one GPU: the PCG's vectors in one allocation (L2 residency window)
    size_t l2Carve = 0, l2Window = 0;
    std::vector<size_t> bodyAllocBytes;      // sizes of bodyAllocs (three per body): re-baking the same bodies keeps them
    size_t bodySampleSlots = 0;              // particle slots the per-body sample arrays are sized for
    float *dStagePos = nullptr, *dStageVel = nullptr; uint32_t stageSlots = 0;   // set_particles' upload staging
    float frameTimeHost = 0.0f;
    bool asyncFrames = false;                 // Simulate() with a frame length on one GPU: the device decides which steps are frames (solver.cu: step)
    uint64_t stepsIssued = 0;
};

} // namespace vfd
