// solver_impl.h — the object behind a VfdDfsph handle.
#pragma once
#include "solver.h"
#include <functional>

namespace vfd {

struct Frame { std::vector<VfdParticleSimple> data; float maxVel2; float dt; };

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
    std::vector<Frame> frames;
    std::mutex frameMutex, dbgMutex, errMutex;
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
    int run_polled_loop(uint32_t maxIt, uint32_t already, int batch, uint32_t* dFlag, const std::function<void()>& enqueueIteration, uint32_t continueValue);

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
    VfdParticleSimple* dFrame = nullptr;
    uint32_t cellEstimate = 27, cellCapacity = 0;
    bool began = false, searched = false;
    float frameTimeHost = 0.0f;
    uint64_t stepsIssued = 0;
};

} // namespace vfd
