// frame_pipe.h — asynchronous capture of baked frames.
//
// Replaces the frame capture at the end of DFSPHImplementation::OnUpdate (reference:
// DFSPHImplementation.cu:148-167: a new std::vector per frame, filled by a synchronous cudaMemcpy on the
// default stream, so the solver idles for every frame) and the host cache of DFSPHParticleBuffer
// (ParticleBuffer/DFSPHParticleBuffer.cu:26-33, format DFSPHParticleSimple, 36 B/particle, original
// particle order).
//
// Here a captured frame flows through a ring of SLOTS device buffers into the frame's own host storage:
//   solver stream : k_export_frame -> dBuf[slot] (+ the frame's (MaxVelocityMagnitude, dt) -> dMeta[slot])   (record `exported`)
//   copy stream   : wait `exported`; D2H dBuf[slot] -> the frame's storage                                    (record `copied`)
//   host worker   : wait `copied`; publish the frame; free the slot
// so the solver stream never waits for PCIe or for the host: the next step's kernels run while the previous frame
// drains.  Frame storage is PINNED host memory taken from a pool that survives clear() — re-baking (the editor's
// "Bake" after a parameter change) allocates nothing and the D2H copy lands where the frame will live, with no
// host-side copy.  Beyond a pinned budget (VFD_FRAME_PINNED_MB, default 8192) frames fall back to pageable storage
// filled by the worker from a pinned staging buffer.  Back-pressure: acquire() blocks only when all SLOTS are in flight.
#pragma once
#include "../../include/vfd_dfsph.h"
#include <cuda_runtime.h>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace vfd {

struct HostBuf {
    VfdParticleSimple* p = nullptr;
    bool pinned = false;
};
struct Frame {
    HostBuf data;
    size_t count = 0;
    float maxVel2 = 0.0f, dt = 0.0f;
};

class FramePipe {
public:
    static constexpr int SLOTS = 3;
    ~FramePipe();
    // (re)size the ring for n particles on `device`; drains first
    cudaError_t configure(int device, uint32_t n);
    // device buffer the next frame must be exported into (blocks while every slot is in flight)
    VfdParticleSimple* acquire();
    // device slot for the frame's scalars (float[4]: MaxVelocityMagnitude, dt, "this step captured a frame" as 1.0 / 0.0, spare)
    // of the buffer returned by acquire()
    float* meta_slot();
    // the export kernel has been enqueued on `solverStream` into the buffer returned by acquire(); fromDevice: the
    // frame's scalars are read from meta_slot() (written by the export kernel), else the host values are used.
    // conditional: whether the step captured a frame at all was decided on the device (FrameTime against FrameLength,
    // DFSPHImplementation.cu:148): the worker publishes the frame only if meta_slot()[2] says so, else the storage goes
    // back to the pool — the host never waits for the decision.
    cudaError_t submit(cudaStream_t solverStream, float maxVel2, float dt, bool fromDevice = false, bool conditional = false);
    // frames published so far and submitted jobs not yet decided / published
    void progress(size_t& publishedFrames, size_t& pendingJobs);
    // blocks until fewer than `below` jobs are pending
    void wait_pending_below(size_t below);
    // wait until every submitted frame is published; returns the first asynchronous error, if any
    cudaError_t drain();
    void clear();                       // drain + drop all published frames
    size_t published();
    // copy a published frame out (false: index not published)
    bool read(uint32_t index, VfdParticleSimple* out, float* maxVel2, float* dt);
    // a published frame where it lies in the store (no copy): valid until the store is cleared (next bake, new particle count, destroy)
    bool view(uint32_t index, const VfdParticleSimple** data, uint32_t* count, float* maxVel2, float* dt);
    uint64_t bytesCopied = 0;

private:
    struct Job { int slot; HostBuf dst; bool metaFromDevice, conditional; float maxVel2, dt; };
    HostBuf take_buffer();
    void release_pool();
    void worker_main();
    void stop_worker();
    int device = -1;
    uint32_t n = 0;
    VfdParticleSimple* dBuf[SLOTS] = {};
    VfdParticleSimple* hBuf[SLOTS] = {};          // pinned staging, only for frames in pageable storage
    float* dMeta = nullptr;                         // SLOTS x float[4]
    float* hMeta = nullptr;                         // pinned mirror
    std::vector<HostBuf> pool;                      // free frame buffers (n particles each)
    size_t pinnedBytes = 0, pinnedBudget = 0;
    cudaEvent_t exported[SLOTS] = {}, copied[SLOTS] = {};
    cudaStream_t copyStream = nullptr;
    bool busy[SLOTS] = {};
    int next = 0, acquired = -1;
    std::deque<Job> jobs;
    size_t inFlight = 0;
    bool quit = false;
    cudaError_t asyncError = cudaSuccess;
    std::thread worker;
    std::mutex m;                       // guards busy, jobs, inFlight, quit, frames
    std::condition_variable cvJob, cvDone;
    std::vector<Frame> frames;
};

} // namespace vfd
