// frame_pipe.cu — see frame_pipe.h.
#include "frame_pipe.h"
#include <cstdlib>
#include <cstring>

namespace vfd {

FramePipe::~FramePipe() {
    stop_worker();
    if (device >= 0) cudaSetDevice(device);
    for (int s = 0; s < SLOTS; s++) {
        if (dBuf[s]) cudaFree(dBuf[s]);
        if (hBuf[s]) cudaFreeHost(hBuf[s]);
        if (exported[s]) cudaEventDestroy(exported[s]);
        if (copied[s]) cudaEventDestroy(copied[s]);
    }
    if (copyStream) cudaStreamDestroy(copyStream);
    for (Frame& f : frames) pool.push_back(f.data);
    frames.clear();
    release_pool();
    if (dMeta) cudaFree(dMeta);
    if (hMeta) cudaFreeHost(hMeta);
}

void FramePipe::release_pool() {
    for (HostBuf& b : pool) {
        if (!b.p) continue;
        if (b.pinned) cudaFreeHost(b.p); else delete[] b.p;
    }
    pool.clear();
    pinnedBytes = 0;
}

HostBuf FramePipe::take_buffer() {
    {
        std::lock_guard<std::mutex> g(m);
        if (!pool.empty()) { HostBuf b = pool.back(); pool.pop_back(); return b; }
    }
    HostBuf b;
    const size_t bytes = (size_t)(n ? n : 1u) * sizeof(VfdParticleSimple);
    if (pinnedBytes + bytes <= pinnedBudget && cudaMallocHost((void**)&b.p, bytes) == cudaSuccess) {
        b.pinned = true;
        pinnedBytes += bytes;
    } else {
        cudaGetLastError();
        b.p = new VfdParticleSimple[n ? n : 1u];
        b.pinned = false;
    }
    return b;
}

void FramePipe::stop_worker() {
    if (!worker.joinable()) return;
    {
        std::lock_guard<std::mutex> g(m);
        quit = true;
    }
    cvJob.notify_all();
    worker.join();
    quit = false;
}

cudaError_t FramePipe::configure(int dev, uint32_t count) {
    cudaError_t e = drain();
    if (e != cudaSuccess) return e;
    if (dev == device && count == n && copyStream) return cudaSuccess;
    stop_worker();
    {
        std::lock_guard<std::mutex> g(m);
        for (Frame& f : frames) pool.push_back(f.data);
        frames.clear();
    }
    if (device >= 0) cudaSetDevice(device);
    release_pool();                                    // buffers of the old particle count
    const char* mb = getenv("VFD_FRAME_PINNED_MB");
    pinnedBudget = (size_t)(mb ? atoll(mb) : 8192) << 20;
    device = dev;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return e;
    for (int s = 0; s < SLOTS; s++) {
        if (dBuf[s]) { cudaFree(dBuf[s]); dBuf[s] = nullptr; }
        if (hBuf[s]) { cudaFreeHost(hBuf[s]); hBuf[s] = nullptr; }
    }
    n = count;
    if (!copyStream && (e = cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking)) != cudaSuccess) return e;
    for (int s = 0; s < SLOTS; s++) {
        if (!exported[s] && (e = cudaEventCreateWithFlags(&exported[s], cudaEventDisableTiming)) != cudaSuccess) return e;
        if (!copied[s] && (e = cudaEventCreateWithFlags(&copied[s], cudaEventDisableTiming)) != cudaSuccess) return e;
        busy[s] = false;
    }
    next = 0; acquired = -1;
    if (!dMeta && (e = cudaMalloc((void**)&dMeta, SLOTS * 4 * sizeof(float))) != cudaSuccess) return e;
    if (!hMeta && (e = cudaMallocHost((void**)&hMeta, SLOTS * 4 * sizeof(float))) != cudaSuccess) return e;
    // buffers are allocated lazily by acquire(): a solver that never bakes a frame pays nothing
    worker = std::thread(&FramePipe::worker_main, this);
    return cudaSuccess;
}

VfdParticleSimple* FramePipe::acquire() {
    const int s = next;
    {
        std::unique_lock<std::mutex> g(m);
        cvDone.wait(g, [&] { return !busy[s]; });
    }
    const size_t bytes = (size_t)(n ? n : 1u) * sizeof(VfdParticleSimple);
    if (!dBuf[s] && cudaMalloc((void**)&dBuf[s], bytes) != cudaSuccess) return nullptr;
    acquired = s;
    return dBuf[s];
}

float* FramePipe::meta_slot() { return acquired >= 0 ? dMeta + 4 * acquired : nullptr; }

cudaError_t FramePipe::submit(cudaStream_t solverStream, float maxVel2, float dt, bool fromDevice, bool conditional) {
    if (acquired < 0) return cudaErrorInvalidValue;
    const int s = acquired;
    acquired = -1;
    const size_t bytes = (size_t)n * sizeof(VfdParticleSimple);
    HostBuf dst = take_buffer();
    cudaError_t e;
    if (!dst.pinned && !hBuf[s] && (e = cudaMallocHost((void**)&hBuf[s], bytes ? bytes : 1)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(exported[s], solverStream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(copyStream, exported[s], 0)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(dst.pinned ? dst.p : hBuf[s], dBuf[s], bytes, cudaMemcpyDeviceToHost, copyStream)) != cudaSuccess) return e;
    if (fromDevice && (e = cudaMemcpyAsync(hMeta + 4 * s, dMeta + 4 * s, 4 * sizeof(float), cudaMemcpyDeviceToHost, copyStream)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(copied[s], copyStream)) != cudaSuccess) return e;
    bytesCopied += (uint64_t)bytes;
    {
        std::lock_guard<std::mutex> g(m);
        busy[s] = true;
        jobs.push_back(Job{ s, dst, fromDevice, conditional, maxVel2, dt });
        inFlight++;
    }
    cvJob.notify_one();
    next = (s + 1) % SLOTS;
    return cudaSuccess;
}

void FramePipe::worker_main() {
    cudaSetDevice(device);
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> g(m);
            cvJob.wait(g, [&] { return quit || !jobs.empty(); });
            if (jobs.empty()) return;          // quit
            j = jobs.front();
            jobs.pop_front();
        }
        Frame f;
        f.count = n; f.maxVel2 = j.maxVel2; f.dt = j.dt;
        f.data = j.dst;
        const cudaError_t e = cudaEventSynchronize(copied[j.slot]);
        const bool keep = !(e == cudaSuccess && j.conditional && hMeta[4 * j.slot + 2] == 0.0f);   // the device decided: no frame this step
        if (e == cudaSuccess && keep && !j.dst.pinned) memcpy(f.data.p, hBuf[j.slot], (size_t)n * sizeof(VfdParticleSimple));
        if (e == cudaSuccess && j.metaFromDevice) { f.maxVel2 = hMeta[4 * j.slot]; f.dt = hMeta[4 * j.slot + 1]; }
        {
            std::lock_guard<std::mutex> g(m);
            if (e != cudaSuccess && asyncError == cudaSuccess) asyncError = e;
            if (keep) frames.push_back(std::move(f)); else pool.push_back(f.data);
            busy[j.slot] = false;
            inFlight--;
        }
        cvDone.notify_all();
    }
}

cudaError_t FramePipe::drain() {
    std::unique_lock<std::mutex> g(m);
    cvDone.wait(g, [&] { return inFlight == 0; });
    const cudaError_t e = asyncError;
    asyncError = cudaSuccess;
    return e;
}

void FramePipe::clear() {
    drain();
    std::lock_guard<std::mutex> g(m);
    for (Frame& f : frames) pool.push_back(f.data);     // the storage is kept for the next bake
    frames.clear();
}

void FramePipe::progress(size_t& publishedFrames, size_t& pendingJobs) {
    std::lock_guard<std::mutex> g(m);
    publishedFrames = frames.size();
    pendingJobs = inFlight;
}

void FramePipe::wait_pending_below(size_t below) {
    std::unique_lock<std::mutex> g(m);
    cvDone.wait(g, [&] { return inFlight < below; });
}

size_t FramePipe::published() {
    std::lock_guard<std::mutex> g(m);
    return frames.size();
}

bool FramePipe::read(uint32_t index, VfdParticleSimple* out, float* maxVel2, float* dt) {
    std::lock_guard<std::mutex> g(m);
    if (index >= frames.size()) return false;
    const Frame& f = frames[index];
    if (out) memcpy(out, f.data.p, f.count * sizeof(VfdParticleSimple));
    if (maxVel2) *maxVel2 = f.maxVel2;
    if (dt) *dt = f.dt;
    return true;
}

bool FramePipe::view(uint32_t index, const VfdParticleSimple** data, uint32_t* count, float* maxVel2, float* dt) {
    std::lock_guard<std::mutex> g(m);
    if (index >= frames.size()) return false;
    const Frame& f = frames[index];
    if (data) *data = f.data.p;
    if (count) *count = (uint32_t)f.count;
    if (maxVel2) *maxVel2 = f.maxVel2;
    if (dt) *dt = f.dt;
    return true;
}

} // namespace vfd
