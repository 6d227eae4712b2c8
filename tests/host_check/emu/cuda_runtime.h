// tests/host_check/emu/cuda_runtime.h — TEST INFRASTRUCTURE: the slice of the CUDA runtime that csrc/frame_pipe.cu uses, emulated
// on the host WITH its asynchrony: a stream is a thread draining a queue of operations in order, an event completes when the
// stream that recorded it gets there, copies take a moment.  So the frame pipe's hazards — a device slot or a staging buffer
// reused before its copy has been consumed, frames published out of order, the worker racing the submitting thread — are live
// in tests/test_frame_pipe_emulation_cpu.py, not just its bookkeeping.
#pragma once
#include "emu_cuda.h"
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>

struct emuStream {
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    bool quit = false, idle = true;
    std::thread th;
    emuStream() : th([this] { run(); }) {}
    void run() {
        for (;;) {
            std::function<void()> op;
            {
                std::unique_lock<std::mutex> g(m);
                idle = q.empty();
                if (idle) cv.notify_all();
                cv.wait(g, [&] { return quit || !q.empty(); });
                if (q.empty()) return;
                op = std::move(q.front());
                q.pop_front();
                idle = false;
            }
            op();
        }
    }
    void push(std::function<void()> op) { { std::lock_guard<std::mutex> g(m); q.push_back(std::move(op)); } cv.notify_all(); }
    void sync() { std::unique_lock<std::mutex> g(m); cv.wait(g, [&] { return q.empty() && idle; }); }
    ~emuStream() { { std::lock_guard<std::mutex> g(m); quit = true; } cv.notify_all(); th.join(); }
};
struct emuEvent {
    std::mutex m;
    std::condition_variable cv;
    uint64_t recorded = 0, completed = 0;
    void wait(uint64_t ticket) { std::unique_lock<std::mutex> g(m); cv.wait(g, [&] { return completed >= ticket; }); }
};
typedef emuStream* cudaStream_t;
typedef emuEvent* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
extern std::atomic<long> g_emuPinnedAllocs, g_emuPinnedBytes;      // cudaMallocHost bookkeeping, for the tests
extern int g_emuCopyMicros;                                          // how long an asynchronous copy takes

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) { *s = new emuStream(); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { if (s) { s->sync(); delete s; } return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { if (s) s->sync(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) { *e = new emuEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    uint64_t ticket;
    { std::lock_guard<std::mutex> g(e->m); ticket = ++e->recorded; }
    s->push([e, ticket] { { std::lock_guard<std::mutex> g(e->m); if (e->completed < ticket) e->completed = ticket; } e->cv.notify_all(); });
    return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned int) {
    uint64_t ticket;
    { std::lock_guard<std::mutex> g(e->m); ticket = e->recorded; }
    s->push([e, ticket] { e->wait(ticket); });
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
    uint64_t ticket;
    { std::lock_guard<std::mutex> g(e->m); ticket = e->recorded; }
    e->wait(ticket);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* src, size_t n, cudaMemcpyKind, cudaStream_t s) {
    s->push([=] { std::this_thread::sleep_for(std::chrono::microseconds(g_emuCopyMicros)); memcpy(d, src, n); });
    return cudaSuccess;
}
template<typename T> inline cudaError_t cudaMallocHost(T** p, size_t n) {
    *p = (T*)malloc(n ? n : 1);
    memset(*p, 0xEE, n ? n : 1);
    g_emuPinnedAllocs++; g_emuPinnedBytes += (long)n;
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
