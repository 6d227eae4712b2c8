// tests/host_check/emu/emu_cuda.h — TEST INFRASTRUCTURE: a CUDA-on-CPU emulation just large enough to EXECUTE the scene-
// preparation kernels of vfd_b200/csrc/volume_map.cu on host cores (tests/test_mesh_emulation_cpu.py), so that their plumbing —
// grids with a ragged last block, shared-memory staging between barriers, the ballot/popc compaction, the warp reduction —
// has run against the reference's outputs before it ever meets a GPU.  Every thread of a block is an OS thread; barriers are
// pthread barriers; __shared__ is a static variable (blocks run one after another).  Nothing of the product links this.
#pragma once
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
inline float3 make_float3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

struct emu_dim3 { unsigned int x = 1, y = 1, z = 1; };
extern thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;

using std::max;
using std::min;
template<typename T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned int v) { return __builtin_popcount(v); }

struct EmuBlock {
    unsigned int threads = 0;
    pthread_barrier_t all;
    pthread_barrier_t warp[32];
    float lanes[32][32];
    std::atomic<unsigned int> ballot[32];
    std::atomic<int> count{0};
};
extern EmuBlock* g_emuBlock;

inline void __syncthreads() { pthread_barrier_wait(&g_emuBlock->all); }
inline int __syncthreads_count(int predicate) {
    EmuBlock& B = *g_emuBlock;
    if (predicate) B.count.fetch_add(1);
    pthread_barrier_wait(&B.all);
    const int r = B.count.load();
    pthread_barrier_wait(&B.all);
    if (threadIdx.x == 0) B.count.store(0);
    pthread_barrier_wait(&B.all);
    return r;
}
inline unsigned int __ballot_sync(unsigned int, int predicate) {
    EmuBlock& B = *g_emuBlock;
    const unsigned int w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (predicate) B.ballot[w].fetch_or(1u << lane);
    pthread_barrier_wait(&B.warp[w]);
    const unsigned int r = B.ballot[w].load();
    pthread_barrier_wait(&B.warp[w]);
    if (lane == 0) B.ballot[w].store(0u);
    pthread_barrier_wait(&B.warp[w]);
    return r;
}
inline float __shfl_down_sync(unsigned int, float v, int delta) {
    EmuBlock& B = *g_emuBlock;
    const unsigned int w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    B.lanes[w][lane] = v;
    pthread_barrier_wait(&B.warp[w]);
    const float r = lane + (unsigned int)delta < 32u ? B.lanes[w][lane + delta] : v;
    pthread_barrier_wait(&B.warp[w]);
    return r;
}

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
template<typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); memset(*p, 0xCD, n ? n : 1); return cudaSuccess; }   // NOT zeroed: like the device
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }

// <<<blocks, threads>>>: the block's threads run concurrently (one OS thread each), the blocks one after another
template<typename K, typename... A>
inline void emu_launch(unsigned int blocks, unsigned int threads, K kernel, A... args) {
    EmuBlock B;
    B.threads = threads;
    pthread_barrier_init(&B.all, nullptr, threads);
    const unsigned int warps = (threads + 31u) / 32u;
    for (unsigned int w = 0; w < warps; w++) { pthread_barrier_init(&B.warp[w], nullptr, std::min(32u, threads - 32u * w)); B.ballot[w].store(0u); }
    g_emuBlock = &B;
    std::vector<std::thread> pool;
    for (unsigned int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            for (unsigned int b = 0; b < blocks; b++) {
                gridDim.x = blocks; blockDim.x = threads; blockIdx.x = b; threadIdx.x = t;
                kernel(args...);
                pthread_barrier_wait(&B.all);          // the next block reuses the block's "shared memory"
            }
        });
    for (std::thread& th : pool) th.join();
    pthread_barrier_destroy(&B.all);
    for (unsigned int w = 0; w < warps; w++) pthread_barrier_destroy(&B.warp[w]);
    g_emuBlock = nullptr;
}
