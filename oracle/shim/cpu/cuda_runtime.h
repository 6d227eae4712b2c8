// oracle/shim/cuda_runtime.h — TEST INFRASTRUCTURE (CPU oracle build only).
// A CUDA-on-CPU emulation just large enough to execute the reference's DFSPH
// kernels (which use no __shared__, no __syncthreads, no textures —
// SURVEY.md F3) thread by thread on host cores.
#ifndef VFD_ORACLE_SHIM_CUDA_RUNTIME_H
#define VFD_ORACLE_SHIM_CUDA_RUNTIME_H
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <cfloat>
#include <climits>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__
#define __restrict__

struct emu_dim3 { unsigned int x = 1, y = 1, z = 1; };
extern thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

template<typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return cudaSuccess; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

inline unsigned int atomicAdd(unsigned int* a, unsigned int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
inline int atomicMin(int* a, int v) {
    int o = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (v < o && !__atomic_compare_exchange_n(a, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
inline int atomicMax(int* a, int v) {
    int o = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (v > o && !__atomic_compare_exchange_n(a, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}

// Kernel launch: blocks in parallel over OpenMP threads, threads of a block serially.
// With VFD_ORACLE_SERIAL_ATOMICS the launch is serial, which makes the atomic rank
// of the reference's counting sort (ParticleSearchKernels.cu:77) deterministic.
extern int g_emu_serial;
template<typename K, typename... A>
inline void cpu_launch(unsigned blocks, unsigned threads, K kernel, A... args) {
    if (g_emu_serial) {
        for (long long b = 0; b < (long long)blocks; b++) {
            gridDim.x = blocks; blockDim.x = threads; blockIdx.x = (unsigned)b;
            for (unsigned t = 0; t < threads; t++) { threadIdx.x = t; kernel(args...); }
        }
        return;
    }
    #pragma omp parallel for schedule(static)
    for (long long b = 0; b < (long long)blocks; b++) {
        gridDim.x = blocks; blockDim.x = threads; blockIdx.x = (unsigned)b;
        for (unsigned t = 0; t < threads; t++) { threadIdx.x = t; kernel(args...); }
    }
}
#endif
