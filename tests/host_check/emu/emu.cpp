// TEST INFRASTRUCTURE: storage of the CUDA-on-CPU emulation (emu_cuda.h).
#include "emu_cuda.h"
thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;
EmuBlock* g_emuBlock = nullptr;
#include "cuda_runtime.h"
std::atomic<long> g_emuPinnedAllocs{0}, g_emuPinnedBytes{0};
int g_emuCopyMicros = 300;
