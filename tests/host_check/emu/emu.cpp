// TEST INFRASTRUCTURE: storage of the CUDA-on-CPU emulation (emu_cuda.h).
#include "emu_cuda.h"
thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;
EmuBlock* g_emuBlock = nullptr;
