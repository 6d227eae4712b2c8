// oracle/shim/emu.cpp — TEST INFRASTRUCTURE: storage for the CUDA-on-CPU emulation.
#include "cuda_runtime.h"
thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;
int g_emu_serial = 0;
