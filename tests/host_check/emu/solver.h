// tests/host_check/emu/solver.h — TEST INFRASTRUCTURE: what volume_map.cu, volume_map.cuh and tables.cpp take from the product's
// solver.h / common.cuh, restated for the CPU emulation build (same names, same expressions).
#pragma once
#include "emu_cuda.h"
#include "vfd_dfsph.h"
#include <string>
#include <vector>

#define VFD_LUT_RES 10000
#define VFD_HALTON_N (16384 * 3)

namespace vfd {

inline float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
inline float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct DevVolumeMap {
    float dmin[3], dmax[3];
    uint32_t res[3];
    float cell[3], cellInv[3];
    uint32_t fieldCount, nodeCount, cellCount, cellMapCount;
    const float* nodes; const uint32_t* cells; const uint32_t* cellMap;
};

struct KernelTables {
    std::vector<float> W, gradW;
    std::vector<float> Wc, Gc;
    float radius, radius2, invStep, wZero, k, l;
    void build(float radius);
};
void build_halton_table(std::vector<float>& out);

} // namespace vfd
