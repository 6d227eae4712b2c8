// volume_map.cu — scene preparation on the GPU: the two-field volume map of an axis-aligned box.
//
// Restates RigidBody::RigidBody (reference: Simulation/DFSPH/RigidBody/RigidBody.cu:10-73) over
// SDF::AddFunction / IndexToNodePosition (Utility/SDF/SDF.cu:45-139, :313-373) and
// GaussQuadrature::Integrate with p = 30 -> 16 points per axis (Core/Math/GaussQuadrature.cpp:5621-5655):
//   field 0 (all 32-node-element nodes): sign * (d_box(x) - (padding - r))     [mesh distance of a box = analytic]
//   field 1: 0.8 * int_{|xi|<h} gamma(phi(x + xi)) dxi,  gamma = 1 (phi<=0) | W(phi)/W(0) (phi<h) | 0,
//            where phi is *interpolated from field 0* (not the exact distance), 0 where phi(x) > 2h.
// The reference evaluates this on the host with OpenMP in 1-8 s per body; here one
// warp integrates one node.  Node numbering and the cell table are the reference's (Discregrid layout),
// so maps built here and maps flattened from the reference's SDF are interchangeable.
#include "solver.h"
#include "volume_map.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace vfd {

struct MapGeom {
    float dmin[3], dmax[3], cell[3], cellInv[3];
    uint32_t res[3];
    uint32_t nv, nex, ney, nez, nodeCount, cellCount;
};

// SDF::IndexToNodePosition (SDF.cu:313-373)
__host__ __device__ inline void node_position(const MapGeom& G, uint32_t i, float out[3]) {
    const uint32_t nx = G.res[0], ny = G.res[1], nz = G.res[2];
    float idx[3];
    if (i < G.nv) {
        idx[2] = (float)(i / ((ny + 1u) * (nx + 1u)));
        const uint32_t t = i % ((ny + 1u) * (nx + 1u));
        idx[1] = (float)(t / (nx + 1u)); idx[0] = (float)(t % (nx + 1u));
        for (int k = 0; k < 3; k++) out[k] = G.dmin[k] + G.cell[k] * idx[k];
    } else if (i < G.nv + 2u * G.nex) {
        i -= G.nv;
        const uint32_t e = i / 2u;
        idx[2] = (float)(e / ((ny + 1u) * nx));
        const uint32_t t = e % ((ny + 1u) * nx);
        idx[1] = (float)(t / nx); idx[0] = (float)(t % nx);
        for (int k = 0; k < 3; k++) out[k] = G.dmin[k] + G.cell[k] * idx[k];
        out[0] += (1.0f + (float)(i % 2u)) / 3.0f * G.cell[0];
    } else if (i < G.nv + 2u * (G.nex + G.ney)) {
        i -= G.nv + 2u * G.nex;
        const uint32_t e = i / 2u;
        idx[0] = (float)(e / ((nz + 1u) * ny));
        const uint32_t t = e % ((nz + 1u) * ny);
        idx[2] = (float)(t / ny); idx[1] = (float)(t % ny);
        for (int k = 0; k < 3; k++) out[k] = G.dmin[k] + G.cell[k] * idx[k];
        out[1] += (1.0f + (float)(i % 2u)) / 3.0f * G.cell[1];
    } else {
        i -= G.nv + 2u * (G.nex + G.ney);
        const uint32_t e = i / 2u;
        idx[1] = (float)(e / ((nx + 1u) * nz));
        const uint32_t t = e % ((nx + 1u) * nz);
        idx[0] = (float)(t / nz); idx[2] = (float)(t % nz);
        for (int k = 0; k < 3; k++) out[k] = G.dmin[k] + G.cell[k] * idx[k];
        out[2] += (1.0f + (float)(i % 2u)) / 3.0f * G.cell[2];
    }
}

// field 0 at every node: signed distance to the box surface (negative inside), shifted and signed
__global__ void k_map_sdf(MapGeom G, float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz,
                          float sign, float tolerance, float* __restrict__ nodes0) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= G.nodeCount) return;
    float x[3];
    node_position(G, l, x);
    const float lo[3] = { bminx, bminy, bminz }, hi[3] = { bmaxx, bmaxy, bmaxz };
    float q[3], outside2 = 0.0f, inside = -FLT_MAX;
    for (int k = 0; k < 3; k++) {
        q[k] = fmaxf(lo[k] - x[k], x[k] - hi[k]);        // > 0 outside the slab
        const float o = fmaxf(q[k], 0.0f);
        outside2 += o * o;
        inside = fmaxf(inside, q[k]);
    }
    const float d = outside2 > 0.0f ? sqrtf(outside2) : inside;
    nodes0[l] = sign * (d - tolerance);
}

// field-0 interpolation only (SDF::Interpolate without gradient, SDF.cu:141-186); FLT_MAX outside the domain
__device__ __forceinline__ float map_phi(const MapGeom& G, const float* __restrict__ nodes0, const uint32_t* __restrict__ cells, float3 x) {
    if (!(G.dmin[0] <= x.x && G.dmin[1] <= x.y && G.dmin[2] <= x.z && G.dmax[0] >= x.x && G.dmax[1] >= x.y && G.dmax[2] >= x.z)) return FLT_MAX;
    uint32_t mi[3];
    mi[0] = (uint32_t)((x.x - G.dmin[0]) * G.cellInv[0]);
    mi[1] = (uint32_t)((x.y - G.dmin[1]) * G.cellInv[1]);
    mi[2] = (uint32_t)((x.z - G.dmin[2]) * G.cellInv[2]);
    for (int k = 0; k < 3; k++) if (mi[k] >= G.res[k]) mi[k] = G.res[k] - 1u;
    const uint32_t ci = G.res[1] * G.res[0] * mi[2] + G.res[0] * mi[1] + mi[0];
    float xi[3];
    const float xv[3] = { x.x, x.y, x.z };
    for (int k = 0; k < 3; k++) {
        const float lo = G.dmin[k] + (float)mi[k] * G.cell[k];
        const float hi = lo + G.cell[k];
        const float den = hi - lo;
        xi[k] = (2.0f / den) * xv[k] - (hi + lo) / den;
    }
    Basis B;
    B.init(xi[0], xi[1], xi[2]);
    const uint32_t* cell = cells + (size_t)ci * 32u;
    float p = 0.0f;
    #pragma unroll
    for (int j = 0; j < 32; j++) {
        float N; float3 dN;
        B.node(j, N, dN);
        p += __ldg(nodes0 + __ldg(cell + j)) * N;
    }
    return p;
}

struct Quad16 { double x[16]; float w[16]; };   // abscissae stay double, weights are narrowed to float (GaussQuadrature.cpp:5636-5648)

// field 1: one warp per node, 4096 quadrature points strided over the lanes
__global__ void __launch_bounds__(256) k_map_volume(MapGeom G, Quad16 Q, const float* __restrict__ nodes0, const uint32_t* __restrict__ cells,
                                                     const float* __restrict__ lutW, float lutInvStep, float wZero, float h, float* __restrict__ nodes1) {
    const uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (l >= G.nodeCount) return;
    float xn[3];
    node_position(G, l, xn);
    const float3 x = f3(xn[0], xn[1], xn[2]);
    const float dx = map_phi(G, nodes0, cells, x);
    if (dx > 2.0f * h) { if (lane == 0) nodes1[l] = 0.0f; return; }
    float acc = 0.0f;
    for (uint32_t q = lane; q < 4096u; q += 32u) {
        const uint32_t i = q >> 8, j = (q >> 4) & 15u, k = q & 15u;
        const float wijk = (Q.w[i] * Q.w[j]) * Q.w[k];
        const float3 xi = f3((float)((double)h * Q.x[i]), (float)((double)h * Q.x[j]), (float)((double)h * Q.x[k]));
        float g = 0.0f;
        if (!(dot3(xi, xi) > h * h)) {
            const float dist = map_phi(G, nodes0, cells, x + xi);
            if (dist <= 0.0f) g = 1.0f;
            else if (dist < h) {
                // kernel.GetW(float r) (Kernel/DFSPHKernels.h:43-52): midpoint lookup, combined table
                const uint32_t pos = min((uint32_t)(dist * lutInvStep), (uint32_t)(VFD_LUT_RES - 2));
                g = __ldg(lutW + pos) / wZero;
            }
        }
        acc += wijk * g;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const double c = (double)(0.5f * (2.0f * h));       // c0 = 0.5 * Diagonal() of [-h, h]^3
        const float res = (float)((double)acc * (c * c * c));
        nodes1[l] = 0.8f * res;
    }
}

// 16-point Gauss-Legendre rule on [-1, 1] (Newton iteration on P_16 in double)
static void gauss_legendre16(double* x, double* w) {
    const int n = 16;
    for (int i = 0; i < n; i++) {
        double z = std::cos(3.14159265358979323846 * (i + 0.75) / (n + 0.5));
        double pp = 0.0;
        for (int it = 0; it < 100; it++) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; j++) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double z1 = z;
            z = z1 - p1 / pp;
            if (std::fabs(z - z1) < 1e-16) break;
        }
        x[n - 1 - i] = z;                       // ascending
        w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

static std::string g_mapError;

} // namespace vfd

using namespace vfd;

extern "C" int vfd_volume_map_build_box(const float bmin[3], const float bmax[3], int inverted, float padding,
                                        const uint32_t resolution[3], float particleRadius, int device, VfdVolumeMap* out) {
    if (!bmin || !bmax || !resolution || !out) return VFD_E_INVALID;
    memset(out, 0, sizeof *out);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) { cudaGetLastError(); return VFD_E_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return VFD_E_CUDA;
    for (int k = 0; k < 3; k++) if (resolution[k] == 0 || resolution[k] > 1024 || !(bmax[k] > bmin[k])) return VFD_E_INVALID;

    const float h = 4.0f * particleRadius;
    const float tolerance = padding - particleRadius;
    const float sign = inverted ? -1.0f : 1.0f;
    MapGeom G;
    // BoundingBox(vertices) starts from min = max = 0, i.e. always contains the origin
    for (int k = 0; k < 3; k++) {
        const float lo = fminf(0.0f, bmin[k]), hi = fmaxf(0.0f, bmax[k]);
        G.dmax[k] = hi + (8.0f * h + tolerance);
        G.dmin[k] = lo - (8.0f * h + tolerance);
        G.res[k] = resolution[k];
        G.cell[k] = (G.dmax[k] - G.dmin[k]) / (float)resolution[k];
        G.cellInv[k] = 1.0f / G.cell[k];
    }
    const uint32_t nx = G.res[0], ny = G.res[1], nz = G.res[2];
    G.nv = (nx + 1) * (ny + 1) * (nz + 1);
    G.nex = nx * (ny + 1) * (nz + 1); G.ney = (nx + 1) * ny * (nz + 1); G.nez = (nx + 1) * (ny + 1) * nz;
    G.nodeCount = G.nv + 2 * (G.nex + G.ney + G.nez);
    G.cellCount = nx * ny * nz;

    // cell -> node table (SDF.cu:81-131)
    std::vector<uint32_t> cells((size_t)G.cellCount * 32);
    for (uint32_t l = 0; l < G.cellCount; l++) {
        const uint32_t k = l / (ny * nx), t = l % (ny * nx), j = t / nx, i = t % nx;
        uint32_t* c = &cells[(size_t)l * 32];
        for (uint32_t b = 0; b < 8; b++) c[b] = (nx + 1) * (ny + 1) * (k + ((b >> 2) & 1)) + (nx + 1) * (j + ((b >> 1) & 1)) + i + (b & 1);
        uint32_t off = G.nv;
        for (uint32_t b = 0; b < 4; b++) { const uint32_t bz = b & 1, by = (b >> 1) & 1; c[8 + 2 * b] = off + 2 * (nx * (ny + 1) * (k + bz) + nx * (j + by) + i); c[9 + 2 * b] = c[8 + 2 * b] + 1; }
        off += 2 * G.nex;
        for (uint32_t b = 0; b < 4; b++) { const uint32_t bx = b & 1, bz = (b >> 1) & 1; c[16 + 2 * b] = off + 2 * (ny * (nz + 1) * (i + bx) + ny * (k + bz) + j); c[17 + 2 * b] = c[16 + 2 * b] + 1; }
        off += 2 * G.ney;
        for (uint32_t b = 0; b < 4; b++) { const uint32_t by = b & 1, bx = (b >> 1) & 1; c[24 + 2 * b] = off + 2 * (nz * (nx + 1) * (j + by) + nz * (i + bx) + k); c[25 + 2 * b] = c[24 + 2 * b] + 1; }
    }

    KernelTables T;
    T.build(h);
    double gx[16], gw[16];
    gauss_legendre16(gx, gw);
    Quad16 Q;
    for (int i = 0; i < 16; i++) { Q.x[i] = gx[i]; Q.w[i] = (float)gw[i]; }
    float *dN0 = nullptr, *dN1 = nullptr, *dW = nullptr; uint32_t* dC = nullptr;
    cudaError_t e = cudaMalloc(&dN0, (size_t)G.nodeCount * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dN1, (size_t)G.nodeCount * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dC, cells.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dW, VFD_LUT_RES * 4);
    if (e == cudaSuccess) e = cudaMemcpy(dC, cells.data(), cells.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dW, T.Wc.data(), VFD_LUT_RES * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_map_sdf<<<(G.nodeCount + 255) / 256, 256>>>(G, bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2], sign, tolerance, dN0);
        const uint64_t threads = (uint64_t)G.nodeCount * 32;
        k_map_volume<<<(uint32_t)((threads + 255) / 256), 256>>>(G, Q, dN0, dC, dW, T.invStep, T.wZero, h, dN1);
        e = cudaDeviceSynchronize();
    }
    float* nodes = (float*)malloc((size_t)G.nodeCount * 2 * 4);
    uint32_t* cellsOut = (uint32_t*)malloc(cells.size() * 2 * 4);
    uint32_t* cmap = (uint32_t*)malloc((size_t)G.cellCount * 2 * 4);
    if (e == cudaSuccess && nodes && cellsOut && cmap) {
        e = cudaMemcpy(nodes, dN0, (size_t)G.nodeCount * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(nodes + G.nodeCount, dN1, (size_t)G.nodeCount * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(dN0); cudaFree(dN1); cudaFree(dC); cudaFree(dW);
    if (e != cudaSuccess || !nodes || !cellsOut || !cmap) { free(nodes); free(cellsOut); free(cmap); cudaGetLastError(); return VFD_E_CUDA; }
    for (int f = 0; f < 2; f++) {
        memcpy(cellsOut + (size_t)f * cells.size(), cells.data(), cells.size() * 4);
        for (uint32_t l = 0; l < G.cellCount; l++) cmap[(size_t)f * G.cellCount + l] = l;
    }
    for (int k = 0; k < 3; k++) { out->domainMin[k] = G.dmin[k]; out->domainMax[k] = G.dmax[k]; out->resolution[k] = G.res[k]; out->cellSize[k] = G.cell[k]; out->cellSizeInverse[k] = G.cellInv[k]; }
    out->fieldCount = 2; out->nodeCount = G.nodeCount; out->cellCount = G.cellCount; out->cellMapCount = G.cellCount;
    out->nodes = nodes; out->cells = cellsOut; out->cellMap = cmap;
    return VFD_OK;
}

extern "C" void vfd_volume_map_free(VfdVolumeMap* m) {
    if (!m) return;
    free((void*)m->nodes); free((void*)m->cells); free((void*)m->cellMap);
    m->nodes = nullptr; m->cells = nullptr; m->cellMap = nullptr;
}
