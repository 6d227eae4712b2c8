// volume_map.cu — scene preparation on the GPU: the two-field volume map of a rigid body
// (an axis-aligned box by its analytic distance, or any closed triangle mesh by the reference's mesh distance), the signed
// distance to a mesh at arbitrary points, and the sampling of a mesh volume with particles.
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
#include "mesh_distance.cuh"
#include "map_geometry.cuh"
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

namespace vfd {

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

// ---- general triangle meshes --------------------------------------------------------------------------------------------
// Signed distance to a mesh at n points (map nodes when `points` is null): every thread owns one point and walks ALL faces —
// staged MESH_CHUNK at a time in shared memory, every lane of a warp reads the same face (a broadcast) — keeping the first
// face with the smallest squared distance; then the sign from that face's closest feature (mesh_distance.cuh).  The
// reference prunes with a sphere tree on the host (MeshDistance.cpp:61-185); here 62 k nodes x 100 k faces is still only
// ~6e9 point-triangle tests.
#define MESH_CHUNK 512
__global__ void __launch_bounds__(256) k_mesh_sdf(meshd::MeshView M, MapGeom G, const float* __restrict__ points, uint32_t n,
                                                   float sign, float tolerance, float* __restrict__ out) {
    __shared__ float4 sTri[3 * MESH_CHUNK];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    meshd::V3 p = meshd::v3(0.0f, 0.0f, 0.0f);
    if (live) {
        if (points) p = meshd::v3(points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2]);
        else { float x[3]; node_position(G, i, x); p = meshd::v3(x[0], x[1], x[2]); }
    }
    const float4* tri4 = reinterpret_cast<const float4*>(M.tri);
    float best = FLT_MAX;
    uint32_t face = 0u;
    for (uint32_t base = 0u; base < M.faceCount; base += MESH_CHUNK) {
        const uint32_t cnt = min((uint32_t)MESH_CHUNK, M.faceCount - base);
        __syncthreads();                                          // the previous chunk has been read by every thread
        for (uint32_t k = threadIdx.x; k < 3u * cnt; k += blockDim.x) sTri[k] = tri4[3 * (size_t)base + k];
        __syncthreads();
        if (live) {
            for (uint32_t k = 0u; k < cnt; k++) {
                const float4 a = sTri[3u * k], b = sTri[3u * k + 1u], c = sTri[3u * k + 2u];
                const float d2 = meshd::closest_on_triangle(p, meshd::v3(a.x, a.y, a.z), meshd::v3(b.x, b.y, b.z), meshd::v3(c.x, c.y, c.z)).d2;
                if (d2 < best) { best = d2; face = base + k; }
            }
        }
    }
    if (live) out[i] = sign * (meshd::signed_distance_on_face(M, face, p) - tolerance);
}

// ParticleSampler::SampleMeshVolume (ParticleSampler.cpp:33-87): lattice candidate idx = (iz ny + iy) nx + ix is a sample where
// the interpolated signed distance is negative.  Pass 1 flags the candidates and counts them per block; the host turns the
// counts into offsets; pass 2 writes the samples in lattice order (the reference's push_back order).
__device__ __forceinline__ meshd::V3 sample_candidate(const meshd::Lattice& L, const float* __restrict__ xs, const float* __restrict__ ys,
                                                      const float* __restrict__ zs, uint64_t idx) {
    const uint32_t ix = (uint32_t)(idx % L.nx), iy = (uint32_t)((idx / L.nx) % L.ny), iz = (uint32_t)(idx / ((uint64_t)L.nx * L.ny));
    return meshd::lattice_position(L, xs[ix], ys[iy], zs[iz], ix, iy);
}

__global__ void __launch_bounds__(256) k_sample_flag(meshd::Lattice L, const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ zs,
                                                      MapGeom G, const float* __restrict__ nodes0, const uint32_t* __restrict__ cells, uint64_t total,
                                                      unsigned char* __restrict__ flags, uint32_t* __restrict__ blockCounts) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int inside = 0;
    if (idx < total) {
        const meshd::V3 p = sample_candidate(L, xs, ys, zs, idx);
        inside = map_phi(G, nodes0, cells, f3(p.x, p.y, p.z)) < 0.0f ? 1 : 0;          // SDF::GetDistance(p, 0) < 0 (FLT_MAX outside the grid)
        flags[idx] = (unsigned char)inside;
    }
    const int c = __syncthreads_count(inside);
    if (threadIdx.x == 0) blockCounts[blockIdx.x] = (uint32_t)c;
}

__global__ void __launch_bounds__(256) k_sample_write(meshd::Lattice L, const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ zs,
                                                       uint64_t total, const unsigned char* __restrict__ flags, const uint64_t* __restrict__ blockOffsets,
                                                       float* __restrict__ out) {
    __shared__ uint32_t warpBase[8];
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool inside = idx < total && flags[idx] != 0;
    const uint32_t ballot = __ballot_sync(0xffffffffu, inside);
    if (lane == 0u) warpBase[warp] = (uint32_t)__popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = 0u; for (int w = 0; w < 8; w++) { const uint32_t c = warpBase[w]; warpBase[w] = run; run += c; } }
    __syncthreads();
    if (inside) {
        const uint64_t slot = blockOffsets[blockIdx.x] + warpBase[warp] + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
        const meshd::V3 p = sample_candidate(L, xs, ys, zs, idx);
        out[3 * slot] = p.x; out[3 * slot + 1] = p.y; out[3 * slot + 2] = p.z;
    }
}

// a prepared mesh in device memory
struct MeshDev {
    meshd::MeshView view;
    void* mem[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    cudaError_t upload(const meshd::MeshHost& H) {
        const void* src[5] = { H.tri.data(), H.faceNormal.data(), H.vertNormal.data(), H.corner.data(), H.across.data() };
        const size_t bytes[5] = { H.tri.size() * 4, H.faceNormal.size() * 4, H.vertNormal.size() * 4, H.corner.size() * 4, H.across.size() * 4 };
        for (int k = 0; k < 5; k++) {
            cudaError_t e = cudaMalloc(&mem[k], bytes[k] ? bytes[k] : 16);
            if (e == cudaSuccess && bytes[k]) e = cudaMemcpy(mem[k], src[k], bytes[k], cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return e;
        }
        view.tri = (const float*)mem[0]; view.faceNormal = (const float*)mem[1]; view.vertNormal = (const float*)mem[2];
        view.corner = (const uint32_t*)mem[3]; view.across = (const int32_t*)mem[4];
        view.faceCount = H.faceCount; view.vertexCount = H.vertexCount;
        return cudaSuccess;
    }
    ~MeshDev() { for (int k = 0; k < 5; k++) if (mem[k]) cudaFree(mem[k]); }
};

static int pick_device(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) { cudaGetLastError(); return VFD_E_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return VFD_E_CUDA;
    return VFD_OK;
}

// cell -> node table (SDF.cu:81-131)
static void map_cell_table(const MapGeom& G, std::vector<uint32_t>& cells) {
    const uint32_t nx = G.res[0], ny = G.res[1], nz = G.res[2];
    cells.resize((size_t)G.cellCount * 32);
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
}

// The two-field map over G: `field0` launches the kernel that fills field 0 at every node (device array), field 1 is
// integrated from it; both come back in malloc'ed host arrays (vfd_volume_map_free).
static int build_two_field_map(const MapGeom& G, float h, const std::function<cudaError_t(float*)>& field0, VfdVolumeMap* out) {
    std::vector<uint32_t> cells;
    map_cell_table(G, cells);
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
    if (e == cudaSuccess) e = field0(dN0);
    if (e == cudaSuccess) {
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

static bool resolution_ok(const uint32_t resolution[3]) {
    for (int k = 0; k < 3; k++) if (resolution[k] == 0 || resolution[k] > 1024) return false;
    // node and cell counts are 32-bit in the interchange format (SDFDeviceData): refuse a grid whose counts would wrap
    const uint64_t nx = resolution[0], ny = resolution[1], nz = resolution[2];
    const uint64_t nodes = (nx + 1) * (ny + 1) * (nz + 1) + 2 * (nx * (ny + 1) * (nz + 1) + (nx + 1) * ny * (nz + 1) + (nx + 1) * (ny + 1) * nz);
    return nodes <= 0x7fffffffull && nx * ny * nz * 32ull <= 0xffffffffull;
}

} // namespace vfd

using namespace vfd;

extern "C" int vfd_volume_map_build_box(const float bmin[3], const float bmax[3], int inverted, float padding,
                                        const uint32_t resolution[3], float particleRadius, int device, VfdVolumeMap* out) {
    if (!bmin || !bmax || !resolution || !out) return VFD_E_INVALID;
    memset(out, 0, sizeof *out);
    if (int rc = pick_device(device)) return rc;
    if (!resolution_ok(resolution)) return VFD_E_INVALID;
    for (int k = 0; k < 3; k++) if (!(bmax[k] > bmin[k])) return VFD_E_INVALID;

    const float h = 4.0f * particleRadius;
    const float tolerance = padding - particleRadius;
    const float sign = inverted ? -1.0f : 1.0f;
    MapGeom G;
    body_map_geometry(bmin, bmax, h, tolerance, resolution, G);
    return build_two_field_map(G, h, [&](float* dN0) -> cudaError_t {
        k_map_sdf<<<(G.nodeCount + 255) / 256, 256>>>(G, bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2], sign, tolerance, dN0);
        return cudaGetLastError();
    }, out);
}

// RigidBody::RigidBody for any closed triangle mesh (RigidBody.cu:10-73): vertices under the body's transform, the map
// domain from their bounds, field 0 = sign (d_mesh(x) - (padding - r)) with the reference's mesh distance.
extern "C" int vfd_volume_map_build_mesh(const float* vertices, uint32_t vertexCount, const uint32_t* triangles, uint32_t triangleCount,
                                         const float* transform16, int inverted, float padding, const uint32_t resolution[3],
                                         float particleRadius, int device, VfdVolumeMap* out) {
    if (!vertices || !triangles || !resolution || !out || vertexCount == 0 || triangleCount == 0) return VFD_E_INVALID;
    memset(out, 0, sizeof *out);
    if (int rc = pick_device(device)) return rc;
    if (!resolution_ok(resolution)) return VFD_E_INVALID;
    meshd::MeshHost H;
    if (!meshd::prepare_mesh(vertices, vertexCount, triangles, triangleCount, transform16, H)) return VFD_E_INVALID;

    const float h = 4.0f * particleRadius;
    const float tolerance = padding - particleRadius;
    const float sign = inverted ? -1.0f : 1.0f;
    MapGeom G;
    body_map_geometry(H.lo, H.hi, h, tolerance, resolution, G);
    MeshDev D;
    if (D.upload(H) != cudaSuccess) { cudaGetLastError(); return VFD_E_CUDA; }
    return build_two_field_map(G, h, [&](float* dN0) -> cudaError_t {
        k_mesh_sdf<<<(G.nodeCount + 255) / 256, 256>>>(D.view, G, nullptr, G.nodeCount, sign, tolerance, dN0);
        return cudaGetLastError();
    }, out);
}

// MeshDistance::SignedDistance (MeshDistance.cpp:187-222) at `count` points
extern "C" int vfd_mesh_signed_distance(const float* vertices, uint32_t vertexCount, const uint32_t* triangles, uint32_t triangleCount,
                                        const float* transform16, const float* points, uint32_t count, int device, float* out) {
    if (!vertices || !triangles || vertexCount == 0 || triangleCount == 0 || (count && (!points || !out))) return VFD_E_INVALID;
    if (int rc = pick_device(device)) return rc;
    if (count == 0) return VFD_OK;
    meshd::MeshHost H;
    if (!meshd::prepare_mesh(vertices, vertexCount, triangles, triangleCount, transform16, H)) return VFD_E_INVALID;
    MeshDev D;
    float *dP = nullptr, *dO = nullptr;
    cudaError_t e = D.upload(H);
    if (e == cudaSuccess) e = cudaMalloc(&dP, (size_t)count * 12);
    if (e == cudaSuccess) e = cudaMalloc(&dO, (size_t)count * 4);
    if (e == cudaSuccess) e = cudaMemcpy(dP, points, (size_t)count * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        MapGeom G;
        memset(&G, 0, sizeof G);
        k_mesh_sdf<<<(count + 255) / 256, 256>>>(D.view, G, dP, count, 1.0f, 0.0f, dO);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dO, (size_t)count * 4, cudaMemcpyDeviceToHost);
    cudaFree(dP); cudaFree(dO);
    if (e != cudaSuccess) { cudaGetLastError(); return VFD_E_CUDA; }
    return VFD_OK;
}

// FluidObject::FluidObject (FluidObject.cpp:6-26) -> ParticleSampler::SampleMeshVolume (ParticleSampler.cpp:7-91): the
// mesh under the object's transform, a signed-distance grid over its bounds (SDF::SDF(mesh, bounds, resolution, inverted),
// SDF.cu:16-37), the lattice of the sample mode, a sample wherever the interpolated distance is negative.
extern "C" int vfd_sample_mesh_volume(const float* vertices, uint32_t vertexCount, const uint32_t* triangles, uint32_t triangleCount,
                                      const float* transform16, float particleRadius, const uint32_t resolution[3], int inverted, int sampleMode,
                                      int device, float** positions, uint32_t* count) {
    if (!vertices || !triangles || !resolution || !positions || !count || vertexCount == 0 || triangleCount == 0) return VFD_E_INVALID;
    *positions = nullptr; *count = 0;
    if (sampleMode < 0 || sampleMode > 2 || !(particleRadius > 0.0f)) return VFD_E_INVALID;
    if (int rc = pick_device(device)) return rc;
    if (!resolution_ok(resolution)) return VFD_E_INVALID;
    meshd::MeshHost H;
    if (!meshd::prepare_mesh(vertices, vertexCount, triangles, triangleCount, transform16, H)) return VFD_E_INVALID;

    MapGeom G;
    sampler_grid_geometry(H.lo, H.hi, resolution, G);
    std::vector<uint32_t> cells;
    map_cell_table(G, cells);

    // the lattice over the ORIGINAL bounds (ParticleSampler.cpp:33-35)
    meshd::Lattice L;
    float stepX, stepY, stepZ;
    meshd::lattice_steps(sampleMode, particleRadius, stepX, stepY, stepZ);
    std::vector<float> xs, ys, zs;
    meshd::lattice_axis(H.lo[0], H.hi[0], stepX, xs);
    meshd::lattice_axis(H.lo[1], H.hi[1], stepY, ys);
    meshd::lattice_axis(H.lo[2], H.hi[2], stepZ, zs);
    L.nx = (uint32_t)xs.size(); L.ny = (uint32_t)ys.size(); L.nz = (uint32_t)zs.size();
    L.mode = sampleMode; L.radius = particleRadius; L.diameter = 2.0f * particleRadius; L.shiftX = stepX;
    const uint64_t total = (uint64_t)L.nx * L.ny * L.nz;
    if (total == 0) return VFD_OK;
    if (total > (1ull << 32)) return VFD_E_CAPACITY;
    const uint32_t blocks = (uint32_t)((total + 255) / 256);

    MeshDev D;
    float *dN0 = nullptr, *dX = nullptr, *dY = nullptr, *dZ = nullptr, *dOut = nullptr;
    uint32_t *dC = nullptr, *dCounts = nullptr; uint64_t* dOffsets = nullptr; unsigned char* dFlags = nullptr;
    std::vector<uint32_t> counts(blocks);
    std::vector<uint64_t> offsets(blocks);
    uint64_t kept = 0;
    float* host = nullptr;
    cudaError_t e = D.upload(H);
    if (e == cudaSuccess) e = cudaMalloc(&dN0, (size_t)G.nodeCount * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dC, cells.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dX, xs.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dY, ys.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dZ, zs.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dFlags, (size_t)total);
    if (e == cudaSuccess) e = cudaMalloc(&dCounts, (size_t)blocks * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dOffsets, (size_t)blocks * 8);
    if (e == cudaSuccess) e = cudaMemcpy(dC, cells.data(), cells.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dX, xs.data(), xs.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dY, ys.data(), ys.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dZ, zs.data(), zs.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const float factor = inverted ? -1.0f : 1.0f;
        k_mesh_sdf<<<(G.nodeCount + 255) / 256, 256>>>(D.view, G, nullptr, G.nodeCount, factor, 0.0f, dN0);
        k_sample_flag<<<blocks, 256>>>(L, dX, dY, dZ, G, dN0, dC, total, dFlags, dCounts);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(counts.data(), dCounts, (size_t)blocks * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
        for (uint32_t b = 0; b < blocks; b++) { offsets[b] = kept; kept += counts[b]; }
        if (kept > 0xffffffffull) e = cudaErrorInvalidValue;
    }
    if (e == cudaSuccess && kept) {
        e = cudaMemcpy(dOffsets, offsets.data(), (size_t)blocks * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&dOut, (size_t)kept * 12);
        if (e == cudaSuccess) {
            k_sample_write<<<blocks, 256>>>(L, dX, dY, dZ, total, dFlags, dOffsets, dOut);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) { host = (float*)malloc((size_t)kept * 12); if (!host) e = cudaErrorMemoryAllocation; }
        if (e == cudaSuccess) e = cudaMemcpy(host, dOut, (size_t)kept * 12, cudaMemcpyDeviceToHost);
    }
    cudaFree(dN0); cudaFree(dC); cudaFree(dX); cudaFree(dY); cudaFree(dZ); cudaFree(dFlags); cudaFree(dCounts); cudaFree(dOffsets); cudaFree(dOut);
    if (e != cudaSuccess) { free(host); cudaGetLastError(); return VFD_E_CUDA; }
    *positions = host; *count = (uint32_t)kept;
    return VFD_OK;
}

extern "C" void vfd_free(void* p) { free(p); }

extern "C" void vfd_volume_map_free(VfdVolumeMap* m) {
    if (!m) return;
    free((void*)m->nodes); free((void*)m->cells); free((void*)m->cellMap);
    m->nodes = nullptr; m->cells = nullptr; m->cellMap = nullptr;
}
