// volume_map.cuh — device-side lookup in a volume map: the 32-node cubic serendipity basis and the
// interpolation of field 0 (signed distance, with gradient) and field 1 (boundary volume).
// Reference: Utility/SDF/SDFDeviceData.cuh:36-369 (ShapeFunction), :372-419 (DetermineShapeFunction),
// :421-470 (Interpolate).  Shared by boundary.cu (per-step lookups) and volume_map.cu (map precompute).
#pragma once
#include "solver.h"

namespace vfd {

struct Basis {
    float X[2], Y[2], Z[2];            // 1 -/+ x, ...
    float T3x[2], T3y[2], T3z[2];      // 1 -/+ 3x, ...
    float omx2, omy2, omz2;            // 1 - x^2, ...
    float facC, facX, facY, facZ;      // corner / edge prefactors
    float Ax, Ay, Az, x18, y18, z18;   // corner-gradient terms
    float ex[2], ey[2], ez[2];         // edge-gradient terms along the edge direction: -/+ (3-9x^2) - 2x
    __device__ __forceinline__ void init(float x, float y, float z) {
        const float x2 = x * x, y2 = y * y, z2 = z * z;
        X[0] = 1.0f - x; X[1] = 1.0f + x; Y[0] = 1.0f - y; Y[1] = 1.0f + y; Z[0] = 1.0f - z; Z[1] = 1.0f + z;
        T3x[0] = 1.0f - 3.0f * x; T3x[1] = 1.0f + 3.0f * x;
        T3y[0] = 1.0f - 3.0f * y; T3y[1] = 1.0f + 3.0f * y;
        T3z[0] = 1.0f - 3.0f * z; T3z[1] = 1.0f + 3.0f * z;
        omx2 = 1.0f - x2; omy2 = 1.0f - y2; omz2 = 1.0f - z2;
        facC = 1.0f / 64.0f * (9.0f * (x2 + y2 + z2) - 19.0f);
        facX = 9.0f / 64.0f * omx2; facY = 9.0f / 64.0f * omy2; facZ = 9.0f / 64.0f * omz2;
        Ax = 9.0f * (3.0f * x2 + y2 + z2) - 19.0f;
        Ay = 9.0f * (x2 + 3.0f * y2 + z2) - 19.0f;
        Az = 9.0f * (x2 + y2 + 3.0f * z2) - 19.0f;
        x18 = 18.0f * x; y18 = 18.0f * y; z18 = 18.0f * z;
        const float a = 3.0f - 9.0f * x2, b = 3.0f - 9.0f * y2, c = 3.0f - 9.0f * z2;
        ex[0] = -a - 2.0f * x; ex[1] = a - 2.0f * x;
        ey[0] = -b - 2.0f * y; ey[1] = b - 2.0f * y;
        ez[0] = -c - 2.0f * z; ez[1] = c - 2.0f * z;
    }
    // value and (unscaled) gradient of node j, j a compile-time constant after unrolling
    __device__ __forceinline__ void node(int j, float& N, float3& dN) const {
        constexpr float rfe = 9.0f / 64.0f;
        if (j < 8) {
            const int bx = j & 1, by = (j >> 1) & 1, bz = (j >> 2) & 1;
            N = facC * (X[bx] * Y[by]) * Z[bz];
            dN.x = ((bx ? x18 + Ax : x18 - Ax) * (Y[by] * Z[bz])) / 64.0f;
            dN.y = ((X[bx] * Z[bz]) * (by ? y18 + Ay : y18 - Ay)) / 64.0f;
            dN.z = ((X[bx] * Y[by]) * (bz ? z18 + Az : z18 - Az)) / 64.0f;
        } else if (j < 16) {
            const int t = j & 1, bz = (j >> 1) & 1, by = (j >> 2) & 1;
            const float e = omx2 * T3x[t];
            N = (facX * T3x[t]) * (Y[by] * Z[bz]);
            dN.x = (ex[t] * (Y[by] * Z[bz])) * rfe;
            dN.y = ((by ? e : -e) * Z[bz]) * rfe;
            dN.z = ((bz ? e : -e) * Y[by]) * rfe;
        } else if (j < 24) {
            const int t = j & 1, bx = (j >> 1) & 1, bz = (j >> 2) & 1;
            const float e = omy2 * T3y[t];
            N = (facY * T3y[t]) * (X[bx] * Z[bz]);
            dN.x = ((bx ? e : -e) * Z[bz]) * rfe;
            dN.y = (ey[t] * (X[bx] * Z[bz])) * rfe;
            dN.z = ((bz ? e : -e) * X[bx]) * rfe;
        } else {
            const int t = j & 1, by = (j >> 1) & 1, bx = (j >> 2) & 1;
            const float e = omz2 * T3z[t];
            N = (facZ * T3z[t]) * (X[bx] * Y[by]);
            dN.x = ((bx ? e : -e) * Y[by]) * rfe;
            dN.y = ((by ? e : -e) * X[bx]) * rfe;
            dN.z = (ez[t] * (X[bx] * Y[by])) * rfe;
        }
    }
};

// Looks up field 0 (phi, grad phi) and field 1 (volume) at x. Returns false where the reference's
// DetermineShapeFunction does (outside the map domain / unmapped cell).
__device__ __forceinline__ bool map_lookup(const DevVolumeMap& M, float3 x, float& phi, float3& grad, float& vol) {
    if (!(M.dmin[0] <= x.x && M.dmin[1] <= x.y && M.dmin[2] <= x.z && M.dmax[0] >= x.x && M.dmax[1] >= x.y && M.dmax[2] >= x.z)) return false;
    uint32_t mi[3];
    mi[0] = (uint32_t)(M.cellInv[0] * (x.x - M.dmin[0]));
    mi[1] = (uint32_t)(M.cellInv[1] * (x.y - M.dmin[1]));
    mi[2] = (uint32_t)(M.cellInv[2] * (x.z - M.dmin[2]));
    #pragma unroll
    for (int k = 0; k < 3; k++) if (mi[k] >= M.res[k]) mi[k] = M.res[k] - 1u;
    const uint32_t ci = M.res[1] * M.res[0] * mi[2] + M.res[0] * mi[1] + mi[0];
    const uint32_t cj = __ldg(M.cellMap + ci);
    if (cj == 0xffffffffu) return false;
    // sub-domain of the cell and the map to local coordinates in [-1,1]^3 (SDFDeviceData.cuh:403-407)
    float xi[3], c0[3];
    const float xv[3] = { x.x, x.y, x.z };
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = M.dmin[k] + (float)mi[k] * M.cell[k];
        const float hi = lo + M.cell[k];
        const float den = hi - lo;
        c0[k] = 2.0f / den;
        const float c1 = (hi + lo) / den;
        xi[k] = c0[k] * xv[k] - c1;
    }
    Basis B;
    B.init(xi[0], xi[1], xi[2]);
    const uint32_t* cell = M.cells + (size_t)cj * 32u;
    const float* n0 = M.nodes;
    const float* n1 = M.nodes + M.nodeCount;
    float p = 0.0f, v = 0.0f;
    float3 g = f3(0.0f, 0.0f, 0.0f);
    #pragma unroll
    for (int j = 0; j < 32; j++) {
        const uint32_t nid = __ldg(cell + j);
        const float c = __ldg(n0 + nid);
        const float cv = __ldg(n1 + nid);
        float N; float3 dN;
        B.node(j, N, dN);
        p += c * N;
        g.x += c * dN.x; g.y += c * dN.y; g.z += c * dN.z;
        v += cv * N;
    }
    phi = p;
    grad = f3(g.x * c0[0], g.y * c0[1], g.z * c0[2]);
    vol = v;
    return true;
}


} // namespace vfd
