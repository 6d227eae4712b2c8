// boundary.cu — rigid-body boundary handling with volume maps (Bender et al. 2019), as the
// reference does it: per particle and body one virtual boundary sample x_b with volume V_b.
//
// Replaces ComputeVolumeAndBoundaryKernel (reference: DFSPHKernels.cu:69-150) and the device
// lookups it inlines (Utility/SDF/SDFDeviceData.cuh:372-470: DetermineShapeFunction / Interpolate
// over the 32-node cubic serendipity basis, :36-369).  The basis is evaluated node by node from
// the sign bits of the node instead of 128 spelled-out expressions; every product keeps the
// reference's association so results agree to rounding of compiler FMA choices.
#include "solver.h"
#include "volume_map.cuh"

namespace vfd {

__global__ void __launch_bounds__(VFD_TPB) k_boundary(Params P, Arrays A, BodySet B) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const float4 x4 = A.pos[p];
    float3 x = f3(x4);
    bool moved = false, nearBody = false;
    for (uint32_t b = 0; b < P.nBodies; b++) {
        float3 xb = f3(0.0f, 0.0f, 0.0f);
        float vb = 0.0f;
        float dist = FLT_MAX, vol = 0.0f;
        float3 nrm = f3(0.0f, 0.0f, 0.0f);
        const bool inside = map_lookup(B.map[b], x, dist, nrm, vol);
        if (!inside) dist = FLT_MAX;
        if (dist > 0.0f && dist < P.h) {
            if (vol > 0.0f && vol != FLT_MAX) {
                const float nl = sqrtf(dot3(nrm, nrm));
                if (nl > 1.0e-9f) {
                    nrm = nrm / nl;
                    const float pd = fmaxf(dist + 0.5f * P.r, P.d);
                    xb = x - pd * nrm;
                    vb = vol;
                }
            }
        } else if (dist <= 0.0f) {
            // penetration: push the particle out along the normal and stop it (DFSPHKernels.cu:124-143)
            const float nl = sqrtf(dot3(nrm, nrm));
            if (nl > 1.0e-5f) {
                nrm = nrm / nl;
                float delta = P.d - dist;
                delta = fminf(delta, 0.1f * P.r);
                x += delta * nrm;
                A.vel[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                moved = true;
            }
        }
        A.bx[b][p] = make_float4(xb.x, xb.y, xb.z, vb);
        nearBody = nearBody || vb > 0.0f;
    }
    if (nearBody) A.cnt[p] |= VFD_NEAR_BODY;             // the search has just written the plain count
    if (moved) A.pos[p] = make_float4(x.x, x.y, x.z, x4.w);
}

void launch_boundary(const LaunchCfg& L, const Params& P, const Arrays& A, const BodySet& B) {
    if (P.nBodies == 0) return;
    LaunchScope ls(L, KID_BOUNDARY);
    k_boundary<<<std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB), VFD_TPB, 0, L.stream>>>(P, A, B);
}

} // namespace vfd
