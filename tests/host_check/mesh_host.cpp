// tests/host_check/mesh_host.cpp — TEST INFRASTRUCTURE.  Compiles the host+device arithmetic of the GPU scene preparation
// (vfd_b200/csrc/mesh_distance.cuh) with g++ so that it can be pinned against the reference's MeshDistance (oracle/_ref)
// on a machine without a GPU.  The loop over points and triangles below stands in for the kernels of volume_map.cu
// (k_mesh_sdf: same order — faces ascending, strict "<", then signed_distance_on_face).  Never linked into the product.
#include "../../vfd_b200/csrc/mesh_distance.cuh"
#include <cfloat>

using namespace vfd::meshd;

extern "C" int hc_mesh_signed_distance(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* transform16,
                                       const float* points, uint32_t n, float* out, uint32_t* nearestFace) {
    MeshHost H;
    if (!prepare_mesh(verts, nv, tris, nt, transform16, H)) return 1;
    const MeshView M = H.view();
    for (uint32_t i = 0; i < n; i++) {
        const V3 p = v3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
        float best = FLT_MAX; uint32_t face = 0;
        for (uint32_t f = 0; f < nt; f++) {
            const float d2 = closest_on_triangle(p, load3(M.tri, 3 * (size_t)f), load3(M.tri, 3 * (size_t)f + 1), load3(M.tri, 3 * (size_t)f + 2)).d2;
            if (d2 < best) { best = d2; face = f; }
        }
        out[i] = signed_distance_on_face(M, face, p);
        if (nearestFace) nearestFace[i] = face;
    }
    return 0;
}

extern "C" void hc_mesh_bounds(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* transform16, float* lo3hi3) {
    MeshHost H;
    prepare_mesh(verts, nv, tris, nt, transform16, H);
    for (int k = 0; k < 3; k++) { lo3hi3[k] = H.lo[k]; lo3hi3[3 + k] = H.hi[k]; }
}
