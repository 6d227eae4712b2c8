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

// ---- whole grids: node positions and field 0 as vfd_volume_map_build_mesh / vfd_sample_mesh_volume lay them out ------------
#include "../../vfd_b200/csrc/map_geometry.cuh"
using namespace vfd;

static void field0(const MeshHost& H, const MapGeom& G, float sign, float tolerance, float* nodes0) {
    const MeshView M = H.view();
    #pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < (int64_t)G.nodeCount; l++) {
        float x[3];
        node_position(G, (uint32_t)l, x);
        const V3 p = v3(x[0], x[1], x[2]);
        float best = FLT_MAX; uint32_t face = 0;
        for (uint32_t f = 0; f < H.faceCount; f++) {
            const float d2 = closest_on_triangle(p, load3(M.tri, 3 * (size_t)f), load3(M.tri, 3 * (size_t)f + 1), load3(M.tri, 3 * (size_t)f + 2)).d2;
            if (d2 < best) { best = d2; face = f; }
        }
        nodes0[l] = sign * (signed_distance_on_face(M, face, p) - tolerance);
    }
}

static void put_geometry(const MapGeom& G, float* geom15, uint32_t* counts) {
    for (int k = 0; k < 3; k++) { geom15[k] = G.dmin[k]; geom15[3 + k] = G.dmax[k]; geom15[6 + k] = G.cell[k]; geom15[9 + k] = G.cellInv[k]; geom15[12 + k] = (float)G.res[k]; }
    counts[0] = G.nodeCount; counts[1] = G.cellCount;
}

// the body map's grid and field 0 (nodes0 may be null: sizes only)
extern "C" int hc_body_map_field0(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* transform16, int inverted, float padding,
                                  const uint32_t* resolution, float particleRadius, float* geom15, uint32_t* counts, float* nodes0) {
    MeshHost H;
    if (!prepare_mesh(verts, nv, tris, nt, transform16, H)) return 1;
    const float h = 4.0f * particleRadius, tolerance = padding - particleRadius, sign = inverted ? -1.0f : 1.0f;
    MapGeom G;
    body_map_geometry(H.lo, H.hi, h, tolerance, resolution, G);
    put_geometry(G, geom15, counts);
    if (nodes0) field0(H, G, sign, tolerance, nodes0);
    return 0;
}

// the sampler's distance grid, field 0 and lattice; axes: up to cap values each, counts[2..4] = axis lengths
extern "C" int hc_sampler_grid(const float* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, const float* transform16, int inverted,
                               const uint32_t* resolution, float particleRadius, int mode, float* geom15, uint32_t* counts, float* nodes0,
                               float* xs, float* ys, float* zs, uint32_t cap) {
    MeshHost H;
    if (!prepare_mesh(verts, nv, tris, nt, transform16, H)) return 1;
    MapGeom G;
    sampler_grid_geometry(H.lo, H.hi, resolution, G);
    put_geometry(G, geom15, counts);
    if (nodes0) field0(H, G, inverted ? -1.0f : 1.0f, 0.0f, nodes0);
    float sx, sy, sz;
    lattice_steps(mode, particleRadius, sx, sy, sz);
    std::vector<float> a[3];
    lattice_axis(H.lo[0], H.hi[0], sx, a[0]); lattice_axis(H.lo[1], H.hi[1], sy, a[1]); lattice_axis(H.lo[2], H.hi[2], sz, a[2]);
    float* dst[3] = { xs, ys, zs };
    for (int k = 0; k < 3; k++) { counts[2 + k] = (uint32_t)a[k].size(); for (size_t i = 0; i < a[k].size() && i < cap; i++) dst[k][i] = a[k][i]; }
    return 0;
}

// candidate positions of the lattice (ix fastest), as the kernels form them
extern "C" void hc_lattice_positions(int mode, float particleRadius, const float* xs, uint32_t nx, const float* ys, uint32_t ny, const float* zs, uint32_t nz, float* out) {
    Lattice L;
    float sx, sy, sz;
    lattice_steps(mode, particleRadius, sx, sy, sz);
    L.nx = nx; L.ny = ny; L.nz = nz; L.mode = mode; L.radius = particleRadius; L.diameter = 2.0f * particleRadius; L.shiftX = sx;
    size_t o = 0;
    for (uint32_t iz = 0; iz < nz; iz++) for (uint32_t iy = 0; iy < ny; iy++) for (uint32_t ix = 0; ix < nx; ix++, o++) {
        const V3 p = lattice_position(L, xs[ix], ys[iy], zs[iz], ix, iy);
        out[3 * o] = p.x; out[3 * o + 1] = p.y; out[3 * o + 2] = p.z;
    }
}
