// mesh_distance.cuh — signed distance from a point to a closed triangle mesh, the arithmetic of the reference's
// MeshDistance (Utility/SDF/MeshDistance.cpp) without its bounding-sphere hierarchy.
//
// The reference answers a query by a depth-first walk over a sphere tree with one traversal stack, one "closest face of
// my previous query" and one result cache PER OpenMP THREAD (MeshDistance.cpp:12-21, :61-122) — a shape that does not
// map onto a GPU.  Here every query point looks at every triangle (the kernels in volume_map.cu stage the triangles in
// shared memory, all lanes of a warp read the same one: a broadcast) and keeps the smallest squared distance; what is
// the reference's, operation for operation, is
//   * the squared distance point -> triangle with the barycentric region it falls into (:289-601: the seven-region
//     minimisation of the quadratic  Q(s,t) = a00 s^2 + 2 a01 s t + a11 t^2 + 2 b0 s + 2 b1 t + c  over the triangle,
//     D. Eberly, "Distance Between Point and Triangle in 3D") — its fp32 expressions decide which face is nearest;
//   * distance = sqrt(Q) of the nearest face, closest point = v0 + s e0 + t e1 (:599);
//   * the sign from the angle-weighted pseudo-normal of the closest FEATURE (:187-222): vertex normals = sum over the
//     incident faces of (corner angle x face normal) accumulated in face order (:26-52), edge normals = the two adjacent
//     face normals added (:260-279), face normals = normalize(cross(v1 - v0, v2 - v0)).
// Ties between faces (equal fp32 squared distances, e.g. a point whose closest feature is a shared edge) go to the lowest
// face index here and to whichever face the tree walk met first in the reference; the features coincide, the distance does
// not depend on the choice.
//
// Everything in this header compiles for the host as well (plain C++): tests/host_check builds it with g++ and compares it
// with the reference's own MeshDistance (oracle/_ref) on the CPU — the one part of the GPU scene preparation that can be
// pinned without a GPU.  The product only ever runs it inside CUDA kernels (volume_map.cu); there is no CPU path in the
// library.  Compile with contraction off (-fmad=false / -ffp-contract=off): the expressions below are the reference's
// roundings.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#if defined(__CUDACC__)
#define VFD_MESH_HD __host__ __device__ __forceinline__
#else
#define VFD_MESH_HD inline
#endif

namespace vfd {
namespace meshd {

struct V3 { float x, y, z; };
VFD_MESH_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
VFD_MESH_HD V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
VFD_MESH_HD V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
VFD_MESH_HD V3 mul(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
VFD_MESH_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }          // glm::dot: (x + y) + z
VFD_MESH_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }

// closest feature of a triangle: its three corners, its three edges (edge k runs from corner k to corner k+1), its interior
enum Feature : int { CORNER0 = 0, CORNER1 = 1, CORNER2 = 2, EDGE01 = 3, EDGE12 = 4, EDGE20 = 5, INTERIOR = 6 };

struct Closest {
    float d2;      // squared distance
    float s, t;    // closest point = v0 + s (v1 - v0) + t (v2 - v0)
    int feature;
};

// minimum of Q along the edge t = 0 (b = b0, a = a00, far corner 1) or s = 0 (b = b1, a = a11, far corner 2)
VFD_MESH_HD int along_edge(float b, float a, int farCorner, int edge) {
    if (b >= 0.0f) return CORNER0;
    return -b >= a ? farCorner : edge;
}

VFD_MESH_HD Closest closest_on_triangle(V3 p, V3 v0, V3 v1, V3 v2) {
    const V3 diff = sub(v0, p), e0 = sub(v1, v0), e1 = sub(v2, v0);
    const float a00 = dot(e0, e0), a01 = dot(e0, e1), a11 = dot(e1, e1);
    const float b0 = dot(diff, e0), b1 = dot(diff, e1), c = dot(diff, diff);
    const float det = fabsf(a00 * a11 - a01 * a01);
    float s = a01 * b1 - a11 * b0, t = a01 * b0 - a00 * b1;      // unnormalised minimiser of Q in the plane

    // 1. the feature: which of the seven regions of the (s, t) plane the minimiser lies in, and where on that region's
    //    boundary Q is smallest
    int f;
    if (s + t <= det) {
        if (s < 0.0f)      f = (t < 0.0f && b0 < 0.0f) ? along_edge(b0, a00, CORNER1, EDGE01) : along_edge(b1, a11, CORNER2, EDGE20);
        else if (t < 0.0f) f = along_edge(b0, a00, CORNER1, EDGE01);
        else               f = INTERIOR;
    } else {
        const float denom = a00 - 2.0f * a01 + a11;
        if (s < 0.0f) {                                   // beyond the edge s = 0, past corner 2
            const float q0 = a01 + b0, q1 = a11 + b1;
            if (q1 > q0) {
                const float numer = q1 - q0;
                if (numer >= denom) f = CORNER1;
                else { f = EDGE12; s = numer / denom; t = 1.0f - s; }
            } else f = q1 <= 0.0f ? CORNER2 : (b1 >= 0.0f ? CORNER0 : EDGE20);
        } else if (t < 0.0f) {                            // beyond the edge t = 0, past corner 1
            const float q0 = a01 + b1, q1 = a00 + b0;
            if (q1 > q0) {
                const float numer = q1 - q0;
                if (numer >= denom) f = CORNER2;
                else { f = EDGE12; t = numer / denom; s = 1.0f - t; }
            } else f = q1 <= 0.0f ? CORNER1 : (b0 >= 0.0f ? CORNER0 : EDGE01);
        } else {                                          // beyond the edge s + t = 1
            const float numer = a11 + b1 - a01 - b0;
            if (numer <= 0.0f) f = CORNER2;
            else if (numer >= denom) f = CORNER1;
            else { f = EDGE12; s = numer / denom; t = 1.0f - s; }
        }
    }

    // 2. Q at that feature
    Closest r;
    r.feature = f;
    switch (f) {
    case CORNER0: s = 0.0f; t = 0.0f; r.d2 = c; break;
    case CORNER1: s = 1.0f; t = 0.0f; r.d2 = a00 + 2.0f * b0 + c; break;
    case CORNER2: s = 0.0f; t = 1.0f; r.d2 = a11 + 2.0f * b1 + c; break;
    case EDGE01:  t = 0.0f; s = -b0 / a00; r.d2 = b0 * s + c; break;
    case EDGE20:  s = 0.0f; t = -b1 / a11; r.d2 = b1 * t + c; break;
    default:
        if (f == INTERIOR) { const float inv = 1.0f / det; s *= inv; t *= inv; }
        r.d2 = s * (a00 * s + a01 * t + 2.0f * b0) + t * (a01 * s + a11 * t + 2.0f * b1) + c;
        break;
    }
    if (r.d2 < 0.0f) r.d2 = 0.0f;                          // round-off
    r.s = s; r.t = t;
    return r;
}

// A mesh prepared for queries.  Device code sees the same arrays through MeshView.
struct MeshView {
    const float* tri;        // 12 floats per face: corner 0, 1, 2 as (x, y, z, 0)
    const float* faceNormal; // 4 floats per face
    const float* vertNormal; // 4 floats per vertex (angle-weighted sum, not normalised)
    const uint32_t* corner;  // 4 per face: vertex ids of the corners (+ pad)
    const int32_t* across;   // 4 per face: the face on the other side of edge 0, 1, 2 (-1: border) (+ pad)
    uint32_t faceCount, vertexCount;
};

VFD_MESH_HD V3 load3(const float* a, size_t i) { return v3(a[4 * i], a[4 * i + 1], a[4 * i + 2]); }

// MeshDistance::SignedDistance (:187-222) once the nearest face is known
VFD_MESH_HD float signed_distance_on_face(const MeshView& M, uint32_t face, V3 p) {
    const V3 v0 = load3(M.tri, 3 * (size_t)face), v1 = load3(M.tri, 3 * (size_t)face + 1), v2 = load3(M.tri, 3 * (size_t)face + 2);
    const Closest h = closest_on_triangle(p, v0, v1, v2);
    const V3 cp = add(add(v0, mul(h.s, sub(v1, v0))), mul(h.t, sub(v2, v0)));
    V3 n;
    if (h.feature <= CORNER2) n = load3(M.vertNormal, M.corner[4 * (size_t)face + h.feature]);
    else {
        n = load3(M.faceNormal, face);
        if (h.feature != INTERIOR) {
            const int32_t o = M.across[4 * (size_t)face + (h.feature - EDGE01)];
            if (o >= 0) n = add(n, load3(M.faceNormal, (size_t)o));
        }
    }
    float d = sqrtf(h.d2);
    if (dot(sub(p, cp), n) < 0.0f) d *= -1.0f;
    return d;
}

// ---- host side: preparing the arrays (once per mesh) ----------------------------------------------------------------
struct MeshHost {
    std::vector<float> tri, faceNormal, vertNormal;
    std::vector<uint32_t> corner;
    std::vector<int32_t> across;
    float lo[3], hi[3];      // BoundingBox(vertices): grown from min = max = 0, so it always holds the origin (BoundingBox.h:16-21,48-49)
    uint32_t faceCount = 0, vertexCount = 0;
    MeshView view() const {
        MeshView v; v.tri = tri.data(); v.faceNormal = faceNormal.data(); v.vertNormal = vertNormal.data(); v.corner = corner.data();
        v.across = across.data(); v.faceCount = faceCount; v.vertexCount = vertexCount; return v;
    }
};

inline V3 normalized(V3 a) { const float k = 1.0f / sqrtf(dot(a, a)); return v3(a.x * k, a.y * k, a.z * k); }   // glm::normalize: v * inversesqrt(dot)

// vertices: 3 floats each, triangles: 3 vertex ids each, transform16: column-major 4x4 applied as glm does
// (RigidBody.cu:19-23, FluidObject.cpp:10-14) or nullptr.  Returns false on an index out of range.
inline bool prepare_mesh(const float* vertices, uint32_t nv, const uint32_t* triangles, uint32_t nf, const float* transform16, MeshHost& H) {
    H.faceCount = nf; H.vertexCount = nv;
    std::vector<V3> x(nv);
    for (int k = 0; k < 3; k++) { H.lo[k] = 0.0f; H.hi[k] = 0.0f; }
    for (uint32_t i = 0; i < nv; i++) {
        V3 v = v3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
        if (transform16) {
            const float* m = transform16;                 // mat4 * vec4(v, 1): (col0 x + col1 y) + (col2 z + col3 w)
            float r[3];
            for (int k = 0; k < 3; k++) r[k] = (m[k] * v.x + m[4 + k] * v.y) + (m[8 + k] * v.z + m[12 + k] * 1.0f);
            v = v3(r[0], r[1], r[2]);
        }
        x[i] = v;
        const float c[3] = { v.x, v.y, v.z };
        for (int k = 0; k < 3; k++) { H.lo[k] = fminf(H.lo[k], c[k]); H.hi[k] = fmaxf(H.hi[k], c[k]); }
    }
    H.tri.assign((size_t)nf * 12, 0.0f); H.faceNormal.assign((size_t)nf * 4, 0.0f); H.vertNormal.assign((size_t)nv * 4, 0.0f);
    H.corner.assign((size_t)nf * 4, 0u); H.across.assign((size_t)nf * 4, -1);
    // half-edge (a -> b) of face f, edge e pairs with the first unpaired half-edge (b -> a) of an earlier face
    // (EdgeMesh.cpp:144-176); what stays unpaired is a border
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> open;    // (from, to) -> half-edge ids 3 f + e, oldest first
    for (uint32_t f = 0; f < nf; f++) {
        const uint32_t id[3] = { triangles[3 * f], triangles[3 * f + 1], triangles[3 * f + 2] };
        for (int k = 0; k < 3; k++) if (id[k] >= nv) return false;
        const V3 x0 = x[id[0]], x1 = x[id[1]], x2 = x[id[2]];
        const V3 c[3] = { x0, x1, x2 };
        for (int k = 0; k < 3; k++) { H.tri[12 * (size_t)f + 4 * k] = c[k].x; H.tri[12 * (size_t)f + 4 * k + 1] = c[k].y; H.tri[12 * (size_t)f + 4 * k + 2] = c[k].z; H.corner[4 * (size_t)f + k] = id[k]; }
        const V3 n = normalized(cross(sub(x1, x0), sub(x2, x0)));
        const V3 d1 = normalized(sub(x1, x0)), d2 = normalized(sub(x2, x1)), d3 = normalized(sub(x0, x2));
        const float alpha[3] = { acosf(dot(d1, v3(-d3.x, -d3.y, -d3.z))), acosf(dot(d2, v3(-d1.x, -d1.y, -d1.z))), acosf(dot(d3, v3(-d2.x, -d2.y, -d2.z))) };
        for (int k = 0; k < 3; k++) {
            float* vn = &H.vertNormal[4 * (size_t)id[k]];
            vn[0] += alpha[k] * n.x; vn[1] += alpha[k] * n.y; vn[2] += alpha[k] * n.z;
        }
        H.faceNormal[4 * (size_t)f] = n.x; H.faceNormal[4 * (size_t)f + 1] = n.y; H.faceNormal[4 * (size_t)f + 2] = n.z;
        for (uint32_t e = 0; e < 3; e++) {
            const uint32_t a = id[e], b = id[(e + 1) % 3];
            auto it = open.find(std::make_pair(b, a));
            if (it != open.end() && !it->second.empty()) {
                const uint32_t other = it->second.front();
                it->second.erase(it->second.begin());
                H.across[4 * (size_t)f + e] = (int32_t)(other / 3u);
                H.across[4 * (size_t)(other / 3u) + other % 3u] = (int32_t)f;
            } else open[std::make_pair(a, b)].push_back(3u * f + e);
        }
    }
    return true;
}
// ---- the sampling lattice of ParticleSampler::SampleMeshVolume (Utility/Sampler/ParticleSampler.cpp:7-91) -----------------
// Three nested float loops (z outermost, x innermost) over the mesh bounds; every axis is its own accumulation
// v0, v0 + step, (v0 + step) + step, ... , so the lattice is the product of three coordinate lists (lattice_axis) and the
// candidate at (ix, iy, iz) depends on the parities of ix and iy only (the reference's counterX / counterY).
enum SampleMode : int { SAMPLE_MIN_DENSITY = 0, SAMPLE_MEDIUM_DENSITY = 1, SAMPLE_MAX_DENSITY = 2 };

struct Lattice {
    uint32_t nx, ny, nz;
    int mode;
    float radius, diameter, shiftX;
};

inline void lattice_axis(float lo, float hi, float step, std::vector<float>& out) {
    out.clear();
    if (!(step > 0.0f)) return;
    for (float v = lo; v <= hi; v += step) {
        out.push_back(v);
        if (out.size() > (size_t)1 << 24) break;          // a degenerate step (v + step == v) would never end
    }
}

// shiftX, shiftY of the mode (:22-31)
inline void lattice_steps(int mode, float radius, float& stepX, float& stepY, float& stepZ) {
    const float diameter = 2.0f * radius;
    stepX = diameter; stepY = diameter; stepZ = diameter;
    if (mode == SAMPLE_MEDIUM_DENSITY) stepY = sqrtf(3.0f) * radius;
    else if (mode == SAMPLE_MAX_DENSITY) { stepX = sqrtf(3.0f) * radius; stepY = sqrtf(6.0f) * diameter / 3.0f; }
}

VFD_MESH_HD V3 lattice_position(const Lattice& L, float x, float y, float z, uint32_t ix, uint32_t iy) {
    const float r = L.radius;
    if (L.mode == SAMPLE_MIN_DENSITY) return v3(x + r, y + r, z + r);
    if (L.mode == SAMPLE_MEDIUM_DENSITY) return (iy & 1u) == 0u ? v3(x, y + r, z + r) : v3(x + r, y + r, z);
    V3 p = v3(x, y + r, z + r);
    float sx = 0.0f, sz = 0.0f;
    if (ix & 1u) sz += L.diameter / (2.0f * ((iy & 1u) ? -1.0f : 1.0f));
    if (iy & 1u) { sx += L.shiftX / 2.0f; sz += L.diameter / 2.0f; }
    return v3(p.x + sx, p.y + 0.0f, p.z + sz);
}

} // namespace meshd
} // namespace vfd
