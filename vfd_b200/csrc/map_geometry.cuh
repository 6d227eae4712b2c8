// map_geometry.cuh — the grid of a volume map / distance field: domain, cell sizes, node numbering (Discregrid layout, the
// reference's SDF class).  Host+device and free of CUDA types, so that tests/host_check can compile it with g++ next to
// mesh_distance.cuh and pin the node positions and field 0 of whole maps against the reference on the CPU.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define VFD_GEOM_HD __host__ __device__ inline
#else
#define VFD_GEOM_HD inline
#endif

namespace vfd {

struct MapGeom {
    float dmin[3], dmax[3], cell[3], cellInv[3];
    uint32_t res[3];
    uint32_t nv, nex, ney, nez, nodeCount, cellCount;
};

// Position of node i (the numbering SDF::IndexToNodePosition decodes, SDF.cu:313-373 — it is the interchange format of the
// map, so the arithmetic is the reference's).  Nodes come in four groups: the (nx+1)(ny+1)(nz+1) cell corners, x fastest;
// then for each axis a = x, y, z the two interior nodes (at 1/3 and 2/3) of every cell edge along a.  The edges along a are
// numbered with the coordinate along a fastest (res[a] values), then axis a+1 (res+1 values), then axis a+2.
VFD_GEOM_HD void node_position(const MapGeom& G, uint32_t i, float out[3]) {
    uint32_t digit[3];
    int along = -1;                 // corner nodes: no edge
    uint32_t third = 0u;            // 0: the node at 1/3 of its edge, 1: at 2/3
    if (i < G.nv) {
        const uint32_t ex = G.res[0] + 1u, ey = G.res[1] + 1u;
        digit[0] = i % ex; digit[1] = (i / ex) % ey; digit[2] = i / (ex * ey);
    } else {
        uint32_t r = i - G.nv;
        const uint32_t perAxis[3] = { 2u * G.nex, 2u * G.ney, 2u * G.nez };
        along = 0;
        while (along < 2 && r >= perAxis[along]) { r -= perAxis[along]; along++; }
        third = r & 1u;
        const uint32_t e = r >> 1;
        const int f = along, m = (along + 1) % 3, s = (along + 2) % 3;
        const uint32_t ef = G.res[f], em = G.res[m] + 1u;
        digit[f] = e % ef; digit[m] = (e / ef) % em; digit[s] = e / (ef * em);
    }
    for (int k = 0; k < 3; k++) out[k] = G.dmin[k] + G.cell[k] * (float)digit[k];
    if (along >= 0) out[along] += (1.0f + (float)third) / 3.0f * G.cell[along];
}

// cell sizes and node counts of a grid whose domain (dmin, dmax) is set (SDF::SDF, SDF.cu:8-14; AddFunction :47-56)
inline void map_finish_geometry(MapGeom& G, const uint32_t resolution[3]) {
    for (int k = 0; k < 3; k++) {
        G.res[k] = resolution[k];
        G.cell[k] = (G.dmax[k] - G.dmin[k]) / (float)resolution[k];
        G.cellInv[k] = 1.0f / G.cell[k];
    }
    const uint32_t nx = G.res[0], ny = G.res[1], nz = G.res[2];
    G.nv = (nx + 1) * (ny + 1) * (nz + 1);
    G.nex = nx * (ny + 1) * (nz + 1); G.ney = (nx + 1) * ny * (nz + 1); G.nez = (nx + 1) * (ny + 1) * nz;
    G.nodeCount = G.nv + 2 * (G.nex + G.ney + G.nez);
    G.cellCount = nx * ny * nz;
}

// the volume-map domain of a body whose vertices span [lo, hi]: BoundingBox(vertices) starts from min = max = 0, i.e. always
// contains the origin, grown by 8 h + tolerance (RigidBody.cu:34-36)
inline void body_map_geometry(const float lo[3], const float hi[3], float h, float tolerance, const uint32_t resolution[3], MapGeom& G) {
    for (int k = 0; k < 3; k++) {
        const float l = fminf(0.0f, lo[k]), u = fmaxf(0.0f, hi[k]);
        G.dmax[k] = u + (8.0f * h + tolerance);
        G.dmin[k] = l - (8.0f * h + tolerance);
    }
    map_finish_geometry(G, resolution);
}

// the distance grid ParticleSampler::SampleMeshVolume samples against (SDF::SDF(mesh, bounds, resolution, inverted),
// SDF.cu:16-37): the bounds grown by a thousandth of their diagonal — first max, then min by the NEW diagonal (:21-22)
inline void sampler_grid_geometry(const float lo[3], const float hi[3], const uint32_t resolution[3], MapGeom& G) {
    for (int k = 0; k < 3; k++) { G.dmin[k] = lo[k]; G.dmax[k] = hi[k]; }
    for (int pass = 0; pass < 2; pass++) {
        const float d[3] = { G.dmax[0] - G.dmin[0], G.dmax[1] - G.dmin[1], G.dmax[2] - G.dmin[2] };
        const float grow = 0.001f * sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        for (int k = 0; k < 3; k++) { if (pass == 0) G.dmax[k] += grow; else G.dmin[k] -= grow; }
    }
    map_finish_geometry(G, resolution);
}

} // namespace vfd
