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

// SDF::IndexToNodePosition (SDF.cu:313-373)
VFD_GEOM_HD void node_position(const MapGeom& G, uint32_t i, float out[3]) {
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
