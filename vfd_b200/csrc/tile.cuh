// tile.cuh — the tile pass: how every neighbour-sum kernel of the solver walks the particles.
//
// The search (search.cu) sorts particles by a tile-major cell key: the grid is cut into tiles of
// 4x4x4 cells, tiles are ordered x-slowest, and the 64 cells of a tile are contiguous (x-fastest
// inside the tile).  A tile's particles therefore form ONE contiguous range of every SoA array,
// and every neighbour of a tile's particle lies in the 6x6x6-cell "halo box" around it.
//
// A tile pass is executed by one CTA per tile:
//   1. the 216 halo cells' particle ranges are read from the cell table and prefix-summed in
//      shared memory, which defines a tile-LOCAL index space (box order: hz, hy, hx; the three
//      x-adjacent cells of a stencil row are contiguous in it);
//   2. the payload the pass gathers per neighbour (position, plus kappa / velocity / pressure
//      acceleration / PCG direction ...) is staged ONCE for all halo particles from HBM/L2 into
//      shared memory with coalesced loads;
//   3. each thread owns one particle of the tile and streams its neighbour list — 16-bit tile-local
//      indices in a warp-blocked ELL layout, so a warp reads one 64-B line per neighbour slot —
//      gathering payloads from shared memory instead of through L1/L2.
// Per pass and particle HBM sees: own fields once + 2 B per neighbour; neighbour fields never.
//
// If a halo box holds more particles than the staging buffer (pathological clumping), the pass
// falls back to translating local indices to global ones (binary search in the 217-entry table)
// and gathers from global memory: slow, but exact.
#pragma once
#include "solver.h"

namespace vfd {

#define TILE_THREADS 512
#define TILE_WARPS (TILE_THREADS / 32)
#define HALO_CELLS 216
#define TILE_CELLS 64
#define STAGE_CAP16 2560      // staged halo particles for 16-B payloads (40 KB)
#define STAGE_CAP32 2304      // ... for 32-B payloads (72 KB)

struct Pay32 { float4 a, b; };

// ---- grid / key helpers ---------------------------------------------------------------------------
// cell slightly larger than h so that two particles closer than h can never be two cells apart
// through fp32 rounding of the cell coordinate
__device__ __forceinline__ float cell_inv(float h) { return (1.0f / h) * (1.0f - 1.0f / 1024.0f); }

__device__ __forceinline__ uint3 cell_of(float4 x, const DevState* S, float invCell) {
    uint3 c;
    c.x = (uint32_t)((x.x - S->gridOrigin[0]) * invCell);
    c.y = (uint32_t)((x.y - S->gridOrigin[1]) * invCell);
    c.z = (uint32_t)((x.z - S->gridOrigin[2]) * invCell);
    // robustness against NaN / escaped particles: clamp into the padded interior
    c.x = min(max(c.x, 1u), S->gridDim[0] - 2u);
    c.y = min(max(c.y, 1u), S->gridDim[1] - 2u);
    c.z = min(max(c.z, 1u), S->gridDim[2] - 2u);
    return c;
}

// tile-major key: tiles x-slowest (a slab of tile columns is one contiguous particle range), cells x-fastest inside
__device__ __forceinline__ uint32_t cell_key(uint32_t cx, uint32_t cy, uint32_t cz, const DevState* S) {
    const uint32_t tile = ((cx >> 2) * S->tileDim[1] + (cy >> 2)) * S->tileDim[2] + (cz >> 2);
    return tile * TILE_CELLS + (((cz & 3u) << 4) | ((cy & 3u) << 2) | (cx & 3u));
}

// ---- shared-memory header of a tile pass ----------------------------------------------------------
struct TileShared {
    uint32_t cellG[HALO_CELLS];       // global index of the first particle of each halo cell
    uint32_t local[HALO_CELLS + 8];   // exclusive prefix of the halo cells' particle counts (local index space)
    uint32_t scan[32];
    double   red[4 * 32];
    uint32_t flag;
};

struct TileInfo {
    uint32_t begin, end;      // the tile's own particles [begin, end) in the global arrays
    uint32_t total;           // particles in the halo box = size of the local index space
    bool staged;
};

// Step 1: halo cell table + prefix sum. All threads of the CTA call this (contains __syncthreads).
__device__ __forceinline__ TileInfo tile_setup(const DevState* __restrict__ S, const uint32_t* __restrict__ cellBegin,
                                               uint32_t tile, TileShared& sh, uint32_t cap) {
    const uint32_t tdy = S->tileDim[1], tdz = S->tileDim[2];
    const uint32_t tz = tile % tdz, ty = (tile / tdz) % tdy, tx = tile / (tdz * tdy);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t cnt = 0, beg = 0;
    if (tid < HALO_CELLS) {
        const int hx = tid % 6, hy = (tid / 6) % 6, hz = tid / 36;
        const int cx = (int)(tx << 2) + hx - 1, cy = (int)(ty << 2) + hy - 1, cz = (int)(tz << 2) + hz - 1;
        if (cx >= 0 && cy >= 0 && cz >= 0 && cx < (int)S->gridDim[0] && cy < (int)S->gridDim[1] && cz < (int)S->gridDim[2]) {
            const uint32_t key = cell_key((uint32_t)cx, (uint32_t)cy, (uint32_t)cz, S);
            beg = __ldg(cellBegin + key);
            cnt = __ldg(cellBegin + key + 1) - beg;
        }
        sh.cellG[tid] = beg;
    }
    // exclusive scan of 216 counts held by the first 7 warps
    uint32_t inc = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31 && warp < 8) sh.scan[warp] = inc;
    __syncthreads();
    if (tid < HALO_CELLS) {
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += sh.scan[w];
        sh.local[tid] = base + inc - cnt;
        if (tid == HALO_CELLS - 1) sh.local[HALO_CELLS] = base + inc;
    }
    __syncthreads();
    TileInfo t;
    t.begin = __ldg(cellBegin + tile * TILE_CELLS);
    t.end = __ldg(cellBegin + tile * TILE_CELLS + TILE_CELLS);
    t.total = sh.local[HALO_CELLS];
    t.staged = t.total <= cap;
    return t;
}

// Step 2: stage the payload of every halo particle; half a warp per cell (cells hold ~8 particles).
template<class Payload, class Load>
__device__ __forceinline__ void tile_stage(const TileShared& sh, Payload* __restrict__ sPay, const Load& load) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    for (int c = warp * 2 + half; c < HALO_CELLS; c += TILE_WARPS * 2) {
        const uint32_t l0 = sh.local[c], cnt = sh.local[c + 1] - l0, g0 = sh.cellG[c];
        for (uint32_t k = l16; k < cnt; k += 16) sPay[l0 + k] = load(g0 + k);
    }
}

// local -> global index (fallback path and list export): binary search in the prefix table
__device__ __forceinline__ uint32_t tile_local_to_global(const TileShared& sh, uint32_t L) {
    int lo = 0, hi = HALO_CELLS;          // find c with local[c] <= L < local[c+1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sh.local[mid] <= L) lo = mid; else hi = mid; }
    return sh.cellG[lo] + (L - sh.local[lo]);
}

template<class Payload> struct StagedAcc {
    const Payload* sPay;
    __device__ __forceinline__ Payload operator()(uint32_t L) const { return sPay[L]; }
};
template<class Payload, class Load> struct GlobalAcc {
    const TileShared* sh; const Load* load;
    __device__ __forceinline__ Payload operator()(uint32_t L) const { return (*load)(tile_local_to_global(*sh, L)); }
};

// ---- neighbour list access (u16 local indices, warp-blocked ELL) ---------------------------------
// the k-th neighbour of particle p sits at list16[((p>>5)*VFD_MAX_NEIGHBORS + k)*32 + (p&31)]
__device__ __forceinline__ size_t ell_base(uint32_t p) { return (size_t)(p >> 5) * (VFD_MAX_NEIGHBORS * 32) + (p & 31); }

// The pass driver.  Op provides:
//   typedef Payload;  static constexpr bool READ_COUNT;  Payload load(uint32_t g) const;
//   template<class Acc> void particle(uint32_t p, uint32_t m, size_t ell, const Acc& acc);
// Warps take 32-aligned particle groups so that list reads are full-line coalesced.
template<class Op>
__device__ __forceinline__ void tile_pass(const DevState* __restrict__ S, const uint32_t* __restrict__ cellBegin, const uint32_t* __restrict__ cnt,
                                          TileShared& sh, typename Op::Payload* sPay, uint32_t cap, Op& op, uint32_t* errorFlags = nullptr) {
    typedef typename Op::Payload Payload;
    const uint32_t nTiles = S->nTiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        // cheap emptiness test before the full setup
        const uint32_t b0 = __ldg(cellBegin + tile * TILE_CELLS), e0 = __ldg(cellBegin + tile * TILE_CELLS + TILE_CELLS);
        if (b0 == e0) continue;
        __syncthreads();                                   // previous tile's readers are done with shared memory
        const TileInfo t = tile_setup(S, cellBegin, tile, sh, cap);
        if (errorFlags && t.total > 65535u && threadIdx.x == 0) atomicOr(errorFlags, 2u);
        if (t.staged) {
            tile_stage<Payload>(sh, sPay, [&](uint32_t g) { return op.load(g); });
            __syncthreads();
            const StagedAcc<Payload> acc{ sPay };
            for (uint32_t grp = (t.begin >> 5) + warp; (grp << 5) < t.end; grp += TILE_WARPS) {
                const uint32_t p = (grp << 5) + lane;
                if (p >= t.begin && p < t.end) op.particle(p, Op::READ_COUNT ? __ldg(cnt + p) : 0u, ell_base(p), acc);
            }
        } else {
            auto ld = [&](uint32_t g) { return op.load(g); };
            const GlobalAcc<Payload, decltype(ld)> acc{ &sh, &ld };
            for (uint32_t grp = (t.begin >> 5) + warp; (grp << 5) < t.end; grp += TILE_WARPS) {
                const uint32_t p = (grp << 5) + lane;
                if (p >= t.begin && p < t.end) op.particle(p, Op::READ_COUNT ? __ldg(cnt + p) : 0u, ell_base(p), acc);
            }
        }
    }
}

// dynamic shared memory carve-up: [TileShared][LUT floats][payload]
__device__ __forceinline__ TileShared& smem_header(unsigned char* raw) { return *reinterpret_cast<TileShared*>(raw); }
__host__ __device__ constexpr size_t smem_header_bytes() { return (sizeof(TileShared) + 127) / 128 * 128; }

__device__ __forceinline__ void load_lut_tile(float* dst, const float* __restrict__ src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < (VFD_LUT_RES / 4); i += blockDim.x) d4[i] = __ldg(s4 + i);
}

} // namespace vfd
