// search.cu — neighbourhood search: tile-major cell hashing, single-pass counting (radix) sort keyed by
// the cell id, physical reorder of the particle SoA, cell tables and neighbour lists.
//
// Replaces ParticleSearch::FindNeighbors (reference: ParticleSearch/ParticleSearch.h:48-61,
// ParticleSearch.cu:29-195, kernels ParticleSearchKernels.cu:40-192).  The parity obligation is
// the neighbour *set* of every particle:  { j : 0 < |x_j - x_i|^2 < h^2 }, capped at 70
// (ParticleSearchKernels.cu:126,131).  Differences by design:
//   * particles are physically reordered every step (the reference only builds an index
//     permutation and gathers 120-B AoS structs through it);
//   * the cell key is tile-major (4x4x4-cell tiles, the reference uses 8^3 Morton blocks), so that a
//     CTA owns one compact tile and stages its 6x6x6-cell halo box in shared memory (tile.cuh);
//   * in-cell order is by persistent particle id (deterministic, and identical on every GPU that
//     holds a copy of the cell), not by atomic arrival;
//   * one traversal writes the list — 16-bit tile-local indices in a warp-blocked ELL — instead of
//     count + scan + fill of 32-bit global ids;
//   * no host round trip: the grid is derived on the device.
#include "solver.h"
#include "tile.cuh"
#include <limits.h>

namespace vfd {

#define SCAN_ITEMS 16
#define SCAN_TILE (VFD_TPB * SCAN_ITEMS)   // 4096 cells per tile

// S1: extrema of floor(x/h) (same definition as ComputeMinMaxKernel, ParticleSearchKernels.cu:40-62),
// then the last block derives the grid.
__global__ void __launch_bounds__(VFD_TPB) k_bounds(Params P, const float4* __restrict__ pos, DevState* S, uint32_t cellCapacity) {
    int mn[3] = { INT_MAX, INT_MAX, INT_MAX }, mx[3] = { INT_MIN, INT_MIN, INT_MIN };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
        const float4 x = pos[i];
        const int cx = (int)floorf(x.x / P.h), cy = (int)floorf(x.y / P.h), cz = (int)floorf(x.z / P.h);
        mn[0] = min(mn[0], cx); mn[1] = min(mn[1], cy); mn[2] = min(mn[2], cz);
        mx[0] = max(mx[0], cx); mx[1] = max(mx[1], cy); mx[2] = max(mx[2], cz);
    }
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
        mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
    }
    if ((threadIdx.x & 31) == 0) {
        #pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(&S->boundsMin[k], mn[k]); atomicMax(&S->boundsMax[k], mx[k]); }
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&S->ticket[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        unsigned long long cells = 1ull;
        for (int k = 0; k < 3; k++) {
            const int lo = *(volatile int*)&S->boundsMin[k], hi = *(volatile int*)&S->boundsMax[k];
            S->gridMinCell[k] = lo;                      // reported bounds (ParticleSearch.cu:50-54)
            long long dim = (long long)hi - lo + 5;         // two pad cells below, two above
            dim = (min(dim, 1ll << 20) + 3) / 4 * 4;        // whole tiles
            S->gridDim[k] = (uint32_t)dim;
            S->tileDim[k] = (uint32_t)(dim / 4);
            S->gridOrigin[k] = (float)(lo - 2) * P.h;
            cells *= (unsigned long long)dim;
        }
        if (cells + 1 > cellCapacity) {
            S->errorFlags |= 1u; cells = 64;
            for (int k = 0; k < 3; k++) { S->gridDim[k] = 4; S->tileDim[k] = 1; }
        }
        S->nCells = (uint32_t)cells;
        S->nTiles = (uint32_t)(cells / 64);
        S->boundsMax[0] = S->boundsMax[1] = S->boundsMax[2] = INT_MIN;   // re-arm for the next step
        // keep the reported maximum cell for GetBounds in the padded slots of boundsMin? no: store in gridMinCell/gridDim
        S->boundsMin[0] = S->boundsMin[1] = S->boundsMin[2] = INT_MAX;
        S->ticket[0] = 0;
    }
}

// S2: cell key + arrival rank (the digit histogram of the single radix pass)
__global__ void __launch_bounds__(VFD_TPB) k_hist(Params P, const float4* __restrict__ pos, const DevState* __restrict__ S,
                                                  uint32_t* __restrict__ key, uint32_t* __restrict__ rank, uint32_t* cellCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint3 c = cell_of(pos[i], S, cell_inv(P.h));
    const uint32_t k = cell_key(c.x, c.y, c.z, S);
    key[i] = k;
    rank[i] = atomicAdd(&cellCount[k], 1u);
}

// S3: exclusive scan of the cell histogram in three phases over 4096-cell tiles.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* sh, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (VFD_TPB / 32) ? sh[lane] : 0u, winc = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += y; }
        sh[32 + lane] = winc - w;
        if (lane == 31) sh[64] = winc;
    }
    __syncthreads();
    total = sh[64];
    const uint32_t r = sh[32 + warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(VFD_TPB) k_scan_tiles(const DevState* __restrict__ S, const uint32_t* __restrict__ cellCount, uint32_t* __restrict__ tileSums) {
    __shared__ uint32_t sh[65];
    const uint32_t nCells = S->nCells, nTiles = (nCells + SCAN_TILE - 1) / SCAN_TILE;
    for (uint32_t t = blockIdx.x; t < nTiles; t += gridDim.x) {
        const uint32_t base = t * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
        uint32_t s = 0;
        #pragma unroll
        for (int q = 0; q < SCAN_ITEMS; q += 4) {
            if (base + q + 3 < nCells) { const uint4 v = *reinterpret_cast<const uint4*>(cellCount + base + q); s += v.x + v.y + v.z + v.w; }
            else for (int u = 0; u < 4; u++) if (base + q + u < nCells) s += cellCount[base + q + u];
        }
        uint32_t total;
        block_exclusive_scan(s, sh, total);
        if (threadIdx.x == 0) tileSums[t] = total;
    }
}

__global__ void __launch_bounds__(1024) k_scan_tile_sums(const DevState* __restrict__ S, uint32_t* __restrict__ tileSums) {
    __shared__ uint32_t sh[33];
    __shared__ uint32_t carry;
    const uint32_t nTiles = (S->nCells + SCAN_TILE - 1) / SCAN_TILE;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nTiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nTiles ? tileSums[i] : 0u;
        uint32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) sh[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = sh[lane], winc = w;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += y; }
            sh[lane] = winc - w;
            if (lane == 31) sh[32] = winc;
        }
        __syncthreads();
        if (i < nTiles) tileSums[i] = carry + sh[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[32];
        __syncthreads();
    }
}

// phase 3 also clears the histogram for the next step (invariant: cellCount is all-zero outside S2..S3)
__global__ void __launch_bounds__(VFD_TPB) k_scan_apply(Params P, const DevState* __restrict__ S, uint32_t* __restrict__ cellCount,
                                                        const uint32_t* __restrict__ tileSums, uint32_t* __restrict__ cellBegin) {
    __shared__ uint32_t sh[65];
    const uint32_t nCells = S->nCells, nTiles = (nCells + SCAN_TILE - 1) / SCAN_TILE;
    for (uint32_t t = blockIdx.x; t < nTiles; t += gridDim.x) {
        const uint32_t base = t * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        uint32_t s = 0;
        #pragma unroll
        for (int q = 0; q < SCAN_ITEMS; q++) { v[q] = (base + q < nCells) ? cellCount[base + q] : 0u; s += v[q]; }
        uint32_t total;
        uint32_t run = block_exclusive_scan(s, sh, total) + tileSums[t];
        #pragma unroll
        for (int q = 0; q < SCAN_ITEMS; q++) {
            if (base + q < nCells) { cellBegin[base + q] = run; cellCount[base + q] = 0u; }
            run += v[q];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cellBegin[nCells] = P.n;
}

// S3b: the non-empty owned tiles, compacted in tile order: tileList[0] = their number K, tileList[1..K] = tile indices.
// A pipelined tile pass (tile.cuh) hands entry i to CTA i mod G: every CTA gets the same number of non-empty tiles
// (+-1) — about a quarter of the grid's tiles are empty in a settled dam break, and a round robin over ALL tiles left
// some CTAs with a third more work than the mean — while CTAs that run side by side still work on neighbouring tiles
// and share their halo boxes in L2 (a contiguous range per CTA was measured 60 % slower for that reason).  The
// assignment is static, so every reduction's summation order is fixed (bit-reproducible), which a dynamic tile queue
// would not give.  One block; thread i owns a contiguous chunk of tiles.
#define PARTITION_THREADS 1024
__global__ void __launch_bounds__(PARTITION_THREADS) k_compact_tiles(Params P, DevState* __restrict__ S, const uint32_t* __restrict__ cellBegin,
                                                                     uint32_t* __restrict__ tileList) {
    __shared__ uint32_t shScan[PARTITION_THREADS / 32 + 1];
    const uint32_t t0 = P.tile0, t1 = min(S->nTiles, P.tile1);
    const uint32_t nT = t1 > t0 ? t1 - t0 : 0u;
    const uint32_t chunk = (nT + PARTITION_THREADS - 1) / PARTITION_THREADS;
    const uint32_t a = min(t0 + threadIdx.x * chunk, t1), b = min(a + chunk, t1);
    auto nonempty = [&](uint32_t t) -> bool { return cellBegin[(size_t)(t + 1) * TILE_CELLS] != cellBegin[(size_t)t * TILE_CELLS]; };
    uint32_t mine = 0;
    for (uint32_t t = a; t < b; t++) mine += nonempty(t) ? 1u : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = mine;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) shScan[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int w = 0; w < PARTITION_THREADS / 32; w++) { const uint32_t v = shScan[w]; shScan[w] = run; run += v; }
        tileList[0] = run;
        S->tileCursor = 0u; S->doneCtas = 0u;
    }
    __syncthreads();
    uint32_t run = shScan[warp] + inc - mine;
    for (uint32_t t = a; t < b; t++) if (nonempty(t)) tileList[1u + run++] = t;
}

// S4: scatter slot indices into cell order (arbitrary order inside a cell)
__global__ void __launch_bounds__(VFD_TPB) k_scatter(Params P, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank,
                                                     const uint32_t* __restrict__ cellBegin, uint32_t* __restrict__ tmpIdx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    tmpIdx[cellBegin[key[i]] + rank[i]] = i;
}

// S5: make the in-cell order deterministic (rank by persistent particle id) and move the persistent
// particle state to its sorted slot.
__global__ void __launch_bounds__(VFD_TPB) k_reorder(Params P, Arrays A) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.n) return;
    const uint32_t i0 = A.tmpIdx[q];
    const uint32_t c = A.key[i0];
    const uint32_t b = A.cellBegin[c], e = A.cellBegin[c + 1];
    const uint32_t myId = A.id[i0];
    uint32_t r = 0;
    for (uint32_t t = b; t < e; t++) r += (A.id[A.tmpIdx[t]] < myId) ? 1u : 0u;
    const uint32_t dst = b + r;
    A.pos2[dst] = A.pos[i0];
    A.vel2[dst] = A.vel[i0];
    A.dv2[dst] = A.dv[i0];
    A.nbar2[dst] = A.nbar[i0];
    A.curv2[dst] = A.curv[i0];
    A.curvS2[dst] = A.curvS[i0];
    A.curvD2[dst] = A.curvD[i0];
    A.id2[dst] = A.id[i0];
}

// S6/S8: one traversal of the 3x3 stencil rows (3 x-adjacent cells = one contiguous local range) over the
// positions staged in shared memory; writes 16-bit tile-local indices.
template<bool FMA>
struct SearchOp {
    static constexpr bool CUSTOM = true;
    static constexpr int NPAY = 1;
    const float4* __restrict__ pos;
    const DevState* __restrict__ S;
    const TileShared* sh;
    uint32_t* __restrict__ cnt;
    uint16_t* __restrict__ list;
    float h2, invCell;
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return pos[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t) const { return make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
    template<class Acc>
    __device__ __forceinline__ void particle(uint32_t p, bool valid, const Acc& acc) {
        if (!valid) return;
        const float4 xi = pos[p];
        const uint3 c = cell_of(xi, S, invCell);
        const uint32_t hx = (c.x & 3u) + 1u, hy = (c.y & 3u) + 1u, hz = (c.z & 3u) + 1u;   // box coordinates of the own cell
        const TileShared& tab = *sh;
        uint2* col = reinterpret_cast<uint2*>(list) + ell_base(p);
        uint32_t m = 0;
        // Slot order.  Every later pass has lane l of a warp gather, at slot s, its s-th neighbour's 16-byte payload from
        // shared memory; an LDS.128 is served eight lanes at a time and is conflict-free when those eight local indices are
        // distinct mod 8 (profiles/r01_ncu_*: in candidate order 8.9 wavefronts per LDS.128 against an ideal 4).  The set is
        // what parity fixes, the order is free: neighbours are bucketed by index mod 8 as they are found, and slot s then
        // takes one of residue (lane + s) mod 8 while that bucket lasts (else from the fullest one), so the lanes of a
        // quarter-warp walk the eight bank groups in step: ~5.6 wavefronts per LDS.128 on a settled fluid.
        // Buckets: 8 x SLOT_BUCKET entries of thread-local scratch; the eight fill counts are the bytes of one 64-bit register.
        constexpr uint32_t SLOT_BUCKET = 12;                 // 96 entries >= 70; a full bucket spills into the next one
        uint16_t bucket[8 * SLOT_BUCKET];
        unsigned long long fill = 0ull;
        for (int dz = -1; dz <= 1 && m < VFD_MAX_NEIGHBORS; dz++) {
            for (int dy = -1; dy <= 1 && m < VFD_MAX_NEIGHBORS; dy++) {
                const uint32_t c0 = ((hz + dz) * 6u + (hy + dy)) * 6u + hx - 1u;
                const uint32_t jb = tab.local[c0], je = tab.local[c0 + 3u];
                for (uint32_t j = jb; j < je; j++) {
                    const float4 xj = acc(j);
                    const float dx = xj.x - xi.x, dyy = xj.y - xi.y, dzz = xj.z - xi.z;
                    float d2;
                    if (FMA) d2 = __fmaf_rn(dzz, dzz, __fmaf_rn(dx, dx, __fmul_rn(dyy, dyy)));   // how nvcc compiles ParticleSearchKernels.cu:124
                    else     d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dyy, dyy)), __fmul_rn(dzz, dzz));
                    if (d2 < h2 && d2 > 0.0f) {
                        uint32_t r = j & 7u;
                        uint32_t f = (uint32_t)(fill >> (8u * r)) & 0xffu;
                        while (f == SLOT_BUCKET) { r = (r + 1u) & 7u; f = (uint32_t)(fill >> (8u * r)) & 0xffu; }
                        bucket[r * SLOT_BUCKET + f] = (uint16_t)j;
                        fill += 1ull << (8u * r);
                        if (++m == VFD_MAX_NEIGHBORS) break;
                    }
                }
            }
        }
        {
            const uint32_t lane = threadIdx.x & 31u;
            uint32_t L[4] = { 0u, 0u, 0u, 0u };
            for (uint32_t sidx = 0; sidx < m; sidx++) {
                uint32_t r = (lane + sidx) & 7u;
                uint32_t f = (uint32_t)(fill >> (8u * r)) & 0xffu;
                if (f == 0u) {                                 // that bucket is used up: take from the fullest
                    #pragma unroll
                    for (uint32_t q = 0; q < 8u; q++) { const uint32_t fq = (uint32_t)(fill >> (8u * q)) & 0xffu; if (fq > f) { f = fq; r = q; } }
                }
                fill -= 1ull << (8u * r);
                const uint32_t j = bucket[r * SLOT_BUCKET + f - 1u];
                const uint32_t q = sidx & 3u;
                L[0] = q == 0u ? j : L[0]; L[1] = q == 1u ? j : L[1]; L[2] = q == 2u ? j : L[2]; L[3] = q == 3u ? j : L[3];
                if (q == 3u) { col[(size_t)(sidx >> 2) * 32] = ell_pack(L); L[0] = L[1] = L[2] = L[3] = 0u; }
            }
            if (m & 3u) col[(size_t)(m >> 2) * 32] = ell_pack(L);
        }
        cnt[p] = m;
    }
};

template<bool FMA>
__global__ void __launch_bounds__(TT_PLAIN) k_build_list(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    TileShared& sh = smem_header(smemRaw);
    SearchOp<FMA> op{ A.pos, S, &sh, A.cnt, A.list16, P.h2, cell_inv(P.h) };
    tile_pass(S, A, sh, smem_pay_a<0>(smemRaw), nullptr, STAGE_CAP, op, P.tile0, P.tile1, true);
}

static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

void launch_search(const LaunchCfg& L, const Params& P, Arrays& A, DevState* S, uint32_t cellCapacity, uint32_t cellEstimate) {
    const uint32_t nb = std::max(1u, div_up(P.n, VFD_TPB));
    const uint32_t gb = std::max(1u, std::min<uint32_t>(nb, (uint32_t)L.numSMs * 8u));
    // several ranks: the grid is the fixed global one (distributed.cu); one GPU: derived from the particles' extent
    if (P.nRanks == 1) { LaunchScope ls(L, KID_BOUNDS); k_bounds<<<gb, VFD_TPB, 0, L.stream>>>(P, A.pos, S, cellCapacity); }
    { LaunchScope ls(L, KID_HIST); k_hist<<<nb, VFD_TPB, 0, L.stream>>>(P, A.pos, S, A.key, A.rank, A.cellCount); }
    const uint32_t est = std::min<uint64_t>((uint64_t)cellCapacity, std::max<uint64_t>(4ull * cellEstimate, 1u << 16));
    const uint32_t st = std::max<uint32_t>(1u, std::min<uint32_t>(div_up(est, SCAN_TILE), (uint32_t)L.numSMs * 8u));
    { LaunchScope ls(L, KID_SCAN); k_scan_tiles<<<st, VFD_TPB, 0, L.stream>>>(S, A.cellCount, A.tileSums); }
    { LaunchScope ls(L, KID_SCAN); k_scan_tile_sums<<<1, 1024, 0, L.stream>>>(S, A.tileSums); }
    { LaunchScope ls(L, KID_SCAN); k_scan_apply<<<st, VFD_TPB, 0, L.stream>>>(P, S, A.cellCount, A.tileSums, A.cellBegin); }
    { LaunchScope ls(L, KID_SCAN); k_compact_tiles<<<1, PARTITION_THREADS, 0, L.stream>>>(P, S, A.cellBegin, A.tileList); }
    { LaunchScope ls(L, KID_SCATTER); k_scatter<<<nb, VFD_TPB, 0, L.stream>>>(P, A.key, A.rank, A.cellBegin, A.tmpIdx); }
    { LaunchScope ls(L, KID_REORDER); k_reorder<<<nb, VFD_TPB, 0, L.stream>>>(P, A); }
    std::swap(A.pos, A.pos2); std::swap(A.vel, A.vel2); std::swap(A.dv, A.dv2); std::swap(A.nbar, A.nbar2);
    std::swap(A.curv, A.curv2); std::swap(A.curvS, A.curvS2); std::swap(A.curvD, A.curvD2); std::swap(A.id, A.id2);
    {
        LaunchScope ls(L, KID_BUILD_LIST);
        const size_t smem = tile_smem_bytes<0, 1>(STAGE_CAP);
        static thread_local int dev = -1;
        if (launch_device_changed(dev)) {
            cudaFuncSetAttribute(k_build_list<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_build_list<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        const uint32_t grid = (uint32_t)L.numSMs * 3u;
        if (P.searchFma) k_build_list<true><<<grid, TT_PLAIN, smem, L.stream>>>(P, A, S);
        else             k_build_list<false><<<grid, TT_PLAIN, smem, L.stream>>>(P, A, S);
    }
}

} // namespace vfd
