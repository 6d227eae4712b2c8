// tile.cuh — the tile pass: how every neighbour-sum kernel of the solver walks the particles.
//
// The search (search.cu) sorts particles by a tile-major cell key: the grid is cut into tiles of
// 4x4x4 cells, tiles are ordered x-slowest, and the 64 cells of a tile are contiguous (x-fastest
// inside the tile).  A tile's particles therefore form ONE contiguous range of every SoA array,
// and every neighbour of a tile's particle lies in the 6x6x6-cell "halo box" around it.
//
// A tile pass is executed by persistent CTAs (one per SM) as an asynchronous pipeline (pipe_pass below):
//   1. a scout warp draws tiles from a queue, reads the 216 halo cells' particle ranges from the cell table and prefix-sums
//      them, which defines a tile-LOCAL index space (box order: hz, hy, hx; see "the tile-local index space" below);
//   2. copy warps stage the payload the pass gathers per neighbour (position, plus kappa / velocity / pressure acceleration /
//      PCG direction ...) ONCE for all halo particles from HBM/L2 into a shared-memory ring with bulk (TMA) copies, ahead of
//      the consumers;
//   3. each consumer thread owns one particle of a 32-particle batch and streams its neighbour list — 16-bit tile-local
//      indices in a warp-blocked ELL layout, so a warp reads one 256-B block per four neighbour slots, groups ahead of their
//      use — and the per-pair coefficient stream next to it, gathering payloads from shared memory instead of through L1/L2.
// Per pass and particle HBM sees: own fields once + 2 B (index) + 4 B (coefficient) per neighbour; neighbour fields never.
//
// If a halo box holds more particles than PIPE_CAP (pathological clumping), the pass falls back to translating local
// indices to global ones (binary search in the 217-entry table) and gathers from global memory: slow, but exact.
#pragma once
#include "solver.h"

namespace vfd {

#define TILE_THREADS_MAX 1024   // CTA sizes are chosen per kernel (shared-memory and register budget differ); the drivers use blockDim
#define TILE_WARPS (blockDim.x >> 5)
#define HALO_CELLS 216
#define TILE_CELLS 64
// A tile holds ~512 particles at rest density (64 cells x 8), so 512 threads give one particle per thread and pass.
// Shared memory per SM is 228 KB; two resident CTAs (one staging while the other computes) need <= 113 KB each.
#define TT_LUT 512             // kernels holding one 40-kB lookup table
#define TT_LUT2 768            // density pass: two tables (80 kB), one CTA per SM
#define TT_MATVEC 512          // PCG mat-vec (no table)
#define TT_PLAIN 512
#define STAGE_CAP 2688         // staged halo particles per tile, kernels without a table (one or two float4 arrays: 38 / 76 KB)
#define STAGE_CAP_LUT 2240     // ... kernels with one table and two payload arrays: 40 + 70 + 3 KB, two CTAs per SM

// ---- grid / key helpers ---------------------------------------------------------------------------
// cell slightly larger than h so that two particles closer than h can never be two cells apart
// through fp32 rounding of the cell coordinate
__device__ __forceinline__ float cell_inv(float h) { return (1.0f / h) * (1.0f - 1.0f / 1024.0f); }

__device__ __forceinline__ uint3 cell_of(float4 x, const DevState* S, float invCell) {
    // global cell (identical on every rank that holds a copy of the particle), then the local grid's offset
    int cx = (int)((x.x - S->gridOrigin[0]) * invCell) - S->cellOffset[0];
    int cy = (int)((x.y - S->gridOrigin[1]) * invCell) - S->cellOffset[1];
    int cz = (int)((x.z - S->gridOrigin[2]) * invCell) - S->cellOffset[2];
    // robustness against NaN / escaped particles: clamp into the padded interior
    uint3 c;
    c.x = (uint32_t)min(max(cx, 1), (int)S->gridDim[0] - 2);
    c.y = (uint32_t)min(max(cy, 1), (int)S->gridDim[1] - 2);
    c.z = (uint32_t)min(max(cz, 1), (int)S->gridDim[2] - 2);
    return c;
}

// tile-major key: tiles x-slowest (a slab of tile columns is one contiguous particle range), cells x-fastest inside
__device__ __forceinline__ uint32_t cell_key(uint32_t cx, uint32_t cy, uint32_t cz, const DevState* S) {
    const uint32_t tile = ((cx >> 2) * S->tileDim[1] + (cy >> 2)) * S->tileDim[2] + (cz >> 2);
    return tile * TILE_CELLS + (((cz & 3u) << 4) | ((cy & 3u) << 2) | (cx & 3u));
}

// ---- shared-memory header of a tile pass ----------------------------------------------------------
// ---- the tile-local index space ------------------------------------------------------------------
// A halo row (six x-adjacent cells) is three contiguous runs of the sorted arrays: the cell of the tile on the left
// (hx = 0), the tile's own four cells (hx = 1..4), the cell of the tile on the right (hx = 5) — 108 segments per box.
// Local indices number the box's particles in box order (hz, hy, hx) with every segment placed so that its particles'
// local index has the parity of their global index, and padded to an even length: a segment [g, g + n) occupies the
// slots [l - (g & 1), l + n + ((g + n) & 1)) where l is its first particle's local index.  8-byte payload arrays can then
// be copied segment by segment in aligned 16-byte granules (bulk copies need 16-byte alignment on both sides and in the
// size) without touching a neighbouring segment's slots; the price is one spare slot per segment on average (~5 %).
__device__ __forceinline__ uint32_t seg_pre(int hx, uint32_t beg) { return (hx == 0 || hx == 1 || hx == 5) ? (beg & 1u) : 0u; }
__device__ __forceinline__ uint32_t seg_post(int hx, uint32_t end) { return (hx == 0 || hx == 4 || hx == 5) ? (end & 1u) : 0u; }

struct TileShared {
    uint32_t cellG[HALO_CELLS];       // global index of the first particle of each halo cell
    uint32_t cellE[HALO_CELLS];       // ... one past its last particle
    uint32_t local[HALO_CELLS + 8];   // local index of the first particle of each halo cell (see above); [HALO_CELLS] = size of the index space
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
    uint32_t cnt = 0, beg = 0, pre = 0;     // cnt: the cell's slots (its particles + the segment padding it carries)
    if (tid < HALO_CELLS) {
        const int hx = tid % 6, hy = (tid / 6) % 6, hz = tid / 36;
        const int cx = (int)(tx << 2) + hx - 1, cy = (int)(ty << 2) + hy - 1, cz = (int)(tz << 2) + hz - 1;
        uint32_t end = 0;
        if (cx >= 0 && cy >= 0 && cz >= 0 && cx < (int)S->gridDim[0] && cy < (int)S->gridDim[1] && cz < (int)S->gridDim[2]) {
            const uint32_t key = cell_key((uint32_t)cx, (uint32_t)cy, (uint32_t)cz, S);
            beg = __ldg(cellBegin + key);
            end = __ldg(cellBegin + key + 1);
        }
        pre = seg_pre(hx, beg);
        cnt = (end - beg) + pre + seg_post(hx, end);
        sh.cellG[tid] = beg; sh.cellE[tid] = end;
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
        sh.local[tid] = base + inc - cnt + pre;
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

// local -> global index: binary search in the prefix table
__device__ __forceinline__ uint32_t tile_local_to_global(const TileShared& sh, uint32_t L) {
    int lo = 0, hi = HALO_CELLS;          // find c with local[c] <= L < local[c+1]
    #pragma unroll
    for (int it = 0; it < 8; it++) { const int mid = (lo + hi) >> 1; if (hi - lo > 1) { if (sh.local[mid] <= L) lo = mid; else hi = mid; } }
    return sh.cellG[lo] + (L - sh.local[lo]);
}

// the same for a slot that may be segment padding: false (and no particle) for those
__device__ __forceinline__ bool tile_slot_to_global(const TileShared& sh, uint32_t L, uint32_t& g) {
    int lo = 0, hi = HALO_CELLS;
    #pragma unroll
    for (int it = 0; it < 8; it++) { const int mid = (lo + hi) >> 1; if (hi - lo > 1) { if (sh.local[mid] <= L) lo = mid; else hi = mid; } }
    g = sh.cellG[lo] + (L - sh.local[lo]);
    return L >= sh.local[lo] && g < sh.cellE[lo];
}

// Step 2: stage the payload of every halo particle.  Flat over the local index space (consecutive threads read
// consecutive particles of a cell: coalesced), four items per thread with all global loads issued before the first
// shared-memory store, so a tile pays one round of L2 latency instead of one per cell.
// The payload is kept as one or two float4 arrays (SoA).
template<int NPAY, class Op>
__device__ __forceinline__ void tile_stage(const TileShared& sh, uint32_t total, float4* __restrict__ sA, float4* __restrict__ sB, const Op& op) {
    constexpr int U = 4;
    for (uint32_t base = threadIdx.x; base < total; base += blockDim.x * U) {
        // items past the end are clamped to the last one (a duplicate store of the same value): no predicates, so the
        // payload registers stay registers
        // padding slots receive a position no particle is near (ops that walk whole local ranges: the neighbour search)
        uint32_t l[U], g[U];
        bool real[U];
        float4 a[U], b[U];
        #pragma unroll
        for (int u = 0; u < U; u++) { l[u] = min(base + u * blockDim.x, total - 1u); real[u] = tile_slot_to_global(sh, l[u], g[u]); }
        #pragma unroll
        for (int u = 0; u < U; u++) {
            a[u] = real[u] ? op.loadA(g[u]) : make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.0f);
            if (NPAY > 1) b[u] = real[u] ? op.loadB(g[u]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        #pragma unroll
        for (int u = 0; u < U; u++) { sA[l[u]] = a[u]; if (NPAY > 1) sB[l[u]] = b[u]; }
    }
}

// ---- neighbour list layout ------------------------------------------------------------------------
// u16 tile-local indices in a warp-blocked ELL of 4-slot groups: slots 4g..4g+3 of particle p are the four u16 of the
// 8-byte word  list16[((p>>5)*ELL_GROUPS + g)*32 + (p&31)]  (unused slots of the last group hold 0); the per-pair
// viscosity coefficient uses the same layout with 16-byte words.  One thread per particle: a warp reads one 256-B
// (list) / 512-B (coefficients) contiguous block per four neighbours with a single LDG.64 / LDG.128 per lane.
#define ELL_GROUPS ((VFD_MAX_NEIGHBORS + 3) / 4)        // 18
#define ELL_SLOTS (ELL_GROUPS * 4)                      // 72 slots allocated per particle
__device__ __forceinline__ size_t ell_base(uint32_t p) { return (size_t)(p >> 5) * (ELL_GROUPS * 32) + (p & 31); }   // in groups; group g at + g*32
__device__ __forceinline__ const uint2* ell_list(const uint16_t* list16, uint32_t p) { return reinterpret_cast<const uint2*>(list16) + ell_base(p); }
__device__ __forceinline__ void ell_unpack(uint2 w, uint32_t (&L)[4]) { L[0] = w.x & 0xffffu; L[1] = w.x >> 16; L[2] = w.y & 0xffffu; L[3] = w.y >> 16; }
__device__ __forceinline__ uint2 ell_pack(const uint32_t (&L)[4]) { return make_uint2(L[0] | (L[1] << 16), L[2] | (L[3] << 16)); }

// staged / fallback gather of payload array A (for ops that run their own loops: search, classifier)
template<bool STAGED, class Op> struct TileAcc {
    const TileShared& sh; const float4* sA; const Op& op;
    __device__ __forceinline__ float4 operator()(uint32_t L) const {
        if (STAGED) return sA[L];
        uint32_t g;                              // a padding slot reads as a position no particle is near, as in the stage
        return tile_slot_to_global(sh, L, g) ? op.loadA(g) : make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.0f);
    }
};

// ---- the barrier-phased pass driver ------------------------------------------------------------------
// setup | stage | compute as phases of one CTA separated by __syncthreads: what every neighbour pass used before the
// asynchronous pipeline below (profiles/r01_ncu_full_step_before_pipeline.txt).  The neighbour-list build still runs on
// it — that kernel is bound by instruction issue, not by latency, and wants the 48 warps per SM it gets here.
// Ops are custom: NPAY == 1, float4 loadA(g), void particle(p, valid, acc) with acc(L) -> float4, called by all 32 lanes.
template<class Op, bool STAGED>
__device__ __forceinline__ void tile_particles(const TileInfo& t, const TileShared& sh, const Arrays& A,
                                               const float4* __restrict__ sA, const float4* __restrict__ sB, Op& op) {
    static_assert(Op::CUSTOM, "pair ops run on the pipeline (pipe_pass)");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nBatch = (t.end - t.begin + 31u) >> 5;
    for (uint32_t b = warp; b < nBatch; b += TILE_WARPS) {
        const uint32_t p = t.begin + (b << 5) + lane;
        // lanes past the tile's end pass valid = false
        const TileAcc<STAGED, Op> acc{ sh, sA, op };
        op.particle(p, p < t.end, acc);
    }
}

template<class Op>
__device__ __forceinline__ void tile_pass(DevState* __restrict__ S, const Arrays& A, TileShared& sh,
                                          float4* sA, float4* sB, uint32_t cap, Op& op, uint32_t tile0, uint32_t tile1, bool checkIndexRange = false) {
    const uint32_t* __restrict__ cellBegin = A.cellBegin;
    // the non-empty owned tiles (search.cu: k_compact_tiles), entry i to CTA i mod G
    const uint32_t nList = __ldg(A.tileList);
    for (uint32_t li = blockIdx.x; li < nList; li += gridDim.x) {
        const uint32_t tile = __ldg(A.tileList + 1u + li);
        __syncthreads();                                   // previous tile's readers are done with shared memory
        const TileInfo t = tile_setup(S, cellBegin, tile, sh, cap);
        // local indices are 16-bit: a halo box beyond 65535 particles cannot be encoded (flagged, caught by the host)
        if (checkIndexRange && t.total > 65535u && threadIdx.x == 0) atomicOr(&S->errorFlags, 2u);
        if (!t.staged && threadIdx.x == 0) atomicAdd(&S->fallbackTiles, 1u);
        if (t.staged) {
            tile_stage<Op::NPAY>(sh, t.total, sA, sB, op);
            __syncthreads();
            tile_particles<Op, true>(t, sh, A, sA, sB, op);
        } else {
            tile_particles<Op, false>(t, sh, A, sA, sB, op);
        }
    }
}

// =====================================================================================================
// The asynchronous tile pipeline (pipe_pass): the same tile pass without a CTA-wide barrier in the loop.
//
// tile_pass above runs  setup | stage | compute  as phases separated by __syncthreads; ncu shows the price
// (profiles/r01_ncu_matvec_before_pipeline.txt: 36 % of the warp-stall samples at barriers, the pair loop only a third of the
// kernel).  pipe_pass splits the CTA into roles that meet only at shared-memory mbarriers:
//   * one SCOUT warp draws tiles from the queue and builds their cell tables (`table` mbarrier per stage);
//   * PIPE_COPY_WARPS copy warps claim ring space and stage the payload HBM/L2 -> shared memory with bulk (TMA) copies, one
//     per contiguous segment and array (a halo row of six x-adjacent cells is three contiguous global segments), tracked by
//     the stage's `full` mbarrier (complete_tx);
//   * the consumer warps take 32-particle batches by ticket, gather from the stage's shared memory, and release the stage
//     through its `empty` mbarrier.
// A warp that finishes its batch takes the CTA's oldest batch not yet started, in the same tile or the next; staging of later
// tiles overlaps the pair loops.  One CTA per SM.  Which warp computes which batch does not matter for the result: every
// batch's partial sums go to the batch's own slot of its tile's record, folded in a fixed order (bit-reproducible).
// (profiles/r02_pipeline_trace.md has the timeline of one CTA before and after these roles were introduced.)
//
// Op interface of a pipelined pass.  All ops:  NPAY (1|2), BBYTES (16: payload B is a float4 array, 4: a float array),
//   const float4* srcA(), const void* srcB()      global payload arrays the producers copy from
//   float4 loadA(g), loadB(g)                     the same payload for the exact fallback path
// Pair ops (CUSTOM == false): NOWN, NSUM, COEF (bit 0: a per-pair coefficient stream is read, bit 1: one is written;
//   both in the list's ELL layout), const float* coef_in(), float* coef_out(), load_own, pair(own, a, b, coef&, acc),
//   finish(p, m, own, sum) — as for tile_pass.  Custom ops: particle(p, valid, acc, H), called by all 32 lanes.
// PIPE_ABLATE (tuning builds only, results are wrong): bit 0: the list/coefficient stream re-reads its first group (no
// HBM streaming), bit 1: no shared-memory gather, bit 2: the producers copy nothing, bit 3: no epilogue
#ifndef PIPE_ABLATE
#define PIPE_ABLATE 0
#endif
#ifndef PIPE_STAGES
#define PIPE_STAGES 6            // tiles in flight per CTA (header slots); their payloads share one ring of particle slots
#endif
#ifndef PIPE_CONSUMER_WARPS
#define PIPE_CONSUMER_WARPS 28
#endif
#ifndef PIPE_PRODUCER_WARPS
#define PIPE_PRODUCER_WARPS 4
#endif
#define PIPE_PRODUCER_THREADS (PIPE_PRODUCER_WARPS * 32)
#define PIPE_CAP 2816          // largest halo box that is staged (larger ones take the exact global-memory path)
// The payload ring takes whatever shared memory the pass has left (pipe_ring_slots below): how many tiles a CTA has in
// flight — staged or being gathered — is what bounds its throughput once the consumers no longer wait for memory (a batch
// lives ~10 us in a warp, a tile is released when its slowest batch is done): 6700 particle slots for the 32-byte payloads,
// 8900 for the mat-vec's 24 bytes, ~13000 for one 16-byte array; a typical halo box holds 1700 - 2000 particles.
#define PIPE_SMEM_MAX 232448u  // 227 KB of dynamic shared memory per CTA on sm_100
#ifndef PIPE_GATHER_WIDTH
#define PIPE_GATHER_WIDTH 2    // payload gathers in flight per lane (1, 2 or 4)
#endif
#ifndef PIPE_LOOKAHEAD
#define PIPE_LOOKAHEAD 2
#endif                         // neighbour groups (of four) in flight per warp

// Shape of a pass, chosen per kernel (every Op names one as Op::Cfg): consumer warps per CTA, neighbour groups in flight
// per warp (slots of its shared-memory stream ring), payload gathers in flight per lane.  Measured on the settled 1M scene
// (round 2): with the list/coefficient stream in shared memory every pair pass fits 64 registers, and 28 consumer warps
// with a two-slot ring do best — deeper rings take shared memory from the payload ring (tiles in flight), which costs more
// than the extra look-ahead gains (ring 2/3/4/6: 0.116/0.120/0.122/0.144 ms per mat-vec).
template<int CW_, int D_, int GW_> struct PipeCfg {
    static constexpr int CW = CW_, D = D_, GW = GW_, THREADS = (CW_ + PIPE_PRODUCER_WARPS) * 32;
};
using PipeCfgWide = PipeCfg<PIPE_CONSUMER_WARPS, PIPE_LOOKAHEAD, PIPE_GATHER_WIDTH>;   // two 16-byte payloads (source terms, Jacobi updates)
#ifndef PIPE_MANY_CW
#define PIPE_MANY_CW 28
#define PIPE_MANY_D 2
#define PIPE_MANY_GW 2
#endif
using PipeCfgMany = PipeCfg<PIPE_MANY_CW, PIPE_MANY_D, PIPE_MANY_GW>;
#ifndef PIPE_MV_CW
#define PIPE_MV_CW 24
#define PIPE_MV_D 2
#define PIPE_MV_GW 2
#endif
using PipeCfgMatvec = PipeCfg<PIPE_MV_CW, PIPE_MV_D, PIPE_MV_GW>;    // the PCG mat-vec (payload 16 + 8 bytes)

// PIPE_TRACE (diagnostic builds): CTA 0 logs (SM clock, event) pairs of its warps' lane 0 into a per-translation-unit buffer
#ifdef PIPE_TRACE
#define PIPE_TRACE_CAP 65536u
static __device__ unsigned long long g_pipeTrace[2 * PIPE_TRACE_CAP];
static __device__ unsigned int g_pipeTraceN, g_pipeTraceOn;
__device__ __forceinline__ void ptrace(uint32_t code, uint32_t a, uint32_t b) {
    if (blockIdx.x != 0 || (threadIdx.x & 31u) != 0u || !g_pipeTraceOn) return;
    const unsigned int i = atomicAdd(&g_pipeTraceN, 1u);
    if (i >= PIPE_TRACE_CAP) return;
    g_pipeTrace[2 * i] = (unsigned long long)clock64();
    g_pipeTrace[2 * i + 1] = ((unsigned long long)code << 56) | ((unsigned long long)(threadIdx.x >> 5) << 48) | ((unsigned long long)(a & 0xffffffu) << 24) | (unsigned long long)(b & 0xffffffu);
}
#else
#define ptrace(code, a, b) do { } while (0)
#endif

__device__ __forceinline__ uint2 ldg_stream(const uint2* p) { return __ldg(p); }
__device__ __forceinline__ float4 ldg_stream(const float4* p) { return __ldg(p); }

struct StageHeader {
    uint32_t cellG[HALO_CELLS];       // as TileShared
    uint32_t local[HALO_CELLS + 8];
    uint32_t begin, end, total, staged;      // begin == 0xffffffff: no more tiles for this CTA
    uint32_t base, li, pad[2];               // first ring slot of this tile's payload; the tile's entry in the tile list
    uint32_t cellE[HALO_CELLS];              // one past the last particle of each halo cell in the sorted arrays (for the copy warps)
};
// Grid-wide sums of a dynamically scheduled pass.  Every batch's warp-level sum goes into the record of its tile
// (bsum[.][batch]); the warp that completes a tile's record folds it in batch order and stores ONE value per tile in
// the global slot of the tile's list entry; the CTA that finishes the pass last folds the tile slots in list order
// (common.cuh: fold_slots).  Every order is fixed, so the result does not depend on which CTA or warp did what.
// Records are reused every PIPE_RED_RECORDS tiles of the CTA: by the time tile k + PIPE_STAGES can be entered every warp
// has finished the epilogues of its batches up to tile k - 1, so 2 * PIPE_STAGES records are plenty.
#define PIPE_MAX_BATCH 64      // a tile holds ~16 batches; beyond 63 (4x rest density) the surplus shares the last entry atomically
#define PIPE_RED_RECORDS (2 * PIPE_STAGES)
struct RedRecord {
    double bsum[2][PIPE_MAX_BATCH];
    uint32_t count, pad[3];
};
// Custom ops with Op::TILE_QUEUE: the same storage, seen as the tile's work queue — items one warp found (for the surface
// classifier: the tile's surface particles) that ANY consumer warp of the CTA may take, so the warps whose batches hold no
// such item help the ones whose batches are full of them instead of running ahead and stalling on the ring.
// Entries are 16-bit offsets from the tile's first particle; PIPE_QUEUE_NONE marks a slot that is not written yet.
#define PIPE_QUEUE_CAP 504
#define PIPE_QUEUE_NONE 0xffffu
struct TileQueue {
    uint32_t head, tail, left, pad;
    unsigned short e[PIPE_QUEUE_CAP];
};
static_assert(sizeof(TileQueue) <= sizeof(RedRecord), "the queue overlays a reduction record");

struct PipeShared {
    unsigned long long full[PIPE_STAGES], empty[PIPE_STAGES];
    unsigned long long table[PIPE_STAGES];   // "the cell table of this stage's tile is complete" (scout -> copy warps)
    uint32_t lastCta, ticket, pad_[2];       // "this CTA finished last"; batches handed out so far (pair passes)
    double red[4 * 32];
    RedRecord rec[PIPE_RED_RECORDS];
    StageHeader hdr[PIPE_STAGES];
};
__host__ __device__ constexpr size_t pipe_header_bytes() { return (sizeof(PipeShared) + 127) / 128 * 128; }
// dynamic shared memory: [PipeShared][per-warp neighbour-stream rings (pair ops)][NLUT tables][payload A ring][payload B ring]
// The stream ring of a consumer warp: Cfg::D slots of one neighbour group each — the 32 lanes' list words (256 B) and, for ops
// that read a per-pair coefficient stream, their coefficient words (512 B) — filled with cp.async by the lanes themselves.
template<class Op> struct PipeLayout {
    static constexpr bool STREAM = !Op::CUSTOM;
    static constexpr uint32_t SLOT = STREAM ? 256u + ((Op::COEF & 1) ? 512u : 0u) : 0u;           // bytes per group
    static constexpr uint32_t STREAM_BYTES = STREAM ? (uint32_t)Op::Cfg::CW * (uint32_t)Op::Cfg::D * SLOT : 0u;
    static constexpr uint32_t BB = Op::NPAY > 1 ? (uint32_t)Op::BBYTES : 0u;
    static constexpr uint32_t LUT_BYTES = (uint32_t)Op::NLUT * (uint32_t)(VFD_LUT_RES * sizeof(float));
    static constexpr uint32_t OFF_STREAM = (uint32_t)pipe_header_bytes();
    static constexpr uint32_t OFF_LUT = OFF_STREAM + STREAM_BYTES;
    static constexpr uint32_t OFF_PAY = OFF_LUT + LUT_BYTES;
    // particle slots of the payload ring: whatever shared memory is left
    static constexpr uint32_t SLOTS = ((PIPE_SMEM_MAX - OFF_PAY) / (16u + BB)) & ~63u;
    static constexpr size_t BYTES = (size_t)OFF_PAY + (size_t)SLOTS * (16u + BB);
    static_assert(SLOTS >= PIPE_CAP, "the payload ring must hold the largest staged halo box");
};
template<class Op> struct PipeRing { static constexpr uint32_t SLOTS = PipeLayout<Op>::SLOTS; };
template<class Op> static inline size_t pipe_smem_bytes() { return PipeLayout<Op>::BYTES; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
// A waiting warp must not eat the issue slots of the working ones (ncu of round 1's mat-vec: a quarter of all executed warp
// instructions were these polls): after a failed try it sleeps before the next one (PIPE_WAIT_NS, 0 = poll flat out).
#ifndef PIPE_COEF_CA
#define PIPE_COEF_CA 0
#endif
#ifndef PIPE_WAIT_NS
#define PIPE_WAIT_NS 100
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (done) break;
        if (PIPE_WAIT_NS) __nanosleep(PIPE_WAIT_NS);
    }
}
__device__ __forceinline__ bool mbar_test(unsigned long long* b, uint32_t parity) {     // non-blocking
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
// cp.async of the lane's own word with zero fill: srcBytes = 0 writes zeros without touching global memory
__device__ __forceinline__ void cp_async8_zfill(uint32_t dstShared, const void* src, uint32_t srcBytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dstShared), "l"(src), "r"(srcBytes) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dstShared, const void* src, uint32_t srcBytes) {
#if PIPE_COEF_CA
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(dstShared), "l"(src), "r"(srcBytes) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dstShared), "l"(src), "r"(srcBytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float4 lds128(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
// the mbarrier receives one (pre-counted) arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive(unsigned long long* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(b)) : "memory");
}
// bulk (TMA) copy global -> shared of `bytes` (a multiple of 16, both addresses 16-byte aligned), completing on the mbarrier
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
// fire-and-forget prefetch of a contiguous byte range into L2 (16-byte granularity; the range is widened to it)
__device__ __forceinline__ void l2_prefetch(const void* base, size_t byteBegin, size_t byteEnd) {
    const size_t a = byteBegin & ~(size_t)15, e = (byteEnd + 15) & ~(size_t)15;
    if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(reinterpret_cast<const unsigned char*>(base) + a), "r"((uint32_t)(e - a)) : "memory");
}

__device__ __forceinline__ uint32_t hdr_local_to_global(const StageHeader& H, uint32_t L) {
    int lo = 0, hi = HALO_CELLS;
    #pragma unroll
    for (int it = 0; it < 8; it++) { const int mid = (lo + hi) >> 1; if (hi - lo > 1) { if (H.local[mid] <= L) lo = mid; else hi = mid; } }
    return H.cellG[lo] + (L - H.local[lo]);
}

// ---- producers: one scout warp and PIPE_COPY_WARPS copy warps ---------------------------------------------------
// Round 1 had all producer warps do everything for one tile at a time: queue draw -> tile id -> cell table (three dependent
// trips to L2) -> prefix sum -> ring space -> copies -> next tile.  A trace of the PCG mat-vec (PIPE_TRACE build,
// profiles/r02_pipeline_trace.md) showed that chain to be the whole kernel: 6.5 - 7.5 us per tile, one tile after the other,
// never waiting for the consumers, while the consumers went through every tile the instant it arrived (20 % of their time
// waiting for it, every batch started cold).  The chain is now split by role:
//   * the SCOUT warp draws tiles from the queue and builds their tables, one tile ahead of the copy warps, and publishes
//     each through the stage's `table` mbarrier;
//   * the COPY warps take the tables in order, claim ring space (the same bookkeeping in every copy thread, no barrier
//     between them), and issue the copies: their loop is nothing but address arithmetic and LDGSTS / bulk copies.
// Op supplies the global source arrays:  const float4* srcA();  const void* srcB()  (BBYTES 16: float4 array, 8 / 4: float2 / float array)
#ifndef PIPE_SCOUT_LAG
#define PIPE_SCOUT_LAG 2u         // the scout draws tile k once tile k - PIPE_SCOUT_LAG is staged
#endif
#define PIPE_COPY_WARPS (PIPE_PRODUCER_WARPS - 1)
#define PIPE_COPY_THREADS (PIPE_COPY_WARPS * 32)
#define PIPE_FULL_ARRIVALS (PIPE_COPY_THREADS + 1u)   // every copy thread's copies have landed (cp.async.mbarrier.arrive.noinc) + the header (release)
template<class Op>
__device__ __forceinline__ void pipe_scout(DevState* __restrict__ S, const Arrays& A, PipeShared& ps, bool checkIndexRange) {
    constexpr int CPL = (HALO_CELLS + 31) / 32;          // cells per lane: lane l owns the box cells [7 l, 7 l + 7)
    constexpr uint32_t NONE = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t* __restrict__ cellBegin = A.cellBegin;
    const uint32_t* __restrict__ tileList = A.tileList;
    const uint32_t tdy = S->tileDim[1], tdz = S->tileDim[2];
    const int gdx = (int)S->gridDim[0], gdy = (int)S->gridDim[1], gdz = (int)S->gridDim[2];
    const uint32_t nList = __ldg(tileList);
    // Tiles come from a queue: the list of non-empty owned tiles (search.cu: k_compact_tiles) is handed out entry by entry
    // through an atomic cursor, so a CTA that drew cheap tiles (surface, boundary) simply draws more of them, and CTAs
    // that run side by side still work on neighbouring tiles and share their halo boxes in L2.  Reductions do not depend
    // on the schedule (common.cuh: fold_slots).
    auto draw = [&]() -> uint32_t {
        uint32_t v = 0;
        if (lane == 0) v = atomicAdd(&S->tileCursor, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };
    uint32_t b0 = 0, e0 = 0, beg[CPL], fin[CPL];          // per cell: first particle, one past the last
    for (uint32_t k = 0;; k++) {
        const uint32_t s = k % PIPE_STAGES, u = k / PIPE_STAGES;
        StageHeader& H = ps.hdr[s];
        // A tile is drawn when the tile two before it has been staged: late enough that a CTA never sits on a queue of tiles
        // while others run dry at the end of the pass (the draw, the tile id and the cell table are three dependent trips to
        // L2, ~2 us — a fraction of what the consumers need for a tile), early enough that the consumers find the table of
        // their next batch's tile when they look ahead.
        if (k >= PIPE_SCOUT_LAG) mbar_wait(&ps.full[(k - PIPE_SCOUT_LAG) % PIPE_STAGES], ((k - PIPE_SCOUT_LAG) / PIPE_STAGES) & 1u);
        const uint32_t li = draw();
        const uint32_t tile = li < nList ? __ldg(tileList + 1u + li) : NONE;
        const bool end = tile == NONE;
        #pragma unroll
        for (int i = 0; i < CPL; i++) { beg[i] = 0; fin[i] = 0; }
        if (!end) {
            b0 = __ldg(cellBegin + tile * TILE_CELLS); e0 = __ldg(cellBegin + tile * TILE_CELLS + TILE_CELLS);
            const uint32_t tz = tile % tdz, ty = (tile / tdz) % tdy, tx = tile / (tdz * tdy);
            #pragma unroll
            for (int i = 0; i < CPL; i++) {
                const int c = (int)lane * CPL + i;
                if (c < HALO_CELLS) {
                    const int hx = c % 6, hy = (c / 6) % 6, hz = c / 36;
                    const int cx = (int)(tx << 2) + hx - 1, cy = (int)(ty << 2) + hy - 1, cz = (int)(tz << 2) + hz - 1;
                    if (cx >= 0 && cy >= 0 && cz >= 0 && cx < gdx && cy < gdy && cz < gdz) {
                        const uint32_t key = cell_key((uint32_t)cx, (uint32_t)cy, (uint32_t)cz, S);
                        beg[i] = __ldg(cellBegin + key);
                        fin[i] = __ldg(cellBegin + key + 1);
                    }
                }
            }
        }
        const uint32_t tb = b0, te = e0, tli = li;
        // slots of each cell: its particles plus the padding of the segment it begins / ends (see "the tile-local index space")
        uint32_t pre[CPL], slots[CPL], mine = 0;
        #pragma unroll
        for (int i = 0; i < CPL; i++) {
            const int hx = ((int)lane * CPL + i) % 6;
            pre[i] = seg_pre(hx, beg[i]);
            slots[i] = (fin[i] - beg[i]) + pre[i] + seg_post(hx, fin[i]);
            mine += slots[i];
        }
        uint32_t inc = mine;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += y; }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        ptrace(1, k, total);
        mbar_wait(&ps.empty[s], (u & 1u) ^ 1u);          // every consumer warp has released the slot's previous tile
        ptrace(2, k, 0);
        uint32_t run = inc - mine;
        bool real0 = false;                              // does a particle sit in local slot 0?
        #pragma unroll
        for (int i = 0; i < CPL; i++) {
            const int c = (int)lane * CPL + i;
            if (c < HALO_CELLS) { H.cellG[c] = beg[i]; H.local[c] = run + pre[i]; H.cellE[c] = fin[i]; }
            real0 = real0 || (run + pre[i] == 0u && fin[i] > beg[i]);
            run += slots[i];
        }
        const bool anyReal0 = __any_sync(0xffffffffu, real0);
        if (lane == 0) {
            // Unused slots of a particle's last neighbour group hold index 0, and the pair passes gather it with a zero
            // coefficient ("a valid read"): if slot 0 is segment padding nothing is copied there, and whatever an earlier
            // kernel left in shared memory (an Inf, a NaN pattern) would turn 0 x garbage into NaN — the copy warps zero it
            H.pad[0] = anyReal0 ? 0u : 1u;
            H.local[HALO_CELLS] = total;
            H.begin = end ? NONE : tb; H.end = end ? NONE : te; H.total = end ? 0u : total; H.li = tli;
            if (!end && checkIndexRange && total > 65535u) atomicOr(&S->errorFlags, 2u);
            if (!end && total > PIPE_CAP) atomicAdd(&S->fallbackTiles, 1u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ps.table[s]);        // release: the table of the CTA's k-th tile is complete
        ptrace(3, k, 0);
        if (end) break;
    }
}

template<class Op>
__device__ __forceinline__ void pipe_copier(PipeShared& ps, unsigned char* pay, const Op& op) {
    constexpr int BBYTES = Op::NPAY > 1 ? Op::BBYTES : 0;
    constexpr uint32_t NONE = 0xffffffffu;
    const uint32_t ct = threadIdx.x - (Op::Cfg::CW + 1) * 32, lane = ct & 31u, cw = ct >> 5;     // copy thread, its warp
    const unsigned char* __restrict__ gA = reinterpret_cast<const unsigned char*>(op.srcA());
    const unsigned char* __restrict__ gB = reinterpret_cast<const unsigned char*>(op.srcB());
    // ring bookkeeping (identical in every copy thread): regions of the tiles k-1 ... and the next free slot
    uint32_t regB[PIPE_STAGES - 1], regE[PIPE_STAGES - 1];
    #pragma unroll
    for (int j = 0; j < PIPE_STAGES - 1; j++) { regB[j] = 0; regE[j] = 0; }
    uint32_t ringNext = 0;
    for (uint32_t k = 0;; k++) {
        const uint32_t s = k % PIPE_STAGES, u = k / PIPE_STAGES;
        StageHeader& H = ps.hdr[s];
        mbar_wait(&ps.table[s], u & 1u);
        if (H.begin == NONE) {                            // the end marker goes straight through
            if (ct == 0) { H.staged = 0u; H.base = 0u; }
            mbar_arrive(&ps.full[s]);
            if (ct == 0) mbar_arrive(&ps.full[s]);
            break;
        }
        const uint32_t total = H.total;
        const bool staged = total <= PIPE_CAP;
        // ring space: the tile's payload takes `total` contiguous slots; tiles still in flight whose region it would
        // overlap are waited for, oldest first (consumers release tiles in order)
        uint32_t base = 0;
        if (staged) {
            base = ringNext + total <= PipeRing<Op>::SLOTS ? ringNext : 0u;
            #pragma unroll
            for (int j = PIPE_STAGES - 2; j >= 0; j--) {            // oldest first: j = 0 is tile k-1
                const uint32_t kk = k - 1u - (uint32_t)j;
                if (k >= 1u + (uint32_t)j && regE[j] > regB[j] && base < regE[j] && regB[j] < base + total)
                    mbar_wait(&ps.empty[kk % PIPE_STAGES], (kk / PIPE_STAGES) & 1u);
            }
            ringNext = base + total;
        }
        #pragma unroll
        for (int j = PIPE_STAGES - 2; j > 0; j--) { regB[j] = regB[j - 1]; regE[j] = regE[j - 1]; }
        regB[0] = base; regE[0] = staged ? base + total : base;
        if (ct == 0) { H.staged = staged ? 1u : 0u; H.base = base; }
        ptrace(4, k, base);
        if (staged && !(PIPE_ABLATE & 4)) {
            unsigned char* sA = pay + (size_t)base * 16;
            unsigned char* sB = pay + (size_t)PipeRing<Op>::SLOTS * 16 + (size_t)base * BBYTES;
            // One bulk (TMA) copy per segment and payload array, 108 segments over the copy threads: no LSU instruction per
            // particle, no per-sector shared-memory write.  16-byte arrays copy the exact range; the 8-byte array copies the
            // enclosing aligned 16-byte granules into the segment's (even, padded) slots.
            unsigned long long* fb = &ps.full[s];
            if (ct == 0 && H.pad[0]) {                    // slot 0 is padding: give it a finite value (see the scout)
                *reinterpret_cast<float4*>(sA) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (BBYTES == 16) *reinterpret_cast<float4*>(sB) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                // (an 8- or 4-byte array's copy may cover the slot as well, with the finite value of a real particle: either is fine)
                if (BBYTES == 8) *reinterpret_cast<float2*>(sB) = make_float2(0.0f, 0.0f);
                if (BBYTES == 4) *reinterpret_cast<float*>(sB) = 0.0f;
            }
            for (uint32_t seg = ct; seg < 108u; seg += PIPE_COPY_THREADS) {
                const uint32_t r = seg / 3u, q = seg - 3u * r, c0 = r * 6u;
                const uint32_t cb = q == 0u ? c0 : (q == 1u ? c0 + 1u : c0 + 5u), cl = q == 0u ? c0 : (q == 1u ? c0 + 4u : c0 + 5u);
                const uint32_t g = H.cellG[cb], n = H.cellE[cl] - g, lb = H.local[cb];
                if (n) {
                    const uint32_t pre_ = g & 1u, ext = n + pre_ + ((g + n) & 1u);
                    mbar_expect_tx(fb, n * 16u + (BBYTES == 16 ? n * 16u : 0u) + (BBYTES == 8 ? ext * 8u : 0u));
                    bulk_copy(sA + (size_t)lb * 16, gA + (size_t)g * 16, n * 16u, fb);
                    if (BBYTES == 16) bulk_copy(sB + (size_t)lb * 16, gB + (size_t)g * 16, n * 16u, fb);
                    if (BBYTES == 8) bulk_copy(sB + (size_t)(lb - pre_) * 8, gB + (size_t)(g - pre_) * 8, ext * 8u, fb);
                }
            }
            if (BBYTES == 4) {
                // a 4-byte array (one pass: the pressure acceleration's kappa) goes particle by particle, a halo row (three
                // segments, contiguous in the local index space up to their padding slots, which receive whatever follows the
                // segment in the sorted array) per warp and turn
                constexpr uint32_t ROWS = (36u + PIPE_COPY_WARPS - 1u) / PIPE_COPY_WARPS;
                #pragma unroll 2
                for (uint32_t j = 0; j < ROWS; j++) {
                    const uint32_t r = cw + j * PIPE_COPY_WARPS;
                    if (r < 36u) {
                        const uint32_t c0 = r * 6u;
                        const uint32_t lA = H.local[c0], lB = H.local[c0 + 1], lC = H.local[c0 + 5], lE = H.local[c0 + 6];
                        const uint32_t g0 = H.cellG[c0], g1 = H.cellG[c0 + 1], g5 = H.cellG[c0 + 5];
                        for (uint32_t l = lA + lane; l < lE; l += 32u) {
                            const uint32_t g = l < lB ? g0 + (l - lA) : (l < lC ? g1 + (l - lB) : g5 + (l - lC));
                            cp_async4(sB + (size_t)l * 4, gB + (size_t)g * 4);
                        }
                    }
                }
            }
        }
        cp_async_arrive(&ps.full[s]);                     // one arrival per copy thread: its copies have landed
        if (ct == 0) mbar_arrive(&ps.full[s]);            // + 1: the header (release)
        ptrace(5, k, 0);
    }
}

template<class Op> struct PipeAcc {
    const StageHeader& H; const float4* sA; const Op& op; bool staged;
    __device__ __forceinline__ float4 operator()(uint32_t L) const { return staged ? sA[L] : op.loadA(hdr_local_to_global(H, L)); }
};

template<class Op>
__device__ __forceinline__ void pipe_consumer(const Arrays& A, PipeShared& ps, const unsigned char* pay, Op& op) {
    constexpr int BBYTES = Op::NPAY > 1 ? Op::BBYTES : 0;
    const uint32_t lane = threadIdx.x & 31u, cw = threadIdx.x >> 5;
    uint32_t rot = 0;                                     // running batch count of this CTA (mod the consumer warps)
    for (uint32_t k = 0;; k++) {
        const uint32_t s = k % PIPE_STAGES, u = k / PIPE_STAGES;
        mbar_wait(&ps.full[s], u & 1u);
        const StageHeader& H = ps.hdr[s];
        const uint32_t begin = H.begin, end = H.end;
        if (begin == 0xffffffffu) break;
        const bool staged = H.staged != 0u;
        const uint32_t base = H.base;
        const float4* sA = reinterpret_cast<const float4*>(pay + (size_t)base * 16);
        const uint32_t nBatch = (end - begin + 31u) >> 5;
        // batch b of this tile goes to warp (rot + b) mod W: consecutive batches of consecutive tiles visit the warps in turn
        const PipeAcc<Op> acc{ H, sA, op, staged };
        if constexpr (Op::TILE_QUEUE) {
            TileQueue& Q = *reinterpret_cast<TileQueue*>(&ps.rec[k % PIPE_RED_RECORDS]);
            for (uint32_t b = (cw + Op::Cfg::CW - rot) % Op::Cfg::CW; b < nBatch; b += Op::Cfg::CW) {
                const uint32_t p = begin + (b << 5) + lane;
                // phase A: the op handles the particle or asks for it to be queued
                const bool want = op.particle(p, p < end, acc, H);
                const uint32_t mask = __ballot_sync(0xffffffffu, want);
                if (mask) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&Q.tail, (uint32_t)__popc(mask));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const uint32_t slot = base + __popc(mask & ((1u << lane) - 1u));
                    if (want && slot < PIPE_QUEUE_CAP) *reinterpret_cast<volatile unsigned short*>(&Q.e[slot]) = (unsigned short)(p - begin);
                    // a tile with more items than the queue holds (never at fluid densities): the finder does the surplus itself
                    uint32_t over = __ballot_sync(0xffffffffu, want && slot >= PIPE_QUEUE_CAP);
                    while (over) { const int src = __ffs(over) - 1; over &= over - 1u; op.item(__shfl_sync(0xffffffffu, p, src), acc, H); }
                }
            }
            // phase B: take items until the queue is empty (whoever queued them)
            for (;;) {
                uint32_t h = 0, got = 0;
                if (lane == 0) {
                    for (;;) {
                        h = *reinterpret_cast<volatile uint32_t*>(&Q.head);
                        const uint32_t t = min(*reinterpret_cast<volatile uint32_t*>(&Q.tail), (uint32_t)PIPE_QUEUE_CAP);
                        if (h >= t) break;
                        if (atomicCAS(&Q.head, h, h + 1u) == h) { got = 1; break; }
                    }
                }
                got = __shfl_sync(0xffffffffu, got, 0);
                if (!got) break;
                h = __shfl_sync(0xffffffffu, h, 0);
                uint32_t off = PIPE_QUEUE_NONE;
                if (lane == 0) {
                    volatile unsigned short* e = reinterpret_cast<volatile unsigned short*>(&Q.e[h]);
                    do { off = *e; } while (off == PIPE_QUEUE_NONE);     // reserved before it is written: a few cycles at most
                    *e = PIPE_QUEUE_NONE;
                }
                off = __shfl_sync(0xffffffffu, off, 0);
                op.item(begin + off, acc, H);
            }
            // the last warp to leave resets the counters for the record's next tile (the slots are PIPE_QUEUE_NONE again)
            if (lane == 0) {
                __threadfence_block();
                if (atomicAdd(&Q.left, 1u) == Op::Cfg::CW - 1u) { Q.head = 0u; Q.tail = 0u; Q.left = 0u; }
            }
        } else {
            for (uint32_t b = (cw + Op::Cfg::CW - rot) % Op::Cfg::CW; b < nBatch; b += Op::Cfg::CW) {
                const uint32_t p = begin + (b << 5) + lane;
                op.particle(p, p < end, acc, H);
            }
        }
        rot = (rot + nBatch) % Op::Cfg::CW;
        __syncwarp();
        if (lane == 0) mbar_arrive(&ps.empty[s]);
    }
}

// ---- consumer of pair ops: one continuous neighbour stream per warp ---------------------------------------
// A warp owns a sequence of 32-particle batches (one particle per lane), possibly spread over several tiles.  ncu of
// the round-1 consumer (profiles/r01_ncu_full_step.txt, source page) put 42 % of the consumer warps' stall samples OUTSIDE
// the gather loop: every batch started cold (neighbour count, own fields, first list and coefficient words: one exposed
// L2/HBM latency per batch, only partly hidden behind the previous epilogue at the price of a second register set and
// spills), and a warp that had finished a tile waited for the next one.  Here
//   * the list/coefficient register ring runs ACROSS batch boundaries: the last round of a batch refills the ring with the
//     first groups of the warp's next batch, whose particle index and neighbour count were fetched a whole batch earlier
//     (the batch cursor walks one batch ahead of the gather, without blocking: a tile that is not staged yet is not waited
//     for while the warp still holds an older one);
//   * a particle's own fields are read from the shared-memory stage (the tile's own cells are part of its halo box), not
//     from global memory: no long-latency load sits between two gather loops;
//   * the group loop has a warp-uniform trip count (lanes past their last group are predicated off) and, for ops whose pair
//     term vanishes with the coefficient (Op::PAD_SAFE: the unused slots of a particle's last group hold index 0 and
//     coefficient 0), no separate code path for partial groups.
// The order of a lane's pair terms is the list order, as before: results are bit-identical to the round-1 consumer up to
// the sign of a zero.
//
// Op interface (pair ops): NOWN, NSUM, NRED, COEF, PAD_SAFE; own_from(p, a, b, own) fills the lane's own fields from its
// own payload entry (a, b) — it may add plain global loads of fields the epilogue wants; pair(own, a, b, coef&, acc);
// finish(p, countWord, own, acc).

// local index of one of the tile's OWN particles: the tile's 16 rows of four x-adjacent cells (row r = 4 z + y, box cell
// 36 z + 6 y + 43) are contiguous runs of both the global and the local index space
__device__ __forceinline__ uint32_t hdr_own_local(const StageHeader& H, uint32_t p) {
    uint32_t r = 0;
    #pragma unroll
    for (uint32_t step = 8u; step >= 1u; step >>= 1) {
        const uint32_t t = r + step;
        if (H.cellG[43u + 6u * (t & 3u) + 36u * (t >> 2)] <= p) r = t;
    }
    const uint32_t c = 43u + 6u * (r & 3u) + 36u * (r >> 2);
    return H.local[c] + (p - H.cellG[c]);
}

template<int BBYTES>
__device__ __forceinline__ float4 pipe_payload_b(const void* __restrict__ sB, uint32_t L) {
    if (BBYTES == 16) return reinterpret_cast<const float4*>(sB)[L];
    if (BBYTES == 8) { const float2 t = reinterpret_cast<const float2*>(sB)[L]; return make_float4(t.x, t.y, 0.0f, 0.0f); }
    if (BBYTES == 4) return make_float4(reinterpret_cast<const float*>(sB)[L], 0.0f, 0.0f, 0.0f);
    return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

template<class Op>
__device__ __forceinline__ void pipe_consumer_pairs(const Arrays& A, PipeShared& ps, const unsigned char* pay, unsigned char* streams, Op& op) {
    static_assert(!Op::CUSTOM, "pair ops only");
    constexpr int BBYTES = Op::NPAY > 1 ? Op::BBYTES : 0;
    constexpr int R = Op::Cfg::D, GW = Op::Cfg::GW;
    constexpr uint32_t CW = Op::Cfg::CW, NONE = 0xffffffffu, SLOT = PipeLayout<Op>::SLOT;
    const uint32_t lane = threadIdx.x & 31u, cw = threadIdx.x >> 5;
    const uint2* __restrict__ listBase = reinterpret_cast<const uint2*>(A.list16);
    const float4* __restrict__ cinBase = reinterpret_cast<const float4*>(op.coef_in());
    float4* __restrict__ coutBase = reinterpret_cast<float4*>(op.coef_out());
    const unsigned char* payB = pay + (size_t)PipeRing<Op>::SLOTS * 16;
    // this lane's words in slot 0 of the warp's stream ring
    const uint32_t ringL = smem_u32(streams) + cw * (uint32_t)R * SLOT + lane * 8u, ringC = ringL - lane * 8u + 256u + lane * 16u;

    // Batches are handed out by a CTA-wide ticket counter (ps.ticket): ticket t is the CTA's t-th batch, counting through its
    // tiles in order, so the warps that are free always take the oldest batches not yet started — a tile's batches run side by
    // side and its ring space comes back as early as possible (with a fixed warp <-> batch assignment a tile waited for
    // whichever warps its batches were bound to; the payload ring holds only about three halo boxes: profiles/r02_pipeline_trace.md).
    // Sums do not depend on who took what: every batch's partial goes to its own slot of the tile's record.
    // The cursor: tile counter of this CTA, its batch count, the ticket of its first batch, the ticket the cursor is looking for
    uint32_t wK = 0, wB = 0, wN = 0, wBase = 0, wTick = 0;
    uint32_t heldK = NONE;                      // tile of the batch being gathered: released after its gather, not by the cursor
    bool wDone = false;
    auto claim = [&]() {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&ps.ticket, 1u);
        wTick = __shfl_sync(0xffffffffu, t, 0);
    };
    auto enter = [&]() {
        const uint32_t s = wK % PIPE_STAGES;
        ptrace(10, wK, 0);
        mbar_wait(&ps.table[s], (wK / PIPE_STAGES) & 1u);   // the tile's table (which particles, how many batches): the scout is well ahead of the payload
        ptrace(11, wK, 0);
        const StageHeader& H = ps.hdr[s];
        const uint32_t begin = H.begin;
        if (begin == NONE) { wDone = true; wN = 0; wB = 0; return; }
        wN = (H.end - begin + 31u) >> 5;
        wB = wTick - wBase;
    };
    // moves the cursor to this warp's next batch, releasing the tiles it leaves behind.  Non-blocking: stops (false) in front
    // of a tile the producers have not staged yet — the caller still holds a tile the producers may be waiting for.
    auto seek = [&](bool blocking) -> bool {
        while (!wDone && wB >= wN) {
            const uint32_t kn = wK + 1u;
            if (!blocking && !__any_sync(0xffffffffu, mbar_test(&ps.table[kn % PIPE_STAGES], (kn / PIPE_STAGES) & 1u))) return false;
            if (wK != heldK) { __syncwarp(); if (lane == 0) mbar_arrive(&ps.empty[wK % PIPE_STAGES]); }
            wBase += wN;
            wK = kn;
            enter();
        }
        return true;
    };
    auto batch_particle = [&]() -> uint32_t {   // the lane's particle of the cursor's batch
        const StageHeader& H = ps.hdr[wK % PIPE_STAGES];
        const uint32_t q = H.begin + (wB << 5) + lane;
        return q < H.end ? q : NONE;
    };
    auto ell_off = [](uint32_t p) -> uint32_t { return (p >> 5) * (uint32_t)(ELL_GROUPS * 32) + (p & 31u); };
    // one neighbour group of one particle into a ring slot (zeros when `valid` is false: nothing is read from global memory)
    auto issue = [&](uint32_t slot, uint32_t off, bool valid, bool cold = false) {
        if ((PIPE_ABLATE & 128) && !cold) return;
        cp_async8_zfill(ringL + slot * SLOT, listBase + off, valid ? 8u : 0u);
        // (the coefficient word as two 8-byte copies would halve its shared-memory write wavefronts — a 16-byte LDGSTS writes sector
        // by sector, 10.6 wavefronts per warp instruction against 2 x 2 — but the L1-allocating 8-byte form measured 10 % slower)
        if (Op::COEF & 1) cp_async16_zfill(ringC + slot * SLOT, cinBase + off, valid ? 16u : 0u);
    };

    claim();
    enter();
    seek(true);
    if (wDone) return;
    uint32_t pN = batch_particle(), mN = 0u;
    if (pN != NONE) mN = __ldg(A.cnt + pN);
    uint32_t q = 0;                             // ring slot of the next group to be consumed
    uint32_t issued = 0;                        // groups of the coming batch that are already in the ring (or on their way)
    for (;;) {
        // ---- the batch the cursor points at; then the cursor moves on, one batch ahead of the gather -----------------
        const uint32_t curK = wK, curB = wB, curN = wN;
        const StageHeader& H = ps.hdr[curK % PIPE_STAGES];
        const uint32_t p = pN, mf = mN;
        const uint32_t curLi = H.li;
        const uint32_t m = mf & VFD_COUNT_MASK, nG = (m + 3u) >> 2;
        const uint32_t cOff = p != NONE ? ell_off(p) : 0u;
        const uint32_t nGw = __reduce_max_sync(0xffffffffu, nG);
        ptrace(12, curK, (curB << 8) | issued);
        // the first groups of this batch: normally prefetched during the previous batch; otherwise (first batch, the next tile
        // was not staged in time, a predecessor with fewer than R groups) they are fetched now
        if (issued < (uint32_t)R) {
            cp_async_wait<0>();                 // nothing older may still be on its way into the slots written next
            #pragma unroll
            for (int j = 0; j < R; j++) if ((uint32_t)j >= issued) issue((q + (uint32_t)j) % (uint32_t)R, cOff + (uint32_t)j * 32u, (uint32_t)j < nG, true);
            cp_async_commit();
            cp_async_wait<0>();
        }
        ptrace(13, curK, curB);
        heldK = curK;
        claim();                                // this warp's next batch
        wB = wTick - wBase;
        const bool ahead = seek(false);
        const bool haveNext = ahead && !wDone;
        pN = NONE; mN = 0u;
        if (haveNext) { pN = batch_particle(); if (pN != NONE) mN = __ldg(A.cnt + pN); }
        const uint32_t nOff = pN != NONE ? ell_off(pN) : 0u;
        issued = 0;

        // everything above needed only the tile's table; the gathers need its payload
        ptrace(17, curK, curB);
        mbar_wait(&ps.full[curK % PIPE_STAGES], (curK / PIPE_STAGES) & 1u);
        ptrace(18, curK, curB);
        const bool staged = H.staged != 0u;
        const float4* __restrict__ sA = reinterpret_cast<const float4*>(pay + (size_t)H.base * 16);
        const void* __restrict__ sB = payB + (size_t)H.base * BBYTES;
        float own[Op::NOWN];
        #pragma unroll
        for (int i = 0; i < Op::NOWN; i++) own[i] = 0.0f;
        if (p != NONE) {
            float4 a, b;
            if (staged) { const uint32_t L = hdr_own_local(H, p); a = sA[L]; b = pipe_payload_b<BBYTES>(sB, L); }
            else { a = op.loadA(p); b = Op::NPAY > 1 ? op.loadB(p) : make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
            op.own_from(p, a, b, own);
        }
        float acc[Op::NSUM];
        #pragma unroll
        for (int s = 0; s < Op::NSUM; s++) acc[s] = 0.0f;

        // one group of four neighbour slots: GW payloads are gathered back to back before their pair terms are accumulated
        auto group = [&](const uint2 wq, const float4 cq, const uint32_t g, auto fetch) {
            uint32_t L[4];
            ell_unpack(wq, L);
            float c[4] = { cq.x, cq.y, cq.z, cq.w };
            if (Op::PAD_SAFE || g * 4u + 4u <= m) {
                #pragma unroll
                for (int u0 = 0; u0 < 4; u0 += GW) {
                    float4 pa[GW], pb[GW];
                    #pragma unroll
                    for (int u = 0; u < GW; u++) fetch(L[u0 + u], pa[u], pb[u]);
                    #pragma unroll
                    for (int u = 0; u < GW; u++) op.pair(own, pa[u], pb[u], c[u0 + u], acc);
                }
            } else {
                #pragma unroll
                for (int u = 0; u < 3; u++) {                    // a partial group holds one to three neighbours
                    if (g * 4u + (uint32_t)u < m) { float4 xa, xb; fetch(L[u], xa, xb); op.pair(own, xa, xb, c[u], acc); }
                    else c[u] = 0.0f;
                }
                c[3] = 0.0f;
            }
            if (Op::COEF & 2) coutBase[cOff + g * 32u] = make_float4(c[0], c[1], c[2], c[3]);
        };
        ptrace(14, curK, (curB << 8) | (ahead ? 1u : 0u));
        if (staged) {
            // PIPE_ABLATE (tuning builds, wrong results): 32: no gathers, 128: no list/coefficient stream
            auto fetch = [&](uint32_t L, float4& a, float4& b) {
                if (PIPE_ABLATE & 32) { a = make_float4(__uint_as_float(L | 0x3f800000u), own[1], own[2], 1.0f); b = a; return; }
                if (PIPE_ABLATE & 128) L = min(L, 1023u);
                a = sA[L]; b = pipe_payload_b<BBYTES>(sB, L);
            };
            // Every consumed group frees its ring slot, which is refilled at once — R groups ahead in this batch, or with the
            // next group of the warp's NEXT batch once this one's are all on their way — so R groups per warp are in flight
            // whatever the consumer is doing: the copy is issued before the group's gathers and arithmetic, and unlike a load
            // into registers it cannot be moved down to its use by the instruction scheduler.
            const uint32_t nGnext = ((mN & VFD_COUNT_MASK) + 3u) >> 2;
            const bool chain = haveNext && nGw >= (uint32_t)R;   // a batch with fewer groups than ring slots hands nothing over
            #pragma unroll 2
            for (uint32_t g = 0; g < nGw; g++) {
                cp_async_wait<R - 1>();         // this lane's words of group g have landed
                const uint2 wq = lds64(ringL + q * SLOT);
                float4 cq = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (Op::COEF & 1) cq = lds128(ringC + q * SLOT);
                const uint32_t gn = g + (uint32_t)R;
                if (gn < nGw) issue(q, cOff + gn * 32u, gn < nG);
                else if (chain) { issue(q, nOff + issued * 32u, issued < nGnext); issued++; }
                cp_async_commit();
                q = q + 1u == (uint32_t)R ? 0u : q + 1u;
                if (Op::PAD_SAFE) group(wq, cq, g, fetch);      // lanes past their last group hold index 0 / coefficient 0: a pair term of exactly zero
                else if (g < nG) group(wq, cq, g, fetch);
            }
            if (nGw < (uint32_t)R) {
                // fewer groups than ring slots: the slots behind them hold this batch's unused groups, not the next batch's
                issued = 0;
                q = (q + (uint32_t)R - nGw) % (uint32_t)R;       // q back to where this batch started: the next batch refills from there
            }
        } else {
            // a halo box beyond PIPE_CAP: exact path through global memory (rare); the ring is refilled by the next batch
            auto fetch = [&](uint32_t L, float4& a, float4& b) {
                const uint32_t gi = hdr_local_to_global(H, L);
                a = op.loadA(gi); b = Op::NPAY > 1 ? op.loadB(gi) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            };
            for (uint32_t g = 0; g < nG; g++) {
                const uint2 wq = __ldg(listBase + cOff + g * 32u);
                float4 cq = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (Op::COEF & 1) cq = __ldg(cinBase + cOff + g * 32u);
                group(wq, cq, g, fetch);
            }
            issued = 0;
        }
        // the gathers of this batch are complete: release its tile if the cursor has left it
        ptrace(15, curK, (curB << 8) | nGw);
        __syncwarp();
        if (wK != curK && lane == 0) mbar_arrive(&ps.empty[curK % PIPE_STAGES]);
        heldK = NONE;

        if (p != NONE && !((PIPE_ABLATE & 8) && acc[0] != 12345.0f)) op.finish(p, mf, own, acc);
        if constexpr (Op::NRED > 0) {
            __syncwarp();
            RedRecord& R_ = ps.rec[curK % PIPE_RED_RECORDS];
            #pragma unroll
            for (int k = 0; k < Op::NRED; k++) {
                double x = p != NONE ? (double)op.red[k] : 0.0;
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
                if (lane == 0) {
                    if (curB < PIPE_MAX_BATCH - 1) R_.bsum[k][curB] = x;
                    else atomicAdd(&R_.bsum[k][PIPE_MAX_BATCH - 1], x);
                }
            }
            uint32_t old = 0;
            if (lane == 0) { __threadfence_block(); old = atomicAdd(&R_.count, 1u); }
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old + 1u == curN) {                       // this warp completed the tile: fold its batches in order
                __threadfence_block();
                const uint32_t nb = min(curN, (uint32_t)PIPE_MAX_BATCH);
                #pragma unroll
                for (int k = 0; k < Op::NRED; k++) {
                    const volatile double* bs = R_.bsum[k];
                    double x = (lane < nb ? bs[lane] : 0.0) + (lane + 32u < nb ? bs[lane + 32u] : 0.0);
                    #pragma unroll
                    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
                    if (lane == 0) A.slotSums[(size_t)k * A.slotStride + curLi] = x;
                }
                __syncwarp();
                if (lane == 0) { R_.bsum[0][PIPE_MAX_BATCH - 1] = 0.0; R_.bsum[1][PIPE_MAX_BATCH - 1] = 0.0; R_.count = 0u; }
            }
        }
        ptrace(16, curK, curB);
        if (!ahead) {
            // the next tile was not staged when this batch began: wait for it now (this warp holds no tile any more)
            seek(true);
            if (!wDone) { pN = batch_particle(); if (pN != NONE) mN = __ldg(A.cnt + pN); }
            issued = 0;
        }
        if (wDone) break;
    }
    cp_async_wait<0>();
}

// Called by all Op::Cfg::THREADS threads of the CTA (after any lookup table has been loaded; contains __syncthreads).
// Returns true in every thread of the CTA that finished last (it has reset the tile queue; kernels with a grid-wide
// sum fold the batch slots there).
template<class Op>
__device__ __forceinline__ bool pipe_pass(DevState* __restrict__ S, const Arrays& A, PipeShared& ps, unsigned char* pay, Op& op,
                                          uint32_t tile0, uint32_t tile1, bool checkIndexRange = false) {
    if (threadIdx.x == 0) {
        ps.ticket = 0u;
        #pragma unroll
        for (int s = 0; s < PIPE_STAGES; s++) { mbar_init(&ps.full[s], PIPE_FULL_ARRIVALS); mbar_init(&ps.empty[s], Op::Cfg::CW); mbar_init(&ps.table[s], 1u); }
    }
    if constexpr (Op::CUSTOM) {
        if constexpr (Op::TILE_QUEUE) {
            for (uint32_t i = threadIdx.x; i < PIPE_RED_RECORDS * PIPE_QUEUE_CAP; i += blockDim.x) {
                TileQueue& Q = *reinterpret_cast<TileQueue*>(&ps.rec[i / PIPE_QUEUE_CAP]);
                Q.e[i % PIPE_QUEUE_CAP] = PIPE_QUEUE_NONE;
                if (i % PIPE_QUEUE_CAP == 0) { Q.head = 0u; Q.tail = 0u; Q.left = 0u; }
            }
        }
    } else if (threadIdx.x < PIPE_RED_RECORDS) {
        RedRecord& R = ps.rec[threadIdx.x];
        R.count = 0u; R.bsum[0][PIPE_MAX_BATCH - 1] = 0.0; R.bsum[1][PIPE_MAX_BATCH - 1] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x >= (Op::Cfg::CW + 1) * 32) pipe_copier(ps, pay, op);
    else if (threadIdx.x >= Op::Cfg::CW * 32) pipe_scout<Op>(S, A, ps, checkIndexRange);
    else if constexpr (Op::CUSTOM) pipe_consumer(A, ps, pay, op);
    else pipe_consumer_pairs(A, ps, pay, pay - PipeLayout<Op>::OFF_PAY + PipeLayout<Op>::OFF_STREAM, op);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                  // this CTA's slot and field stores before its ticket
        const bool last = atomicAdd(&S->doneCtas, 1u) == gridDim.x - 1u;
        if (last) { S->tileCursor = 0u; S->doneCtas = 0u; }
        ps.lastCta = last ? 1u : 0u;
    }
    __syncthreads();
    return ps.lastCta != 0u;
}

__device__ __forceinline__ PipeShared& pipe_header(unsigned char* raw) { return *reinterpret_cast<PipeShared*>(raw); }
template<class Op> __device__ __forceinline__ float* pipe_lut(unsigned char* raw) { return reinterpret_cast<float*>(raw + PipeLayout<Op>::OFF_LUT); }
template<class Op> __device__ __forceinline__ unsigned char* pipe_pay(unsigned char* raw) { return raw + PipeLayout<Op>::OFF_PAY; }
template<class Op> __device__ __forceinline__ unsigned char* pipe_stream(unsigned char* raw) { return raw + PipeLayout<Op>::OFF_STREAM; }

// The particles this rank owns (multi-GPU: the tiles [tile0, tile1) of the local grid; the tile columns before and
// after hold ghost copies of the neighbour slabs' edge particles): a contiguous range because tiles are x-slowest.
__device__ __forceinline__ void owned_range(const Params& P, const uint32_t* __restrict__ cellBegin, uint32_t& b, uint32_t& e) {
    b = P.tile0 ? __ldg(cellBegin + (size_t)P.tile0 * TILE_CELLS) : 0u;
    e = P.tile1 != 0xffffffffu ? __ldg(cellBegin + (size_t)P.tile1 * TILE_CELLS) : P.n;
}

// dynamic shared memory carve-up: [TileShared][NLUT lookup tables][payload A][payload B]
__device__ __forceinline__ TileShared& smem_header(unsigned char* raw) { return *reinterpret_cast<TileShared*>(raw); }
__host__ __device__ constexpr size_t smem_header_bytes() { return (sizeof(TileShared) + 127) / 128 * 128; }
template<int NLUT> __device__ __forceinline__ float* smem_lut(unsigned char* raw) { return reinterpret_cast<float*>(raw + smem_header_bytes()); }
template<int NLUT> __device__ __forceinline__ float4* smem_pay_a(unsigned char* raw) {
    return reinterpret_cast<float4*>(raw + smem_header_bytes() + (size_t)NLUT * VFD_LUT_RES * sizeof(float));
}
template<int NLUT> __device__ __forceinline__ float4* smem_pay_b(unsigned char* raw, uint32_t cap) { return smem_pay_a<NLUT>(raw) + cap; }
template<int NLUT, int NPAY> static inline size_t tile_smem_bytes(uint32_t cap) {
    return smem_header_bytes() + (size_t)NLUT * VFD_LUT_RES * sizeof(float) + (size_t)cap * NPAY * sizeof(float4);
}

__device__ __forceinline__ void load_lut_tile(float* dst, const float* __restrict__ src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < (VFD_LUT_RES / 4); i += blockDim.x) d4[i] = __ldg(s4 + i);
}

} // namespace vfd
