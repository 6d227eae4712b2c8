// distributed.cu — spatial decomposition of one simulation over the GPUs of an NVSwitch box: one process (rank) per
// GPU, slabs of whole tile columns along x, NCCL point-to-point for the particle state and the halos, NCCL
// all-reduce for the solver's scalars.  The reference has nothing of the kind (single device 0, default stream:
// SURVEY.md F10); the obligation is that an N-rank run reproduces the 1-rank run of this library.
//
// Geometry: every rank uses the SAME global grid (origin, cell size, tiles) over a fixed domain, so a particle's
// cell — and with it the in-cell order by persistent id — is identical on every rank that holds a copy of it.
// Rank r owns the tile columns [colLo, colHi); its local arrays additionally hold a ghost copy of the last owned
// column of rank r-1 and of the first owned column of rank r+1.  Because tiles are ordered x-slowest the local
// arrays read   [ ghost L | first owned column ... last owned column | ghost R ]   and every halo exchange is a
// pair of contiguous copies with no packing: "my first owned column -> left neighbour's ghost R" and
// "my last owned column -> right neighbour's ghost L".
//
// Per step:
//   1. state exchange (exchange_state): owned particles are classified by the tile column of their new position;
//      those in [colLo-1, colHi] stay in the local arrays (as owned or ghost), those in columns <= colLo are sent
//      left and those in columns >= colHi-1 are sent right (80-B records: position, velocity, velocity change,
//      smoothed normal, curvatures, id) — migration and ghost refresh in one message per neighbour;
//   2. the usual sort / search over owned + ghost particles; neighbour sums run over the owned tiles only;
//   3. after every kernel that produces a field a later kernel gathers from neighbours (density, kappa, pressure
//      acceleration, PCG direction, normals, velocity) the two edge columns are exchanged (halo);
//   4. every reduction (Jacobi residual, CFL maximum, PCG dot products) is all-reduced and the control decision
//      (control.cuh) is then taken by a one-thread kernel, identically on all ranks.
#include "solver_impl.h"
#include "tile.cuh"
#include "control.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <cstring>

namespace vfd {

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail_cuda(e__, #call, __LINE__); } while (0)
#define RC_(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)
#define NK(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return fail(VFD_E_NCCL, std::string("NCCL error: ") + dist->api.GetErrorString(r__) + " in " #call); } while (0)

// NCCL is resolved at run time (the library has no link-time dependency on it; a single-GPU user never loads it).
bool NcclApi::load(std::string& err) {
    if (handle) return true;
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);            // the copy torch already mapped, if any
    if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define SYM(name) *(void**)(&name) = dlsym(handle, "nccl" #name); if (!name) { err = "libnccl lacks nccl" #name; return false; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(GetErrorString) SYM(GroupStart) SYM(GroupEnd) SYM(Send) SYM(Recv) SYM(AllReduce) SYM(AllGather)
#undef SYM
    return true;
}

struct Record { float4 pos, vel, dv, nbar; float curv, curvS, curvD; uint32_t id; };   // 80 B
static_assert(sizeof(Record) == 80, "state record");

// tile column (global) of a position
__device__ __forceinline__ int tile_column(float x, float originX, float invCell, int gdimX) {
    int c = (int)floorf((x - originX) * invCell);
    c = min(max(c, 1), gdimX - 2);
    return c >> 2;
}

// step 1 of the state exchange: classify the owned particles, compact the ones that stay, pack the ones that leave
__global__ void __launch_bounds__(VFD_TPB) k_migrate(Params P, Arrays A, uint32_t ownB, uint32_t ownE, float originX, float invCell, int gdimX,
                                                     int colLo, int colHi, int hasLeft, int hasRight,
                                                     Record* __restrict__ sendL, Record* __restrict__ sendR, uint32_t sendCap, uint32_t* __restrict__ counters) {
    const uint32_t p = ownB + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ownE) return;
    const float4 x = A.pos[p];
    const int col = tile_column(x.x, originX, invCell, gdimX);
    const bool keep = col >= colLo - 1 && col <= colHi;
    const bool toL = hasLeft && col <= colLo, toR = hasRight && col >= colHi - 1;
    Record r{ x, A.vel[p], A.dv[p], A.nbar[p], A.curv[p], A.curvS[p], A.curvD[p], A.id[p] };
    if (keep) {
        const uint32_t q = atomicAdd(&counters[0], 1u);
        A.pos2[q] = r.pos; A.vel2[q] = r.vel; A.dv2[q] = r.dv; A.nbar2[q] = r.nbar;
        A.curv2[q] = r.curv; A.curvS2[q] = r.curvS; A.curvD2[q] = r.curvD; A.id2[q] = r.id;
    }
    if (toL) { const uint32_t q = atomicAdd(&counters[1], 1u); if (q < sendCap) sendL[q] = r; }
    if (toR) { const uint32_t q = atomicAdd(&counters[2], 1u); if (q < sendCap) sendR[q] = r; }
}

// step 2: append the received records behind the kept particles; the receiver keeps only what lies in its local grid
__global__ void __launch_bounds__(VFD_TPB) k_unpack(Arrays A, const Record* __restrict__ recv, uint32_t count, float originX, float invCell, int gdimX,
                                                    int colLo, int colHi, uint32_t capacity, uint32_t* __restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Record r = recv[i];
    const int col = tile_column(r.pos.x, originX, invCell, gdimX);
    if (col < colLo - 1 || col > colHi) return;
    const uint32_t q = atomicAdd(&counters[0], 1u);
    if (q >= capacity) return;
    A.pos2[q] = r.pos; A.vel2[q] = r.vel; A.dv2[q] = r.dv; A.nbar2[q] = r.nbar;
    A.curv2[q] = r.curv; A.curvS2[q] = r.curvS; A.curvD2[q] = r.curvD; A.id2[q] = r.id;
}

// ---- peer-memory slab --------------------------------------------------------------------------------------
// [ PeerCtl (4 KB) | seven float4 arrays | two float2 arrays | two float arrays ], each array np particle slots long
enum SlabRegion { SR_POSRHO = 0, SR_VEL_A, SR_VEL_B, SR_PACC, SR_NRM, SR_CGXG, SR_CGXP, SR_CGGYZ, SR_CGPYZ, SR_KAPPA, SR_KAPPAV, SR_COUNT };
static const uint32_t kRegionBytes[SR_COUNT] = { 16, 16, 16, 16, 16, 16, 16, 8, 8, 4, 4 };
static size_t slab_offset(int region, uint64_t np) {
    size_t off = 4096;
    for (int r = 0; r < region; r++) off += (size_t)np * kRegionBytes[r];
    return off;
}
static size_t slab_bytes(uint64_t np) { return slab_offset(SR_COUNT, np) + 256; }     // + slack: granule-aligned bulk copies read up to one element past a range (tile.cuh)

// One halo exchange of one array: this rank's first / last owned tile column straight into the ghost range of the left /
// right neighbour's copy of the array (peer stores over NVLink), and theirs into ours.
//   1. "ready": every earlier kernel of this rank has finished with its ghost ranges (stream order) — told to both neighbours;
//   2. wait for the neighbours' ready, copy, fence;
//   3. the last block tells both neighbours that the data has landed and waits for theirs: when the kernel ends the ghost
//      ranges of this rank are current.  One launch per exchange; nothing on the host.
struct HaloArray { const unsigned char* mine; unsigned char* dstL; unsigned char* dstR; uint32_t eb, pad; };
struct HaloJob { HaloArray a[2]; uint32_t count; };          // arrays exchanged by one launch (the PCG direction travels as two)
__global__ void __launch_bounds__(VFD_TPB) k_halo_exchange(const __grid_constant__ HaloJob J, uint32_t ownB, uint32_t edgeLEnd, uint32_t edgeRBegin, uint32_t ownE,
                                                           PeerCtl* me, PeerCtl* ctlL, PeerCtl* ctlR, DevState* S) {
    __shared__ uint32_t shSeq;
    if (threadIdx.x == 0) {
        const uint32_t seq = *(volatile uint32_t*)&me->haloSeq + 1u;          // the last block advances it once every block has read it
        if (blockIdx.x == 0) {
            if (ctlL) *(volatile uint32_t*)&ctlL->haloReady[1] = seq;           // I am my left neighbour's right neighbour
            if (ctlR) *(volatile uint32_t*)&ctlR->haloReady[0] = seq;
        }
        if (ctlL) while (*(volatile uint32_t*)&me->haloReady[0] != seq) { }
        if (ctlR) while (*(volatile uint32_t*)&me->haloReady[1] != seq) { }
        __threadfence_system();
        shSeq = seq;
    }
    __syncthreads();
    const uint32_t nL = ctlL ? edgeLEnd - ownB : 0u, nR = ctlR ? ownE - edgeRBegin : 0u;
    for (uint32_t k = 0; k < J.count; k++) {
        const HaloArray& H = J.a[k];
        const uint32_t eb = H.eb;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nL + nR; i += gridDim.x * blockDim.x) {
            const bool left = i < nL;
            const size_t so = (size_t)(left ? ownB + i : edgeRBegin + (i - nL)) * eb, dof = (size_t)(left ? i : i - nL) * eb;
            unsigned char* dst = (left ? H.dstL : H.dstR) + dof;
            if (eb == 16) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(H.mine + so);
            else if (eb == 8) *reinterpret_cast<float2*>(dst) = *reinterpret_cast<const float2*>(H.mine + so);
            else *reinterpret_cast<float*>(dst) = *reinterpret_cast<const float*>(H.mine + so);
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&S->ticket[7], 1u) == gridDim.x - 1u) {
        const uint32_t seq = shSeq;
        S->ticket[7] = 0u;
        __threadfence_system();
        if (ctlL) *(volatile uint32_t*)&ctlL->haloData[1] = seq;
        if (ctlR) *(volatile uint32_t*)&ctlR->haloData[0] = seq;
        if (ctlL) while (*(volatile uint32_t*)&me->haloData[0] != seq) { }
        if (ctlR) while (*(volatile uint32_t*)&me->haloData[1] != seq) { }
        __threadfence_system();
        *(volatile uint32_t*)&me->haloSeq = seq;
    }
}

__global__ void k_control(Params P, DevState* S, int site) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // iterations enqueued beyond convergence are no-ops in the kernels; their (stale) reduction must not be acted upon
    if ((site == SITE_VISC_PQ || site == SITE_VISC_UPDATE) && S->viscActive != 1u) return;
    if (site == SITE_DIV && !S->divActive) return;
    if (site == SITE_PRESS && !S->pressActive) return;
    control_site(site, P, S, &S->red[site * 2]);
}

// owned particles in local order, with their persistent ids (frame / dump gathering on the host)
__global__ void k_export_owned(Params P, Arrays A, uint32_t ownB, uint32_t ownE, uint32_t* __restrict__ ids, VfdParticle* __restrict__ out) {
    const uint32_t p = ownB + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ownE) return;
    VfdParticle q;
    const float4 x = A.pos[p], v = A.vel[p], a = A.acc[p], pa = A.pacc[p], dv = A.dv[p], n = A.nrm[p], nb = A.nbar[p];
    q.Position[0] = x.x; q.Position[1] = x.y; q.Position[2] = x.z;
    q.Velocity[0] = v.x; q.Velocity[1] = v.y; q.Velocity[2] = v.z;
    q.Acceleration[0] = a.x; q.Acceleration[1] = a.y; q.Acceleration[2] = a.z;
    q.PressureAcceleration[0] = pa.x; q.PressureAcceleration[1] = pa.y; q.PressureAcceleration[2] = pa.z;
    q.PressureResiduum = A.res[p]; q.Density = A.rho[p]; q.DensityAdvection = A.rhoAdv[p];
    q.PressureRho2 = A.kappa[p]; q.PressureRho2V = A.kappaV[p]; q.Factor = A.alpha[p];
    q.VelocityDifference[0] = dv.x; q.VelocityDifference[1] = dv.y; q.VelocityDifference[2] = dv.z;
    q.MonteCarloSurfaceNormal[0] = n.x; q.MonteCarloSurfaceNormal[1] = n.y; q.MonteCarloSurfaceNormal[2] = n.z;
    q.MonteCarloSurfaceNormalSmooth[0] = nb.x; q.MonteCarloSurfaceNormalSmooth[1] = nb.y; q.MonteCarloSurfaceNormalSmooth[2] = nb.z;
    q.MonteCarloSurfaceCurvature = A.curv[p]; q.MonteCarloSurfaceCurvatureSmooth = A.curvS[p]; q.DeltaFinalCurvature = A.curvD[p];
    out[p - ownB] = q;
    ids[p - ownB] = A.id[p];
}

// ---------------------------------------------------------------------------------------------------------
int dist_unique_id(char* out128, std::string& err) {
    NcclApi api;
    if (!api.load(err)) return VFD_E_NCCL;
    ncclUniqueId id;
    ncclResult_t r = api.GetUniqueId(&id);
    if (r != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + api.GetErrorString(r); return VFD_E_NCCL; }
    memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return VFD_OK;
}

// the global grid of a domain: cell = h (1 + 1/1023) like the single-GPU search, two pad cells below, whole tiles
void dist_grid(const float* dmin, const float* dmax, float h, float origin[3], uint32_t dim[3]) {
    const double cell = (double)h / (1.0 - 1.0 / 1024.0);
    for (int k = 0; k < 3; k++) {
        origin[k] = dmin[k] - 2.0f * (float)cell;
        const int cells = (int)std::ceil(((double)dmax[k] - (double)origin[k]) / cell) + 2;
        dim[k] = (uint32_t)((cells + 3) / 4 * 4);
    }
}

int Solver::dist_init(int rank, int nranks, const char* id128, const float* dmin, const float* dmax) {
    CK(cudaSetDevice(device));
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(VFD_E_INVALID, "bad rank / world size");
    if (info.ParticleCount) return fail(VFD_E_INVALID, "init_distributed must precede set_particles");
    dist.reset(new Dist());
    dist->rank = rank; dist->nranks = nranks;
    std::string err;
    if (!dist->api.load(err)) return fail(VFD_E_NCCL, err);
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    NK(dist->api.CommInitRank(&dist->comm, nranks, id, rank));
    dist_grid(dmin, dmax, info.SupportRadius, dist->origin, dist->gdim);
    for (int k = 0; k < 3; k++) dist->gtiles[k] = dist->gdim[k] / 4;
    dist->colLo = 0; dist->colHi = dist->gtiles[0];
    CK(cudaMalloc(&dist->dCounters, 64 * sizeof(uint32_t)));
    { const char* e = getenv("VFD_DIST_REBALANCE"); if (e) dist->rebalanceEvery = atoi(e); }
    { const char* e = getenv("VFD_DIST_FUSED_HALO"); dist->fusedHalo = e && atoi(e) == 1; }
    CK(cudaMallocHost(&dist->hCounters, 16 * sizeof(uint32_t)));
    return VFD_OK;
}

int Solver::dist_get_grid(float* origin3, float* cellSize, uint32_t* tiles3) {
    if (!dist) return fail(VFD_E_INVALID, "not distributed");
    for (int k = 0; k < 3; k++) { origin3[k] = dist->origin[k]; tiles3[k] = dist->gtiles[k]; }
    *cellSize = (float)((double)info.SupportRadius / (1.0 - 1.0 / 1024.0));
    return VFD_OK;
}

int Solver::dist_set_slab(uint32_t colLo, uint32_t colHi) {
    if (!dist) return fail(VFD_E_INVALID, "not distributed");
    if (colLo >= colHi || colHi > dist->gtiles[0]) return fail(VFD_E_INVALID, "bad slab (tile columns)");
    if (info.ParticleCount) return fail(VFD_E_INVALID, "set_slab must precede set_particles");
    dist->colLo = colLo; dist->colHi = colHi;
    return VFD_OK;
}

// local grid of this rank: its owned tile columns plus one ghost column towards each existing neighbour
void Solver::dist_apply_grid(DevState& s) {
    const Dist& D = *dist;
    const bool hasL = D.rank > 0, hasR = D.rank + 1 < D.nranks;
    const uint32_t c0 = D.colLo - (hasL ? 1u : 0u), c1 = D.colHi + (hasR ? 1u : 0u);
    s.gridOrigin[0] = D.origin[0]; s.gridOrigin[1] = D.origin[1]; s.gridOrigin[2] = D.origin[2];
    s.cellOffset[0] = (int32_t)(c0 * 4u); s.cellOffset[1] = 0; s.cellOffset[2] = 0;
    s.gridDim[0] = (c1 - c0) * 4u; s.gridDim[1] = D.gdim[1]; s.gridDim[2] = D.gdim[2];
    s.tileDim[0] = c1 - c0; s.tileDim[1] = D.gtiles[1]; s.tileDim[2] = D.gtiles[2];
    s.nTiles = s.tileDim[0] * s.tileDim[1] * s.tileDim[2];
    s.nCells = s.nTiles * 64u;
    s.gridMinCell[0] = s.gridMinCell[1] = s.gridMinCell[2] = 0;
}

void Solver::dist_params(Params& P) const {
    const Dist& D = *dist;
    const uint32_t T = D.gtiles[1] * D.gtiles[2];
    P.nRanks = (uint32_t)D.nranks; P.rank = (uint32_t)D.rank;
    P.nGlobal = D.nGlobal;
    P.n = D.nLocal;
    P.tile0 = (D.rank > 0 ? 1u : 0u) * T;
    P.tile1 = P.tile0 + (D.colHi - D.colLo) * T;
    for (int r = 0; r < 8; r++) P.peerCtl[r] = (D.p2p && r < D.nranks) ? (void*)D.peerSlab[r] : nullptr;
    dist_halo_targets(P);
}

void Solver::dist_count_fused_halo(uint64_t bytesPerItem) {
    Dist& D = *dist;
    const bool hasL = D.rank > 0, hasR = D.rank + 1 < D.nranks;
    D.halos += 1;
    D.bytesHalo += (uint64_t)((hasL ? D.edgeLEnd - D.ownB : 0u) + (hasR ? D.ownE - D.edgeRBegin : 0u)) * bytesPerItem;
}

// where this step's edge columns go in the neighbours' copies of the PCG direction (the ranges move with every sort)
void Solver::dist_halo_targets(Params& P) const {
    const Dist& D = *dist;
    const bool hasL = D.rank > 0, hasR = D.rank + 1 < D.nranks;
    for (int k = 0; k < 2; k++) {
        const int region = k == 0 ? SR_CGXP : SR_CGPYZ;
        const size_t eb = kRegionBytes[region];
        P.haloP[0][k] = (D.fusedNow && hasL) ? D.peerSlab[D.rank - 1] + slab_offset(region, D.slabNp[D.rank - 1]) + (size_t)D.leftOwnE * eb : nullptr;
        P.haloP[1][k] = (D.fusedNow && hasR) ? D.peerSlab[D.rank + 1] + slab_offset(region, D.slabNp[D.rank + 1]) : nullptr;
    }
    P.haloRange[0] = D.ownB; P.haloRange[1] = D.edgeLEnd; P.haloRange[2] = D.edgeRBegin; P.haloRange[3] = D.ownE;
}

// The slab of this rank (control block + the arrays halo exchanges touch) and its mapping into every other rank: called by
// alloc_particles on every rank at the same point (set_particles_distributed is collective).
int Solver::dist_alloc_slab(size_t np) {
    Dist& D = *dist;
    Arrays& A = arrays;
    dist_free_slab();
    const char* off = getenv("VFD_DIST_P2P");
    const bool want = !(off && atoi(off) == 0) && D.nranks <= 8;
    D.slabBytes = slab_bytes(np);
    CK(cudaMalloc(&D.slab, D.slabBytes));
    CK(cudaMemset(D.slab, 0, D.slabBytes));
    unsigned char* b = D.slab;
    A.posRho = (float4*)(b + slab_offset(SR_POSRHO, np)); A.vel = (float4*)(b + slab_offset(SR_VEL_A, np)); A.vel2 = (float4*)(b + slab_offset(SR_VEL_B, np));
    A.pacc = (float4*)(b + slab_offset(SR_PACC, np)); A.nrm = (float4*)(b + slab_offset(SR_NRM, np));
    A.cgXG = (float4*)(b + slab_offset(SR_CGXG, np)); A.cgXP = (float4*)(b + slab_offset(SR_CGXP, np));
    A.cgGyz = (float2*)(b + slab_offset(SR_CGGYZ, np)); A.cgPyz = (float2*)(b + slab_offset(SR_CGPYZ, np));
    A.kappa = (float*)(b + slab_offset(SR_KAPPA, np)); A.kappaV = (float*)(b + slab_offset(SR_KAPPAV, np));
    D.p2p = false;
    for (int r = 0; r < 8; r++) { D.peerSlab[r] = nullptr; D.slabNp[r] = 0; }
    // every rank's IPC handle and slot count, gathered through the communicator that exists anyway
    struct Info { cudaIpcMemHandle_t h; uint64_t np; uint64_t ok; };
    Info mine; memset(&mine, 0, sizeof mine);
    mine.np = np;
    mine.ok = want && cudaIpcGetMemHandle(&mine.h, D.slab) == cudaSuccess ? 1u : 0u;
    cudaGetLastError();
    Info* dInfo = nullptr;
    CK(cudaMalloc(&dInfo, sizeof(Info) * (size_t)(D.nranks + 1)));
    CK(cudaMemcpyAsync(dInfo + D.nranks, &mine, sizeof mine, cudaMemcpyHostToDevice, stream));
    NK(D.api.AllGather(dInfo + D.nranks, dInfo, sizeof(Info), ncclChar, D.comm, stream));
    std::vector<Info> all((size_t)D.nranks);
    CK(cudaMemcpyAsync(all.data(), dInfo, sizeof(Info) * (size_t)D.nranks, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    cudaFree(dInfo);
    bool ok = true;
    for (int r = 0; r < D.nranks; r++) ok = ok && all[(size_t)r].ok;
    if (ok) {
        for (int r = 0; r < D.nranks && ok; r++) {
            D.slabNp[r] = all[(size_t)r].np;
            if (r == D.rank) { D.peerSlab[r] = D.slab; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            D.peerSlab[r] = (unsigned char*)ptr;
        }
    }
    // all ranks take the same path: agree on the outcome
    int* dOk = nullptr;
    CK(cudaMalloc(&dOk, sizeof(int)));
    int hOk = ok ? 1 : 0;
    CK(cudaMemcpyAsync(dOk, &hOk, sizeof(int), cudaMemcpyHostToDevice, stream));
    NK(D.api.AllReduce(dOk, dOk, 1, ncclInt32, ncclMin, D.comm, stream));
    CK(cudaMemcpyAsync(&hOk, dOk, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    cudaFree(dOk);
    D.p2p = hOk == 1;
    if (!D.p2p) for (int r = 0; r < D.nranks; r++) if (r != D.rank && D.peerSlab[r]) { cudaIpcCloseMemHandle(D.peerSlab[r]); D.peerSlab[r] = nullptr; }
    return VFD_OK;
}

void Solver::dist_free_slab() {
    if (!dist) return;
    Dist& D = *dist;
    Arrays& A = arrays;
    if (!D.slab) return;
    for (int r = 0; r < 8; r++) { if (r != D.rank && D.peerSlab[r]) cudaIpcCloseMemHandle(D.peerSlab[r]); D.peerSlab[r] = nullptr; }
    A.posRho = A.vel = A.vel2 = A.pacc = A.nrm = A.cgXG = A.cgXP = nullptr;
    A.cgGyz = A.cgPyz = nullptr; A.kappa = A.kappaV = nullptr;
    cudaFree(D.slab);
    D.slab = nullptr; D.slabBytes = 0; D.p2p = false;
}

// migration + ghost refresh (see the header of this file)
// the device's copy of the local grid (DevState: gridMinCell .. gridOrigin, cellOffset) after the slab's bounds have changed
int Solver::dist_upload_grid() {
    static_assert(offsetof(DevState, gridOrigin) > offsetof(DevState, gridMinCell), "grid fields are contiguous in DevState");
    DevState& g = hState[2];
    memset(&g, 0, sizeof g);
    dist_apply_grid(g);
    const size_t a = offsetof(DevState, gridMinCell), b = offsetof(DevState, gridOrigin) + sizeof(g.gridOrigin);
    CK(cudaMemcpyAsync((char*)dState + a, (char*)&g + a, b - a, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync((char*)dState + offsetof(DevState, cellOffset), (char*)&g + offsetof(DevState, cellOffset), sizeof(g.cellOffset), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));        // hState[2] is reused
    cellEstimate = g.nCells;
    return VFD_OK;
}

int Solver::dist_exchange_state() {
    Dist& D = *dist;
    const bool hasL = D.rank > 0, hasR = D.rank + 1 < D.nranks;
    if (D.pendL || D.pendR) {
        // a tile column changes hands: the new bounds classify this step's particles, the column itself travels with the
        // migrants below (what was an edge column becomes the neighbour's first owned one, and stays here as the ghost copy)
        D.colLo = (uint32_t)((int)D.colLo + D.pendL);
        D.colHi = (uint32_t)((int)D.colHi + D.pendR);
        D.shifts += (D.pendL != 0) + (D.pendR != 0);
        D.pendL = D.pendR = 0;
        RC_(dist_upload_grid());
    }
    Params P = params;
    const float invCell = (1.0f / info.SupportRadius) * (1.0f - 1.0f / 1024.0f);
    CK(cudaMemsetAsync(D.dCounters, 0, 16 * sizeof(uint32_t), stream));
    const uint32_t nOwned = D.ownE - D.ownB;
    if (nOwned) {
        k_migrate<<<(nOwned + VFD_TPB - 1) / VFD_TPB, VFD_TPB, 0, stream>>>(P, arrays, D.ownB, D.ownE, D.origin[0], invCell, (int)D.gdim[0],
            (int)D.colLo, (int)D.colHi, hasL ? 1 : 0, hasR ? 1 : 0, (Record*)D.sendL, (Record*)D.sendR, D.haloCap, D.dCounters);
        launches += 1;
    }
    // counts: mine to the host, and to the neighbours
    NK(D.api.GroupStart());
    if (hasL) { NK(D.api.Send(D.dCounters + 1, 1, ncclUint32, D.rank - 1, D.comm, stream)); NK(D.api.Recv(D.dCounters + 4, 1, ncclUint32, D.rank - 1, D.comm, stream)); }
    if (hasR) { NK(D.api.Send(D.dCounters + 2, 1, ncclUint32, D.rank + 1, D.comm, stream)); NK(D.api.Recv(D.dCounters + 5, 1, ncclUint32, D.rank + 1, D.comm, stream)); }
    NK(D.api.GroupEnd());
    CK(cudaMemcpyAsync(D.hCounters, D.dCounters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const uint32_t nKeep = D.hCounters[0], nSendL = D.hCounters[1], nSendR = D.hCounters[2];
    const uint32_t nRecvL = hasL ? D.hCounters[4] : 0u, nRecvR = hasR ? D.hCounters[5] : 0u;
    if (nSendL > D.haloCap || nSendR > D.haloCap || nRecvL > D.haloCap || nRecvR > D.haloCap)
        return fail(VFD_E_CAPACITY, "halo buffer too small for one tile column of particles (raise the capacity passed to set_particles)");
    if ((uint64_t)nKeep + nRecvL + nRecvR > D.capacity) return fail(VFD_E_CAPACITY, "slab outgrew its particle capacity (rebalance the slabs or raise the capacity)");
    NK(D.api.GroupStart());
    if (hasL) {
        if (nSendL) NK(D.api.Send(D.sendL, (size_t)nSendL * 20, ncclFloat32, D.rank - 1, D.comm, stream));
        if (nRecvL) NK(D.api.Recv(D.recvL, (size_t)nRecvL * 20, ncclFloat32, D.rank - 1, D.comm, stream));
    }
    if (hasR) {
        if (nSendR) NK(D.api.Send(D.sendR, (size_t)nSendR * 20, ncclFloat32, D.rank + 1, D.comm, stream));
        if (nRecvR) NK(D.api.Recv(D.recvR, (size_t)nRecvR * 20, ncclFloat32, D.rank + 1, D.comm, stream));
    }
    NK(D.api.GroupEnd());
    if (nRecvL) { k_unpack<<<(nRecvL + VFD_TPB - 1) / VFD_TPB, VFD_TPB, 0, stream>>>(arrays, (const Record*)D.recvL, nRecvL, D.origin[0], invCell, (int)D.gdim[0], (int)D.colLo, (int)D.colHi, D.capacity, D.dCounters); launches += 1; }
    if (nRecvR) { k_unpack<<<(nRecvR + VFD_TPB - 1) / VFD_TPB, VFD_TPB, 0, stream>>>(arrays, (const Record*)D.recvR, nRecvR, D.origin[0], invCell, (int)D.gdim[0], (int)D.colLo, (int)D.colHi, D.capacity, D.dCounters); launches += 1; }
    CK(cudaMemcpyAsync(D.hCounters, D.dCounters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    Arrays& A = arrays;
    std::swap(A.pos, A.pos2); std::swap(A.vel, A.vel2); std::swap(A.dv, A.dv2); std::swap(A.nbar, A.nbar2);
    std::swap(A.curv, A.curv2); std::swap(A.curvS, A.curvS2); std::swap(A.curvD, A.curvD2); std::swap(A.id, A.id2);
    D.nLocal = D.hCounters[0];
    D.bytesState += (uint64_t)(nSendL + nSendR) * sizeof(Record);
    params.n = D.nLocal;
    return VFD_OK;
}

// after the sort: where the ghost / edge / owned ranges of the local arrays lie; cross-checked with the neighbours
int Solver::dist_read_ranges() {
    Dist& D = *dist;
    const bool hasL = D.rank > 0, hasR = D.rank + 1 < D.nranks;
    const uint32_t T = D.gtiles[1] * D.gtiles[2];
    const uint32_t t0 = params.tile0, t1 = params.tile1;
    const uint32_t* cb = arrays.cellBegin;
    uint32_t* h = D.hCounters + 8;
    CK(cudaMemcpyAsync(h + 0, cb + (size_t)t0 * 64, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h + 1, cb + (size_t)(t0 + T) * 64, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h + 2, cb + (size_t)(t1 - T) * 64, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h + 3, cb + (size_t)t1 * 64, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    D.ownB = h[0]; D.edgeLEnd = h[1]; D.edgeRBegin = h[2]; D.ownE = h[3];
    // my ghost ranges must be exactly my neighbours' edge columns: verify before the first halo message of the step
    uint32_t* d = D.dCounters + 8;
    // to each neighbour: the size of the edge column it mirrors, and where my own range ends (= where my ghost-R range begins:
    // the right neighbour writes its halos there)
    // to each neighbour: the size of the edge column it mirrors, where my own range ends (= where my ghost-R range begins: the
    // right neighbour writes its halos there), how many particles and tile columns I own (slab re-balancing)
    const uint32_t nOwn = D.ownE - D.ownB, cols = D.colHi - D.colLo;
    const uint32_t mine[8] = { D.edgeLEnd - D.ownB, D.ownE, nOwn, cols, D.ownE - D.edgeRBegin, D.ownE, nOwn, cols };
    uint32_t* d2 = D.dCounters + 16;            // [0..7] mine, [8..11] from the left, [12..15] from the right
    CK(cudaMemcpyAsync(d2, mine, 32, cudaMemcpyHostToDevice, stream));
    NK(D.api.GroupStart());
    if (hasL) { NK(D.api.Send(d2 + 0, 4, ncclUint32, D.rank - 1, D.comm, stream)); NK(D.api.Recv(d2 + 8, 4, ncclUint32, D.rank - 1, D.comm, stream)); }
    if (hasR) { NK(D.api.Send(d2 + 4, 4, ncclUint32, D.rank + 1, D.comm, stream)); NK(D.api.Recv(d2 + 12, 4, ncclUint32, D.rank + 1, D.comm, stream)); }
    NK(D.api.GroupEnd());
    uint32_t got[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    CK(cudaMemcpyAsync(got, d2 + 8, 32, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    h[4] = got[0]; h[5] = got[4];
    D.leftOwnE = got[1];
    D.maxOwned = nOwn;
    if (D.fusedHalo && D.p2p) {
        // Whether the fused vector kernel carries the step must be the SAME decision on every rank (it replaces an exchange
        // kernel that the neighbours would otherwise wait in): it depends on the largest owned count over all ranks
        uint32_t* dm = D.dCounters + 48;
        CK(cudaMemcpyAsync(dm, &nOwn, 4, cudaMemcpyHostToDevice, stream));
        NK(D.api.AllReduce(dm, dm + 1, 1, ncclUint32, ncclMax, D.comm, stream));
        uint32_t mx = 0;
        CK(cudaMemcpyAsync(&mx, dm + 1, 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        D.maxOwned = mx;
    }
    D.fusedNow = D.fusedHalo && D.p2p && (uint64_t)D.maxOwned <= (uint64_t)numSMs * 8u * 1024u - 32ull * (uint64_t)numSMs;
    dist_halo_targets(params);
    D.stepsDone++;
    if (D.rebalanceEvery > 0 && D.stepsDone % (uint64_t)D.rebalanceEvery == 0) {
        // boundary between ranks a (left) and b (right): a's last column goes to b when a is heavier by more than that column
        // (the move then shrinks the difference), b's first column to a in the opposite case; only a slab of four or more columns gives one away (it may lose one on
        // either side in the same step).  Same numbers, same decision on both sides.
        auto decide = [](uint32_t nA, uint32_t nB, uint32_t edgeA, uint32_t edgeB, uint32_t colsA, uint32_t colsB) -> int {
            if (nA > nB && (uint64_t)(nA - nB) > (uint64_t)edgeA && colsA > 3u) return -1;
            if (nB > nA && (uint64_t)(nB - nA) > (uint64_t)edgeB && colsB > 3u) return +1;
            return 0;
        };
        if (hasL) D.pendL = decide(got[2], nOwn, got[0], D.edgeLEnd - D.ownB, got[3], cols);
        if (hasR) D.pendR = decide(nOwn, got[6], D.ownE - D.edgeRBegin, got[4], cols, got[7]);
    }
    const uint32_t ghostL = D.ownB, ghostR = D.nLocal - D.ownE;
    if ((hasL && h[4] != ghostL) || (hasR && h[5] != ghostR) || (!hasL && ghostL) || (!hasR && ghostR)) {
        char buf[256];
        snprintf(buf, sizeof buf, "rank %d: ghost ranges (%u, %u) do not match the neighbours' edge columns (%u, %u)", D.rank, ghostL, ghostR, hasL ? h[4] : 0u, hasR ? h[5] : 0u);
        return fail(VFD_E_NCCL, buf);
    }
    return VFD_OK;
}

// halo of one or two per-particle arrays (elemFloats floats per particle each)
int Solver::dist_halo(void* base, uint32_t elemFloats, void* base2, uint32_t elemFloats2) {
    Dist& D = *dist;
    const bool hasL = D.rank > 0, hasR = D.rank + 1 < D.nranks;
    if (D.p2p) {
        HaloJob J; memset(&J, 0, sizeof J);
        void* bases[2] = { base, base2 }; const uint32_t efs[2] = { elemFloats, elemFloats2 };
        uint64_t bytes = 0;
        const uint32_t items = (hasL ? D.edgeLEnd - D.ownB : 0u) + (hasR ? D.ownE - D.edgeRBegin : 0u);
        for (int k = 0; k < 2 && bases[k]; k++) {
            // which array of the slab this is (velocity changes sides with every sort, on every rank alike)
            const size_t myOff = (size_t)((unsigned char*)bases[k] - D.slab);
            int region = -1;
            for (int r = 0; r < SR_COUNT; r++) if (slab_offset(r, D.slabNp[D.rank]) == myOff) region = r;
            if (region < 0 || kRegionBytes[region] != efs[k] * 4u) return fail(VFD_E_INVALID, "halo exchange of an array outside the peer-memory slab");
            HaloArray& H = J.a[J.count++];
            H.eb = kRegionBytes[region];
            H.mine = (const unsigned char*)bases[k];
            H.dstL = hasL ? D.peerSlab[D.rank - 1] + slab_offset(region, D.slabNp[D.rank - 1]) + (size_t)D.leftOwnE * H.eb : nullptr;
            H.dstR = hasR ? D.peerSlab[D.rank + 1] + slab_offset(region, D.slabNp[D.rank + 1]) : nullptr;
            bytes += (uint64_t)items * H.eb;
        }
        const uint32_t blocks = std::max(1u, std::min<uint32_t>((items + VFD_TPB - 1) / VFD_TPB, (uint32_t)numSMs * 4u));
        k_halo_exchange<<<blocks, VFD_TPB, 0, stream>>>(J, D.ownB, D.edgeLEnd, D.edgeRBegin, D.ownE,
            (PeerCtl*)D.slab, hasL ? (PeerCtl*)D.peerSlab[D.rank - 1] : nullptr, hasR ? (PeerCtl*)D.peerSlab[D.rank + 1] : nullptr, dState);
        launches += 1;
        D.halos += 1;
        D.bytesHalo += bytes;
        return VFD_OK;
    }
    void* bases[2] = { base, base2 }; const uint32_t efs[2] = { elemFloats, elemFloats2 };
    NK(D.api.GroupStart());
    for (int k = 0; k < 2 && bases[k]; k++) {
        float* f = (float*)bases[k];
        const size_t e = efs[k];
        if (hasL) {
            if (D.edgeLEnd > D.ownB) NK(D.api.Send(f + D.ownB * e, (D.edgeLEnd - D.ownB) * e, ncclFloat32, D.rank - 1, D.comm, stream));
            if (D.ownB) NK(D.api.Recv(f, D.ownB * e, ncclFloat32, D.rank - 1, D.comm, stream));
        }
        if (hasR) {
            if (D.ownE > D.edgeRBegin) NK(D.api.Send(f + D.edgeRBegin * e, (D.ownE - D.edgeRBegin) * e, ncclFloat32, D.rank + 1, D.comm, stream));
            if (D.nLocal > D.ownE) NK(D.api.Recv(f + D.ownE * e, (D.nLocal - D.ownE) * e, ncclFloat32, D.rank + 1, D.comm, stream));
        }
        D.bytesHalo += (uint64_t)((D.edgeLEnd - D.ownB) * (hasL ? 1 : 0) + (D.ownE - D.edgeRBegin) * (hasR ? 1 : 0)) * e * 4;
    }
    NK(D.api.GroupEnd());
    D.halos += 1;
    return VFD_OK;
}

// combine the ranks' partial reduction results of one site and take the control decision on every rank
int Solver::dist_reduce(int site, bool isMax) {
    Dist& D = *dist;
    if (D.p2p) { D.reductions += 1; return VFD_OK; }        // all-reduced by the reducing kernel itself (control.cuh: peer_allreduce)
    NK(D.api.AllReduce(&dState->red[site * 2], &dState->red[site * 2], 2, ncclFloat64, isMax ? ncclMax : ncclSum, D.comm, stream));
    k_control<<<1, 32, 0, stream>>>(params, dState, site);
    launches += 1;
    D.reductions += 1;
    return VFD_OK;
}

// ---- frame capture of a decomposed run (DFSPHImplementation.cu:148-167 across ranks) ------------------------------------
// Every rank packs its owned particles as 40-byte records (persistent id + DFSPHParticleSimple); rank 0 receives them and
// scatters them by id into the frame (original particle order, what the reference's renderer indexes across frames:
// Scene.cpp:361-368), which then flows through rank 0's frame pipe like a single-GPU frame.  The other ranks bake nothing.
struct FrameRec { uint32_t id; float v[9]; };
static_assert(sizeof(FrameRec) == 40, "frame record");

__global__ void k_export_owned_frame(Arrays A, uint32_t ownB, uint32_t ownE, FrameRec* __restrict__ out) {
    const uint32_t p = ownB + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ownE) return;
    const float4 x = A.pos[p], v = A.vel[p], a = A.acc[p];
    FrameRec r;
    r.id = A.id[p];
    r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = v.x; r.v[4] = v.y; r.v[5] = v.z; r.v[6] = a.x; r.v[7] = a.y; r.v[8] = a.z;
    out[p - ownB] = r;
}

__global__ void k_scatter_frame(const FrameRec* __restrict__ in, uint32_t count, uint32_t nGlobal, VfdParticleSimple* __restrict__ out,
                                const DevState* __restrict__ S, float* __restrict__ meta) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && meta) { meta[0] = S->vmax2; meta[1] = S->dt; }
    if (i >= count) return;
    const FrameRec r = in[i];
    if (r.id >= nGlobal) return;
    VfdParticleSimple q;
    q.Position[0] = r.v[0]; q.Position[1] = r.v[1]; q.Position[2] = r.v[2];
    q.Velocity[0] = r.v[3]; q.Velocity[1] = r.v[4]; q.Velocity[2] = r.v[5];
    q.Acceleration[0] = r.v[6]; q.Acceleration[1] = r.v[7]; q.Acceleration[2] = r.v[8];
    out[r.id] = q;
}

int Solver::dist_capture_frame(bool metaFromDevice, float vmax2, float dt) {
    Dist& D = *dist;
    const uint32_t nOwn = D.ownE - D.ownB;
    if (!D.frameSend) CK(cudaMalloc(&D.frameSend, (size_t)std::max(D.capacity, 1u) * sizeof(FrameRec)));
    if (nOwn) { k_export_owned_frame<<<(nOwn + VFD_TPB - 1) / VFD_TPB, VFD_TPB, 0, stream>>>(arrays, D.ownB, D.ownE, (FrameRec*)D.frameSend); launches += 1; }
    // every rank's record count, on every rank
    uint32_t* d = D.dCounters + 32;             // [0] mine, [8..15] all
    CK(cudaMemcpyAsync(d, &nOwn, 4, cudaMemcpyHostToDevice, stream));
    NK(D.api.AllGather(d, d + 8, 1, ncclUint32, D.comm, stream));
    uint32_t cnt[8] = {};
    CK(cudaMemcpyAsync(cnt, d + 8, 4 * (size_t)D.nranks, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    uint64_t total = 0;
    for (int r = 0; r < D.nranks; r++) total += cnt[r];
    if (total != D.nGlobal) return fail(VFD_E_NCCL, "frame gather: the ranks' owned particles do not add up to the global count");
    if (D.rank == 0 && D.frameRecvCap < D.nGlobal) {
        if (D.frameRecv) cudaFree(D.frameRecv);
        D.frameRecv = nullptr;
        CK(cudaMalloc(&D.frameRecv, (size_t)D.nGlobal * sizeof(FrameRec)));
        D.frameRecvCap = D.nGlobal;
    }
    NK(D.api.GroupStart());
    if (D.rank == 0) {
        size_t off = cnt[0];
        for (int r = 1; r < D.nranks; r++) { if (cnt[r]) NK(D.api.Recv((FrameRec*)D.frameRecv + off, (size_t)cnt[r] * 10, ncclFloat32, r, D.comm, stream)); off += cnt[r]; }
    } else if (nOwn) {
        NK(D.api.Send(D.frameSend, (size_t)nOwn * 10, ncclFloat32, 0, D.comm, stream));
    }
    NK(D.api.GroupEnd());
    if (D.rank == 0) {
        if (nOwn) CK(cudaMemcpyAsync(D.frameRecv, D.frameSend, (size_t)nOwn * sizeof(FrameRec), cudaMemcpyDeviceToDevice, stream));
        VfdParticleSimple* dst = pipe.acquire();
        if (!dst) { cudaGetLastError(); return fail(VFD_E_CUDA, "frame capture: cannot allocate the frame ring buffers"); }
        k_scatter_frame<<<(D.nGlobal + VFD_TPB - 1) / VFD_TPB, VFD_TPB, 0, stream>>>((const FrameRec*)D.frameRecv, D.nGlobal, D.nGlobal, dst, dState, metaFromDevice ? pipe.meta_slot() : nullptr);
        launches += 1;
        CK(pipe.submit(stream, vmax2, dt, metaFromDevice));
    }
    return VFD_OK;
}

// the frame section of Solver::step for a rank of a decomposition: the same decisions on every rank (the time step is global)
int Solver::dist_frame_step() {
    if (frameIndexHost >= desc.FrameCount) return VFD_OK;
    if (desc.FrameLength <= 0.0f && state == VFD_STATE_SIMULATING) {
        RC_(dist_capture_frame(true, 0.0f, 0.0f));
        frameTimeHost = 0.0f;
        frameIndexHost++;
        return VFD_OK;
    }
    DevState s;
    RC_(read_state(s));
    frameTimeHost += s.dt;
    update_debug(s, false);
    if (frameTimeHost >= desc.FrameLength) {
        RC_(dist_capture_frame(false, s.vmax2, s.dt));
        frameTimeHost = 0.0f;
        frameIndexHost++;
        std::lock_guard<std::mutex> g(dbgMutex);
        debug.FrameTime = 0.0f; debug.FrameIndex = frameIndexHost;
    }
    return VFD_OK;
}

int Solver::dist_get_owned(uint32_t capacity, uint32_t* count, uint32_t* ids, VfdParticle* out) {
    CK(cudaSetDevice(device));
    if (!dist) return fail(VFD_E_INVALID, "not distributed");
    Dist& D = *dist;
    const uint32_t n = D.ownE - D.ownB;
    if (count) *count = n;
    if (!ids || !out) return VFD_OK;
    if (capacity < n) return fail(VFD_E_INVALID, "get_owned: capacity too small");
    if (!n) return VFD_OK;
    uint32_t* dIds = nullptr; VfdParticle* dOut = nullptr;
    CK(cudaMalloc(&dIds, (size_t)n * 4)); CK(cudaMalloc(&dOut, (size_t)n * sizeof(VfdParticle)));
    k_export_owned<<<(n + VFD_TPB - 1) / VFD_TPB, VFD_TPB, 0, stream>>>(params, arrays, D.ownB, D.ownE, dIds, dOut);
    launches += 1;
    cudaError_t e = cudaMemcpyAsync(ids, dIds, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dOut, (size_t)n * sizeof(VfdParticle), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(dIds); cudaFree(dOut);
    if (e != cudaSuccess) return fail_cuda(e, "get_owned", __LINE__);
    return VFD_OK;
}

Dist::~Dist() {
    for (int r = 0; r < 8; r++) if (r != rank && peerSlab[r]) cudaIpcCloseMemHandle(peerSlab[r]);
    if (slab) cudaFree(slab);
    if (comm && api.CommDestroy) api.CommDestroy(comm);
    cudaFree(sendL); cudaFree(sendR); cudaFree(recvL); cudaFree(recvR); cudaFree(dCounters); cudaFree(frameSend); cudaFree(frameRecv);
    if (hCounters) cudaFreeHost(hCounters);
}

} // namespace vfd
