"""TEST INFRASTRUCTURE: an independent restatement of the two simplest neighbour sums of the path — density and the DFSPH
factor (reference: Simulation/DFSPH/DFSPHKernels.cu:152-187 ComputeDensityKernel, :188-232 ComputeDFSPHFactorKernel; the
lookup Kernel/DFSPHKernels.h:43-80) — with the per-pair TERMS formed exactly as the reference forms them (fp32 differences,
glm::dot's (x + y) + z, IEEE sqrt, the midpoint table lookup) and the SUM over the neighbours taken in float64.

What it is for: the reference accumulates these sums in fp32 in neighbour-list order (an order that differs from thread to
thread count and from its CPU to its CUDA build), the CUDA path of this repo in another order; both are "the reference's
result" only up to summation round-off.  The float64 sum of the same fp32 terms is the value both approximate, so it says
how far each of them is from it (VERDICT round 1, weak point 3) — and, being written from the kernel's published formula
rather than compiled from the reference's sources, it pins the oracle's Density and Factor independently of the oracle.

Also `density_analytic`: the cubic spline itself in float64 (no table), the published SPH density; the table's
piecewise-constant lookup moves a term by up to ~1e-3 of itself, so this one is a coarse check (1e-3) that the table, the
volume and the rest density are the ones of the formula."""
import numpy as np

F = np.float32
RES = 10000


def parse_kernel(blob):
    """PrecomputedDFSPHCubicKernel as the reference lays it out (DFSPHKernels.h:126-134): W[RES], gradW[RES + 1], radius,
    radius^2, 1 / step, W(0), k, l."""
    f = np.frombuffer(np.ascontiguousarray(blob).tobytes(), np.float32)
    assert len(f) == 2 * RES + 7
    return {"W": f[:RES], "G": f[RES:2 * RES + 1], "radius": f[2 * RES + 1], "radius2": f[2 * RES + 2], "inv_step": f[2 * RES + 3],
            "w_zero": f[2 * RES + 4], "k": f[2 * RES + 5], "l": f[2 * RES + 6]}


def info_scalars(info_bytes):
    """Volume (offset 40) and Density0 (offset 44) of the 128-byte DFSPHSimulationInfo (include/vfd_dfsph.h)."""
    b = np.ascontiguousarray(info_bytes).tobytes()
    return np.frombuffer(b[40:44], np.float32)[0], np.frombuffer(b[44:48], np.float32)[0]


def _pairs(counts, offsets, ids):
    i = np.repeat(np.arange(len(counts), dtype=np.int64), counts.astype(np.int64))
    start = np.repeat(offsets.astype(np.int64), counts.astype(np.int64))
    k = np.arange(len(i), dtype=np.int64) - np.repeat(np.cumsum(counts.astype(np.int64)) - counts.astype(np.int64), counts.astype(np.int64))
    return i, ids[start + k].astype(np.int64)


def _r(d):
    """fp32 length the reference's way: glm::dot = (x x + y y) + z z, then sqrt."""
    r2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return r2.astype(F), np.sqrt(r2.astype(F))


def _lookup_w(K, d):
    r2, rl = _r(d)
    pos = np.minimum((rl * K["inv_step"]).astype(np.uint32), np.uint32(RES - 2))
    w = F(0.5) * (K["W"][pos] + K["W"][pos + 1])
    return np.where(r2 <= K["radius2"], w, F(0.0)).astype(F)


def _lookup_grad(K, d):
    _, rl = _r(d)
    pos = np.minimum((rl * K["inv_step"]).astype(np.uint32), np.uint32(RES - 2))
    g = (F(0.5) * (K["G"][pos] + K["G"][pos + 1]))[:, None] * d
    return np.where((rl <= K["radius"])[:, None], g, F(0.0)).astype(F)


def density_terms_f64(pos, counts, offsets, ids, bxj, bvol, K, volume, rho0):
    """rho0 x [ V W(0) + sum_j fl(V W_ij) + fl(V_b W_ib) ] with fp32 terms and a float64 sum."""
    pos = pos.astype(F)
    i, j = _pairs(counts, offsets, ids)
    t = (F(volume) * _lookup_w(K, pos[i] - pos[j])).astype(F)
    s = np.full(len(pos), float(F(volume) * K["w_zero"]), np.float64)
    np.add.at(s, i, t.astype(np.float64))
    bxj = np.where((bvol > 0)[:, None], bxj.astype(F), pos)           # (a body that is out of reach leaves no sample: whatever is stored there is not read)
    tb = np.where(bvol > 0, bvol.astype(F) * _lookup_w(K, pos - bxj), F(0.0)).astype(F)
    return (s + tb.astype(np.float64)) * float(rho0)


def factor_terms_f64(pos, counts, offsets, ids, bxj, bvol, K, volume, eps=1.0e-6):
    """1 / (sum_j |g_j|^2 + |sum_j g_j + g_b|^2), g_j = -V gradW_ij in fp32, the sums in float64 (0 where the sum is <= EPS)."""
    pos = pos.astype(F)
    i, j = _pairs(counts, offsets, ids)
    g = (-F(volume) * _lookup_grad(K, pos[i] - pos[j])).astype(F)
    gg = ((g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1]) + g[:, 2] * g[:, 2]).astype(F)
    spk = np.zeros(len(pos), np.float64)
    np.add.at(spk, i, gg.astype(np.float64))
    gi = np.zeros((len(pos), 3), np.float64)
    np.add.at(gi, i, -g.astype(np.float64))
    bxj = np.where((bvol > 0)[:, None], bxj.astype(F), pos)
    gb = np.where((bvol > 0)[:, None], -bvol.astype(F)[:, None] * _lookup_grad(K, pos - bxj), F(0.0)).astype(F)
    gi -= gb.astype(np.float64)
    tot = spk + (gi * gi).sum(1)
    return np.where(tot > eps, 1.0 / np.where(tot > eps, tot, 1.0), 0.0)


def density_analytic(pos, counts, offsets, ids, bxj, bvol, h, volume, rho0):
    """The cubic spline W(q) = 8 / (pi h^3) { 6 q^3 - 6 q^2 + 1 (q <= 1/2), 2 (1 - q)^3 (q <= 1) } in float64, no table."""
    def W(r):
        q = r / h
        k = 8.0 / (np.pi * h ** 3)
        return np.where(q <= 0.5, k * (6 * q ** 3 - 6 * q ** 2 + 1), np.where(q <= 1.0, k * 2 * (1 - q) ** 3, 0.0))
    p = pos.astype(np.float64)
    i, j = _pairs(counts, offsets, ids)
    s = np.full(len(p), volume * W(np.zeros(1))[0], np.float64)
    np.add.at(s, i, volume * W(np.sqrt(((p[i] - p[j]) ** 2).sum(1))))
    s += np.where(bvol > 0, bvol.astype(np.float64) * W(np.sqrt(((p - bxj.astype(np.float64)) ** 2).sum(1))), 0.0)
    return s * rho0
