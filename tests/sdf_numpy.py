"""TEST INFRASTRUCTURE: float32 numpy restatement of the reference's SDF::Interpolate (value only; Utility/SDF/SDF.cu:141-186
with ShapeFunction :403-470 and the cell -> node table of AddFunction :81-131), operation for operation, vectorised over
points.  Used on the CPU to check that the grids, node values and lattice candidates the GPU sampler is built from select
the reference's samples; the GPU kernels themselves interpolate with csrc/volume_map.cuh."""
import numpy as np

F = np.float32
FLT_MAX = np.finfo(np.float32).max


def cell_table(res):
    nx, ny, nz = (int(r) for r in res)
    l = np.arange(nx * ny * nz, dtype=np.int64)
    k, t = l // (ny * nx), l % (ny * nx)
    j, i = t // nx, t % nx
    c = np.zeros((len(l), 32), np.int64)
    nv = (nx + 1) * (ny + 1) * (nz + 1)
    nex, ney = nx * (ny + 1) * (nz + 1), (nx + 1) * ny * (nz + 1)
    for b in range(8):
        c[:, b] = (nx + 1) * (ny + 1) * (k + ((b >> 2) & 1)) + (nx + 1) * (j + ((b >> 1) & 1)) + i + (b & 1)
    off = nv
    for b in range(4):
        bz, by = b & 1, (b >> 1) & 1
        c[:, 8 + 2 * b] = off + 2 * (nx * (ny + 1) * (k + bz) + nx * (j + by) + i)
        c[:, 9 + 2 * b] = c[:, 8 + 2 * b] + 1
    off += 2 * nex
    for b in range(4):
        bx, bz = b & 1, (b >> 1) & 1
        c[:, 16 + 2 * b] = off + 2 * (ny * (nz + 1) * (i + bx) + ny * (k + bz) + j)
        c[:, 17 + 2 * b] = c[:, 16 + 2 * b] + 1
    off += 2 * ney
    for b in range(4):
        by, bx = b & 1, (b >> 1) & 1
        c[:, 24 + 2 * b] = off + 2 * (nz * (nx + 1) * (j + by) + nz * (i + bx) + k)
        c[:, 25 + 2 * b] = c[:, 24 + 2 * b] + 1
    return c


def shape_functions(xi):
    """32 cubic serendipity shape functions at local coordinates xi (n, 3) float32 -> (n, 32) float32."""
    x, y, z = xi[:, 0], xi[:, 1], xi[:, 2]
    one, three, nine = F(1.0), F(3.0), F(9.0)
    x2, y2, z2 = x * x, y * y, z * z
    X = [one - x, one + x]; Y = [one - y, one + y]; Z = [one - z, one + z]
    T3x = [one - three * x, one + three * x]; T3y = [one - three * y, one + three * y]; T3z = [one - three * z, one + three * z]
    N = np.zeros((len(x), 32), F)
    fac = F(1.0 / 64.0) * (nine * (x2 + y2 + z2) - F(19.0))
    for j in range(8):
        bx, by, bz = j & 1, (j >> 1) & 1, (j >> 2) & 1
        N[:, j] = fac * (X[bx] * Y[by]) * Z[bz]
    fx, fy, fz = F(9.0 / 64.0) * (one - x2), F(9.0 / 64.0) * (one - y2), F(9.0 / 64.0) * (one - z2)
    for j in range(8, 16):
        t, bz, by = j & 1, (j >> 1) & 1, (j >> 2) & 1
        N[:, j] = (fx * T3x[t]) * (Y[by] * Z[bz])
    for j in range(16, 24):
        t, bx, bz = j & 1, (j >> 1) & 1, (j >> 2) & 1
        N[:, j] = (fy * T3y[t]) * (X[bx] * Z[bz])
    for j in range(24, 32):
        t, by, bx = j & 1, (j >> 1) & 1, (j >> 2) & 1
        N[:, j] = (fz * T3z[t]) * (X[bx] * Y[by])
    return N


def interpolate(dmin, dmax, res, cell, cell_inv, nodes0, points, cells=None):
    """phi at points (n, 3) float32; FLT_MAX outside the domain."""
    p = np.ascontiguousarray(points, F)
    dmin, dmax, cell, cell_inv = (np.asarray(a, F) for a in (dmin, dmax, cell, cell_inv))
    res = np.asarray(res, np.int64)
    inside = np.all((dmin <= p) & (dmax >= p), axis=1)
    out = np.full(len(p), FLT_MAX, F)
    q = p[inside]
    mi = ((q - dmin) * cell_inv).astype(np.int64)           # float -> unsigned truncation (the values are >= 0)
    mi = np.minimum(mi, res - 1)
    ci = res[1] * res[0] * mi[:, 2] + res[0] * mi[:, 1] + mi[:, 0]
    lo = dmin + mi.astype(F) * cell
    hi = lo + cell
    den = hi - lo
    xi = (F(2.0) / den) * q - (hi + lo) / den
    N = shape_functions(xi.astype(F))
    if cells is None:
        cells = cell_table(res)
    c = np.asarray(nodes0, F)[cells[ci]]
    phi = np.zeros(len(q), F)
    for k in range(32):
        phi = phi + c[:, k] * N[:, k]
    out[inside] = phi
    return out
