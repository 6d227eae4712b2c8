"""Comparison helpers for the parity tests."""
import numpy as np

VEC_FIELDS = ["Position", "Velocity", "Acceleration", "PressureAcceleration", "VelocityDifference",
              "MonteCarloSurfaceNormal", "MonteCarloSurfaceNormalSmooth"]
SCALAR_FIELDS = ["PressureResiduum", "Density", "DensityAdvection", "PressureRho2", "PressureRho2V", "Factor",
                 "MonteCarloSurfaceCurvature", "MonteCarloSurfaceCurvatureSmooth", "DeltaFinalCurvature"]
ALL_FIELDS = VEC_FIELDS + SCALAR_FIELDS


def field_errors(a, b, fields=ALL_FIELDS):
    """Per field: max |a-b| / max |b|  (error relative to the field's scale; 0/0 -> 0) and the worst
    element-wise relative error among elements that are not tiny against that scale."""
    out = {}
    for f in fields:
        x = np.asarray(a[f], np.float64)
        y = np.asarray(b[f], np.float64)
        scale = np.abs(y).max()
        d = np.abs(x - y)
        rel_scale = 0.0 if scale == 0.0 and d.max() == 0.0 else d.max() / max(scale, 1e-300)
        big = np.abs(y) > 1e-3 * scale if scale > 0 else np.zeros_like(y, bool)
        elem = (d[big] / np.abs(y[big])).max() if big.any() else 0.0
        out[f] = (rel_scale, elem, scale)
    return out


def format_errors(errs):
    return "\n".join("  %-34s rel-to-scale %.3e   elementwise %.3e   (scale %.4g)" % (k, v[0], v[1], v[2]) for k, v in errs.items())


def csr_rows(counts, offsets, ids):
    return [np.sort(ids[int(o):int(o) + int(c)]) for c, o in zip(counts, offsets)]


def neighbor_mismatches(a, b):
    """Indices of particles whose neighbour *sets* differ between two CSR lists (a, b = (counts, offsets, ids))."""
    ra, rb = csr_rows(*a), csr_rows(*b)
    return [i for i, (x, y) in enumerate(zip(ra, rb)) if len(x) != len(y) or not np.array_equal(x, y)]
