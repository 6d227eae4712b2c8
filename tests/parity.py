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


# Fields that are a clamped *difference* of O(1) quantities: PressureResiduum = min(1 - DensityAdvection - dt^2 * sum, 0)
# (reference: DFSPHKernels.cu:329-337).  Its values are ~1e-3 while its operands are ~1, so one ulp of an operand (any
# change of summation order, including the reference's own from run to run) is already 3e-5 of the residuum's scale.
# Such a field is held to 1e-6 of its OPERAND's scale (ten times tighter than the 1e-5 gate on the operand itself).
CANCELLING = {"PressureResiduum": "DensityAdvection"}


def beyond_tolerance(errs, tol, out, ref):
    """Fields whose error exceeds tol[field] (relative to the field's scale); a cancelling field also passes when its
    absolute error is within 1e-6 of its operand's scale."""
    bad = {}
    for f, v in errs.items():
        if v[0] <= tol[f]:
            continue
        op = CANCELLING.get(f)
        if op is not None:
            abs_err = np.abs(np.asarray(out[f], np.float64) - np.asarray(ref[f], np.float64)).max()
            if abs_err <= 1e-6 * np.abs(np.asarray(ref[op], np.float64)).max():
                continue
        bad[f] = v
    return bad


def format_errors(errs):
    return "\n".join("  %-34s rel-to-scale %.3e   elementwise %.3e   (scale %.4g)" % (k, v[0], v[1], v[2]) for k, v in errs.items())


def csr_rows(counts, offsets, ids):
    return [np.sort(ids[int(o):int(o) + int(c)]) for c, o in zip(counts, offsets)]


def neighbor_mismatches(a, b):
    """Indices of particles whose neighbour *sets* differ between two CSR lists (a, b = (counts, offsets, ids))."""
    ra, rb = csr_rows(*a), csr_rows(*b)
    return [i for i, (x, y) in enumerate(zip(ra, rb)) if len(x) != len(y) or not np.array_equal(x, y)]
