"""The oracle's density and DFSPH factor against an independent restatement (tests/exact_sums.py): the same fp32 per-pair
terms, summed in float64, on the four committed fixtures — the reference's results lie within fp32 summation round-off of the
exactly summed value, and within the kernel table's discretisation of the analytic cubic spline.  (The hardware twin of this
file, tests/test_gpu_y_exact_sums.py, measures the same distance for the CUDA path.)"""
import os

import numpy as np
import pytest

import exact_sums as E
import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def inputs(g):
    K = E.parse_kernel(g["kernel"])
    V, rho0 = E.info_scalars(g["info_out"])
    return K, V, rho0, g["state_in"]["Position"], (g["nbr_counts"], g["nbr_offsets"], g["nbr_ids"])


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_reference_density_is_the_exact_sum_up_to_fp32_roundoff(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    K, V, rho0, pos, nb = inputs(g)
    exact = E.density_terms_f64(pos, *nb, g["boundary_xj"], g["boundary_vol"], K, V, rho0)
    ref = g["state_out"]["Density"].astype(np.float64)
    err = np.abs(ref - exact).max() / np.abs(exact).max()
    # <= 60 fp32 additions of same-signed terms: a few 1e-7 of the sum (measured 1.9e-7 ... 3.4e-7)
    assert err < 1.0e-6, err
    # and the published formula itself (float64 spline, no table): the table's midpoint lookup is piecewise constant in r
    analytic = E.density_analytic(pos, *nb, g["boundary_xj"], g["boundary_vol"], float(K["radius"]), float(V), float(rho0))
    assert np.abs(ref - analytic).max() / np.abs(analytic).max() < 1.0e-3


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_reference_factor_is_the_exact_sum_up_to_fp32_roundoff(name):
    """Factor leaves the step scaled by 1/dt (DFSPHKernels.cu:526), dt (:601) and 1/dt_new^2 (:274): three more roundings."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    K, V, rho0, pos, nb = inputs(g)
    exact = E.factor_terms_f64(pos, *nb, g["boundary_xj"], g["boundary_vol"], K, V)
    ref = g["state_out"]["Factor"].astype(np.float64) * float(g["dt_out"]) ** 2
    live = exact > 0
    assert np.array_equal(live, ref > 0)
    rel = np.abs(ref[live] - exact[live]) / exact[live]
    assert rel.max() < 2.0e-6, rel.max()            # measured 4e-7
