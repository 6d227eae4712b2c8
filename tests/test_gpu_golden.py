"""GPU parity against the committed golden fixtures (outputs of the reference's own solver sources,
tests/golden/make_golden.py): one solver step from an identical, evolved full state.

Tolerance (north star): neighbour sets bit-exact; per-step outputs within 1e-5 relative in fp32.  "Relative"
is measured against the field's scale (max |reference|): every output is a sum of ~30 signed fp32 terms, and
the reference itself is not reproducible below that level (its neighbour order and reduction order change
from run to run: SURVEY.md Appendix E)."""
import os

import numpy as np
import pytest

import parity
import scenes

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1.0e-5


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def field_tolerances(g):
    """1e-5 of the field's scale, or three times the reference's own run-to-run deviation on this very step where that is
    larger (recorded in the fixture by make_golden.py: the reference is not reproducible below it)."""
    noise = dict(zip([str(f) for f in g["noise_fields"]], g["noise_values"]))
    return {f: max(TOL, 3.0 * float(noise.get(f, 0.0))) for f in parity.ALL_FIELDS}


def make_sim(name, g, fma=0):
    from vfd_b200 import api
    sc = scenes.SCENES[name]
    desc = api.DFSPHSimulationDescription(FrameCount=0, **sc["desc"])
    sim = api.DFSPHSimulation(desc)
    sim.set_option(api.VFD_OPT_SEARCH_FMA, fma)
    sim.SetFluidObjects([api.FluidObject(g["pos0"])])
    vm = api.VolumeMap(g["map_domain_min"], g["map_domain_max"], g["map_resolution"], g["map_cell_size"], g["map_cell_size_inv"],
                       int(g["map_field_count"]), int(g["map_node_count"]), int(g["map_cell_count"]), int(g["map_cell_map_count"]),
                       g["map_nodes"], g["map_cells"], g["map_cell_map"])
    sim.SetRigidBodies([vm])
    return sim


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_one_step_from_golden_state(name, lib_built):
    g = load(name)
    sim = make_sim(name, g)
    sim.set_particles_full(g["state_in"])
    sim.set_time_step(float(g["dt_in"]))
    sim.set_surface_tension_state(int(g["st_in"][0]), float(g["st_in"][1]))
    sim.OnUpdate()
    out = sim.particles()
    dbg = sim.GetDebugInfo()

    # neighbour sets: bit-exact (host-compiler arithmetic of the distance test, like the CPU build of the reference)
    mism = parity.neighbor_mismatches(sim.neighbors(), (g["nbr_counts"], g["nbr_offsets"], g["nbr_ids"]))
    assert not mism, "neighbour sets differ for %d particles, e.g. %s" % (len(mism), mism[:5])

    errs = parity.field_errors(out, g["state_out"])
    print("\n[%s] one step from golden state, n = %d\n%s" % (name, len(out), parity.format_errors(errs)))
    print("  dt %.9g vs %.9g   its (div, press, visc) = (%d, %d, %d) vs %s" % (
        sim.GetCurrentTimeStepSize(), float(g["dt_out"]), dbg.DivergenceSolverIterationCount, dbg.PressureSolverIterationCount,
        dbg.ViscositySolverIterationCount, g["its_out"]))
    assert abs(sim.GetCurrentTimeStepSize() - float(g["dt_out"])) <= 1e-6 * float(g["dt_out"])
    assert dbg.DivergenceSolverIterationCount == g["its_out"][0]
    assert dbg.PressureSolverIterationCount == g["its_out"][1]
    assert abs(int(dbg.ViscositySolverIterationCount) - int(g["its_out"][2])) <= 1
    tol = field_tolerances(g)
    bad = parity.beyond_tolerance(errs, tol, out, g["state_out"])
    assert not bad, "fields beyond tolerance %s:\n%s" % ({k: "%.1e" % tol[k] for k in bad}, parity.format_errors(bad))


@pytest.mark.parametrize("name", ["dfsph", "viscous"])
def test_boundary_samples_match_golden(name, lib_built):
    g = load(name)
    sim = make_sim(name, g)
    sim.set_particles_full(g["state_in"])
    sim.set_time_step(float(g["dt_in"]))
    sim.OnUpdate()
    xj, vol = sim.boundary(0)
    assert np.array_equal(vol > 0, g["boundary_vol"] > 0)
    s = max(np.abs(g["boundary_vol"]).max(), 1e-30)
    assert np.abs(vol - g["boundary_vol"]).max() / s < TOL
    assert np.abs(xj - g["boundary_xj"]).max() / np.abs(g["boundary_xj"]).max() < TOL
