"""CPU-side checks of the C-ABI library (no compute calls): it loads, exports every symbol include/vfd_dfsph.h
declares, its structs have the reference's layouts, the host-built lookup tables are bit-identical to the
reference's, and there is no CPU fallback."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vfd_dfsph.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vfd_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib_built):
    L = C.CDLL(lib_built)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_binding_covers_the_header(lib_built):
    from vfd_b200 import api
    bound = {s[0] for s in api.SYMBOLS}
    assert set(declared_symbols()) == bound


def test_struct_layouts_match_the_reference(lib_built):
    from vfd_b200 import api
    # DFSPHSimulationInfo is 128 B with `bool TemporalSmoothing` at 96 (Structures/DFSPHSimulationInfo.h:5-44)
    assert C.sizeof(api.VfdDfsphInfo) == 128
    assert api.VfdDfsphInfo.TemporalSmoothing.offset == 96
    assert api.VfdDfsphInfo.SurfaceTensionSampleCount.offset == 84
    assert api.VfdDfsphInfo.MonteCarloFactor.offset == 112
    # DFSPHParticle 120 B, DFSPHParticleSimple 36 B (Structures/DFSPHParticle.h:8-32, DFSPHParticleSimple.h:8-13)
    assert api.PARTICLE_DTYPE.itemsize == 120 and api.PARTICLE_DTYPE.fields["VelocityDifference"][1] == 72
    assert api.PARTICLE_DTYPE.fields["DeltaFinalCurvature"][1] == 116
    assert api.PARTICLE_SIMPLE_DTYPE.itemsize == 36
    assert C.sizeof(api.DFSPHSimulationDescription) == 116


def test_default_description_is_the_references(lib_built):
    from vfd_b200 import api
    d = api.DFSPHSimulationDescription().as_dict()
    # Structures/DFSPHSimulationDescription.h:9-50
    assert d["TimeStepSize"] == pytest.approx(0.001) and d["MaxTimeStepSize"] == pytest.approx(0.005)
    assert d["FrameCount"] == 200 and d["MinPressureSolverIterations"] == 0 and d["MaxPressureSolverIterations"] == 100
    assert d["Viscosity"] == 10.0 and d["BoundaryViscosity"] == 10.0 and d["CSD"] == 10000 and d["CSDFix"] == -1
    assert d["ParticleRadius"] == pytest.approx(0.025) and d["Gravity"] == pytest.approx((0.0, -9.81, 0.0))


def test_kernel_tables_bit_identical_to_the_reference(lib_built):
    """The golden fixture holds the reference's PrecomputedDFSPHCubicKernel bytes: W[10000], gradW[10001], then
    radius, radius2, invStep, k, l, wZero in struct order (Kernel/DFSPHKernels.h:136-143)."""
    from vfd_b200 import api
    k = np.load(os.path.join(GOLDEN, "dfsph.npz"))["kernel"]
    W, G, sc = api.kernel_tables(0.1)
    f = k.view(np.float32)
    assert np.array_equal(f[:10000], W)
    assert np.array_equal(f[10000:20001], G)
    tail = f[20001:20007]
    assert sorted(tail.tolist()) == sorted(sc.tolist()), (tail, sc)
    # known answers: h = 0.1 -> W(0) = 2546.47876
    assert sc[3] == np.float32(2546.47876)


def test_halton_table_bit_identical_to_the_reference(lib_built):
    from vfd_b200 import api
    ref = json.load(open(os.path.join(GOLDEN, "halton_ref.json")))
    t = api.halton_table()
    assert t.size == ref["count"]
    for i, v in zip(ref["probe_index"], ref["probe_value"]):
        assert t[i] == np.float32(v)
    assert hashlib.sha256(t.astype("<f4").tobytes()).hexdigest() == ref["sha256_f32le"]
    p = t.reshape(-1, 3).astype(np.float64)
    assert np.abs((p * p).sum(1) - 1.0).max() < 1e-6      # unit sphere


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from vfd_b200 import api
    with pytest.raises(api.VfdError) as e:
        api.DFSPHSimulation()
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(api.VfdError):
        api.VolumeMap.build_box((0, 0, 0), (1, 1, 1))
