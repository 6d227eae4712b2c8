"""What the compiled library contains (cuobjdump on the in-tree .so; no GPU needed): the neighbour passes really are the
asynchronous sm_100a pipeline — cp.async (LDGSTS), mbarrier (SYNCS), bulk copies (UBLKCP) — and nothing was built for
another architecture."""
import re
import shutil
import subprocess

import pytest


def sass(lib):
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not installed")
    return subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def functions(text):
    out, name = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name is not None:
            out[name].append(line)
    return {k: "\n".join(v) for k, v in out.items()}


def test_library_is_sm_100a_only(lib_built):
    r = subprocess.run(["cuobjdump", "-lelf", lib_built], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout if shutil.which("cuobjdump") else ""
    if not r:
        pytest.skip("cuobjdump not installed")
    archs = set(re.findall(r"sm_\d+a?", r))
    assert archs == {"sm_100a"}, archs


def test_neighbour_passes_are_the_asynchronous_pipeline(lib_built):
    fn = functions(sass(lib_built))
    pipe = {k: v for k, v in fn.items() if re.search(r"k_visc_matvec_pipe|k_solve_iteration|k_source|k_pressure_accel|k_density_factor|k_visc_setup|k_st_classify|k_st_smooth", k)}
    assert len(pipe) >= 14, sorted(pipe)
    for name, body in pipe.items():
        assert "SYNCS" in body, "%s: no mbarrier operations" % name                      # full/empty stage barriers
        assert "LDGSTS" in body or "UBLKCP" in body, "%s: no asynchronous global->shared copies" % name
    assert any("UBLKCP" in b for b in pipe.values())                                     # TMA bulk copies (one-payload passes)
    assert any("LDGSTS" in b for b in pipe.values())
