import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity, scenes
from vfd_b200 import api
from test_gpu_golden import make_sim, load
np.set_printoptions(linewidth=200, precision=7)
for name in scenes.SCENES:
    g = load(name)
    sim = make_sim(name, g)
    sim.set_particles_full(g["state_in"]); sim.set_time_step(float(g["dt_in"]))
    sim.set_surface_tension_state(int(g["st_in"][0]), float(g["st_in"][1]))
    sim.OnUpdate()
    out = sim.particles()
    xj, vol = sim.boundary(0)
    gv, gx = g["boundary_vol"], g["boundary_xj"]
    print("==", name, "n", len(out), "boundary particles ours/golden", (vol > 0).sum(), (gv > 0).sum())
    print("  vol max abs err", np.abs(vol - gv).max(), "of", np.abs(gv).max(), " xj max abs err", np.abs(xj - gx).max())
    d = np.abs(out["Density"] - g["state_out"]["Density"])
    worst = np.argsort(-d)[:8]
    print("  density worst idx", worst, "err", d[worst], "vol", gv[worst], "our vol", vol[worst])
    print("  density err among vol==0:", d[gv == 0].max() if (gv == 0).any() else None, " among vol>0:", d[gv > 0].max() if (gv > 0).any() else None)
    c, o, ids = sim.neighbors()
    print("  nbr count of worst", c[worst], "golden", g["nbr_counts"][worst])
    for f in ["PressureRho2V", "Factor", "DensityAdvection", "PressureRho2", "PressureAcceleration", "Velocity"]:
        a = np.asarray(out[f], np.float64); b = np.asarray(g["state_out"][f], np.float64)
        e = np.abs(a - b); e = e.max(axis=1) if e.ndim > 1 else e
        w = np.argsort(-e)[:5]
        print("  %-22s worst idx %s err %s vol %s" % (f, w, e[w], gv[w]))
