"""Synthetic scenes shared by the golden generator and the parity tests:
a lattice block of particles inside an inverted-box rigid body."""
import numpy as np

R = 0.025


def jitter(pos, seed, amp=0.2 * R):
    """Deterministic jitter so that no pair sits exactly at the support radius."""
    rng = np.random.RandomState(seed)
    return (pos + rng.uniform(-amp, amp, pos.shape)).astype(np.float32)


def block(nx, ny, nz, origin):
    d = np.float32(2.0 * R)
    i = (np.arange(nx, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[0])
    j = (np.arange(ny, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[1])
    k = (np.arange(nz, dtype=np.float32) + np.float32(0.5)) * d + np.float32(origin[2])
    Z, Y, X = np.meshgrid(k, j, i, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)


# name -> (block dims, block origin, box min, box max, map resolution, description overrides)
SCENES = {
    # config 1 of BASELINE.json in miniature: DFSPH only, pinned Jacobi iteration counts
    "dfsph": dict(dims=(10, 10, 10), origin=(0.15, 0.10, 0.15), box=((0, 0, 0), (0.8, 0.9, 0.8)), res=(8, 8, 8),
                  desc=dict(EnableViscositySolver=0, EnableSurfaceTensionSolver=0,
                            MinPressureSolverIterations=2, MaxPressureSolverIterations=2,
                            MinDivergenceSolverIterations=2, MaxDivergenceSolverIterations=2)),
    # reference defaults (zero Jacobi iterations execute) but without viscosity / surface tension
    "dfsph_default": dict(dims=(10, 10, 10), origin=(0.15, 0.10, 0.15), box=((0, 0, 0), (0.8, 0.9, 0.8)), res=(8, 8, 8),
                          desc=dict(EnableViscositySolver=0, EnableSurfaceTensionSolver=0)),
    # config 2 in miniature: honey-like implicit viscosity (PCG), incl. boundary friction
    "viscous": dict(dims=(10, 8, 10), origin=(0.15, 0.08, 0.15), box=((0, 0, 0), (0.8, 0.9, 0.8)), res=(8, 8, 8),
                    desc=dict(EnableSurfaceTensionSolver=0,
                              MinPressureSolverIterations=2, MaxPressureSolverIterations=2,
                              MinDivergenceSolverIterations=2, MaxDivergenceSolverIterations=2)),
    # config 3 in miniature: DFSPH + viscosity + surface tension
    "full": dict(dims=(10, 8, 10), origin=(0.15, 0.08, 0.15), box=((0, 0, 0), (0.8, 0.9, 0.8)), res=(8, 8, 8),
                 desc=dict(MinPressureSolverIterations=2, MaxPressureSolverIterations=2,
                           MinDivergenceSolverIterations=2, MaxDivergenceSolverIterations=2, CSDFix=24)),
}


def scene_positions(name, seed=1234):
    s = SCENES[name]
    return jitter(block(*s["dims"], s["origin"]), seed)
