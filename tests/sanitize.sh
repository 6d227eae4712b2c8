#!/bin/bash
# compute-sanitizer over the smoke step (800 particles: DFSPH + viscosity + surface tension, every pipeline kernel) and a
# 27k-particle step with the tile queue, ring wrap-around and reduction records in play.
# usage (GPU box): bash tests/sanitize.sh [out-dir]      -> <out-dir>/sanitizer_{memcheck,racecheck,initcheck,synccheck}.log + summary
OUT=${1:-gpurun_out}
mkdir -p $OUT
cat > /tmp/vfd_sanitize_case.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import __graft_entry__ as g
g.smoke()
if os.environ.get("VFD_SANITIZE_BIG"):
    from vfd_b200 import api
    R, D = 0.025, 0.05
    pos = api.block_positions(30, 30, 30, R, origin=(4 * D, 4 * D, 4 * D))
    pos = (pos + np.random.RandomState(3).uniform(-0.2 * R, 0.2 * R, pos.shape)).astype(np.float32)
    box = ((0.0, 0.0, 0.0), (46 * D, 42 * D, 38 * D))
    vm = api.VolumeMap.build_box(box[0], box[1], inverted=True, padding=0.0, resolution=(12, 12, 10), particle_radius=R)
    sim = api.DFSPHSimulation(api.DFSPHSimulationDescription(FrameCount=3, CSDFix=16, MinPressureSolverIterations=2, MaxPressureSolverIterations=2,
                                                             MinDivergenceSolverIterations=2, MaxDivergenceSolverIterations=2))
    sim.SetFluidObjects([api.FluidObject(pos, velocities=np.random.RandomState(4).uniform(-0.3, 0.3, pos.shape).astype(np.float32))])
    sim.SetRigidBodies([vm])
    sim.Simulate()                      # three steps, the default frame length: the device-side frame decision runs too
    print("27k case: frames", sim.GetFrameCount(), "PCG iterations", sim.GetDebugInfo().ViscositySolverIterationCount)
    sim.close()
PY
for tool in memcheck racecheck initcheck synccheck; do
  big=1; [ $tool == racecheck ] && big=""          # racecheck replays every shared-memory access: the 800-particle step only
  VFD_SANITIZE_BIG=$big timeout 900 compute-sanitizer --tool $tool --print-limit 30 python /tmp/vfd_sanitize_case.py > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_$tool.log | tail -n 1)"
done
grep -hE "ERROR SUMMARY|RACECHECK SUMMARY|smoke:|27k case" $OUT/sanitizer_*.log > $OUT/sanitizer_summary.txt
cat $OUT/sanitizer_summary.txt
