#!/bin/bash
# A/B of the tile pass variants on one box: parity tests, then the bench line per variant (env VFD_TILE_PIPELINE).
mkdir -p gpurun_out
for mode in "$@"; do
  export VFD_TILE_PIPELINE=$mode
  timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$mode.log 2>&1; echo "pytest[$mode] rc=$?" | tee -a gpurun_out/pytest_$mode.log
  tail -n 3 gpurun_out/pytest_$mode.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench[$mode] rc=$?"
  tail -n 2 gpurun_out/bench_$mode.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$mode.json'))
    print('[$mode] value %.3e  ms/step %.3f  pcg %d  launches %d slow-path %d'%(d['value'], d['ms_per_step'], d['config']['pcg_iterations_last_step'], d['gpu_launches'], d['config']['tile_passes_on_slow_path']))
    for k,v in list(d['config']['kernels'].items())[:14]:
        print('  %-20s %8.4f ms/step  x%-5.1f avg %.4f ms  %6s GB/s  frac %s'%(k, v['ms_per_step'], v['launches_per_step'], v['avg_ms_active'], v['algo_GBps'], v['frac_of_peak']))
except Exception as e: print('no bench', e)
PY
done
