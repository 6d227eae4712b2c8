#!/bin/bash
# usage: tools_gpu_tune.sh [--no-tests] "ENV1=a ENV2=b" "ENV1=c" ...   — parity tests once, then one short bench per environment setting
mkdir -p gpurun_out
if [ "$1" == "--no-tests" ]; then shift; else
  timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_tune.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_tune.log
fi
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_tune$i.json 2> gpurun_out/bench_tune$i.err || tail -n 3 gpurun_out/bench_tune$i.err
  python - "$cfg" gpurun_out/bench_tune$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    k = d['config']['kernels']
    print('[%s] ms/step %.3f pcg %d | %s' % (sys.argv[1], d['ms_per_step'], d['config']['pcg_iterations_last_step'],
          '  '.join('%s %.4f' % (n, k[n]['avg_ms_active']) for n in ('visc_matvec', 'visc_update', 'visc_direction', 'div_solve', 'press_accel', 'div_source', 'density_factor', 'visc_setup', 'search_build_list', 'st_classify', 'st_smooth') if n in k)))
except Exception as e:
    print('[%s] failed: %r' % (sys.argv[1], e))
PY
done
