#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 4 gpurun_out/pytest.log; tail -n 2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.3e  ms/step %.3f  pcg %d  launches %d'%(d['value'], d['ms_per_step'], d['config']['pcg_iterations_last_step'], d['gpu_launches']))
for k,v in list(d['config']['kernels'].items())[:16]:
    print('  %-20s %8.4f ms/step  x%-5.1f avg %.4f ms  %6s GB/s  frac %s'%(k, v['ms_per_step'], v['launches_per_step'], v['avg_ms_active'], v['algo_GBps'], v['frac_of_peak']))
PY
