#!/bin/bash
# usage: tools_gpu_check.sh ["ENV=.." ...] — smoke + parity tests under tight timeouts (a hung kernel must not eat the GPU budget), then short benches
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 gpurun_out/smoke.log
[ $rc -ne 0 ] && exit 1
timeout 500 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_tune.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 4 gpurun_out/pytest_tune.log
[ $rc -ne 0 ] && exit 1
timeout 400 bash tools_gpu_tune.sh --no-tests "$@"
