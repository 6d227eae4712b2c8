#!/bin/bash
# usage: tools_ncu.sh <kernel-regex> <out-name> [skip] [count]  — one `ncu --set full` capture of matching launches of a bench step
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$1" -s ${3:-4} -c ${4:-1} -f -o gpurun_out/$2 \
    python bench.py --steps 1 --warmup 3 --ncu --ncu-visc-it 8 > gpurun_out/ncu_$2.log 2>&1
tail -n 3 gpurun_out/ncu_$2.log | cut -c1-300
ls -la gpurun_out | grep $2
