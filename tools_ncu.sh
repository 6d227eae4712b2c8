#!/bin/bash
# ncu: launch list of 2 steps + full capture of the PCG mat-vec
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --ncu > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_visc_matvec -s 5 -c 2 -f -o gpurun_out/prof_matvec python bench.py --steps 1 --warmup 3 --ncu > gpurun_out/ncu_matvec.log 2>&1
ls -la gpurun_out
