#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_visc_matvec" -s 4 -c 1 -f -o gpurun_out/prof_tpp2 python bench.py --steps 1 --warmup 3 --ncu > gpurun_out/ncu_tile.log 2>&1
ls -la gpurun_out | grep prof
