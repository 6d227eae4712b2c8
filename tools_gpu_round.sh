#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench line, the ncu launch list and one `ncu --set full` capture
# of a step's neighbour-sum kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -s --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --ncu > gpurun_out/ncu_launches.log 2>&1; echo "ncu-launches rc=$?" >> gpurun_out/ncu_launches.log
# one launch of each neighbour-sum kernel class (-c caps the capture; --import-source makes the report ~3 MB per launch, and
# gpurun_out/ is only copied back below 64 MiB: the text summary is written here on the box, the report is dropped if too big)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_visc_matvec|k_visc_setup|k_density_factor|k_source|k_pressure_accel|k_solve_iteration|k_build_list|k_st_classify|k_st_smooth|k_visc_update" \
    -c 14 -f -o gpurun_out/full_step python bench.py --steps 1 --warmup 3 --ncu --ncu-visc-it 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu-full rc=$?" >> gpurun_out/ncu_full.log
python tools_ncu_summary.py gpurun_out/full_step.ncu-rep gpurun_out/ncu_full_summary.txt gpurun_out/ncu_traffic.json 1000000 > /dev/null 2>&1
[ "$(du -sm gpurun_out | cut -f1)" -gt 55 ] && rm -f gpurun_out/full_step.ncu-rep
tail -n 5 gpurun_out/pytest.log gpurun_out/smoke.log gpurun_out/bench.err gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
cat gpurun_out/bench.json | cut -c1-1500
ls -la gpurun_out
