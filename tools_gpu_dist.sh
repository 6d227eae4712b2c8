#!/bin/bash
# usage: tools_gpu_dist.sh N [bench args]  — N-rank correctness check, then the N-GPU bench line with peer memory and with NCCL only
N=${1:-2}; shift
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run tests/dist_check.py 6 > gpurun_out/dc_p2p.log 2>&1; echo "dist_check p2p rc=$?"; grep -E "DIST_CHECK|PCG iterations|owned per rank" gpurun_out/dc_p2p.log
VFD_DIST_P2P=0 run tests/dist_check.py 6 > gpurun_out/dc_nccl.log 2>&1; echo "dist_check nccl rc=$?"; grep -E "DIST_CHECK|PCG iterations" gpurun_out/dc_nccl.log
run bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/bench_n${N}_p2p.json 2> gpurun_out/bench_n${N}_p2p.err; echo "bench p2p rc=$?"
VFD_DIST_P2P=0 run bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "bench nccl rc=$?"
python - $N <<'PY'
import json, sys
for tag in ("p2p", "nccl"):
    f = "gpurun_out/bench_n%s_%s.json" % (sys.argv[1], tag)
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(tag, "ms/step %.3f  value %.4g  pcg %s  %s" % (d["ms_per_step"], d["value"], d["config"].get("pcg_iterations_last_step"), d["config"].get("per_step_rank0")))
    except Exception as e:
        print(tag, "failed", repr(e)); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
