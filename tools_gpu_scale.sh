#!/bin/bash
# usage: tools_gpu_scale.sh "8 4 2" [bench args] — the N-GPU bench line for each N (run under gpurun --gpus max(N))
mkdir -p gpurun_out
for N in $1; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "${@:2}" > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "N=$N rc=$?"
  python - $N <<'PY'
import json, sys
f = "gpurun_out/scale_n%s.json" % sys.argv[1]
try:
    d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
    c = d["config"]
    print("N=%s ms/step %.3f value %.4g pcg %s owned %s slab0 %s per-step %s" % (sys.argv[1], d["ms_per_step"], d["value"], c.get("pcg_iterations_last_step"), c.get("owned_per_rank"), c.get("slab_rank0_now"), c.get("per_step_rank0")))
    k = c.get("kernels_rank0", {})
    print("   rank0 kernel ms/step: total %.3f | %s" % (sum(v["ms_per_step"] for v in k.values()), "  ".join("%s %.3f" % (n, k[n]["ms_per_step"]) for n in list(k)[:6])))
except Exception as e:
    print("failed", repr(e)); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
done
