#!/bin/bash
N=${1:-2}
cd /root/repo
export SPASM_B200_PROFILE=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 4 --warmup 2 --e2e-steps 1 > gpurun_out/la3_n$N.json 2> gpurun_out/la3_n$N.err
echo "bench rc=$?"
grep "^\[dense\]" gpurun_out/la3_n$N.err | tail -$N
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/la3_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "e2e", d["e2e"]["value"], "steps", d.get("step_s"), "phases", {k:round(v,3) for k,v in d["phases_last_step_s"].items() if k in ("total","greedy","density","tail")}, "gemm_ms", d["roofline"].get("kernel_ms_per_step"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/la3_n$N.err").read()[-2500:])
PY
