#!/bin/bash
# look-ahead validation on one GPU: parity of the dense tail, then the bench with one and with two streams
cd /root/repo
timeout 600 python -m pytest tests/test_z_c_driver.py tests/test_gpu_echelonize.py -x -q -m gpu -k "deferred or mid_tail or bit_exact or dense_tail or low_rank" 2>&1 | tail -4
for mode in two one; do
  if [ $mode = one ]; then export SPASM_B200_ONE_STREAM=1; else unset SPASM_B200_ONE_STREAM; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu --e2e-steps 1 > gpurun_out/la_$mode.json 2> gpurun_out/la_$mode.err
  echo "mode=$mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/la_$mode.json").read().strip().splitlines()[-1])
    print("$mode", "value", d["value"], "e2e", d["e2e"]["value"], "steps", d.get("step_s"), "tail", d["phases_last_step_s"]["tail"], "gemm_ms", d["roofline"].get("kernel_ms_per_step"), "achieved", d["roofline"]["achieved"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/la_$mode.err").read()[-1500:])
PY
done
