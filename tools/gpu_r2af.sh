#!/bin/bash
# r2af: precomputed inter-pass twiddles in the TMA row kernel (K2) at 4096: on / off
mkdir -p gpurun_out
T=r2af
for v in 1 0 1 0; do
  GFB_K2_PRETW=$v timeout 300 python bench.py --steps 20 --warmup 5 --repeats 5 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_pretw$v.json 2> gpurun_out/${T}_pretw$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_pretw$v.json").read().strip().splitlines()[-1])
    print("K2_PRETW=$v sqg4096 ms/step %.4f"%d["ms_per_step"],{k:round(x["ms_per_step"],4) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
except Exception as e: print("no result", e); print(open("gpurun_out/${T}_pretw$v.err").read()[-600:])
PY
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "4096 or config4" 2>&1 | tail -3
