#!/bin/bash
# r2u: start stagger of the persistent 4096 kernels (K1 k_col_inv_ws, K2 k_row_product_tma): does chip-wide lock step cost anything?
mkdir -p gpurun_out
T=r2u
for cfg in "0 0" "8000 0" "0 8000" "8000 8000" "30000 30000" "2000 2000"; do
  set -- $cfg
  GFB_STAGGER_K1=$1 GFB_STAGGER_K2=$2 timeout 300 python bench.py --steps 20 --warmup 5 --repeats 5 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_k1s$1_k2s$2.json 2> gpurun_out/${T}_k1s$1_k2s$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_k1s$1_k2s$2.json").read().strip().splitlines()[-1])
    print("stagger K1=$1 K2=$2 sqg4096 ms/step %.4f"%d["ms_per_step"],{k:round(x["ms_per_step"],4) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
except Exception as e: print("no result", e); print(open("gpurun_out/${T}_k1s$1_k2s$2.err").read()[-600:])
PY
done
