#!/bin/bash
# r2t: K1 at ny = 16384: load batch size (2 rows / all 16 rows before first use) x start stagger (chip-wide lock step)
mkdir -p gpurun_out
T=r2t
for cfg in "2 0" "2 16000" "16 16000" "2 50000" "16 50000"; do
  set -- $cfg
  GFB_K1_16K_LB=$1 GFB_K1_16K_STAGGER=$2 timeout 600 python bench.py --workload twodns16384 --steps 3 --warmup 3 --repeats 2 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_lb$1_s$2.json 2> gpurun_out/${T}_lb$1_s$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_lb$1_s$2.json").read().strip().splitlines()[-1])
    print("lb=$1 stagger=$2 twodns16384 N=1 ms/step %.3f"%d["ms_per_step"],{k:round(x["ms_per_step"],3) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
except Exception as e: print("no result", e); print(open("gpurun_out/${T}_lb$1_s$2.err").read()[-600:])
PY
done
