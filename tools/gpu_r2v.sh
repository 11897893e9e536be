#!/bin/bash
# r2v: K1 k_col_inv_ws helper variants at 4096: per column (0), 16-row batches (4), pair form (16), pair + x64 TMEM load (24), pair + register twiddles (17)
mkdir -p gpurun_out
T=r2v
for v in 0 16 4 24 17; do
  GFB_K1WS_VAR=$v timeout 300 python bench.py --steps 20 --warmup 5 --repeats 5 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_var$v.json 2> gpurun_out/${T}_var$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_var$v.json").read().strip().splitlines()[-1])
    print("K1WS_VAR=$v sqg4096 ms/step %.4f"%d["ms_per_step"],{k:round(x["ms_per_step"],4) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
except Exception as e: print("no result", e); print(open("gpurun_out/${T}_var$v.err").read()[-600:])
PY
done
GFB_K1WS_VAR=16 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -p no:cacheprovider -k "4096 or config4 or sqg" 2>&1 | tail -3
