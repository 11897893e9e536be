#!/bin/bash
# r2q: lane-paired K1 (k_col_inv16k_dit) and K3 (k_col_fwd16k_lp) at ny = 16384: parity tests, bench A/B, ncu of the two new kernels
mkdir -p gpurun_out
T=r2q
timeout 300 python -m pytest tests/test_zz_gpu_16384.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for v in 1 0; do
  GFB_K3_16K_LP=$v timeout 600 python bench.py --workload twodns16384 --steps 3 --warmup 3 --repeats 2 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_16k_lp$v.json 2> gpurun_out/${T}_16k_lp$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_16k_lp$v.json").read().strip().splitlines()[-1])
    print("lp=$v twodns16384 N=1 ms/step %.3f"%d["ms_per_step"],{k:round(x["ms_per_step"],3) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
except Exception as e: print("no result", e); print(open("gpurun_out/${T}_16k_lp$v.err").read()[-600:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_col_inv16k_dit|k_col_fwd16k_lp' -s 8 -c 2 \
  -o gpurun_out/${T}_16k python bench.py --workload twodns16384 --steps 1 --warmup 3 --repeats 1 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
