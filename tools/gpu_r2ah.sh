#!/bin/bash
# r2ah: K1 at ny = 16384 with sub-lines 1 and 2 staged by cp.async behind the previous transform: on / off, parity tests
mkdir -p gpurun_out
T=r2ah
timeout 400 python -m pytest tests/test_zz_gpu_16384.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for v in 1 0; do
  GFB_K1_16K_STAGE=$v timeout 600 python bench.py --workload twodns16384 --steps 3 --warmup 3 --repeats 2 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_stage$v.json 2> gpurun_out/${T}_stage$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_stage$v.json").read().strip().splitlines()[-1])
    print("stage=$v twodns16384 N=1 ms/step %.3f"%d["ms_per_step"],{k:round(x["ms_per_step"],3) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
except Exception as e: print("no result", e); print(open("gpurun_out/${T}_stage$v.err").read()[-600:])
PY
done
