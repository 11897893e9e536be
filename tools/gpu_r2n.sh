#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
  GFB_K1_16K_PAIR=$v timeout 600 python bench.py --workload twodns16384 --steps 3 --warmup 3 --repeats 2 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/r2n_16k_pair$v.json 2> gpurun_out/r2n_16k_pair$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n_16k_pair$v.json").read().strip().splitlines()[-1])
    print("pair=$v twodns16384 N=1 ms/step %.3f"%d["ms_per_step"],{k:round(x["ms_per_step"],3) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
except Exception as e: print("no result", e); print(open("gpurun_out/r2n_16k_pair$v.err").read()[-600:])
PY
done
timeout 300 python -m pytest tests/test_zz_gpu_16384.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
