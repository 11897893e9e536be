#!/bin/bash
# round 2, pass j (1 GPU): the 16384^2 forced TwoDNavierStokes problem on ONE GPU (46 GB): kernel times without NVLink, ncu of the 16k kernels
mkdir -p gpurun_out
T=r2j
timeout 900 python bench.py --workload twodns16384 --steps 4 --warmup 3 --repeats 2 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_twodns16384_1gpu.json 2> gpurun_out/${T}_twodns16384_1gpu.err
tail -3 gpurun_out/${T}_twodns16384_1gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2j_twodns16384_1gpu.json").read().strip().splitlines()[-1])
    print("twodns16384 N=1 ms/step %.3f value %.3e setup %.1fs"%(d["ms_per_step"],d["value"],d["config"]["setup_s"]),{k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, d["config"]["state_finite"], d["step_roofline"]["frac_of_measured_peak"])
except Exception as e: print("no result", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_col_inv16k|k_row_product|k_col_fwd16k' -s 12 -c 3 \
  -o gpurun_out/${T}_16k python bench.py --workload twodns16384 --steps 1 --warmup 3 --repeats 1 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "bqgql" 2>&1 | tail -5
