#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cufft --e2e-steps 2 --repeats 3"
for v in 0 2; do
  GFB_K3_L2PF=$v timeout 300 $B > gpurun_out/r2m_pf$v.json 2> gpurun_out/r2m_pf$v.err
done
python - <<'PY'
import json
for v in (0,2):
    try:
        d=json.loads(open("gpurun_out/r2m_pf%d.json"%v).read().strip().splitlines()[-1])
        print("K3 L2PF", v, "ms/step %.4f"%d["ms_per_step"], "observed %.4f"%d["ms_per_step_observed_every_step"], {k:round(x["ms_per_step"],4) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
    except Exception as e:
        print(v, "no result", e); print(open("gpurun_out/r2m_pf%d.err"%v).read()[-800:])
PY
GFB_K3_L2PF=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "4096 or config4 or config5" 2>&1 | tail -3
