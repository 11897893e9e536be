#!/bin/bash
mkdir -p gpurun_out
T=r2g
timeout 120 ./tools/ws_fft_bench > gpurun_out/${T}_ws_fft_bench.txt 2>&1; cat gpurun_out/${T}_ws_fft_bench.txt
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2"
for v in 0 1 2 4 8; do
  GFB_K1WS_VAR=$v GFB_NO_ROW_WS=1 GFB_NO_COL_WS3=1 timeout 300 $B > gpurun_out/${T}_k1var$v.json 2> gpurun_out/${T}_k1var$v.err
done
python - <<'PY'
import json
for v in (0,1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/r2g_k1var%d.json"%v).read().strip().splitlines()[-1])
        print("K1WS var", v, "ms/step %.4f"%d["ms_per_step"], {k:round(x["ms_per_step"],4) for k,x in d["kernels"].items()}, d["config"]["state_finite"])
    except Exception as e:
        print(v, "no result", e); print(open("gpurun_out/r2g_k1var%d.err"%v).read()[-800:])
PY
