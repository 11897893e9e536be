#!/bin/bash
mkdir -p gpurun_out
T=r2f
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2"
timeout 300 $B > gpurun_out/${T}_bench_all.json 2> gpurun_out/${T}_bench_all.err
GFB_NO_ROW_WS=1 timeout 300 $B > gpurun_out/${T}_bench_noRowWS.json 2> gpurun_out/${T}_bench_noRowWS.err
python - <<'PY'
import json
for tag in ("all","noRowWS"):
    try:
        d=json.loads(open("gpurun_out/r2f_bench_%s.json"%tag).read().strip().splitlines()[-1])
        print(tag, "ms/step %.4f"%d["ms_per_step"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()}, d["config"]["state_finite"])
    except Exception as e:
        print(tag, "no result", e); print(open("gpurun_out/r2f_bench_%s.err"%tag).read()[-1500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_col_inv_ws|k_row_product_ws|k_col_fwd_ws' -s 36 -c 3 \
  -o gpurun_out/${T}_ws python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -p no:cacheprovider -k "4096 or config5 or golden" 2>&1 | tail -3
