#!/bin/bash
# round 2: TMA-staged K2 + warp-specialised K1/K3 on hardware: parity, then A/B benches
mkdir -p gpurun_out
T=r2d
timeout 900 python -m pytest tests/test_abi.py tests/test_golden.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2"
timeout 300 $B > gpurun_out/${T}_bench_all.json 2> gpurun_out/${T}_bench_all.err
GFB_NO_COL_WS3=1 timeout 300 $B > gpurun_out/${T}_bench_noWS3.json 2> gpurun_out/${T}_bench_noWS3.err
GFB_NO_COL_WS=1 GFB_NO_COL_WS3=1 timeout 300 $B > gpurun_out/${T}_bench_tmaonly.json 2> gpurun_out/${T}_bench_tmaonly.err
GFB_NO_ROW_TMA=1 GFB_NO_COL_WS=1 GFB_NO_COL_WS3=1 timeout 300 $B > gpurun_out/${T}_bench_none.json 2> gpurun_out/${T}_bench_none.err
python - <<'PY'
import json
for tag in ("all","noWS3","tmaonly","none"):
    try:
        d=json.loads(open("gpurun_out/r2d_bench_%s.json"%tag).read().strip().splitlines()[-1])
        print(tag, "ms/step %.4f"%d["ms_per_step"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()}, d["config"]["state_finite"])
    except Exception as e:
        print(tag, "no result", e); print(open("gpurun_out/r2d_bench_%s.err"%tag).read()[-1500:])
PY
