#!/bin/bash
# round 2: TMA-staged K2 on hardware: parity (4096^2 tests + C ABI test), then A/B bench
mkdir -p gpurun_out
T=r2d
timeout 600 python -m pytest tests/test_abi.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "abi or c_program or 4096 or config5 or native" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${T}_bench_tma.json 2> gpurun_out/${T}_bench_tma.err
GFB_NO_ROW_TMA=1 GFB_NO_COL_WS=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${T}_bench_notma.json 2> gpurun_out/${T}_bench_notma.err
GFB_NO_COL_WS=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${T}_bench_tmaonly.json 2> gpurun_out/${T}_bench_tmaonly.err
python - <<'PY'
import json
for tag in ("tma","tmaonly","notma"):
    try:
        d=json.loads(open("gpurun_out/r2d_bench_%s.json"%tag).read().strip().splitlines()[-1])
        print(tag, "ms/step %.4f"%d["ms_per_step"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()}, d["config"]["state_finite"])
    except Exception as e:
        print(tag, "no result", e); print(open("gpurun_out/r2d_bench_%s.err"%tag).read()[-1500:])
PY
