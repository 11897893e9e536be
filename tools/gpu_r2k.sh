#!/bin/bash
# round 2, pass k (1 GPU): whole GPU suite incl. the new modules/steppers and the full-size configs; bench lines of configs 1-2
mkdir -p gpurun_out
T=r2k
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=12 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -22 gpurun_out/${T}_pytest.log
for wl in slqg1024 mqg2_1024 mqg5_2048; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/${T}_bench_$wl.json 2> gpurun_out/${T}_bench_$wl.err
  tail -2 gpurun_out/${T}_bench_$wl.err
done
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
for tag in ("bench_slqg1024","bench_mqg2_1024","bench_mqg5_2048","bench"):
    try:
        d=json.loads(open("gpurun_out/r2k_%s.json"%tag).read().strip().splitlines()[-1])
        print(tag, "ms/step %.4f"%d["ms_per_step"], "value %.3e"%d["value"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()}, "frac", round(d["step_roofline"]["frac_of_measured_peak"],3))
    except Exception as e:
        print(tag, "no result", e); print(open("gpurun_out/r2k_%s.err"%tag).read()[-1500:])
PY
