#!/bin/bash
# round 2, pass x (1 GPU): GPU suite + smoke at HEAD, default bench (cuFFT + CPU baselines), launch list, ncu of the default kernels,
# full-size lines of the other configurations
mkdir -p gpurun_out
T=r2x
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=8 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -14 gpurun_out/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
tail -3 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.err
timeout 600 python bench.py --steps 100 --e2e-steps 30 > gpurun_out/${T}_bench100.json 2> gpurun_out/${T}_bench100.err
timeout 300 python bench.py --workload twodns4096 --steps 20 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${T}_bench_twodns4096.json 2> gpurun_out/${T}_bench_twodns4096.err
timeout 600 python bench.py --workload twodns16384 --steps 4 --warmup 3 --repeats 3 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_twodns16384_1gpu.json 2> gpurun_out/${T}_twodns16384_1gpu.err
python - <<'PY'
import json
for tag in ("bench","bench100","bench_twodns4096","twodns16384_1gpu"):
    try:
        d=json.loads(open("gpurun_out/r2x_%s.json"%tag).read().strip().splitlines()[-1])
        print(tag, "ms/step %.4f"%d["ms_per_step"], "observed %.4f"%d["ms_per_step_observed_every_step"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()},
              d["config"]["state_finite"], "e2e", d["e2e"]["ms_per_step"], "cufft", (d.get("cufft_baseline") or {}).get("ms_per_step"),
              "cpu", (d.get("cpu_baseline") or {}).get("ms_per_step"), "roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "step", d["step_roofline"]["frac_of_measured_peak"])
    except Exception as e:
        print(tag, "no result", e); print(open("gpurun_out/r2x_%s.err"%tag).read()[-1500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_col_inv_ws|k_row_product_tma|k_col_fwd' -s 36 -c 3 \
  -o gpurun_out/${T}_hot python bench.py --steps 4 --warmup 3 --repeats 1 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference.json 2> gpurun_out/${T}_reference.err
tail -c 400 gpurun_out/${T}_reference.json
