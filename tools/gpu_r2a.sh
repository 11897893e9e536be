#!/bin/bash
# round 2, first hardware pass at HEAD: GPU suite (per-test timeouts), smoke, bench, ncu launch list, sanitizer on the small cases
mkdir -p gpurun_out
T=r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt
timeout 1000 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=15 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -25 gpurun_out/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
tail -3 gpurun_out/${T}_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 1500 gpurun_out/${T}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_golden.py -m gpu -q -p no:cacheprovider > gpurun_out/${T}_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log
tail -5 gpurun_out/${T}_memcheck.log
