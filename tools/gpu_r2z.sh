#!/bin/bash
# r2z: after the S0/ALT ping-pong of the layered one-stage steppers: the diagnostic again, the whole GPU suite, a launch list that
# really contains the step kernels (filtered by name), the new 16384^2 invariants test
mkdir -p gpurun_out
T=r2z
timeout 600 python tools/diag_nonpow2.py > gpurun_out/${T}_diag.txt 2>&1
tail -9 gpurun_out/${T}_diag.txt | cut -c1-180
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -12 gpurun_out/${T}_pytest.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(col|row|diag|repack|dealias|matvec)' -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_ncu_launches.log 2>&1
grep -c "k_col\|k_row" gpurun_out/${T}_launches.csv
