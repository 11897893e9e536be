#!/bin/bash
# r2ag: last check of the committed tree on hardware: GPU suite, smoke, default bench line
mkdir -p gpurun_out
T=r2ag
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=4 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log | cut -c1-160
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ag_bench.json").read().strip().splitlines()[-1])
print("bench ms/step %.4f e2e %.3f value %.4e frac %.3f step %.3f launches %d clocks %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d["roofline"]["frac"], d["step_roofline"]["frac_of_measured_peak"], d["gpu_launches"], d["clocks"]))
PY
