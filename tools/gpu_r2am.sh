#!/bin/bash
# r2am: the committed tree after the K2-half twiddle work: GPU suite, smoke, 16384^2 one-GPU line
mkdir -p gpurun_out
T=r2am
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=3 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log | cut -c1-160
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --workload twodns16384 --steps 4 --warmup 3 --repeats 3 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/${T}_twodns16384_1gpu.json 2> gpurun_out/${T}_twodns16384_1gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2am_twodns16384_1gpu.json").read().strip().splitlines()[-1])
print("twodns16384 N=1 ms/step %.3f value %.4e"%(d["ms_per_step"], d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, d["config"]["state_finite"], d["config"]["kernels"])
PY
