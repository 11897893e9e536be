#!/bin/bash
# r2aa (N GPUs): sharded GPU tests (incl. the 16384-line cases), then the north_star multi-GPU workload and the strong-scaled 4096^2 headline problem
N=${1:-2}
mkdir -p gpurun_out
T=${TAG:-r2aa}
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/${T}_pytest_sharded_${N}.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${T}_pytest_sharded_${N}.log
  tail -8 gpurun_out/${T}_pytest_sharded_${N}.log | cut -c1-200
fi
for wl in ${WLS:-twodns16384 sqg4096}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --workload $wl --steps 8 --warmup 3 --repeats 3 --e2e-steps 1 > gpurun_out/${T}_${wl}_${N}gpu.json 2> gpurun_out/${T}_${wl}_${N}gpu.err
  tail -3 gpurun_out/${T}_${wl}_${N}gpu.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_${wl}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$wl N=$N ms/step %.3f value %.3e e2e_ms %.2f setup %.1fs"%(d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"],d["config"]["setup_s"]),{k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, d["config"]["state_finite"], d.get("nvlink"), d.get("strong_scaling"))
except Exception as e: print("no result", e)
PY
done
