#!/bin/bash
# round 2, pass i (2 GPUs): sharded GPU tests, then the north_star multi-GPU workload (TwoDNavierStokes forced 16384^2) and 4096^2/8192^2 lines
N=${1:-2}
mkdir -p gpurun_out
T=r2i
true


for wl in ${WLS:-twodns16384 sqg4096 twodns8192}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --workload $wl --steps 8 --warmup 3 --repeats 3 --e2e-steps 1 > gpurun_out/${T}_${wl}_${N}gpu.json 2> gpurun_out/${T}_${wl}_${N}gpu.err
  cat gpurun_out/${T}_${wl}_${N}gpu.err | tail -5 | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_${wl}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$wl N=$N ms/step %.3f value %.3e e2e_ms %.2f setup %.1fs"%(d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"],d["config"]["setup_s"]),{k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, d["config"]["state_finite"], d.get("nvlink"))
except Exception as e: print("no result", e)
PY
done
nvidia-smi --query-gpu=memory.used --format=csv | head -3
