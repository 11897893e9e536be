#!/bin/bash
# usage: tools/multigpu_bench.sh N TAG   -- the two sharded benchmarks on N GPUs of this box (results under gpurun_out/)
N=$1; TAG=$2
mkdir -p gpurun_out
for wl in sqg4096 twodns8192; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --workload $wl --steps 8 --warmup 3 --e2e-steps 1 > gpurun_out/${TAG}_${wl}_${N}gpu.json 2> gpurun_out/${TAG}_${wl}_${N}gpu.err
  tail -2 gpurun_out/${TAG}_${wl}_${N}gpu.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${wl}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$wl N=$N ms/step %.3f value %.3e e2e_ms %.2f"%(d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"]),{k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, d["config"]["state_finite"])
except Exception as e: print("no result", e)
PY
done
