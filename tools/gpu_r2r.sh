#!/bin/bash
# r2r (N GPUs): sharded 16384 parity test, then the north_star multi-GPU workload with the lane-paired K1
N=${1:-2}
mkdir -p gpurun_out
T=r2r
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -p no:cacheprovider -k "16384 or p2p" 2>&1 | tail -4
for wl in ${WLS:-twodns16384}; do
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
