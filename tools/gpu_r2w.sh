#!/bin/bash
# r2w: ncu --set full of the pair-form K1 (k_col_inv_ws<4096,16>) alone, for the per-line stall picture of its two groups
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_col_inv_ws' -s 12 -c 1 \
  -o gpurun_out/r2w_k1 python bench.py --steps 4 --warmup 3 --repeats 1 --no-cpu-baseline --no-cufft --e2e-steps 1 > gpurun_out/r2w_ncu.log 2>&1
tail -2 gpurun_out/r2w_ncu.log
