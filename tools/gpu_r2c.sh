#!/bin/bash
# round 2: ncu --set full of the three hot kernels at HEAD (pruned stage launches), + TMA bench
mkdir -p gpurun_out
T=r2c
timeout 120 ./tools/tma_bench > gpurun_out/${T}_tma_bench.txt 2>&1
cat gpurun_out/${T}_tma_bench.txt
# launches per step: 12 (K1,K2,K3 x 4 stages) after 1 import; skip warm-up 3 steps = 1 + 36 launches, capture the next 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_col_inv_single|k_row_product|k_col_fwd' -s 36 -c 3 \
  -o gpurun_out/${T}_hot python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
ls -la gpurun_out/${T}_hot.ncu-rep
