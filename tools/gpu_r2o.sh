#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2o_smi.txt
./tools/tlb_probe > gpurun_out/r2o_tlb_probe.txt 2>&1
cat gpurun_out/r2o_tlb_probe.txt
