#!/bin/bash
# r2ae: compute-sanitizer memcheck and racecheck over small problems that reach every round-2 kernel family
mkdir -p gpurun_out
timeout 300 python tools/sanitize_cases.py > gpurun_out/r2ae_plain.log 2>&1; echo "rc=$?" >> gpurun_out/r2ae_plain.log; tail -7 gpurun_out/r2ae_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_cases.py > gpurun_out/r2ae_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2ae_memcheck.log; tail -4 gpurun_out/r2ae_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_cases.py > gpurun_out/r2ae_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2ae_racecheck.log; tail -4 gpurun_out/r2ae_racecheck.log
