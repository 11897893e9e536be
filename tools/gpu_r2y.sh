#!/bin/bash
# r2y: diagnose test_nonpow2_ny_vs_oracle[32-250-FilteredAB3] (MultiLayerQG part, 2.6e-9 on the r2x box) and run the rest of the suite without -x
mkdir -p gpurun_out
timeout 600 python tools/diag_nonpow2.py > gpurun_out/r2y_diag.txt 2>&1
tail -12 gpurun_out/r2y_diag.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/r2y_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -12 gpurun_out/r2y_pytest.log | cut -c1-200
