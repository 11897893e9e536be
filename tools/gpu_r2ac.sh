#!/bin/bash
# r2ac: bench.py at HEAD after the last edits (default line), the FilteredRK4 / RK4 forms of the 4096^2 step, the cufft arm as its own line
mkdir -p gpurun_out
T=r2ac
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 300 gpurun_out/${T}_bench.err
for st in FilteredRK4 RK4 FilteredETDRK4; do
  timeout 300 python bench.py --stepper $st --steps 20 --e2e-steps 2 > gpurun_out/${T}_bench_$st.json 2> gpurun_out/${T}_bench_$st.err
done
timeout 300 python bench.py --impl cufft --steps 10 > gpurun_out/${T}_cufft.json 2> gpurun_out/${T}_cufft.err
python - <<'PY'
import json
for tag in ("bench","bench_FilteredRK4","bench_RK4","bench_FilteredETDRK4","cufft"):
    try:
        d=json.loads(open("gpurun_out/r2ac_%s.json"%tag).read().strip().splitlines()[-1])
        print(tag, "ms/step %.4f"%d["ms_per_step"], d["config"]["workload"][:60], {k:round(v["ms_per_step"],4) for k,v in d.get("kernels",{}).items()},
              "step frac", d.get("step_roofline",{}).get("frac_of_measured_peak"), "K1 fn", d.get("roofline",{}).get("all_kernels",{}).get("k_col_inv",{}).get("function"), d.get("roofline",{}).get("all_kernels",{}).get("k_col_inv",{}).get("traffic"))
    except Exception as e:
        print(tag, "no result", e); print(open("gpurun_out/r2ac_%s.err"%tag).read()[-800:])
PY
