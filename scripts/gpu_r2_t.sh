#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_pytest.log 2>&1
tail -4 gpurun_out/t_pytest.log
for mode in none serial parallel; do
  python scripts/sweep.py --dtype f32 --mode $mode --batch 1000000 --iters 4 --ns 24,25,26,27,28,29,30,31,32 > gpurun_out/t_sweep_f32_$mode.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/t_sweep_f32_$mode.log"):
    try: d=json.loads(l)
    except Exception: continue
    print("$mode", d["n"], round(d["ms"],3), round(d["frac_measured_peak"],3))
PY
done
