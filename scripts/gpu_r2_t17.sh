#!/bin/bash
mkdir -p gpurun_out
for mode in none serial parallel; do
  timeout 900 python scripts/sweep.py --dtype f32 --mode $mode --ns 8,12,16 --ab --iters 6 --out gpurun_out/t17_f32_$mode.json > gpurun_out/t17_f32_$mode.log 2>&1
  python - <<PY
import json
d=json.load(open("gpurun_out/t17_f32_$mode.json"))
print("f32 $mode", " ".join("%d:%.3f/%.3f" % (r["n"], r["ms"], r["ms_lsu_staging"]) for r in d["rows"]))
PY
done
