#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t10_pytest.log 2>&1
tail -5 gpurun_out/t10_pytest.log
for mode in none serial parallel; do
  timeout 900 python scripts/sweep.py --dtype f64 --mode $mode --iters 4 --out gpurun_out/t10_f64_$mode.json > gpurun_out/t10_f64_$mode.log 2>&1
  python - <<PY
import json
d=json.load(open("gpurun_out/t10_f64_$mode.json"))
print("f64 $mode", " ".join("%d:%.3f" % (r["n"], r["ms"]) for r in d["rows"]))
PY
done
