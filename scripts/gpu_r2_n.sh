#!/bin/bash
# round 2: bulk-copy staged kernel in the product -- parity tests, then A/B sweeps (bulk vs LSU staging) of the sizes it serves
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1
tail -15 gpurun_out/n_pytest.log
NS=5,6,7,9,10,11,13,14,15,17,18,19,21,22,23,25,26,27,29,30,31
for mode in none serial parallel; do
  timeout 600 python scripts/sweep.py --dtype f32 --mode $mode --ns $NS --ab --iters 4 --out gpurun_out/n_ab_f32_$mode.json > gpurun_out/n_ab_f32_$mode.log 2>&1
  python - <<PY
import json
d=json.load(open("gpurun_out/n_ab_f32_$mode.json"))
print("$mode", " ".join("%d:%.3f/%.3f" % (r["n"], r["ms"], r["ms_lsu_staging"]) for r in d["rows"]))
PY
done
