#!/bin/bash
# round 2: bulk-copy staged kernel, fp64 included -- full parity suite, then A/B sweeps (library choice vs LSU staging)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t2_pytest.log 2>&1
tail -5 gpurun_out/t2_pytest.log
for dt in f64 f32; do
for mode in none serial parallel; do
  timeout 900 python scripts/sweep.py --dtype $dt --mode $mode --ab --iters 4 --out gpurun_out/t2_ab_${dt}_$mode.json > gpurun_out/t2_ab_${dt}_$mode.log 2>&1
  python - <<PY
import json
d=json.load(open("gpurun_out/t2_ab_${dt}_$mode.json"))
print("$dt $mode", " ".join("%d:%.3f/%.3f" % (r["n"], r["ms"], r["ms_lsu_staging"]) for r in d["rows"]))
PY
done
done
