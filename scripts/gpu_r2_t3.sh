#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t3_pytest.log 2>&1
tail -3 gpurun_out/t3_pytest.log
timeout 900 python scripts/sweep.py --dtype f32 --mode parallel --ns 8,12,16,20,24,28,32 --ab --iters 4 --out gpurun_out/t3_ab_f32_parallel.json > gpurun_out/t3_ab_f32_parallel.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/t3_ab_f32_parallel.json"))
print("f32 parallel", " ".join("%d:%.3f/%.3f" % (r["n"], r["ms"], r["ms_lsu_staging"]) for r in d["rows"]))
PY
for mode in none parallel; do
timeout 900 python scripts/sweep.py --dtype f64 --mode $mode --ns 8,9,13,18,20 --ab --iters 4 --out gpurun_out/t3_ab_f64_$mode.json > gpurun_out/t3_ab_f64_$mode.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/t3_ab_f64_$mode.json"))
print("f64 $mode", " ".join("%d:%.3f/%.3f" % (r["n"], r["ms"], r["ms_lsu_staging"]) for r in d["rows"]))
PY
done
