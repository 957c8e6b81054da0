#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/p_pytest.log 2>&1
tail -3 gpurun_out/p_pytest.log
cd scripts/tune
timeout 600 python run.py --threads 256,384 --only bulk --iters 4 > ../../gpurun_out/p_tune_bulk.jsonl 2> ../../gpurun_out/p_tune_bulk.err
tail -3 ../../gpurun_out/p_tune_bulk.err
cd ../..
python - <<'PY'
import json
for l in open("gpurun_out/p_tune_bulk.jsonl"):
    d=json.loads(l)
    if d.get("ok"): print(d["variant"], d["threads"], d["ms"], "occ", d["occ_blocks"], d["piv_equal"], d["values_close"], d["matrices_differing_1e-6"])
PY
NS=5,6,7,9,10,11,13,14,15,17,18,19,21,22,23,25,26,27,29,30,31
for mode in none serial parallel; do
  timeout 600 python scripts/sweep.py --dtype f32 --mode $mode --ns $NS --ab --iters 4 --out gpurun_out/p_ab_f32_$mode.json > gpurun_out/p_ab_f32_$mode.log 2>&1
  python - <<PY
import json
d=json.load(open("gpurun_out/p_ab_f32_$mode.json"))
print("$mode", " ".join("%d:%.3f/%.3f" % (r["n"], r["ms"], r["ms_lsu_staging"]) for r in d["rows"]))
PY
done
