#!/bin/bash
# round 2: systematic tuning run of the bulk-copy staged kernel (scripts/tune/gen_bulk_variants.py)
mkdir -p gpurun_out
cd scripts/tune
timeout 1200 python run.py --threads 384,416,448,480 --only bulk --iters 4 > ../../gpurun_out/t12_tune_maxt.jsonl 2> ../../gpurun_out/t12_tune_maxt.err
tail -3 ../../gpurun_out/t12_tune_maxt.err
cd ../..
python - <<'PY'
import json
for l in open("gpurun_out/t12_tune_maxt.jsonl"):
    d=json.loads(l)
    if d.get("ok"): print(d["variant"], d["threads"], d["ms"], "occ", d["occ_blocks"], d["piv_equal"], d["values_close"], d["matrices_differing_1e-6"])
PY
