#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/o_pytest.log 2>&1
tail -3 gpurun_out/o_pytest.log
cd scripts/tune
timeout 600 python run.py --threads 256,384 --only bulk --iters 4 > ../../gpurun_out/o_tune_bulk.jsonl 2> ../../gpurun_out/o_tune_bulk.err
tail -3 ../../gpurun_out/o_tune_bulk.err
cd ../..
python - <<'PY'
import json
for l in open("gpurun_out/o_tune_bulk.jsonl"):
    d=json.loads(l)
    if d.get("ok"): print(d["variant"], d["threads"], d["ms"], "occ", d["occ_blocks"], d["piv_equal"], d["values_close"], d["matrices_differing_1e-6"])
PY
