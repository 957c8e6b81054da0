#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
tail -3 gpurun_out/j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/j_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('reference_gpu'), d['details'].get('precheck_vs_oracle'), d['cublas'])
PY
