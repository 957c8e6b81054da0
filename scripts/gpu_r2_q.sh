#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/q_pytest.log 2>&1
tail -12 gpurun_out/q_pytest.log
python bench.py --steps 10 --warmup 3 --n 32 --dtype f64 --no-e2e > gpurun_out/q_bench_cfg5.json 2> gpurun_out/q_bench_cfg5.err
python -c "
import json; d=json.loads(open('gpurun_out/q_bench_cfg5.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d.get('reference_gpu'), d.get('cublas'), d['details'].get('precheck_vs_oracle'))"
