#!/bin/bash
# records: full sweeps with both comparators (ours / cuBLAS / reference kernels rebuilt for sm_100), bench line
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/t13_bench.json 2> gpurun_out/t13_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/t13_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])"

