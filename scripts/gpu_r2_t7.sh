#!/bin/bash
# final-candidate: GPU suite, bench, full sweeps with both comparators
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t7_pytest.log 2>&1
tail -3 gpurun_out/t7_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/t7_bench.json 2> gpurun_out/t7_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/t7_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline'], d.get('reference_gpu'))"
bash scripts/gpu_r2_sweeps.sh r02d
