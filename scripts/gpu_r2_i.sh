#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1
tail -15 gpurun_out/i_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
cut -c1-600 gpurun_out/i_bench.json
