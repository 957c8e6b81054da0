#!/bin/bash
# full GPU suite + headline bench after the perm hardening and the new bulk tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t6_pytest.log 2>&1
tail -15 gpurun_out/t6_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/t6_bench.json 2> gpurun_out/t6_bench.err
cat gpurun_out/t6_bench.json | cut -c1-1500
