#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/final_pytest.log 2>&1; tail -4 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
LU_NS=6,7,8 bash scripts/gpu_r2_lu.sh | cut -c1-330
bash scripts/gpu_r2_il.sh > /dev/null 2>&1
