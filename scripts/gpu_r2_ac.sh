#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/ac_pytest.log 2>&1; tail -8 gpurun_out/ac_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ac_smoke.log 2>&1; tail -3 gpurun_out/ac_smoke.log
LU_NS=7 bash scripts/gpu_r2_lu.sh
