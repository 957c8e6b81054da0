#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_lapack_layout.py -m gpu -x -q > gpurun_out/m3small_pytest.log 2>&1; tail -6 gpurun_out/m3small_pytest.log
LU_NS=2,3,4,5,6,7,8 bash scripts/gpu_r2_lu.sh
