#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lu_only" > gpurun_out/lu2_pytest.log 2>&1; tail -12 gpurun_out/lu2_pytest.log
bash scripts/gpu_r2_lu.sh
