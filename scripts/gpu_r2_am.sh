#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_lapack_layout.py tests/test_gpu_parity.py -m gpu -x -q -k "lapack or interleaved or tie_heavy or lu_only" > gpurun_out/am_pytest.log 2>&1; tail -4 gpurun_out/am_pytest.log
python scripts/tune/run.py --threads 256,384 --iters 5 2>/dev/null | grep -v '"ok": false' | cut -c1-120
LU_NS=4,5,6,7,8 bash scripts/gpu_r2_lu.sh | cut -c1-330
bash scripts/gpu_r2_il.sh > /dev/null 2>&1
