#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/x_pytest.log 2>&1; tail -3 gpurun_out/x_pytest.log
for mode in none serial parallel; do
  python scripts/sweep.py --dtype f64 --mode $mode --batch 1000000 --iters 4 --cublas --refgpu --ns 32 --out gpurun_out/x_f64_n32_${mode}.json 2>&1 | cut -c1-200
done
python scripts/dmma_accuracy.py > gpurun_out/x_dmma_accuracy.json 2>&1; head -5 gpurun_out/x_dmma_accuracy.json
