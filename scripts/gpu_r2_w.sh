#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/w_pytest.log 2>&1; tail -3 gpurun_out/w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for mode in none serial parallel; do
  python scripts/sweep.py --dtype f64 --mode $mode --batch 1000000 --iters 4 --cublas --refgpu --out gpurun_out/sweep_r02c_f64_${mode}.json > gpurun_out/sweep_r02c_f64_${mode}.log 2>&1
  tail -1 gpurun_out/sweep_r02c_f64_${mode}.log | cut -c1-160
done
for mode in none serial parallel; do
  python scripts/sweep.py --dtype f32 --mode $mode --batch 1000000 --iters 4 --cublas --refgpu --out gpurun_out/sweep_r02c_f32_${mode}.json > gpurun_out/sweep_r02c_f32_${mode}.log 2>&1
  tail -1 gpurun_out/sweep_r02c_f32_${mode}.log | cut -c1-160
done
