#!/bin/bash
# full N = 1..32 sweeps, batch 1M, ours / cuBLAS / reference kernels rebuilt for sm_100
mkdir -p gpurun_out
TAG=${1:-r02a}
for dt in f32 f64; do
  for mode in none serial parallel; do
    python scripts/sweep.py --dtype $dt --mode $mode --batch 1000000 --iters 4 --cublas --refgpu --out gpurun_out/sweep_${TAG}_${dt}_${mode}.json > gpurun_out/sweep_${TAG}_${dt}_${mode}.log 2>&1
    tail -1 gpurun_out/sweep_${TAG}_${dt}_${mode}.log | cut -c1-200
  done
done
