#!/bin/bash
# pivot_mode 3 (true partial pivoting) against cuBLAS getrfBatched + getriBatched, N = 2..32, both dtypes, inverse and factors
mkdir -p gpurun_out
NS=2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32
python scripts/sweep.py --dtype f32 --mode lapack --batch 1000000 --iters 3 --cublas --lu --ns $NS --out gpurun_out/sweep_f32_lapack.json 2>&1 | cut -c1-220 | tail -8
python scripts/sweep.py --dtype f64 --mode lapack --batch 1000000 --iters 3 --cublas --lu --ns $NS --out gpurun_out/sweep_f64_lapack.json 2>&1 | cut -c1-220 | tail -8
