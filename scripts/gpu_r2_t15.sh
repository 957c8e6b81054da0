#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lub_lapack2 -s 1 -c 1 -f -o gpurun_out/t15_lapack2_n32_f32 python scripts/prof_one.py --n 32 --dtype f32 --mode lapack --iters 2 > gpurun_out/t15_ncu.log 2>&1
tail -2 gpurun_out/t15_ncu.log
python scripts/ncu_summary.py gpurun_out/t15_lapack2_n32_f32.ncu-rep > gpurun_out/t15_sum.txt 2>&1
cat gpurun_out/t15_sum.txt
