#!/bin/bash
# ncu capture of the two-phase pivot_mode 3 kernel (N = 32 fp32)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lub_bulk -s 1 -c 1 -f -o gpurun_out/m3_n32_f32 python scripts/prof_one.py --n 32 --mode lapack --iters 2 > gpurun_out/m3p_ncu.log 2>&1
tail -3 gpurun_out/m3p_ncu.log; ls -la gpurun_out/*.ncu-rep
