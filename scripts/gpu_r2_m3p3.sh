#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lub_tma -s 1 -c 1 -f -o gpurun_out/lu_n32_f32_parallel python scripts/prof_lu_one.py > gpurun_out/lu_ncu.log 2>&1
tail -2 gpurun_out/lu_ncu.log; ls -la gpurun_out/*.ncu-rep
