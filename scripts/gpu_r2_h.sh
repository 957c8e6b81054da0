#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lub_tma -s 2 -c 1 -f -o gpurun_out/h_headline_bs48 python scripts/tune/run.py --threads 384 --iters 2 > gpurun_out/h_ncu.log 2>&1
tail -3 gpurun_out/h_ncu.log
