#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-200; wc -l gpurun_out/launches_r02.csv
