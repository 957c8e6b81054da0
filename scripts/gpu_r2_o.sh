#!/bin/bash
mkdir -p gpurun_out
python scripts/tune/run.py --threads 256 --iters 3 > gpurun_out/o_tune_dmma.jsonl 2> gpurun_out/o_tune_dmma.err
cat gpurun_out/o_tune_dmma.jsonl
ncu --set full --clock-control none --import-source on -k regex:lub_dmma -s 1 -c 1 -f -o gpurun_out/o_dmma_mode2 python scripts/tune/run.py --threads 256 --iters 1 --only mode2 > gpurun_out/o_ncu.log 2>&1
tail -2 gpurun_out/o_ncu.log
