#!/bin/bash
# packed-key position-aware search (one reduction per step) against the two-reduction search, bulk kernel, fp32 parallel
mkdir -p gpurun_out
python scripts/tune/run.py --threads 256 --only maxt256 --iters 6 > gpurun_out/z_tune_packed.jsonl 2> gpurun_out/z_tune.err
python scripts/tune/run.py --threads 384 --only maxt384 --iters 6 >> gpurun_out/z_tune_packed.jsonl 2>> gpurun_out/z_tune.err
cut -c1-230 gpurun_out/z_tune_packed.jsonl; tail -3 gpurun_out/z_tune.err
