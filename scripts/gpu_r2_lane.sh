#!/bin/bash
mkdir -p gpurun_out
python scripts/tune/run.py --threads 256,384 --iters 5 > gpurun_out/lane_tune.jsonl 2> gpurun_out/lane_tune.err
grep -v '"ok": false' gpurun_out/lane_tune.jsonl | cut -c1-175; tail -3 gpurun_out/lane_tune.err
