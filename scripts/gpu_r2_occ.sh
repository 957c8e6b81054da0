#!/bin/bash
# factors-only kernels: one image per warp at 16 / 20 / 24 warps per SM against two images at 12 (bulk kernel, lane = row LU)
mkdir -p gpurun_out
python scripts/tune/run.py --threads 256,384,512,640,768 --iters 4 > gpurun_out/occ_tune.jsonl 2> gpurun_out/occ_tune.err
grep -v '"ok": false' gpurun_out/occ_tune.jsonl | cut -c1-140; tail -3 gpurun_out/occ_tune.err
