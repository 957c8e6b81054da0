#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
python scripts/tune/run.py --threads 256,384 --iters 6 2>/dev/null | grep -v '"ok": false' | sed 's/^/zero  /' | cut -c1-110
TUNE_LIB=libtune_nozero.so python scripts/tune/run.py --threads 256,384 --iters 6 2>/dev/null | grep -v '"ok": false' | sed 's/^/nozero/' | cut -c1-110
done
