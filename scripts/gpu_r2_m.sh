#!/bin/bash
# round 2, bulk-copy staged kernel: first run of the tuning variants
mkdir -p gpurun_out
cd scripts/tune
timeout 600 python run.py --threads 256,384 --only bulk --iters 4 > ../../gpurun_out/m_tune_bulk.jsonl 2> ../../gpurun_out/m_tune_bulk.err
tail -5 ../../gpurun_out/m_tune_bulk.err
cat ../../gpurun_out/m_tune_bulk.jsonl | cut -c1-220
