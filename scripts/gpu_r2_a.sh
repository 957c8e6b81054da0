#!/bin/bash
# round 2, GPU call A: instruction-cost micro-benchmarks, step floor, headline variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt
./scripts/micro/issue_costs > gpurun_out/a_issue_costs.jsonl 2>&1
./scripts/micro/step_floor > gpurun_out/a_step_floor.jsonl 2>&1
python scripts/tune/run.py --threads 192,256,384,416 --iters 5 > gpurun_out/a_tune.jsonl 2> gpurun_out/a_tune.err
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
tail -3 gpurun_out/a_pytest.log
cat gpurun_out/a_step_floor.jsonl
cat gpurun_out/a_tune.jsonl | cut -c1-200
