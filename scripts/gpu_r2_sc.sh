#!/bin/bash
mkdir -p gpurun_out
./scripts/micro/search_costs > gpurun_out/search_costs.jsonl 2>&1
cat gpurun_out/search_costs.jsonl
