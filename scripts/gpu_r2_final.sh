#!/bin/bash
# final state of round 2: GPU suite, smoke, racecheck / memcheck over the late kernels, the six sweeps with both comparators, bench
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/final_pytest.log 2>&1; tail -4 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_new_paths.py > gpurun_out/san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error" gpurun_out/san_$tool.log | head -4
done
bash scripts/gpu_r2_sweeps.sh r02z
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-250 gpurun_out/final_bench.json
