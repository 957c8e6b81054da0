#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/san2_pytest.log 2>&1; tail -3 gpurun_out/san2_pytest.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_new_paths.py > gpurun_out/san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error" gpurun_out/san_$tool.log | head -5
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_lapack_layout.py -m gpu -q -k "two_phase or info or factors" > gpurun_out/san_memcheck_tests.log 2>&1
echo "== memcheck over the pivot_mode 3 tests"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck_tests.log | head
