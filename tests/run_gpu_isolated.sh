#!/usr/bin/env bash
# Runs every GPU test in its own process (a CUDA fault is sticky for the whole process, so
# one bad kernel would otherwise fail every later test) and logs to gpurun_out/tests_isolated.log
mkdir -p gpurun_out
LOG=gpurun_out/tests_isolated.log
: > "$LOG"
for t in $(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep '::'); do
  echo "=== $t" >> "$LOG"
  timeout 400 python -m pytest "$t" -x -q 2>&1 | tail -${TAIL:-30} >> "$LOG"
done
grep -E "^(=== |FAILED|[0-9]+ (passed|failed))" "$LOG"
