#!/bin/bash
# round-2 re-entry check: GPU tests, smoke, both bench arms on the tree as committed
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/y_pytest.log 2>&1; tail -4 gpurun_out/y_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1; tail -2 gpurun_out/y_smoke.log
timeout 600 python bench.py > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err; cut -c1-400 gpurun_out/y_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/y_bench_ref.json 2> gpurun_out/y_bench_ref.err; cut -c1-300 gpurun_out/y_bench_ref.json
