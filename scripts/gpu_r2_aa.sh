#!/bin/bash
# after the two-phase pivot_mode 3 kernels: the whole GPU suite, smoke, bench
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/aa_pytest.log 2>&1; tail -8 gpurun_out/aa_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/aa_smoke.log 2>&1; tail -3 gpurun_out/aa_smoke.log
timeout 600 python bench.py --steps 20 > gpurun_out/aa_bench.json 2> gpurun_out/aa_bench.err; cut -c1-300 gpurun_out/aa_bench.json
