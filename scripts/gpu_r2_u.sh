#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_smoke.log 2>&1; tail -2 gpurun_out/u_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/u_pytest.log 2>&1; tail -3 gpurun_out/u_pytest.log
