#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t18_smoke.log 2>&1; tail -3 gpurun_out/t18_smoke.log
