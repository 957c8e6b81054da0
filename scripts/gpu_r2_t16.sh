#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t16_pytest.log 2>&1
tail -5 gpurun_out/t16_pytest.log
