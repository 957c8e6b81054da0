#!/bin/bash
mkdir -p gpurun_out
cd scripts/micro
for v in 1 2 3; do
  ncu --set full --clock-control none --import-source on -k regex:elim_only -s 1 -c 1 -f -o ../../gpurun_out/b_step_v$v ./step_floor $v > ../../gpurun_out/b_ncu_v$v.log 2>&1
done
./step_floor > ../../gpurun_out/b_step_floor.jsonl 2>&1
cd ../..
ls -la gpurun_out/*.ncu-rep
