#!/bin/bash
# ncu full capture of the bulk-copy staged kernel: N = 29 and N = 18 fp32 parallel
mkdir -p gpurun_out
for n in 29 18; do
ncu --set full --clock-control none --import-source on -k regex:lub_bulk -s 2 -c 1 -f -o gpurun_out/t4_bulk_n${n}_f32_parallel python scripts/prof_one.py --n $n --dtype f32 --mode parallel --iters 3 > gpurun_out/t4_ncu_$n.log 2>&1
tail -2 gpurun_out/t4_ncu_$n.log
python scripts/ncu_summary.py gpurun_out/t4_bulk_n${n}_f32_parallel.ncu-rep > gpurun_out/t4_sum_$n.txt 2>&1
python scripts/ncu_phases.py gpurun_out/t4_bulk_n${n}_f32_parallel.ncu-rep >> gpurun_out/t4_sum_$n.txt 2>&1
cat gpurun_out/t4_sum_$n.txt
done
