#!/bin/bash
# ncu captures: two-phase pivot_mode 3 on the TMA image (N = 32 fp32, inverse), and the factors-only kernel (N = 32 fp32 parallel)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lub_tma -s 1 -c 1 -f -o gpurun_out/m3_n32_f32_tma python scripts/prof_one.py --n 32 --mode lapack --iters 2 > gpurun_out/m3p2_ncu.log 2>&1
tail -2 gpurun_out/m3p2_ncu.log
cat > /tmp/lu_one.py <<'PY'
import torch, matrixinversion_b200 as lub
g = torch.Generator(device="cuda").manual_seed(32)
A0 = torch.rand((1_000_000, 32, 32), generator=g, device="cuda")
for i in range(3):
    A = A0.clone(); lub.lu_batched_factor_inplace(A, None, "parallel")
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:lub_tma -s 1 -c 1 -f -o gpurun_out/lu_n32_f32_parallel python /tmp/lu_one.py > gpurun_out/lu_ncu.log 2>&1
tail -2 gpurun_out/lu_ncu.log; ls -la gpurun_out/*.ncu-rep
