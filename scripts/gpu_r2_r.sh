#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, torch
sys.path.insert(0, "/root/repo")
import matrixinversion_b200 as lub
n, dt, mode = int(sys.argv[1]), sys.argv[2], sys.argv[3]
tdt = torch.float32 if dt == "f32" else torch.float64
g = torch.Generator(device="cuda").manual_seed(n)
A0 = torch.rand((1_000_000, n, n), generator=g, device="cuda", dtype=tdt)
A = A0.clone()
for _ in range(3):
    A.copy_(A0)
    lub.lu_batched_inplace(A, None, mode)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:lub_ -s 2 -c 1 -f -o gpurun_out/r_cfg5_dmma python /tmp/one.py 32 f64 parallel > gpurun_out/r_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lub_ -s 2 -c 1 -f -o gpurun_out/r_cfg3 python /tmp/one.py 18 f32 parallel > gpurun_out/r_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lub_ -s 2 -c 1 -f -o gpurun_out/r_n31_par python /tmp/one.py 31 f32 parallel > gpurun_out/r_ncu3.log 2>&1
tail -1 gpurun_out/r_ncu1.log gpurun_out/r_ncu2.log gpurun_out/r_ncu3.log
