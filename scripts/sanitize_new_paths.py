#!/usr/bin/env python3
"""Exercise the round-2 late kernels (two-phase pivot_mode 3, factors-only on the staged image) on small ragged batches -- run under
compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import matrixinversion_b200 as lub
rng = np.random.default_rng(3)
for dt in (np.float32, np.float64):
    for n in (5, 6, 7, 8, 9, 16, 18, 21, 27, 31, 32):   # n <= 8: the one-lane-per-matrix paths
        for batch in (1, 37, 1301):
            A = rng.uniform(0, 1, size=(batch, n, n)).astype(dt)
            for mode in ("none", "serial", "parallel", "lapack"):
                src = A + (n * np.eye(n, dtype=dt) if mode == "none" else 0)
                dA = torch.from_numpy(src).cuda()
                piv = torch.zeros((batch, n), dtype=torch.int32, device="cuda")
                info = torch.zeros((batch,), dtype=torch.int32, device="cuda")
                lub.lu_batched_factor_inplace(dA, piv, mode, info=info) if mode == "lapack" else lub.lu_batched_factor_inplace(dA, piv, mode)
                if mode == "lapack" or n <= 8:
                    dB = torch.from_numpy(src).cuda()
                    lub.lu_batched_inplace(dB, piv, mode, info=info)
    torch.cuda.synchronize()
print("done")
