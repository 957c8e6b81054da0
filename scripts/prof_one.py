#!/usr/bin/env python3
"""Launch the hot-path kernel a few times for one configuration (for ncu / quick timing).

    python scripts/prof_one.py --n 32 --dtype f32 --mode parallel --batch 1000000 --iters 5
"""
import argparse, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import matrixinversion_b200 as lub

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--mode", default="parallel")
ap.add_argument("--batch", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--piv", action="store_true")
a = ap.parse_args()
tdt = torch.float32 if a.dtype == "f32" else torch.float64
if a.threads:
    lub.set_num_threads(a.threads)
g = torch.Generator(device="cuda").manual_seed(a.n)
A = torch.rand((a.batch, a.n, a.n), generator=g, device="cuda", dtype=tdt)
if a.mode == "none":
    A += a.n * torch.eye(a.n, device="cuda", dtype=tdt)
piv = torch.empty((a.batch, a.n), dtype=torch.int32, device="cuda") if a.piv else None
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
lub.lu_batched_inplace(A, piv, a.mode)
torch.cuda.synchronize()
ev[0].record()
for i in range(a.iters):
    lub.lu_batched_inplace(A, piv, a.mode)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]
es = 4 if a.dtype == "f32" else 8
best = min(ms)
print(json.dumps({"n": a.n, "dtype": a.dtype, "mode": a.mode, "batch": a.batch, "ms_best": best, "ms_med": float(np.median(ms)),
                  "Mmat_s": a.batch / best / 1e3, "GBps": 2 * a.n * a.n * es * a.batch / best / 1e6,
                  "geometry": vars(lub.geometry(a.n, a.batch, a.mode, np.float32 if a.dtype == "f32" else np.float64))}))
