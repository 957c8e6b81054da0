#!/usr/bin/env python3
"""Launch the factors-only entry point a few times for one configuration (for ncu)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import matrixinversion_b200 as lub
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--mode", default="parallel")
ap.add_argument("--batch", type=int, default=1_000_000)
a = ap.parse_args()
tdt = torch.float32 if a.dtype == "f32" else torch.float64
g = torch.Generator(device="cuda").manual_seed(a.n)
A0 = torch.rand((a.batch, a.n, a.n), generator=g, device="cuda", dtype=tdt)
if a.mode == "none":
    A0 += a.n * torch.eye(a.n, device="cuda", dtype=tdt)
for i in range(3):
    A = A0.clone()
    lub.lu_batched_factor_inplace(A, None, a.mode)
torch.cuda.synchronize()
