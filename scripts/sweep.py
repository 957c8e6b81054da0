#!/usr/bin/env python3
"""Kernel-time sweep over N (warm, best of k) with roofline fractions and the cuBLAS baseline.

    python scripts/sweep.py --dtype f32 --mode parallel --batch 1000000 [--ns 2,4,8,...] [--cublas]
"""
import argparse, ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import matrixinversion_b200 as lub
from matrixinversion_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f32")
ap.add_argument("--mode", default="parallel")
ap.add_argument("--batch", type=int, default=1_000_000)
ap.add_argument("--ns", default=",".join(str(i) for i in range(1, 33)))
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--cublas", action="store_true")
ap.add_argument("--refgpu", action="store_true", help="also time the reference's own kernel rebuilt for sm_100 (oracle/_ref)")
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--out", default="")
ap.add_argument("--staging", type=int, default=0, help="LUB_OPT_STAGING: 0 = library choice, 1 = LSU staging (no TMA / bulk copies)")
ap.add_argument("--ab", action="store_true", help="time every size with both staging settings, side by side")
ap.add_argument("--lu", action="store_true", help="also time the factors-only entry point (lu_batched_factor_inplace) of the mode")
a = ap.parse_args()
tdt = torch.float32 if a.dtype == "f32" else torch.float64
es = 4 if a.dtype == "f32" else 8
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
if a.threads:
    lub.set_num_threads(a.threads)
if a.staging:
    lub.set_option("staging", a.staging)
rows = []
for n in [int(x) for x in a.ns.split(",")]:
    g = torch.Generator(device="cuda").manual_seed(n)
    A = torch.rand((a.batch, n, n), generator=g, device="cuda", dtype=tdt)
    if a.mode == "none":
        A += n * torch.eye(n, device="cuda", dtype=tdt)
    orig = A.clone()
    times = []
    for i in range(a.iters + 1):
        A.copy_(orig)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lub.lu_batched_inplace(A, None, a.mode); e1.record()
        torch.cuda.synchronize()
        if i: times.append(e0.elapsed_time(e1))
    ms = min(times)
    ms_lsu = None
    if a.ab:
        lub.set_option("staging", 1)
        t2 = []
        for i in range(a.iters + 1):
            A.copy_(orig)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lub.lu_batched_inplace(A, None, a.mode); e1.record()
            torch.cuda.synchronize()
            if i: t2.append(e0.elapsed_time(e1))
        lub.set_option("staging", 0)
        ms_lsu = min(t2)
    row = {"n": n, "ms": ms, "Gmat_s": a.batch / ms / 1e6, "GBps": 2 * n * n * es * a.batch / ms / 1e6}
    row["frac_measured_peak"] = row["GBps"] / peak
    if ms_lsu is not None:
        row["ms_lsu_staging"] = ms_lsu
    row["gflops_2n3"] = 2 * n ** 3 * a.batch / ms / 1e6
    if a.lu:
        t3 = []
        for i in range(a.iters + 1):
            A.copy_(orig)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lub.lu_batched_factor_inplace(A, None, a.mode); e1.record()
            torch.cuda.synchronize()
            if i: t3.append(e0.elapsed_time(e1))
        row["lu_only_ms"] = min(t3)
    if a.cublas:
        C = _lib.cublas_lib()
        dst = torch.empty_like(A)
        t1, t2 = ctypes.c_float(), ctypes.c_float()
        best = 1e30
        for i in range(3):
            A.copy_(orig)
            rc = C.lu_batched_cublas_baseline(A.data_ptr(), dst.data_ptr(), n, a.batch, 0 if a.dtype == "f32" else 1,
                                              0 if a.mode == "none" else 1, ctypes.byref(t1), ctypes.byref(t2))
            assert rc == 0
            if t1.value + t2.value < best:
                best = t1.value + t2.value
                row["cublas_getrf_ms"] = t1.value
        row["cublas_ms"] = best
        row["speedup_vs_cublas"] = best / ms
        del dst
    if a.refgpu:
        from oracle import oracle as O  # checker side: dev script, not the product
        mode_code = {"none": 0, "serial": 1, "parallel": 2}[a.mode]
        A.copy_(orig); torch.cuda.synchronize()
        h = O.sweep_compare_hook(A.data_ptr(), n, a.batch, mode_code, "float32" if a.dtype == "f32" else "float64")
        row.update(h)
        if "reference_gpu_ms_warm" in h:
            row["speedup_vs_reference_gpu"] = h["reference_gpu_ms_warm"] * (a.batch / max(h["reference_gpu_matrices"], 1)) / ms
    rows.append(row)
    print(json.dumps(row), flush=True)
    del A, orig
if a.out:
    json.dump({"dtype": a.dtype, "mode": a.mode, "batch": a.batch, "peak_gbps": peak, "rows": rows}, open(a.out, "w"), indent=1)
