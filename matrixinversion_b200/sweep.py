#!/usr/bin/env python3
"""Sweep driver -- successor of the reference's per-variant `run.py`
(/root/reference/templated/run.py:169-261), re-targeted from "recompile ./custom per
configuration and parse its stdout" to calls into the C ABI via ctypes.

What is kept:
  * the loop: every matrix size in 1..32 x batch in {100000, 500000, 1000000} x 10 runs
    (run.py:171-176); `--sizes/--batches/--runs` narrow it;
  * one cold single launch per run, kernel time only from a CUDA event pair -- the
    reference's timing convention (templated/luBatchedInplace.cu:71-82); `--warm` adds the
    warm numbers next to it;
  * the result schema and file names (run.py:236-250): benchmark_results/runtime_results/<100k|500k|1M>/
    benchmark_results_<name>.json with key "m{N}_n{batch}_t{T}" ->
    {matrix_size, num_matrices, num_threads, runtimes[ms], runtime_avg, variance, std_dev,
    incorrect_inversions}, rewritten after every configuration, plus
    benchmark_results_incomplete.json on an exception (run.py:252-260) -- so plot.py-style
    consumers keep working;
  * the NUMTHREADS table (run.py:201-223) is recorded as `reference_num_threads`; the
    library picks its own launch geometry (`num_threads` is what it actually used).
What is fixed: the reference's parser `break`s on the "Kernel execution time" line, which is
printed before "Incorrect inversions", so its sweep never records correctness.
Here `incorrect_inversions` is the verifyInv-compatible count of every run.
  * the per-configuration Nsight Compute capture (run.py:90-119) with `--ncu`: the same sections, one report per
    configuration under benchmark_results/ncu_profiles/<100k|500k|1M>/profile_m{N}_n{batch}_t{T}.ncu-rep; the profiled
    process is this module in `--one-launch` mode (one cold launch, what ./custom did).  Clocks are NOT locked
    (`--clock-control none`): the reference's `base` lock changes what is measured and is not allowed on shared boxes.
Extra keys (additive): gbps, matrices_per_s, gflops_2n3, frac_hbm_roofline, cublas_ms (--cublas), and whatever a
`--compare module:function` hook returns (used by the test tree to put the reference's own kernels, rebuilt for
sm_100, next to ours: function(device_ptr, n, batch, mode, dtype_name) -> dict).
Several GPUs (BASELINE config 4 "at 1/2/4/8 B200"): launched under
`python -m torch.distributed.run --nproc-per-node G -m matrixinversion_b200.sweep ...` every rank
inverts its contiguous slice of each batch (sharding.shard_range: strong scaling, the batch sizes of
the reference's loop stay what they are), a run's time is the maximum of the ranks' kernel times, the
incorrect counts are summed (sharding.reduce_verdict: the only collective, two scalars), rank 0 writes
the files and records `n_gpus`.
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import shutil
import subprocess
import sys

import numpy as np

from . import _lib, api
from .sharding import reduce_verdict, shard_range

BATCH_NAMES = {100000: "100k", 500000: "500k", 1000000: "1M"}


MARKER = ".lubatched_sweep"


def create_output_directories(base="benchmark_results", force=False):
    """run.py:9-22, made safe for a user-chosen path: the reference wipes its fixed `benchmark_results`
    directory; here only the two sub-directories the sweep itself creates are ever removed, and only when
    the directory carries the marker a previous sweep left (or --force is given)."""
    ncu_dir, runtime_dir = os.path.join(base, "ncu_profiles"), os.path.join(base, "runtime_results")
    stale = [d for d in (ncu_dir, runtime_dir) if os.path.exists(d)]
    if stale and not (force or os.path.exists(os.path.join(base, MARKER))):
        raise SystemExit("sweep: %s holds results this sweep did not create (no %s marker); move them away or pass --force"
                         % (base, MARKER))
    for d in stale:
        shutil.rmtree(d)
    os.makedirs(ncu_dir, exist_ok=True)
    os.makedirs(runtime_dir, exist_ok=True)
    with open(os.path.join(base, MARKER), "w") as f:
        f.write("created by matrixinversion_b200.sweep\n")
    return base, ncu_dir, runtime_dir


def run_ncu_profile(n, batch, num_threads, ncu_dir, variant, dtype, input_path):
    """The reference's capture (templated/run.py:90-119): same sections and sampling options, one report per
    configuration; ./custom is replaced by this module's --one-launch mode."""
    profile_path = os.path.join(ncu_dir, "profile_m%d_n%d_t%d" % (n, batch, num_threads))
    cmd = ["ncu", "--set=full", "--import-source", "yes", "--target-processes", "all", "--replay-mode", "kernel",
           "--section", "InstructionStats", "--section", "LaunchStats", "--section", "MemoryWorkloadAnalysis",
           "--section", "SchedulerStats", "--section", "SourceCounters", "--section", "SpeedOfLight",
           "--sampling-interval", "auto", "--sampling-max-passes", "5", "--sampling-buffer-size", "33554432",
           "--clock-control", "none", "--kernel-name", "regex:lub_", "--launch-count", "1", "-f", "-o", profile_path,
           sys.executable, "-m", "matrixinversion_b200.sweep", "--one-launch", "--variant", variant, "--dtype", str(np.dtype(dtype)),
           "--sizes", str(n), "--batches", str(batch), "--input", input_path]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    print("NCU profile saved to %s" % profile_path)
    return profile_path + ".ncu-rep"


def result_key(matrix_size, num_matrices, num_threads):
    return "m%d_n%d_t%d" % (matrix_size, num_matrices, num_threads)


def make_entry(matrix_size, num_matrices, num_threads, runtimes, incorrect, extra=None):
    """The reference's per-configuration record (run.py:236-246)."""
    e = {
        "matrix_size": matrix_size,
        "num_matrices": num_matrices,
        "num_threads": num_threads,
        "runtimes": [float(x) for x in runtimes],
        "runtime_avg": float(sum(runtimes) / len(runtimes)),
        "variance": float(np.var(runtimes)),
        "std_dev": float(np.std(runtimes)),
        "incorrect_inversions": [int(x) for x in incorrect if x > 0],
    }
    if extra:
        e.update(extra)
    return e


def save_results(results, filepath):
    with open(filepath, "w") as f:
        json.dump(results, f, indent=4)


def _max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_config(n, batch, mode, dtype, template, runs, warm=False, cublas=False, peak_gbps=None, rank=0, world=1, compare=None):
    """`runs` cold single launches of one configuration; returns the JSON entry.  With world > 1 this
    rank works on its slice of the batch; times are maxima over ranks, incorrect counts sums."""
    import torch

    total_batch = batch
    lo, hi = shard_range(total_batch, rank, world)
    batch = hi - lo
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    dT = torch.from_numpy(np.ascontiguousarray(template)).cuda()
    orig = dT.unsqueeze(0).expand(batch, n, n).contiguous()   # main()'s replicate loop (this rank's slice)
    geo = api.geometry(n, max(batch, 1), mode, dtype)
    runtimes, incorrect, warm_ms = [], [], []
    A = torch.empty_like(orig)
    api.enable_timing(True)
    try:
        for r in range(runs):
            A.copy_(orig)
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            api.lu_batched_inplace(A, None, mode)
            runtimes.append(_max_over_ranks(api.last_kernel_ms() if batch else 0.0, world))
            _, bad, dev = api.verify_inv(orig, A) if batch else (0, 0, 0.0)
            incorrect.append(reduce_verdict(bad, 0.0 if dev != dev else dev)[0] if world > 1 else bad)
        if warm:
            for r in range(runs):
                A.copy_(orig)
                api.lu_batched_inplace(A, None, mode)
                warm_ms.append(_max_over_ranks(api.last_kernel_ms() if batch else 0.0, world))
    finally:
        api.enable_timing(False)
    es = np.dtype(dtype).itemsize
    best = min(runtimes)
    batch = total_batch
    extra = {"n_gpus": world, "reference_num_threads": api.default_num_threads(n), "threads_per_matrix": geo.threads_per_matrix,
             "matrices_per_block": geo.matrices_per_block, "num_blocks": geo.num_blocks,
             "matrices_per_s": batch / (best * 1e-3), "gbps": 2.0 * n * n * es * batch / (best * 1e-3) / 1e9,
             "gflops_2n3": 2.0 * n ** 3 * batch / (best * 1e-3) / 1e9}
    if peak_gbps:
        extra["frac_hbm_roofline"] = extra["gbps"] / peak_gbps
    if warm_ms:
        extra["warm_runtimes"] = warm_ms
    if cublas and world == 1:
        C = _lib.cublas_lib()
        dst = torch.empty_like(orig)
        t1, t2 = ctypes.c_float(), ctypes.c_float()
        A.copy_(orig)
        rc = C.lu_batched_cublas_baseline(A.data_ptr(), dst.data_ptr(), n, batch, 0 if tdt == torch.float32 else 1,
                                          0 if api._mode(mode) == 0 else 1, ctypes.byref(t1), ctypes.byref(t2))
        if rc == 0:
            extra["cublas_ms"] = t1.value + t2.value
            extra["cublas_getri_share"] = t2.value / (t1.value + t2.value)
            extra["speedup_vs_cublas"] = (t1.value + t2.value) / best
    if compare is not None and world == 1:
        A.copy_(orig)
        torch.cuda.synchronize()
        extra.update(compare(A.data_ptr(), n, batch, api._mode(mode), str(np.dtype(dtype))))
    return make_entry(n, batch, geo.threads_per_block, runtimes, incorrect, extra)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--variant", default="parallel_pivot", help="templated | serial_pivot | parallel_pivot (or none/serial/parallel)")
    ap.add_argument("--input", default="mtrand32_new1.txt", help="template matrix text file (first N*N tokens are used)")
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--sizes", default="1-32")
    ap.add_argument("--batches", default="100000,500000,1000000")
    ap.add_argument("--runs", type=int, default=10)
    ap.add_argument("--out", default="benchmark_results")
    ap.add_argument("--warm", action="store_true")
    ap.add_argument("--cublas", action="store_true")
    ap.add_argument("--peak-gbps", type=float, default=None)
    ap.add_argument("--ncu", action="store_true", help="one Nsight Compute report per configuration (templated/run.py:90-119)")
    ap.add_argument("--one-launch", action="store_true", help="one cold launch of the first size / batch and exit (what --ncu profiles)")
    ap.add_argument("--compare", default="", help="module:function timing hook, see the module docstring")
    ap.add_argument("--force", action="store_true", help="replace result directories this sweep did not create")
    a = ap.parse_args(argv)

    lo, _, hi = a.sizes.partition("-")
    sizes = range(int(lo), int(hi or lo) + 1)
    dtype = np.dtype(a.dtype)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if a.one_launch:
        import torch
        n, batch = sizes[0], int(a.batches.split(",")[0])
        T = api.read_template(a.input, n, dtype)
        A = torch.from_numpy(T).cuda().unsqueeze(0).expand(batch, n, n).contiguous()
        api.lu_batched_inplace(A, None, a.variant)
        torch.cuda.synchronize()
        return 0
    compare = None
    if a.compare:
        mod, _, fn = a.compare.partition(":")
        compare = getattr(importlib.import_module(mod), fn)
    if world > 1:  # one process per GPU (torch.distributed.run); NCCL only for the scalar reductions
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        base, ncu_dir, runtime_dir = create_output_directories(a.out, a.force)
    else:
        base, ncu_dir, runtime_dir = a.out, os.path.join(a.out, "ncu_profiles"), os.path.join(a.out, "runtime_results")
    results = {}
    try:
        for batch in [int(b) for b in a.batches.split(",")]:
            name = BATCH_NAMES.get(batch, str(batch))
            out_dir = os.path.join(runtime_dir, name)
            if rank == 0:
                os.makedirs(os.path.join(ncu_dir, name), exist_ok=True)
                os.makedirs(out_dir, exist_ok=True)
            results = {}
            for n in sizes:
                if rank == 0:
                    print("\nTesting configuration: Matrix Size=%d, Num Matrices=%d" % (n, batch))
                template = api.read_template(a.input, n, dtype)
                entry = run_config(n, batch, a.variant, dtype, template, a.runs, a.warm, a.cublas, a.peak_gbps, rank, world, compare)
                if rank == 0:
                    print("Number of threads: %d" % entry["num_threads"])
                    if a.ncu and world == 1:
                        entry["ncu_report"] = run_ncu_profile(n, batch, entry["num_threads"], os.path.join(ncu_dir, name), a.variant,
                                                              dtype, a.input)
                    results[result_key(n, batch, entry["num_threads"])] = entry
                    save_results(results, os.path.join(out_dir, "benchmark_results_%s.json" % name))
    except Exception as e:  # keep what we have, like the reference (run.py:252-260)
        print("Unexpected error: %s" % e)
        if rank == 0:
            save_results(results, os.path.join(runtime_dir, "benchmark_results_incomplete.json"))
        raise
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
