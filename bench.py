#!/usr/bin/env python3
"""bench.py -- headline benchmark of the luBatchedInplace hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--n 32] [--batch 1000000] [--dtype f32|f64] [--mode parallel|serial|none]

Workload (BASELINE.json `metric`): matrices inverted per second, N=32, batch 1,000,000
per GPU, fp32, parallel pivoting; synthetic distinct uniform(0,1) matrices.  A "step" is
one in-place inversion of the whole batch (one kernel launch).  Every timed step sees the
stated input distribution: before each step the batch is restored from a pristine copy and
L2 is flushed (both outside the timed interval, which is the CUDA-event time of the launch
alone); the 4 GB working set is far larger than L2 anyway.

One JSON line is printed by rank 0 (see the keys below).  `--impl reference` times the
reference algorithm's CPU restatement (oracle/, OpenMP over all host cores) on a bounded
sample of the same workload -- the reference has no CPU implementation of its own for
LU/inverse (verify.hpp only checks), and its GPU kernels are not a CPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "matrices_inverted_per_sec"
UNIT = "matrices/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--batch", type=int, default=1_000_000, help="matrices per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--mode", default="parallel", choices=["none", "serial", "parallel"])
    ap.add_argument("--threads", type=int, default=0, help="NUMTHREADS knob (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline sample time")
    ap.add_argument("--no-cublas", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-refgpu", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return "N=%d batch=%d/GPU %s pivot=%s in-place inverse (BASELINE metric config)" % (a.n, a.batch, a.dtype, a.mode)


def config_of(a, world):
    """The workload identity, the same dict in both arms (the driver compares them)."""
    return {"workload": workload_name(a), "n": a.n, "batch_per_gpu": a.batch, "pivot_mode": a.mode,
            "parallelism": "batch sharded by contiguous slices, %d rank(s), no collective" % world,
            "l2": "GPU arm: working set per GPU >> 126 MB L2, and the input is restored from a pristine copy and L2 "
                  "flushed (256 MB write) before every timed step, outside the event-timed interval"}


def algorithmic_bytes(n, batch, esize, with_piv=False):
    """one read + one in-place write of every matrix (+ 4N if piv stored)."""
    return batch * (2 * n * n * esize + (4 * n if with_piv else 0))


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------

def cpu_baseline(a, target_s):
    from oracle import oracle as O  # cpu_baseline leg: the one place bench.py may run oracle/

    dt = np.float32 if a.dtype == "f32" else np.float64
    mode = {"none": 0, "serial": 1, "parallel": 2}[a.mode]
    # every host thread we may use -- explicitly, because torch.distributed.run exports
    # OMP_NUM_THREADS=1 to its workers
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    rng = np.random.default_rng(32)
    probe = rng.uniform(0, 1, size=(2000, a.n, a.n)).astype(dt)
    t0 = time.perf_counter()
    O.lu_batched_inplace_timed(probe, mode, cores)
    rate = 2000 / max(time.perf_counter() - t0, 1e-6)
    sample = int(min(a.batch, max(2000, rate * target_s)))
    X = rng.uniform(0, 1, size=(sample, a.n, a.n)).astype(dt)
    t0 = time.perf_counter()
    used = O.lu_batched_inplace_timed(X, mode, cores)
    dt_s = time.perf_counter() - t0
    return {"value": sample / dt_s, "unit": UNIT, "cores": used, "kind": "port",
            "sample": "%d distinct uniform(0,1) %dx%d %s matrices, pivot=%s, oracle/lu_oracle.c (C restatement of the "
                      "reference algorithm, OpenMP static over the batch), %.1f s" % (sample, a.n, a.n, a.dtype, a.mode, dt_s),
            "seconds": dt_s, "host_cores_total": os.cpu_count(), "sample_matrices": sample}


def precheck_vs_oracle(a, A_host, X_gpu, piv_gpu):
    """The pre-timing slice (first 4096 matrices of the timed buffer) against the CPU oracle: permutation vectors
    bit for bit, and the verifyInv verdict counts of both results (templated/verify.hpp:50-103)."""
    from oracle import oracle as O  # checker leg
    mode = {"none": 0, "serial": 1, "parallel": 2}[a.mode]
    with np.errstate(all="ignore"):
        Xo, po = O.lu_batched(A_host, mode)
    thr = 1e-3 if a.dtype == "f32" else 1e-8
    _, bad_o, _ = O.verify_inv(A_host, Xo, thr)
    _, bad_g, _ = O.verify_inv(A_host, X_gpu, thr)
    return {"matrices": int(A_host.shape[0]), "pivots_bit_exact": bool(np.array_equal(piv_gpu, po)),
            "verifyInv_incorrect_ours": int(bad_g), "verifyInv_incorrect_oracle": int(bad_o)}


def reference_gpu(a, A, our_ms):
    """The reference's own kernels (templated / serial_pivot / parallel_pivot luBatchedInplace.cuh), rebuilt for
    sm_100 from /root/reference by oracle/build_ref.sh, timed on THIS box on the same device buffer with the
    reference's convention (CUDA events around one launch, templated/luBatchedInplace.cu:71-82): cold = first
    launch of the process, warm = best of the following ones.  Outside the product path, like the cuBLAS leg."""
    try:
        from oracle import oracle as O  # checker leg
        mode = {"none": 0, "serial": 1, "parallel": 2}[a.mode]
        dt = np.float32 if a.dtype == "f32" else np.float64
        cold, warm, done = O.ref_gpu_time(A.data_ptr(), a.n, a.batch, mode, dt, reps=4)
        return {"ms_cold": cold, "ms_warm": warm, "matrices": int(done), "matrices_per_s": done / (warm * 1e-3),
                "speedup_ours_over_reference_gpu": warm / our_ms * (a.batch / max(done, 1)),
                "what": "reference kernel batched_lu_subwarp<%s> rebuilt for sm_100 (-O3 --use_fast_math, NUMTHREADS per "
                        "templated/run.py:201-223), same box, same device buffer" % a.mode}
    except Exception as e:  # the comparison is optional, the product is not
        return {"error": str(e)}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(1, a.steps), max(0, a.warmup)
    # bounded: whole run (warm-up + steps) within a few minutes
    per_step = max(1.0, min(a.cpu_seconds, 150.0 / (steps + warm)))
    vals = []
    cb = None
    for i in range(warm + steps):
        cb = cpu_baseline(a, per_step)
        if i >= warm:
            vals.append((cb["sample_matrices"], cb["seconds"]))
    tot_m = sum(v[0] for v in vals)
    tot_s = sum(v[1] for v in vals)
    value = tot_m / tot_s
    cb["value"] = value
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * tot_s / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic", "impl": "reference",
            "config": config_of(a, max(1, a.gpus)),
            "details": {"note": "CPU arm: each step is a bounded sample of the workload"},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------

class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(rows)}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------

def main():
    a = parse()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist

    import matrixinversion_b200 as lub

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    esize = 4 if a.dtype == "f32" else 8
    n, batch = a.n, a.batch
    if a.threads:
        lub.set_num_threads(a.threads)

    g = torch.Generator(device=dev).manual_seed(1000 * n + rank)
    A = torch.rand((batch, n, n), generator=g, device=dev, dtype=tdt)  # shard of rank: its own batch (weak scaling)
    pristine_head = A[:4096].clone()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness guard on this very buffer before timing (device-side verifyInv on a slice); the reference's
    # pivot rule fails its own 1e-3 predicate on 5-8 % of uniform(0,1) 32 x 32 matrices, so the
    # guard is a ceiling here and an exact comparison with the oracle in the CPU leg below (rank 0, N = 1)
    chk = pristine_head.clone()
    chk_piv = torch.empty((chk.shape[0], n), dtype=torch.int32, device=dev)
    lub.lu_batched_inplace(chk, chk_piv, a.mode)
    ok, bad, _ = lub.verify_inv(pristine_head, chk, 1e-3 if a.dtype == "f32" else 1e-8)
    if bad > 0.25 * (ok + bad):
        raise SystemExit("bench.py: precheck failed, %d of %d inverses miss the verifyInv predicate" % (bad, ok + bad))

    pristine = A.clone()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def restore():
        A.copy_(pristine)
        flush.zero_()

    for _ in range(max(a.warmup, 0)):
        restore()
        lub.lu_batched_inplace(A, None, a.mode)
    barrier()
    vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
    clocks = Clocks(vis[local] if local < len(vis) else local)
    clocks.start()
    time.sleep(0.3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(a.steps):
        restore()                                  # untimed: pristine input + L2 flush
        evs[i][0].record()
        lub.lu_batched_inplace(A, None, a.mode)   # launched on torch's current stream
        evs[i][1].record()
    barrier()
    t_wall1 = time.perf_counter()
    ck = clocks.stop(t_wall0, t_wall1)
    per = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = float(sum(per))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * batch * a.steps / (total_ms_max * 1e-3)

    # ---- e2e: public API with HOST (pinned) buffers, copies inside the timed region ------
    # The rank first binds itself next to its GPU, so that the pinned buffer it allocates (first touch) lives on
    # that GPU's NUMA node: with 8 ranks the node's memory controllers and socket links, not the library, set the
    # ceiling otherwise (round 1: 1.4x at 8 GPUs).  The ceiling itself -- the same buffer copied H2D and D2H at once
    # on two streams, nothing else -- is measured next to the number.
    e2e = None
    if not a.no_e2e:
        bound = lub.bind_thread_near_device(local)
        nbytes = batch * n * n * esize
        hbuf = torch.empty((batch, n, n), dtype=tdt, pin_memory=True)
        hbuf.copy_(pristine)
        H = hbuf.numpy()
        lub.lu_batched_inplace(H, None, a.mode)  # warm-up (allocates the pipeline's device chunks)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            lub.lu_batched_inplace(H, None, a.mode)  # H2D chunks -> kernels -> D2H chunks, synchronous
        barrier()
        dt_e = time.perf_counter() - t0
        te = torch.tensor([dt_e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # copy ceiling: H2D of the buffer on one stream while the previous contents go D2H on another
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dst = torch.empty_like(pristine)
        hout = torch.empty((batch, n, n), dtype=tdt, pin_memory=True)
        barrier()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            with torch.cuda.stream(s_in):
                dst.copy_(hbuf, non_blocking=True)
            with torch.cuda.stream(s_out):
                hout.copy_(pristine, non_blocking=True)
        barrier()
        dt_c = time.perf_counter() - t0
        tc = torch.tensor([dt_c], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        ceil_gbps = reps * nbytes / float(tc.item()) / 1e9            # per rank, each direction
        ceil_mats = world * reps * batch / float(tc.item())
        e2e_val = world * batch * a.e2e_steps / float(te.item())
        e2e = {"value": e2e_val, "unit": UNIT,
               "h2d_bytes_per_step": world * nbytes, "d2h_bytes_per_step": world * nbytes,
               "bytes_per_step_per_rank": nbytes, "steps": a.e2e_steps,
               "api": "matrixinversion_b200.lu_batched_inplace(numpy pinned) -> lu_batched_inplace_host",
               "host_buffer": "pinned, allocated after lu_batched_bind_thread_near_device(%d) -> %s" % (local, "bound" if bound else "topology not exposed, unbound"),
               "copy_ceiling": {"what": "the same pinned buffers copied H2D and D2H concurrently on two streams, max over ranks",
                                "GBps_each_way_per_rank": ceil_gbps, "matrices_per_s_all_ranks": ceil_mats},
               "frac_of_copy_ceiling": e2e_val / ceil_mats}
        del hbuf, H, dst, hout

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline for the one kernel of the step ---------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    avg_ms = float(np.mean(per))
    abytes = algorithmic_bytes(n, batch, esize)
    achieved = abytes / (avg_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("n%d_%s_%s" % (n, a.dtype, a.mode), {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "%s<%s,N=%d,mode=%s>" % (lub.kernel_name(n, a.mode, np.float32 if a.dtype == "f32" else np.float64), a.dtype, n, a.mode),
                "algorithmic_bytes_per_launch": abytes, "avg_launch_ms": avg_ms,
                "frac_of_8TBps_nominal": achieved / 8000.0,
                "gflops_2n3": 2.0 * n ** 3 * batch / (avg_ms * 1e-3) / 1e9}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic", "impl": "ours",
            "config": config_of(a, world),
            "details": {"working_set_gb_per_gpu": batch * n * n * esize / 1e9,
                       "timing": "sum of per-step CUDA-event intervals around the launch, max over ranks",
                       "geometry": vars(lub.geometry(n, batch, a.mode, np.float32 if a.dtype == "f32" else np.float64)),
                       "precheck_verifyInv_first4096": {"correct": ok, "incorrect": bad}},
            "clocks": ck, "e2e": e2e, "gpu_launches": a.steps, "roofline": roofline}

    if not a.no_cublas:
        try:
            import ctypes
            from matrixinversion_b200 import _lib
            C = _lib.cublas_lib()
            src = torch.rand((batch, n, n), generator=g, device=dev, dtype=tdt)
            dst = torch.empty_like(src)
            t1, t2 = ctypes.c_float(), ctypes.c_float()
            best = None
            for _ in range(3):
                src.copy_(torch.rand((batch, n, n), generator=g, device=dev, dtype=tdt))  # getrf overwrites its input
                rc = C.lu_batched_cublas_baseline(src.data_ptr(), dst.data_ptr(), n, batch, 0 if a.dtype == "f32" else 1, 1,
                                                  ctypes.byref(t1), ctypes.byref(t2))
                if rc != 0:
                    raise RuntimeError("cublas rc=%d" % rc)
                tot = t1.value + t2.value
                if best is None or tot < best[0]:
                    best = (tot, t1.value, t2.value)
            line["cublas"] = {"value": batch / (best[0] * 1e-3), "unit": UNIT, "ms_getrf": best[1], "ms_getri": best[2],
                              "what": "cublas%sgetrfBatched(pivoting)+getriBatched, same box, best of 3" % ("S" if a.dtype == "f32" else "D"),
                              "speedup_ours_over_cublas": (batch / (avg_ms * 1e-3)) / (batch / (best[0] * 1e-3))}
            del src, dst
        except Exception as e:  # the comparison baseline is optional, the product is not
            line["cublas"] = {"error": str(e)}

    if world == 1 and not a.no_cpu:
        # checker legs (the only places bench.py runs oracle/): the precheck slice against the CPU oracle, the
        # reference's own kernels rebuilt for sm_100 timed on this box, then the CPU baseline
        line["details"]["precheck_vs_oracle"] = precheck_vs_oracle(a, pristine_head.cpu().numpy(), chk.cpu().numpy(), chk_piv.cpu().numpy())
        if not line["details"]["precheck_vs_oracle"]["pivots_bit_exact"]:
            raise SystemExit("bench.py: precheck failed, permutation vectors differ from the oracle's")
        if not a.no_refgpu:
            restore()
            line["reference_gpu"] = reference_gpu(a, A, avg_ms)
        line["cpu_baseline"] = cpu_baseline(a, a.cpu_seconds)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
