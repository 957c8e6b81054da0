#!/usr/bin/env python3
"""Pinned-memory copy ceiling of the box (dev / evidence script): every rank copies a 2 GiB pinned host buffer to its GPU
and another one back AT THE SAME TIME (two streams), with and without binding the rank next to its GPU before the
buffers are allocated.  Run under torch.distributed.run at 1 / 2 / 4 / 8 ranks; rank 0 prints one JSON line.
The end-to-end (host buffer) number of bench.py cannot exceed the bound line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import matrixinversion_b200 as lub

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
nbytes = 2 << 30
out = {"n_gpus": world, "bytes_each_way_per_rank": nbytes}
for label in ("unbound", "bound"):
    if label == "bound":
        out["bind_ok"] = bool(lub.bind_thread_near_device(local))
    hin = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); hin.fill_(1)
    hout = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); hout.fill_(2)
    din = torch.empty(nbytes, dtype=torch.uint8, device=dev); dout = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def sync():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        sync(); t0 = time.perf_counter()
        for _ in range(3):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
        sync(); dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode + "_GBps_per_rank_each_way"] = 3 * nbytes / float(t.item()) / 1e9
    out[label] = res
    del hin, hout, din, dout
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
