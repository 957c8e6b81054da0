"""CPU tests: the sweep/plot successors keep the reference's JSON schema, and the N > 1
path (contiguous batch shards + scalar verdict reduction, no data-path collective) works
under torch.distributed with the gloo backend, world_size 2."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from matrixinversion_b200 import plot, sweep


def test_sweep_entry_schema_matches_reference_run_py(tmp_path):
    e = sweep.make_entry(18, 1000000, 256, [1.5, 1.7, 1.6], [0, 0, 3], {"gbps": 1.0})
    # templated/run.py:236-246
    for k in ("matrix_size", "num_matrices", "num_threads", "runtimes", "runtime_avg", "variance", "std_dev", "incorrect_inversions"):
        assert k in e
    assert e["runtime_avg"] == pytest.approx(1.6) and e["variance"] == pytest.approx(np.var([1.5, 1.7, 1.6]))
    assert e["incorrect_inversions"] == [3]          # only non-zero counts are kept, like the reference
    assert sweep.result_key(18, 1000000, 18) == "m18_n1000000_t18"
    base, ncu, rt = sweep.create_output_directories(str(tmp_path / "benchmark_results"))
    assert os.path.isdir(ncu) and os.path.isdir(rt)
    p = tmp_path / "benchmark_results_1M.json"
    sweep.save_results({sweep.result_key(18, 1000000, 256): e, sweep.result_key(4, 1000000, 256): dict(e, matrix_size=4)}, p)
    d = plot.load(p)                                    # parallel_pivot/plot.py:30-35
    assert d["matrix_sizes"] == [4, 18] and d["incorrect_inversions"] == [3.0, 3.0]
    assert "avg_ms" in plot.summarise(d)


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from matrixinversion_b200.sharding import shard_range, reduce_verdict
from oracle import oracle as O                     # CPU stand-in for the kernel (tests only)
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
B, n = 1001, 6
rng = np.random.default_rng(5)
A = rng.uniform(0, 1, size=(B, n, n)) + n * np.eye(n)
lo, hi = shard_range(B, rank, world)
X, _ = O.lu_batched(A[lo:hi], 2)
if rank == 1: X[0] += 0.5                           # one bad inverse on rank 1
ok, bad, dev = O.verify_inv(A[lo:hi], X)
tot_bad, max_dev = reduce_verdict(bad, dev)
counts = torch.tensor([hi - lo]); dist.all_reduce(counts)
assert int(counts.item()) == B, counts
assert tot_bad == 1 and max_dev >= 0.4, (tot_bad, max_dev)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok", lo, hi)
"""


def test_two_rank_gloo_shards_and_reduces(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "ok 0 501" in outs[0] and "ok 501 1001" in outs[1]
