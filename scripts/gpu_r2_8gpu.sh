#!/bin/bash
# 8-GPU box: pinned-copy ceiling, bench scaling (device-timed and end-to-end), BASELINE config 4 sweep at 2/4/8 GPUs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/g8_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/g8_lscpu.txt
python - <<'PY'
import numpy as np
d = np.load("tests/golden/inputs.npz")
with open("/tmp/mtrand32_new1.txt", "w") as f:
    t = d["mtrand32_new1_f32"]
    for i in range(0, len(t), 32):
        f.write(" ".join("%.9g" % v for v in t[i:i + 32]) + "\n")
PY
P=29600
for g in 1 2 4 8; do
  P=$((P+1))
  if [ $g = 1 ]; then
    python scripts/memcpy_ceiling.py > gpurun_out/g8_ceiling_$g.json 2>gpurun_out/g8_ceiling_$g.err
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-cublas > gpurun_out/g8_bench_$g.json 2>gpurun_out/g8_bench_$g.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $P scripts/memcpy_ceiling.py > gpurun_out/g8_ceiling_$g.json 2>gpurun_out/g8_ceiling_$g.err
    P=$((P+1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $P bench.py --gpus $g --steps 10 --warmup 3 --no-cublas > gpurun_out/g8_bench_$g.json 2>gpurun_out/g8_bench_$g.err
    P=$((P+1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $P -m matrixinversion_b200.sweep --variant parallel_pivot --input /tmp/mtrand32_new1.txt --sizes 2-32 --batches 1000000 --runs 3 --warm --out gpurun_out/g8_sweep_cfg4_${g}gpu > gpurun_out/g8_sweep_$g.log 2>&1
  fi
  tail -1 gpurun_out/g8_ceiling_$g.json | cut -c1-400
  tail -1 gpurun_out/g8_bench_$g.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['frac_of_copy_ceiling'], d['e2e']['copy_ceiling'])"
done
# host-multi entry point from ONE process over all 8 GPUs
python - <<'PY' > gpurun_out/g8_host_multi.json 2>&1
import json, time, numpy as np, torch
import matrixinversion_b200 as lub
n, per = 32, 1_000_000
out = {}
for nd in (1, 2, 4, 8):
    batch = per * nd
    H = torch.empty((batch, n, n), dtype=torch.float32, pin_memory=True)
    H.uniform_(0, 1)
    A = H.numpy()
    lub.lu_batched_inplace_host_multi(A, None, "parallel", n_devices=nd)
    t0 = time.perf_counter()
    lub.lu_batched_inplace_host_multi(A, None, "parallel", n_devices=nd)
    dt = time.perf_counter() - t0
    out[str(nd)] = {"matrices_per_s": batch / dt, "GBps_each_way_total": batch * n * n * 4 / dt / 1e9}
    del H, A
print(json.dumps(out))
PY
cat gpurun_out/g8_host_multi.json
