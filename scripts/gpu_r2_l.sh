#!/bin/bash
# 2-GPU box: host multi-device test, copy ceiling at 1 and 2 ranks, bench at 2 ranks
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/l_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/l_lscpu.txt
timeout 600 python -m pytest tests/test_gpu_lapack_layout.py -m gpu -q > gpurun_out/l_pytest.log 2>&1; tail -5 gpurun_out/l_pytest.log
python scripts/memcpy_ceiling.py > gpurun_out/l_ceiling_1.json 2>&1; cat gpurun_out/l_ceiling_1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/memcpy_ceiling.py > gpurun_out/l_ceiling_2.json 2>gpurun_out/l_ceiling_2.err; tail -1 gpurun_out/l_ceiling_2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/l_bench_2.json 2>gpurun_out/l_bench_2.err; tail -1 gpurun_out/l_bench_2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
