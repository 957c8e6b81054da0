#!/bin/bash
# pivot_mode 3 on the 2-D lane grid: parity tests, then timing against the lane = row kernel (LUB_OPT_STAGING = 1) and mode 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lapack_layout.py -m gpu -x -q > gpurun_out/t14_pytest.log 2>&1
tail -30 gpurun_out/t14_pytest.log
python - <<'PY' > gpurun_out/t14_mode3_perf.jsonl 2>&1
import json, torch, numpy as np
import matrixinversion_b200 as lub
def t(fn, it=4):
    best=1e9
    for i in range(it+1):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        fn(True)
        e0.record(); fn(False); e1.record(); torch.cuda.synchronize()
        if i: best=min(best,e0.elapsed_time(e1))
    return best
B=1_000_000
for dt in (torch.float32, torch.float64):
    for n in (17, 18, 20, 24, 27, 31, 32):
        g=torch.Generator(device="cuda").manual_seed(n)
        A0=torch.rand((B,n,n),generator=g,device="cuda",dtype=dt)
        A=A0.clone()
        row={"dtype":str(dt),"n":n}
        for mode in ("parallel","lapack"):
            def run(pre):
                if pre: A.copy_(A0); return
                lub.lu_batched_inplace(A,None,mode)
            row[mode+"_ms"]=t(run)
        lub.set_option("staging",1)
        row["lapack_lane_row_ms"]=t(run)
        lub.set_option("staging",0)
        def runf(pre):
            if pre: A.copy_(A0); return
            lub.lu_batched_factor_inplace(A,None,"lapack")
        row["lapack_lu_only_ms"]=t(runf)
        print(json.dumps(row), flush=True)
PY
cat gpurun_out/t14_mode3_perf.jsonl
