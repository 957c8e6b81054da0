#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lapack_layout.py -m gpu -q > gpurun_out/k_pytest.log 2>&1
tail -30 gpurun_out/k_pytest.log
python - <<'PY' > gpurun_out/k_mode3_perf.jsonl 2>&1
import json, torch, numpy as np
import matrixinversion_b200 as lub
def t(fn, it=4):
    best=1e9
    for i in range(it+1):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        pre=fn(None)
        e0.record(); fn(pre); e1.record(); torch.cuda.synchronize()
        if i: best=min(best,e0.elapsed_time(e1))
    return best
B=1_000_000
for dt in (torch.float32, torch.float64):
    for n in (4, 8, 16, 18, 24, 32):
        g=torch.Generator(device="cuda").manual_seed(n)
        A0=torch.rand((B,n,n),generator=g,device="cuda",dtype=dt)
        A=A0.clone()
        row={"dtype":str(dt),"n":n}
        for mode in ("parallel","lapack"):
            def run(pre):
                if pre is None: A.copy_(A0); return 1
                lub.lu_batched_inplace(A,None,mode)
            row[mode+"_ms"]=t(run)
        if n<=8:
            I0=A0.permute(1,2,0).contiguous(); I=I0.clone()
            for mode in ("none","parallel","lapack"):
                def run(pre):
                    if pre is None: I.copy_(I0); return 1
                    lub.lu_batched_inplace(I,None,mode,layout="interleaved")
                row["interleaved_"+mode+"_ms"]=t(run)
            def run(pre):
                if pre is None: A.copy_(A0); return 1
                lub.lu_batched_inplace(A,None,"none")
            row["none_ms"]=t(run)
            es=4 if dt==torch.float32 else 8
            row["interleaved_parallel_frac_of_6491"]=2*n*n*es*B/row["interleaved_parallel_ms"]/1e6/6491.2
        print(json.dumps(row),flush=True)
        del A,A0
PY
cat gpurun_out/k_mode3_perf.jsonl
