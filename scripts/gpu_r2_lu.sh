#!/bin/bash
# factors-only entry point (lu_batched_factor_inplace), modes 0-3 against the inverse of the same mode
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/lu_only_perf.jsonl 2>&1
import json, torch, numpy as np
import matrixinversion_b200 as lub
def t(fn, it=3):
    best=1e9
    for i in range(it+1):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        fn(True)
        e0.record(); fn(False); e1.record(); torch.cuda.synchronize()
        if i: best=min(best,e0.elapsed_time(e1))
    return best
B=1_000_000
for dt in (torch.float32, torch.float64):
    for n in [int(x) for x in __import__("os").environ.get("LU_NS", "8,16,18,24,31,32").split(",")]:
        g=torch.Generator(device="cuda").manual_seed(n)
        A0=torch.rand((B,n,n),generator=g,device="cuda",dtype=dt)
        A0d=A0 + n*torch.eye(n,device="cuda",dtype=dt)
        A=A0.clone()
        row={"dtype":str(dt),"n":n}
        for mode in ("none","serial","parallel","lapack"):
            src = A0d if mode=="none" else A0
            def run(pre):
                if pre: A.copy_(src); return
                lub.lu_batched_inplace(A,None,mode)
            def runf(pre):
                if pre: A.copy_(src); return
                lub.lu_batched_factor_inplace(A,None,mode)
            row[mode+"_inv_ms"]=round(t(run),3); row[mode+"_lu_ms"]=round(t(runf),3)
        print(json.dumps(row), flush=True)
PY
cat gpurun_out/lu_only_perf.jsonl
