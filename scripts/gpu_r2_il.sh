#!/bin/bash
# batch-interleaved layout against matrix-major, N = 2..8, after the literal-tree fix of the in-register inversion
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/il_perf.jsonl 2>&1
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
peak=json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
for dt in (torch.float32, torch.float64):
    es = 4 if dt==torch.float32 else 8
    for n in (2,3,4,5,6,7,8):
        g=torch.Generator(device="cuda").manual_seed(n)
        A0=torch.rand((B,n,n),generator=g,device="cuda",dtype=dt)
        I0=A0.permute(1,2,0).contiguous()
        A=A0.clone(); I=I0.clone()
        row={"dtype":str(dt),"n":n}
        for mode in ("none","serial","parallel","lapack"):
            src = A0 + (n*torch.eye(n,device="cuda",dtype=dt) if mode=="none" else 0)
            isrc = src.permute(1,2,0).contiguous()
            def run(pre):
                if pre: A.copy_(src); return
                lub.lu_batched_inplace(A,None,mode)
            def runi(pre):
                if pre: I.copy_(isrc); return
                lub.lu_batched_inplace(I,None,mode,layout="interleaved")
            a=t(run); b=t(runi)
            gb=2*n*n*es*B/1e6
            row[mode]={"matrix_major_ms":round(a,4),"interleaved_ms":round(b,4),"matrix_major_frac":round(gb/a/peak,3),"interleaved_frac":round(gb/b/peak,3)}
        print(json.dumps(row), flush=True)
PY
cut -c1-400 gpurun_out/il_perf.jsonl
