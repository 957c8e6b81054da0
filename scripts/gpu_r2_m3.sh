#!/bin/bash
# pivot_mode 3, two-phase kernel (prepass_getrf + lub_bulk_kernel<kModeLapack>): parity tests, then timing against the one-phase
# lane = row kernel (LUB_OPT_STAGING = 1) and mode 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lapack_layout.py -m gpu -x -q > gpurun_out/m3_pytest.log 2>&1
tail -15 gpurun_out/m3_pytest.log
python - <<'PY' > gpurun_out/m3_mode3_perf.jsonl 2>&1
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
import os
ns = [int(x) for x in os.environ.get("M3_NS", "5,8,9,10,12,14,16,17,18,20,24,27,31,32").split(",")]
for dt in (torch.float32, torch.float64):
    for n in ns:
        g=torch.Generator(device="cuda").manual_seed(n)
        A0=torch.rand((B,n,n),generator=g,device="cuda",dtype=dt)
        A=A0.clone()
        row={"dtype":str(dt),"n":n}
        for mode in ("parallel","lapack"):
            def run(pre):
                if pre: A.copy_(A0); return
                lub.lu_batched_inplace(A,None,mode)
            row[mode+"_ms"]=round(t(run),4)
        row["kernel"]=lub.kernel_name(n, "lapack", np.float32 if dt==torch.float32 else np.float64)
        lub.set_option("staging",1)
        row["lapack_lane_row_ms"]=round(t(run),4)
        lub.set_option("staging",0)
        def runf(pre):
            if pre: A.copy_(A0); return
            lub.lu_batched_factor_inplace(A,None,"lapack")
        row["lapack_lu_only_ms"]=round(t(runf),4)
        lub.set_option("staging",0)
        print(json.dumps(row), flush=True)
PY
cat gpurun_out/m3_mode3_perf.jsonl
