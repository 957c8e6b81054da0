"""Accuracy of the fp64 N = 32 DMMA (blocked) kernel against the DFMA (unblocked) one on the same matrices: residual
||A X - I||_F in fp64 on the device, per matrix, for both; evidence for profiles/r02_dmma.md."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import matrixinversion_b200 as lub
B, n = 200_000, 32
out = {}
for mode in ("parallel", "serial", "none"):
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.rand((B, n, n), generator=g, device="cuda", dtype=torch.float64)
    if mode == "none":
        A += n * torch.eye(n, device="cuda", dtype=torch.float64)
    eye = torch.eye(n, device="cuda", dtype=torch.float64)
    res = {}
    for name, opt in (("dmma", 2), ("dfma", 1)):   # LUB_OPT_FP64_TENSOR: 2 = always DMMA, 1 = never
        X = A.clone()
        lub.set_option("fp64_tensor", opt)
        try:
            lub.lu_batched_inplace(X, None, mode)
        finally:
            lub.set_option("fp64_tensor", 0)
        r = torch.linalg.matrix_norm(A @ X - eye)
        res[name] = torch.nan_to_num(r, nan=float("inf"))
    ratio = res["dmma"] / res["dfma"].clamp_min(1e-300)
    q = torch.tensor([0.5, 0.9, 0.99, 0.999], device="cuda", dtype=torch.float64)
    out[mode] = {"matrices": B,
                 "residual_quantiles_50_90_99_99.9_dmma": [float(v) for v in torch.quantile(res["dmma"][:100000], q)],
                 "residual_quantiles_50_90_99_99.9_dfma": [float(v) for v in torch.quantile(res["dfma"][:100000], q)],
                 "max_dmma": float(res["dmma"].max()), "max_dfma": float(res["dfma"].max()),
                 "ratio_quantiles_50_90_99_99.9": [float(v) for v in torch.quantile(ratio[:100000], q)],
                 "dmma_10x_worse": int((ratio > 10).sum()), "dfma_10x_worse": int((ratio < 0.1).sum()),
                 "dmma_above_1e-8": int((res["dmma"] > 1e-8).sum()), "dfma_above_1e-8": int((res["dfma"] > 1e-8).sum())}
print(json.dumps(out, indent=1))
