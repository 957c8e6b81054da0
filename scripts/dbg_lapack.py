import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
import matrixinversion_b200 as lub
rng = np.random.default_rng(7)
for n, dtype in ((6, np.float32), (17, np.float32), (32, np.float32)):
    A = rng.uniform(0, 1, size=(40, n, n)).astype(dtype)
    A0 = A.copy()
    A[3] = 0; A[5][:, 2] = 0
    if n > 4: A[7][4] = A[7][1]
    for name, M in (("plain", A0), ("with singular", A)):
        dA = torch.from_numpy(M.copy()).cuda()
        piv = torch.zeros((40, n), dtype=torch.int32, device="cuda"); info = torch.zeros(40, dtype=torch.int32, device="cuda")
        lub.lu_batched_inplace(dA, piv, "lapack", info=info)
        X = dA.cpu().numpy(); inf = info.cpu().numpy()
        with np.errstate(all="ignore"):
            res = np.abs(M.astype(np.float64) @ X.astype(np.float64) - np.eye(n)).max(axis=(1, 2))
        print(n, name, "info", inf.tolist())
        print("   res", np.array2string(res, precision=1, max_line_width=250))
