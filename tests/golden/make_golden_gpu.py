#!/usr/bin/env python3
"""Regenerate tests/golden/golden_gpu_ref.npz ON A GPU BOX (needs oracle/_ref, i.e. the
reference's own CUDA kernels rebuilt for sm_100 from /root/reference by oracle/build_ref.sh):

    gpurun -- 'python tests/golden/make_golden_gpu.py'   ->  gpurun_out/golden_gpu_ref.npz
    cp gpurun_out/golden_gpu_ref.npz tests/golden/

The reference repository holds no golden output vectors.  These are outputs of
the reference ITSELF -- `batched_lu_subwarp` of templated/, serial_pivot/, parallel_pivot/
(unmodified for the inverse; the one-store pivot-exporting patch for the permutation vector) --
run on a B200 on the reference's input files (token-stream prefixes, Q4), so that value parity
is pinned to something the reference produced:

  inv/<file>/<dtype>/<mode>/<N>   T[N,N]     the reference kernel's in-place result
  piv/<file>/<dtype>/<mode>/<N>   int32[N]   its shared-memory pivots[] (modes 1, 2)
fp64 parallel is stored for even N only: upstream faults with "misaligned address" for odd N.
"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

NS = [1, 2, 3, 4, 5, 8, 12, 16, 17, 18, 20, 24, 27, 31, 32]

def main():
    z = np.load(os.path.join(ROOT, "tests/golden/inputs.npz"))
    out = {}
    for name in ("mtrand32", "mtrand32_new1", "mtrand64"):
        for dt, suf in ((np.float32, "f32"), (np.float64, "f64")):
            for n in NS:
                A = z[name + "_" + suf][: n * n].reshape(1, n, n).astype(dt)
                for mode in (0, 1, 2):
                    if mode == 2 and suf == "f64" and n % 2:
                        continue
                    X, _, _ = O.ref_gpu_invert(A, mode)
                    out["inv/%s/%s/%d/%d" % (name, suf, mode, n)] = X[0]
                    if mode:
                        Xp, piv, _ = O.ref_gpu_invert(A, mode, want_piv=True)
                        assert np.array_equal(Xp, X, equal_nan=True)
                        out["piv/%s/%s/%d/%d" % (name, suf, mode, n)] = piv[0]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "golden_gpu_ref.npz"), **out)
    print("wrote", len(out), "arrays")

if __name__ == "__main__":
    main()
