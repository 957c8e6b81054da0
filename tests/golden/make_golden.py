#!/usr/bin/env python3
"""Regenerate tests/golden/golden_cpu.npz (build container only: needs oracle/_ref built
from /root/reference by oracle/build_ref.sh).

The reference holds no golden output vectors; what it does hold
is executable: pivotedA (exact serial pivot order) and verifyInv (the 1e-3 predicate).
This script runs THOSE -- the reference's own code through oracle/_ref/libref_verify.so --
on the reference's input files for every N in 1..32 and stores the answers, together with
the CPU oracle's outputs (so that oracle drift is detected without the reference):

  ref_serial_perm/<file>/<dtype>/<N>   int32[N]   pivots[] from the reference's pivotedA
  ref_verify/<file>/<mode>/<N>         int64[2]   (correct, incorrect) from the reference's
                                                  verifyInv on the oracle's fp32 inverse
  ref_l1/<file>/<N>                    float64    the reference's calc_cond_num output
  orc_perm|orc_steps/<file>/<dtype>/<mode>/<N>    oracle permutation vector / per-step pivots
  orc_inv/<file>/<mode>/<N>            float32[N,N] oracle inverse (fp32, FMA)
"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

FILES32 = ["mtrand32", "mtrand32_new1", "mtrand32_new", "mtrand64", "matrix"]

def main():
    assert O.have_ref("ref_verify"), "run oracle/build_ref.sh first"
    z = np.load(os.path.join(ROOT, "tests/golden/inputs.npz"))
    out = {}
    for name in FILES32:
        for N in range(1, 33):
            for dt, suf in ((np.float32, "f32"), (np.float64, "f64")):
                A = z[name + "_" + suf][: N * N].reshape(N, N).astype(dt)
                _, piv = O.ref_pivotedA(A)
                out["ref_serial_perm/%s/%s/%d" % (name, suf, N)] = piv
                for mode in (1, 2):
                    _, perm, steps = O.lu_batched(A[None], mode, lu_only=True, want_steps=True)
                    out["orc_perm/%s/%s/%d/%d" % (name, suf, mode, N)] = perm[0]
                    out["orc_steps/%s/%s/%d/%d" % (name, suf, mode, N)] = steps[0]
            A = z[name + "_f32"][: N * N].reshape(N, N)
            out["ref_l1/%s/%d" % (name, N)] = np.float64(O.ref_calc_cond_num(A))
            for mode in (0, 1, 2):
                with np.errstate(all="ignore"):
                    X, _ = O.lu_batched(A[None], mode)
                out["orc_inv/%s/%d/%d" % (name, mode, N)] = X[0]
                ok, bad = O.ref_verify_inv(A[None], X)
                out["ref_verify/%s/%d/%d" % (name, mode, N)] = np.array([ok, bad], dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/golden_cpu.npz"), **out)
    print("wrote", len(out), "arrays")

if __name__ == "__main__":
    main()
