#!/usr/bin/env python3
"""Regenerate tests/golden/golden_verify_lu.npz (build container only: needs oracle/_ref built from
/root/reference by oracle/build_ref.sh).

The reference's verifyLUwithPivoting (parallel_pivot/verify.hpp:157-242) is disabled in its main()
because the kernel overwrites the factors with the inverse; it is still the only statement of what a
correct factorisation is.  This script runs THAT function (through oracle/_ref/libref_verify.so) on
the CPU oracle's LU factors of the reference's input files and stores its verdicts:

  ref_verify_lu/<file>/<dtype>/<mode>/<N>   int64[2]  (correct, incorrect) for [LU, LU with one entry
                                                      off by 0.01] -> the pair (1, 0) and (0, 1) when
                                                      the factorisation passes the 1e-3 predicate
  orc_lu/<file>/<mode>/<N>                  float32[N,N]  the oracle's factors (fp32, FMA), for drift detection
"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

FILES = ["mtrand32", "mtrand32_new1", "mtrand64", "matrix"]
SIZES = [1, 2, 3, 5, 8, 12, 16, 18, 20, 24, 31, 32]


def main():
    assert O.have_ref("ref_verify"), "run oracle/build_ref.sh first"
    z = np.load(os.path.join(ROOT, "tests/golden/inputs.npz"))
    out = {}
    for name in FILES:
        for N in SIZES:
            for dt, suf in ((np.float32, "f32"), (np.float64, "f64")):
                A = z[name + "_" + suf][: N * N].reshape(N, N).astype(dt)
                for mode in (0, 1, 2):
                    with np.errstate(all="ignore"):
                        LU, perm = O.lu_batched(A[None], mode, lu_only=True)
                    if not np.isfinite(LU).all():
                        continue
                    bad_lu = LU.copy()
                    bad_lu[0, N - 1, N - 1] += 0.01
                    both = np.concatenate([LU, bad_lu])
                    PA = A[perm[0]]
                    out["ref_verify_lu/%s/%s/%d/%d" % (name, suf, mode, N)] = np.array(
                        [O.ref_verify_lu_piv(PA, both[:1]), O.ref_verify_lu_piv(PA, both[1:])], dtype=np.int64).reshape(-1)
                    if dt == np.float32:
                        out["orc_lu/%s/%d/%d" % (name, mode, N)] = LU[0]
    np.savez_compressed(os.path.join(ROOT, "tests/golden/golden_verify_lu.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
