"""CPU tests: the oracle (oracle/lu_oracle.c) pinned against the reference.

Pins, in order of strength:
  1. the reference's own code run in the build container (pivotedA, verifyInv, calc_cond_num
     through oracle/_ref/libref_verify.so) -> committed as tests/golden/golden_cpu.npz;
  2. the known-answer pivot sequences listed in SURVEY.md section 8(c);
  3. when oracle/_ref is present (it travels with the repo), the live reference code;
  4. an independent numpy twin and numpy.linalg.
"""
import numpy as np
import pytest

from conftest import FILES, synthetic, template
from oracle import oracle as O

MODES = (0, 1, 2)


def test_oracle_matches_reference_pivotedA_golden(inputs, golden):
    """Serial pivot permutation == the reference's pivotedA for every file, N, dtype."""
    for name in FILES:
        for n in range(1, 33):
            for dt, suf in ((np.float32, "f32"), (np.float64, "f64")):
                A = template(inputs, name, n, dt)
                _, perm = O.lu_batched(A[None], O.MODE_SERIAL, lu_only=True)
                ref = golden["ref_serial_perm/%s/%s/%d" % (name, suf, n)]
                assert perm[0].tolist() == ref.tolist(), (name, n, suf)
                assert O.pivotedA(A)[1].tolist() == ref.tolist(), (name, n, suf)


def test_known_answer_sequences_from_survey(inputs):
    A = template(inputs, "mtrand32", 32)
    _, perm = O.lu_batched(A[None], 1, lu_only=True)
    assert perm[0].tolist() == [7, 21, 11, 16, 3, 10, 12, 17, 28, 4, 6, 23, 14, 18, 22, 24, 26, 1, 13, 31, 20,
                                8, 5, 27, 9, 2, 29, 0, 25, 15, 30, 19]
    A = template(inputs, "mtrand32", 18)
    assert O.lu_batched(A[None], 1, lu_only=True, want_steps=True)[2][0].tolist() == \
        [9, 12, 10, 3, 11, 17, 8, 15, 15, 12, 17, 14, 15, 17, 17, 16, 16, 17]
    assert O.lu_batched(A[None], 2, lu_only=True, want_steps=True)[2][0].tolist() == \
        [4, 12, 10, 3, 11, 17, 8, 15, 9, 12, 17, 15, 12, 17, 17, 16, 17, 17]
    A = template(inputs, "mtrand32_new1", 20)
    assert O.lu_batched(A[None], 1, lu_only=True, want_steps=True)[2][0].tolist() == \
        [7, 3, 4, 8, 17, 11, 17, 14, 14, 11, 14, 14, 17, 14, 14, 18, 18, 18, 18, 19]
    assert O.lu_batched(A[None], 2, lu_only=True, want_steps=True)[2][0].tolist() == \
        [7, 3, 4, 16, 17, 11, 17, 14, 14, 11, 16, 15, 15, 16, 16, 18, 19, 19, 19, 19]


def test_serial_and_parallel_agree_for_powers_of_two(inputs):
    """the tree drops nothing when TPM is a power of two and values are distinct."""
    for name in ("mtrand32", "mtrand32_new1"):
        for n in (1, 2, 4, 8, 16, 32):
            A = template(inputs, name, n)
            s = O.lu_batched(A[None], 1, lu_only=True, want_steps=True)[2]
            p = O.lu_batched(A[None], 2, lu_only=True, want_steps=True)[2]
            assert s.tolist() == p.tolist()
    # ... and differ where the survey says they do
    for name, ns in (("mtrand32", (18, 20, 24, 31)), ("mtrand32_new1", (20, 24, 31))):
        for n in ns:
            A = template(inputs, name, n)
            s = O.lu_batched(A[None], 1, lu_only=True, want_steps=True)[2]
            p = O.lu_batched(A[None], 2, lu_only=True, want_steps=True)[2]
            assert s.tolist() != p.tolist(), (name, n)


def test_oracle_outputs_match_golden(inputs, golden):
    """Oracle drift check: permutations, steps and fp32 inverses are what was committed."""
    for name in FILES:
        for n in range(1, 33):
            A = template(inputs, name, n)
            for mode in MODES:
                with np.errstate(all="ignore"):
                    X, perm, steps = O.lu_batched(A[None], mode, want_steps=True)
                g = golden["orc_inv/%s/%d/%d" % (name, mode, n)]
                assert np.array_equal(X[0], g, equal_nan=True), (name, n, mode)
                if mode:
                    assert perm[0].tolist() == golden["orc_perm/%s/f32/%d/%d" % (name, mode, n)].tolist()
                    assert steps[0].tolist() == golden["orc_steps/%s/f32/%d/%d" % (name, mode, n)].tolist()


def test_verify_inv_matches_reference_verifyInv_golden(inputs, golden):
    """The restated predicate gives the reference verifyInv's verdict on every golden inverse."""
    for name in FILES:
        for n in range(1, 33):
            A = template(inputs, name, n)
            for mode in MODES:
                X = golden["orc_inv/%s/%d/%d" % (name, mode, n)]
                ok, bad, _ = O.verify_inv(A[None], X[None])
                assert [ok, bad] == golden["ref_verify/%s/%d/%d" % (name, mode, n)].tolist(), (name, n, mode)


def test_config1_passes_reference_check(inputs, golden):
    """BASELINE config 1: N=32, mtrand32.txt, no pivoting -> verifyInv reports correct."""
    assert golden["ref_verify/mtrand32/0/32"].tolist() == [1, 0]
    assert golden["ref_verify/mtrand32_new1/0/20"].tolist() == [1, 0]   # config 2 template
    assert golden["ref_verify/mtrand32_new1/2/18"].tolist() == [1, 0]   # config 3 template


def test_calc_cond_num_is_the_l1_norm(inputs, golden):
    """the reference's calc_cond_num returns ||A||_1 (no augmented block)."""
    from matrixinversion_b200.api import l1_norm
    for name in ("mtrand32", "mtrand32_new1"):
        for n in (1, 5, 18, 32):
            A = template(inputs, name, n)
            assert golden["ref_l1/%s/%d" % (name, n)] == pytest.approx(l1_norm(A), rel=1e-6)
    assert golden["ref_l1/mtrand32/32"] == pytest.approx(2063.09, rel=1e-5)


@pytest.mark.skipif(not O.have_ref("ref_verify"), reason="oracle/_ref not built")
def test_live_reference_host_code(inputs):
    """When oracle/_ref travels with the repo: the reference's pivotedA / verifyInv, live."""
    rng = np.random.default_rng(7)
    for n in (1, 2, 3, 7, 16, 18, 31, 32):
        for dt in (np.float32, np.float64):
            A = rng.uniform(-1, 1, size=(n, n)).astype(dt)
            assert O.ref_pivotedA(A)[1].tolist() == O.lu_batched(A[None], 1, lu_only=True)[1][0].tolist()
            A = rng.integers(-2, 3, size=(n, n)).astype(dt)  # heavy ties, zeros
            assert O.ref_pivotedA(A)[1].tolist() == O.lu_batched(A[None], 1, lu_only=True)[1][0].tolist()
    A = synthetic(12, 50, np.float32, dominant=True)
    X, _ = O.lu_batched(A, 0)
    X[::7] *= 1.01  # break some
    assert O.ref_verify_inv(A, X) == O.verify_inv(A, X)[:2]


def test_numpy_twin_agrees_bitwise(inputs):
    for n in (1, 2, 3, 5, 8, 11):
        for name in ("mtrand32_new1", "mtrand32_new"):
            A = template(inputs, name, n)
            for mode in MODES:
                with np.errstate(all="ignore"):
                    Xn, pn, sn = O.numpy_invert_one(A, mode)
                    Xc, pc, sc = O.lu_batched(A[None], mode, use_fma=False, want_steps=True)
                assert np.array_equal(Xn, Xc[0], equal_nan=True)
                assert pn.tolist() == pc[0].tolist() and sn.tolist() == sc[0].tolist()


def test_oracle_inverse_against_lapack():
    for n in (1, 4, 9, 16, 25, 32):
        for dt, tol in ((np.float32, 2e-3), (np.float64, 1e-10)):
            A = synthetic(n, 8, dt, dominant=True)
            for mode in MODES:
                X, perm = O.lu_batched(A, mode)
                ref = np.linalg.inv(A.astype(np.float64))
                assert np.abs(X - ref).max() <= tol * np.abs(ref).max()
                assert sorted(perm[0].tolist()) == list(range(n))


def test_oracle_edge_cases():
    X, perm = O.lu_batched(np.zeros((0, 4, 4), np.float32), 2)      # empty batch
    assert X.shape == (0, 4, 4)
    X, perm = O.lu_batched(np.full((3, 1, 1), 4.0, np.float64), 1)  # N = 1
    assert np.allclose(X, 0.25) and perm.tolist() == [[0]] * 3
    with np.errstate(all="ignore"):                                  # singular: inf/NaN, no status (Q7)
        X, _ = O.lu_batched(np.zeros((1, 3, 3), np.float32), 0)
    assert not np.isfinite(X).any()
    # literal matrices left in the reference's main() (parallel_pivot/luBatchedInplace.cu:20-25)
    for lit in ([2, 3, 4, 5], [4, 11, 3, 40, 10, 4, 2, 40, 2], [4, 11, 3, 7, 4, 10, 4, 9, 2, 4, 2, 1, 7, 9, 18, 30],
                [2, 7, 1, 5, 3, -2, 0, 1, 1, 5, 3, 4, 7, 3, 2, 8]):
        n = int(round(len(lit) ** 0.5))
        A = np.array(lit, np.float64).reshape(1, n, n)
        for mode in (1, 2):
            X, _ = O.lu_batched(A, mode)
            assert np.allclose(A[0] @ X[0], np.eye(n), atol=1e-9)
    A = 10 * np.eye(4, dtype=np.float32)[None]
    assert np.allclose(O.lu_batched(A, 2)[0], 0.1 * np.eye(4))


def test_tree_rank_rule_equals_literal_tree():
    """The kernel does not run the reference's shared-memory tree; it orders candidates by
    (|value|, static slot rank) (csrc/lub_kernel.cuh: tree_slot_rank).  Check that rule --
    restated here in Python -- against the oracle's literal tree on tie-heavy inputs."""
    def slot_rank(t, tpm):
        loc, rank, lvl, s = t, 0, 0, tpm // 2
        while s > 0:
            if loc >= 2 * s:
                return -1
            if loc >= s:
                loc -= s
                rank |= 1 << lvl
            s >>= 1
            lvl += 1
        return rank

    def rule(A, mode):
        n = A.shape[0]
        perm, steps = list(range(n)), []
        for k in range(n):
            best = (abs(A[perm[k], k]), 0, k)
            for t in range(n - 1 - k):
                rk = t if mode == 1 else slot_rank(t, n)
                if rk < 0:
                    continue
                v = abs(A[perm[k + 1 + t], k])
                if v > best[0] or (v == best[0] and rk + 1 < best[1]):
                    best = (v, rk + 1, k + 1 + t)
            steps.append(best[2])
            perm[k], perm[best[2]] = perm[best[2]], perm[k]
        return perm, steps

    rng = np.random.default_rng(0)
    for n in range(1, 33):
        for trial in range(12):
            A = rng.integers(-3, 4, size=(n, n)).astype(np.float32)
            for mode in (1, 2):
                _, po, so = O.lu_batched(A[None], mode, lu_only=True, want_steps=True)
                p, s = rule(A, mode)
                assert s == so[0].tolist() and p == po[0].tolist(), (n, mode, trial)


def test_oracle_matches_reference_gpu_kernel_goldens(inputs):
    """tests/golden/golden_gpu_ref.npz = outputs of the reference's OWN CUDA kernels run on a
    B200 (make_golden_gpu.py).  The oracle must reproduce their permutation vectors exactly --
    this pins the restated find_pivot_parallel tree, dropped slots included (N = 17, 18, 20,
    24, 27, 31), to the real kernel -- and their inverses to rounding (the kernels were built
    with --use_fast_math, the oracle divides exactly)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "golden_gpu_ref.npz"))
    n_piv = n_val = 0
    for key in g.files:
        kind, name, suf, mode, n = key.split("/")
        mode, n = int(mode), int(n)
        dt = np.float32 if suf == "f32" else np.float64
        A = template(inputs, name, n, dt)
        with np.errstate(all="ignore"):
            X, perm = O.lu_batched(A[None], mode)
        if kind == "piv":
            assert perm[0].tolist() == g[key].tolist(), key
            n_piv += 1
        else:
            ref = g[key]
            if not np.isfinite(ref).all():
                continue
            eps = 2.0 ** -23 if dt == np.float32 else 2.0 ** -52
            kappa = np.linalg.cond(A.astype(np.float64))
            Xt = np.linalg.inv(A.astype(np.float64))
            err_ref = np.abs(ref - Xt).max()
            tol = max(16 * n * eps * kappa * np.abs(Xt).max(), 8 * err_ref)
            assert np.abs(X[0].astype(np.float64) - ref).max() <= tol, key
            n_val += 1
    assert n_piv >= 150 and n_val >= 200
