"""GPU tests (run with `-m gpu`) of the two round-2 extensions of the hot path, both through the C ABI:

  * pivot_mode 3 -- true partial pivoting with LAPACK getrf semantics: ipiv against
    LAPACK's own getrf (scipy), `info` for exactly-zero pivots, residual constant c = 8, the reference's
    verifyLUwithPivoting predicate (parallel_pivot/verify.hpp:157-242) through lu_batched_verify_lu;
  * the batch-interleaved layout (north star; templated/luBatchedInplace.cuh:89-97 is the default layout):
    bitwise equal to the matrix-major kernels after transposition, every mode, n = 1..8.
"""
import numpy as np
import pytest

import matrixinversion_b200 as lub
from conftest import synthetic
from oracle import oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EPS = {np.dtype(np.float32): 2.0 ** -23, np.dtype(np.float64): 2.0 ** -52}
C_LAPACK = 8.0   # c = 8 for a true partial-pivoting mode


def gpu_lapack(A, lu_only=False):
    dA = torch.from_numpy(np.ascontiguousarray(A)).cuda()
    b, n = A.shape[0], A.shape[1]
    piv = torch.full((b, n), -7, dtype=torch.int32, device="cuda")
    info = torch.full((b,), -7, dtype=torch.int32, device="cuda")
    (lub.lu_batched_factor_inplace if lu_only else lub.lu_batched_inplace)(dA, piv, "lapack", info=info)
    torch.cuda.synchronize()
    return dA.cpu().numpy(), piv.cpu().numpy(), info.cpu().numpy()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lapack_pivoting_matches_getrf_and_inverts(dtype):
    eps = EPS[np.dtype(dtype)]
    mismatching = 0
    for n in range(1, 33):
        A = synthetic(n, 203, dtype)                       # distinct, NOT dominant: real pivoting
        X, ipiv, info = gpu_lapack(A)
        _, ipiv_ref, info_ref = O.lapack_getrf(A)
        assert np.array_equal(info, info_ref) and not info.any(), n
        same = (ipiv == ipiv_ref).all(axis=1)
        # LAPACK's recursive getrf sums the updates in another order: two candidates within rounding distance of each
        # other may be ranked differently (expected ~1e-6 per comparison in fp32).  Such a matrix must be rare and its
        # factorisation still has to be a partial-pivoting one (checked through the residual below like all the others).
        mismatching += int((~same).sum())
        assert (~same).sum() <= 1, (n, int((~same).sum()))
        A64 = A.astype(np.float64)
        kappa = np.linalg.cond(A64)
        res = np.linalg.norm(A64 @ X.astype(np.float64) - np.eye(n), axis=(1, 2))
        ok = kappa < 0.001 / eps
        assert np.all(res[ok] <= C_LAPACK * n * eps * kappa[ok]), (n, float((res[ok] / (n * eps * kappa[ok])).max()))
        # the reference's own predicate: nothing a partial-pivoting inverse of a uniform(0,1) matrix should miss often
        good, bad, _ = lub.verify_inv(A, X)
        assert bad <= 2, (n, bad)
    assert mismatching <= 3, mismatching


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lapack_factors_pass_the_references_check(dtype):
    """lu_batched_factor_inplace(..., "lapack"): P A = L U as getrf stores it; |L| <= 1; componentwise backward
    error bound; verifyLUwithPivoting through lu_batched_verify_lu with the permutation made from ipiv."""
    eps = EPS[np.dtype(dtype)]
    for n in (1, 2, 3, 5, 8, 13, 16, 17, 24, 31, 32):
        A = synthetic(n, 101, dtype)
        LU, ipiv, info = gpu_lapack(A, lu_only=True)
        LUr, ipiv_ref, _ = O.lapack_getrf(A)
        assert not info.any()
        perm = lub.ipiv_to_perm(ipiv)
        L = np.tril(LU.astype(np.float64), -1) + np.eye(n)
        U = np.triu(LU.astype(np.float64))
        assert np.abs(np.tril(LU, -1)).max(initial=0.0) <= 1.0 + 4 * eps     # multipliers bounded: the point of the mode
        PA = np.take_along_axis(A.astype(np.float64), perm[:, :, None].astype(np.int64), axis=1)
        bound = 2.0 * n * eps * (np.abs(L) @ np.abs(U)) + 1e-300
        assert np.all(np.abs(PA - L @ U) <= bound), (n, float((np.abs(PA - L @ U) / bound).max()))
        same = (ipiv == ipiv_ref).all(axis=1)
        assert (~same).sum() <= 1
        # same pivots -> the same factors up to rounding
        assert np.allclose(LU[same], LUr[same], rtol=0, atol=64 * n * eps * np.abs(LUr).max())
        ok, bad, _ = lub.verify_lu(A, LU, perm)
        assert (ok, bad) == (101, 0), (n, ok, bad)


def test_lapack_info_reports_the_first_zero_pivot():
    rng = np.random.default_rng(7)
    for n, dtype in ((6, np.float32), (17, np.float32), (32, np.float32), (9, np.float64), (32, np.float64)):
        A = rng.uniform(0, 1, size=(40, n, n)).astype(dtype)
        A[3] = 0                                  # zero matrix: info = 1
        A[5][:, 2] = 0                            # zero column 2: the first zero pivot is U(3,3)
        if n > 4:
            A[7][4] = A[7][1]                     # two equal rows: singular, but NOT an exactly-zero pivot in floating
                                                  # point (a - p * fl(a / p) != 0): LAPACK reports info = 0 too
        X, ipiv, info = gpu_lapack(A)
        _, ipiv_ref, info_ref = O.lapack_getrf(A)
        assert info[3] == 1 and info[5] == 3
        # matrix 7 (two equal rows) hangs on whether fl(p * fl(1 / p)) == 1 for the pivot p the equal rows meet at: a last
        # pivot of exactly zero (info = n) and a tiny one (info = 0) are both what a getrf can return for it
        keep = np.arange(40) != 7
        assert np.array_equal((info != 0)[keep], (info_ref != 0)[keep])
        assert info[7] in (0, n)
        first = np.where((info != 0) & keep)[0]
        assert np.array_equal(info[first], info_ref[first])
        for b in range(40):
            k = info[b] if info[b] else n          # pivots are comparable up to and including the zero pivot's step
            assert np.array_equal(ipiv[b, :k], ipiv_ref[b, :k]), (n, b)
        good = (info == 0) & (np.linalg.cond(A.astype(np.float64)) < 1e6)
        assert good.sum() == 37
        res = np.abs(A[good].astype(np.float64) @ X[good].astype(np.float64) - np.eye(n)).max()
        assert res < (1e-2 if dtype == np.float32 else 1e-9)
    # first-maximum tie rule (isamax) on exact ties in the first column
    A = np.array([[[2.0, 1, 0], [-2, 0, 1], [2, 3, 5]]], dtype=np.float32)
    _, ipiv, _ = gpu_lapack(A)
    assert ipiv[0, 0] == 1
    # the reference's modes have no status: the array reads 0
    dA = torch.zeros((4, 5, 5), device="cuda")
    info = torch.full((4,), 9, dtype=torch.int32, device="cuda")
    lub.lu_batched_inplace(dA, None, "parallel", info=info)
    assert info.cpu().tolist() == [0, 0, 0, 0]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_two_phase_lapack_kernels_agree_with_the_one_phase_kernels(dtype):
    """pivot_mode 3 from n = 9 on runs the two-phase kernels (prepass_getrf: LU factorisation in the lane = row layout for the
    permutation, then the permuted-load Gauss-Jordan of modes 1 / 2; on the bulk-copy image, or the swizzled TMA image where rows
    are whole 128-byte lines).  LUB_OPT_STAGING = 1 selects the one-phase lane = row kernel (lub_lapack.cuh): same getf2
    recurrence, so ipiv and info must be identical -- ragged batches, singular matrices included -- the factors equal to
    rounding of the reciprocal, and the inverses within the c = 8 residual bound of each other's matrix."""
    eps = EPS[np.dtype(dtype)]
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    for n in (9, 12, 16, 17, 20, 23, 27, 31, 32):
        name = lub.kernel_name(n, "lapack", dtype)
        lines = (n * np.dtype(dtype).itemsize) % 128 == 0
        assert name == ("lub_tma_kernel<getrf>" if lines else "lub_bulk_kernel<getrf>"), (n, name)
        for batch in (1003, 2):                     # ragged last tiles; fewer matrices than one warp tile holds
            A = synthetic(n, batch, dtype)
            A[1] = 0                                # info = 1
            if batch > 5:
                A[5][:, 3] = 0                      # a zero column
            for lu_only in (False, True):
                X, ipiv, info = gpu_lapack(A, lu_only=lu_only)
                lub.set_option("staging", 1)
                try:
                    assert "getrf" not in lub.kernel_name(n, "lapack", dtype)
                    X1, ipiv1, info1 = gpu_lapack(A, lu_only=lu_only)
                finally:
                    lub.set_option("staging", 0)
                assert np.array_equal(info, info1), (n, batch, lu_only)
                good = info == 0
                assert good.sum() >= batch - 2
                assert np.array_equal(ipiv[good], ipiv1[good]), (n, batch, lu_only)
                scale = np.abs(X1[good]).max(axis=(1, 2), keepdims=True)
                if lu_only:
                    assert np.all(np.abs(X[good] - X1[good]) <= 16 * n * eps * scale), (n, batch)
                else:
                    A64 = A[good].astype(np.float64)
                    kappa = np.linalg.cond(A64)
                    res = np.linalg.norm(A64 @ X[good].astype(np.float64) - np.eye(n), axis=(1, 2))
                    ok = kappa < 0.001 / eps
                    assert np.all(res[ok] <= C_LAPACK * n * eps * kappa[ok]), (n, batch, float((res[ok] / (n * eps * kappa[ok])).max()))
    # full size: 200,000 distinct matrices through the reference's own predicate, and a view that starts at an odd matrix index
    n = 32
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.rand((200_000, n, n), generator=g, device="cuda", dtype=tdt)
    dA = A0.clone()
    piv = torch.zeros((200_000, n), dtype=torch.int32, device="cuda")
    info = torch.full((200_000,), -1, dtype=torch.int32, device="cuda")
    lub.lu_batched_inplace(dA, piv, "lapack", info=info)
    torch.cuda.synchronize()
    assert int(info.abs().sum()) == 0
    assert int(piv.min()) >= 1 and int(piv.max()) <= n and bool((piv >= torch.arange(1, n + 1, device="cuda", dtype=torch.int32)).all())
    good, bad, _ = lub.verify_inv(A0, dA)
    assert good + bad == 200_000 and bad <= (1000 if dtype == np.float32 else 0), bad      # reference rule on this input: ~8 % bad
    n = 27                                         # odd n, scalar rows: a view 27 * 27 elements into the buffer is 4 / 8 bytes off 16
    A = synthetic(n, 301, dtype)
    buf = torch.zeros((302, n, n), device="cuda", dtype=tdt)
    buf[1:].copy_(torch.from_numpy(A))
    view = buf[1:]
    pv = torch.zeros((301, n), dtype=torch.int32, device="cuda")
    lub.lu_batched_inplace(view, pv, "lapack")
    ref, pref, _ = gpu_lapack(A)
    torch.cuda.synchronize()
    assert np.array_equal(pv.cpu().numpy(), pref) and np.array_equal(view.cpu().numpy(), ref) and float(buf[0].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_interleaved_layout_is_bitwise_equal_to_matrix_major(dtype):
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    for n in range(1, 9):
        for mode in (0, 1, 2, 3):
            for batch in (1024, 1003, 1):          # vector path, scalar path (odd batch), a single matrix
                A = synthetic(n, batch, dtype, dominant=(mode == 0))
                dA = torch.from_numpy(A).cuda()
                piv = torch.full((batch, n), -1, dtype=torch.int32, device="cuda")
                info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
                lub.lu_batched_inplace(dA, piv, mode, info=info)
                dI = torch.from_numpy(A).cuda().permute(1, 2, 0).contiguous()      # [n, n, batch]
                pivI = torch.full((batch, n), -2, dtype=torch.int32, device="cuda")
                infoI = torch.full((batch,), -2, dtype=torch.int32, device="cuda")
                lub.lu_batched_inplace(dI, pivI, mode, info=infoI, layout="interleaved")
                torch.cuda.synchronize()
                assert torch.equal(pivI, piv), (n, mode, batch)
                assert torch.equal(infoI, info), (n, mode, batch)
                back = dI.permute(2, 0, 1).contiguous()
                assert back.dtype == tdt and torch.equal(back, dA), (n, mode, batch, float((back - dA).abs().max()))
    # tie-heavy small-integer matrices (exact ties and zeros in every column, singular ones included): the pivot vectors of the
    # interleaved kernel must still be the matrix-major kernels' -- which are the oracle's (test_tie_heavy_integer_matrices_pivots_exact)
    rng = np.random.default_rng(43)
    for n in range(2, 9):
        A = rng.integers(-3, 4, size=(512, n, n)).astype(dtype)
        for mode in (1, 2, 3):
            dA = torch.from_numpy(A).cuda()
            piv = torch.full((512, n), -1, dtype=torch.int32, device="cuda")
            lub.lu_batched_inplace(dA, piv, mode)
            dI = torch.from_numpy(A).cuda().permute(1, 2, 0).contiguous()
            pivI = torch.full((512, n), -2, dtype=torch.int32, device="cuda")
            lub.lu_batched_inplace(dI, pivI, mode, layout="interleaved")
            torch.cuda.synchronize()
            assert torch.equal(pivI, piv), (n, mode)
            if mode != 3:
                _, po = O.lu_batched(A, mode, lu_only=True)
                assert np.array_equal(piv.cpu().numpy(), po), (n, mode)
    # a view that is not aligned for vector access still works (scalar instantiation)
    A = synthetic(4, 1024, np.float32, dominant=True)
    flat = torch.zeros(4 * 4 * 1024 + 1, device="cuda")
    view = flat[1:].view(4, 4, 1024)
    view.copy_(torch.from_numpy(A).cuda().permute(1, 2, 0))
    lub.lu_batched_inplace(view, None, "none", layout="interleaved")
    ref = torch.from_numpy(A).cuda()
    lub.lu_batched_inplace(ref, None, "none")
    assert torch.equal(view.permute(2, 0, 1).contiguous(), ref) and flat[0].item() == 0.0
    with pytest.raises(lub.LubError):
        lub.lu_batched_inplace(torch.zeros((9, 9, 64), device="cuda"), None, "none", layout="interleaved")


def test_host_multi_device_entry_point_equals_the_single_device_path():
    """lu_batched_inplace_host_multi: contiguous shards over every visible GPU (one on a single-GPU box), pageable
    buffer registered for the call, workers bound next to their GPU -- bitwise the results of the device-pointer path."""
    ndev = torch.cuda.device_count()
    for n, dtype, mode in ((32, np.float32, "parallel"), (18, np.float32, "parallel"), (7, np.float64, "serial"), (32, np.float64, "lapack")):
        A = synthetic(n, 4099, dtype)                       # odd batch: ragged shards and tiles
        dA = torch.from_numpy(A).cuda()
        piv = torch.zeros((4099, n), dtype=torch.int32, device="cuda")
        lub.lu_batched_inplace(dA, piv, mode)
        torch.cuda.synchronize()
        want, want_piv = dA.cpu().numpy(), piv.cpu().numpy()
        for nd in sorted({1, ndev, 0}):
            for register in (False, True):
                H = A.copy()
                hp = np.full((4099, n), -3, np.int32)
                lub.lu_batched_inplace_host_multi(H, hp, mode, n_devices=nd, register=register)
                assert np.array_equal(H, want, equal_nan=True) and np.array_equal(hp, want_piv), (n, mode, nd, register)
    with pytest.raises(lub.LubError):
        lub.lu_batched_inplace_host_multi(np.zeros((4, 3, 3), np.float32), n_devices=ndev + 1)
    assert lub.bind_thread_near_device(0) in (True, False)


def test_ablation_options_change_the_kernel_not_the_result():
    """lu_batched_set_option: LSU staging instead of TMA / bulk copies is the same arithmetic (bitwise equal results); DFMA instead of DMMA for
    fp64 N = 32 is another operation order (pivots bit-exact, values to rounding)."""
    for n, dtype in ((32, np.float32), (24, np.float32), (16, np.float64)):
        A = synthetic(n, 517, dtype)
        dA = torch.from_numpy(A).cuda(); piv = torch.zeros((517, n), dtype=torch.int32, device="cuda")
        lub.lu_batched_inplace(dA, piv, "parallel")
        name0 = lub.kernel_name(n, "parallel", dtype)
        lub.set_option("staging", 1)
        try:
            name1 = lub.kernel_name(n, "parallel", dtype)
            dB = torch.from_numpy(A).cuda(); pivB = torch.zeros_like(piv)
            lub.lu_batched_inplace(dB, pivB, "parallel")
        finally:
            lub.set_option("staging", 0)
        assert name0 == ("lub_bulk_kernel" if n == 24 else "lub_tma_kernel") and name1 in ("lub_v3_kernel", "lub_v4_kernel")
        assert torch.equal(pivB, piv) and torch.equal(dB, dA), (n, dtype)
    # fp64 N = 32: DMMA is the library's choice without pivoting only; forcing it in a pivot mode keeps the pivots
    assert lub.kernel_name(32, "none", np.float64) == "lub_dmma_kernel"
    assert lub.kernel_name(32, "parallel", np.float64) == "lub_tma_kernel"
    A = synthetic(32, 517, np.float64)
    dA = torch.from_numpy(A).cuda(); piv = torch.zeros((517, 32), dtype=torch.int32, device="cuda")
    lub.lu_batched_inplace(dA, piv, "parallel")
    lub.set_option("fp64_tensor", 2)
    try:
        assert lub.kernel_name(32, "parallel", np.float64) == "lub_dmma_kernel"
        dB = torch.from_numpy(A).cuda(); pivB = torch.zeros_like(piv)
        lub.lu_batched_inplace(dB, pivB, "parallel")
    finally:
        lub.set_option("fp64_tensor", 0)
    assert torch.equal(pivB, piv)
    X, Y = dA.cpu().numpy(), dB.cpu().numpy()
    good = np.linalg.cond(A) < 1e4
    assert np.all(np.abs(X - Y)[good].max(axis=(1, 2)) <= 1e-6 * np.abs(X)[good].max(axis=(1, 2)))
    D = synthetic(32, 517, np.float64, dominant=True)
    dA = torch.from_numpy(D).cuda(); lub.lu_batched_inplace(dA, None, "none")
    lub.set_option("fp64_tensor", 1)
    try:
        assert lub.kernel_name(32, "none", np.float64) == "lub_v4_kernel"
        dB = torch.from_numpy(D).cuda(); lub.lu_batched_inplace(dB, None, "none")
    finally:
        lub.set_option("fp64_tensor", 0)
    assert float((dA - dB).abs().max()) <= 1e-15
    with pytest.raises(lub.LubError):
        lub.set_option(99, 1)
