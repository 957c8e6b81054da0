"""CPU tests of the boundary: the C-ABI library loads and exports every symbol that
include/*.h declares, argument checking, fail-loudly behaviour without a GPU, and the host
steps of main() (text input, replicate, verify.hpp-compatible check).  No compute call
that needs a GPU is expected to succeed here."""
import ctypes
import os
import re

import numpy as np
import pytest

import matrixinversion_b200 as lub
from conftest import ROOT, synthetic, template
from matrixinversion_b200 import _lib
from oracle import oracle as O

try:
    import torch
    HAVE_CUDA = torch.cuda.is_available()
except Exception:  # pragma: no cover
    HAVE_CUDA = False


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lu_batched_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared("lubatched.h")
    assert len(names) >= 16
    L = ctypes.CDLL(_lib.LIB_PATH)
    for nme in names:
        assert hasattr(L, nme), nme
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table mirrors the header
    C = ctypes.CDLL(_lib.CUBLAS_LIB_PATH)
    for nme in _declared("lubatched_cublas.h"):
        assert hasattr(C, nme), nme
    assert b"sm_100a" in _lib.lib().lu_batched_version()


def test_product_never_links_the_oracle():
    """The oracle is test infrastructure: nothing under matrixinversion_b200/ may mention it."""
    pkg = os.path.join(ROOT, "matrixinversion_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("test-only oracle", ""), os.path.join(dirpath, f)


def test_argument_errors_are_codes_not_exits():
    L = _lib.lib()
    assert L.lu_batched_inplace(None, None, 0, 1, 0, 0) == -1     # n out of range
    assert L.lu_batched_inplace(None, None, 33, 1, 0, 0) == -1
    assert L.lu_batched_inplace(None, None, 4, 1, 4, 0) == -2     # mode (3 = LAPACK partial pivoting is valid)
    assert L.lu_batched_inplace(None, None, 4, 1, -1, 0) == -2
    assert L.lu_batched_inplace_ex(None, None, None, 4, 1, 0, 0, 7, None) == -4   # layout
    assert L.lu_batched_inplace_ex(None, None, None, 9, 1, 0, 0, 1, None) == -1   # interleaved layout: n <= 8
    assert L.lu_batched_inplace(None, None, 4, 1, 0, 2) == -3     # dtype
    assert L.lu_batched_inplace(None, None, 4, -1, 0, 0) == -4    # batch
    assert b"batch" in L.lu_batched_last_error()
    assert L.lu_batched_set_threads(48) == -4 and L.lu_batched_set_threads(512) == -4
    assert L.lu_batched_set_threads(64) == 0 and L.lu_batched_get_threads(8, 0) == 64
    assert L.lu_batched_set_threads(0) == 0
    with pytest.raises(lub.LubError):
        lub.lu_batched_inplace(np.zeros((2, 3, 4), np.float32))
    with pytest.raises(lub.LubError):
        lub.lu_batched_inplace(np.zeros((2, 3, 3), np.int32))
    with pytest.raises(lub.LubError):
        lub.lu_batched_inplace(np.zeros((2, 3, 3), np.float32), pivot_mode="no such mode")


@pytest.mark.skipif(HAVE_CUDA, reason="box has a GPU")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    """There is no CPU implementation of the path in the product: every compute entry
    point must fail when no device is present."""
    A = synthetic(4, 3, np.float32, dominant=True)
    before = A.copy()
    with pytest.raises(lub.LubError) as e:
        lub.lu_batched_inplace(A, pivot_mode="none")
    assert e.value.code in (-5, -6)
    assert np.array_equal(A, before)
    with pytest.raises(lub.LubError):
        lub.geometry(8, 1000)


def test_read_template_is_a_stream_prefix(tmp_path, inputs):
    """first N*N tokens, not the top-left block; tokens parse like `file >> T`."""
    toks = inputs["mtrand32_new1_f64"]
    path = tmp_path / "mtrand32_new1.txt"
    with open(path, "w") as f:
        for r in range(32):
            f.write(" ".join(repr(float(v)) for v in toks[r * 32:(r + 1) * 32]) + " \n")
    for n in (1, 5, 18, 32):
        A = lub.read_template(path, n, np.float32)
        assert np.array_equal(A, template(inputs, "mtrand32_new1", n))
        if n not in (1, 32):
            assert not np.array_equal(A, toks.reshape(32, 32)[:n, :n].astype(np.float32))
        assert np.array_equal(lub.read_template(path, n, np.float64), template(inputs, "mtrand32_new1", n, np.float64))
    # numpy-style %.18e file without trailing newline (matrix.txt)
    path2 = tmp_path / "matrix.txt"
    with open(path2, "w") as f:
        f.write("\n".join(" ".join("%.18e" % v for v in inputs["matrix_f64"][r * 100:(r + 1) * 100]) for r in range(100)))
    assert np.array_equal(lub.read_template(path2, 20, np.float64), template(inputs, "matrix", 20, np.float64))
    with pytest.raises(lub.LubError) as e:
        lub.read_template(tmp_path / "missing.txt", 4)
    assert e.value.code == -7
    with pytest.raises(lub.LubError):
        lub.read_template(path, 33)  # 1089 tokens wanted, 1024 present


def test_replicate(inputs):
    T = template(inputs, "mtrand32", 7)
    A = lub.replicate(T, 11)
    assert A.shape == (11, 7, 7) and all(np.array_equal(A[i], T) for i in range(11))
    assert lub.replicate(T, 0).shape == (0, 7, 7)


def test_host_verify_inv_equals_oracle_and_golden(inputs, golden):
    for name in ("mtrand32", "mtrand32_new1"):
        for n in (1, 16, 22, 29, 32):
            A = template(inputs, name, n)
            for mode in (0, 1, 2):
                X = golden["orc_inv/%s/%d/%d" % (name, mode, n)]
                ok, bad, dev = lub.verify_inv(A[None], X[None])
                assert [ok, bad] == golden["ref_verify/%s/%d/%d" % (name, mode, n)].tolist()
                assert dev == pytest.approx(O.verify_inv(A[None], X[None])[2], rel=1e-5, abs=1e-7, nan_ok=True)
    A = synthetic(9, 40, np.float64, dominant=True)
    X, _ = O.lu_batched(A, 0)
    X[3] += 0.01
    X[17, 0, 0] = np.nan
    assert lub.verify_inv(A, X)[:2] == (38, 2) == O.verify_inv(A, X)[:2]


def test_host_verify_lu_equals_reference_verdicts(inputs):
    """lu_batched_verify_lu = verifyLU / verifyLUwithPivoting (templated/verify.hpp:105-186,
    parallel_pivot/verify.hpp:157-242): same verdicts as the reference's own function gave (golden file
    made by tests/golden/make_golden_lu.py through oracle/_ref) on the oracle's factors of the reference's
    inputs, for an intact factorisation and for one with a single entry off by 0.01; the oracle's factors
    themselves have not drifted; and, when oracle/_ref is present, the live reference agrees too."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_verify_lu.npz"))
    seen = 0
    for key in g.files:
        if not key.startswith("ref_verify_lu/"):
            continue
        _, name, suf, mode, n = key.split("/")
        mode, n = int(mode), int(n)
        dt = np.float32 if suf == "f32" else np.float64
        A = template(inputs, name, n, dt)
        with np.errstate(all="ignore"):
            LU, perm = O.lu_batched(A[None], mode, lu_only=True)
        if suf == "f32":
            assert np.array_equal(LU[0], g["orc_lu/%s/%d/%d" % (name, mode, n)]), key
        bad_lu = LU.copy()
        bad_lu[0, n - 1, n - 1] += 0.01
        ours = [*lub.verify_lu(A[None], LU, perm)[:2], *lub.verify_lu(A[None], bad_lu, perm)[:2]]
        assert ours == g[key].tolist(), (key, ours, g[key].tolist())
        if O.have_ref("ref_verify") and seen % 7 == 0:
            assert list(O.ref_verify_lu_piv(A[perm[0]], LU)) == ours[:2], key
        seen += 1
    assert seen > 200
    # no pivoting = verifyLU: piv may be NULL; several distinct matrices, one broken, one NaN
    A = synthetic(9, 40, np.float64, dominant=True)
    LU, _ = O.lu_batched(A, 0, lu_only=True)
    LU[3, 2, 5] += 0.01
    LU[17, 0, 0] = np.nan
    assert lub.verify_lu(A, LU)[:2] == (38, 2)
    assert np.isnan(lub.verify_lu(A, LU)[2])


def test_write_to_file_matches_the_reference_text(tmp_path, inputs, capfd):
    """writeToFile / printMatrices (templated/verify.hpp:12-48): first matrix only, default ostream
    formatting; byte-identical to the reference's own function when oracle/_ref is present."""
    for dt in (np.float32, np.float64):
        A = np.stack([template(inputs, "mtrand32_new1", 5, dt), template(inputs, "mtrand32", 5, dt)])
        A[0, 0, 0] = 1e-7
        A[0, 1, 1] = 123456789.0
        p = tmp_path / "ours.txt"
        lub.write_to_file(A, str(p))
        text = p.read_text()
        rows = text.split("\n")
        assert len(rows) == 6 and rows[-1] == "" and all(r.endswith(" ") and len(r.split()) == 5 for r in rows[:5])
        assert rows[0].split()[0] == "1e-07" and rows[1].split()[1] == "1.23457e+08"
        assert np.allclose(np.array(text.split(), dtype=np.float64).reshape(5, 5), A[0], rtol=1e-5)
        if O.have_ref("ref_verify"):
            import ctypes as C
            ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_verify.so"))
            f = getattr(ref, "ref_write_to_file_" + ("f32" if dt == np.float32 else "f64"))
            f.restype = None
            f.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
            q = tmp_path / "ref.txt"
            Ac = np.ascontiguousarray(A)
            f(Ac.ctypes.data, str(q).encode(), 5, 2)
            assert q.read_bytes() == p.read_bytes()
    lub.print_matrices(A)
    out = capfd.readouterr().out
    assert out == text + "\n"


def test_default_num_threads_table():
    """templated/run.py:201-223."""
    table = {1: 32, 2: 32, 3: 30, 4: 32, 5: 30, 6: 30, 7: 28, 8: 32, 9: 27, 10: 30, 11: 22, 12: 24, 13: 26,
             14: 28, 15: 30, 16: 32, 17: 17, 20: 20, 31: 31, 32: 32}
    for n, t in table.items():
        assert lub.default_num_threads(n) == t


def test_shard_range_partitions_the_batch():
    for batch in (0, 1, 7, 8, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [lub.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-batch // world)


def test_ipiv_to_perm_host_helper():
    """LAPACK swap lists -> permutation vectors (lu_batched_ipiv_to_perm), against a literal replay."""
    rng = np.random.default_rng(3)
    n, b = 11, 50
    ipiv = np.stack([np.array([rng.integers(k, n) + 1 for k in range(n)], np.int32) for _ in range(b)])
    perm = lub.ipiv_to_perm(ipiv)
    for i in range(b):
        p = list(range(n))
        for k in range(n):
            q = ipiv[i, k] - 1
            p[k], p[q] = p[q], p[k]
        assert perm[i].tolist() == p
    bad = ipiv.copy(); bad[0, 3] = 2            # points above the diagonal: not a getrf swap list
    with pytest.raises(lub.LubError):
        lub.ipiv_to_perm(bad)


def test_every_staged_configuration_fits_the_shared_memory_of_an_sm(tmp_path):
    """Host-only (no GPU): scripts/check_smem_budget.cu instantiates the launch tables of the bulk-copy staged kernels -- inverse,
    factors only, pivot_mode 3; n = 2..32, both dtypes -- and checks that none asks for more than the 227 KB of dynamic shared
    memory an SM has (a configuration that did once: fp64 n = 6 with two pivot vectors per matrix in a 384-thread block)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "check_smem")
    subprocess.check_call([nvcc, "-std=c++17", "-I" + os.path.join(root, "matrixinversion_b200", "csrc"), "-I" + os.path.join(root, "include"),
                           "-gencode", "arch=compute_100a,code=sm_100a", os.path.join(root, "scripts", "check_smem_budget.cu"), "-o", exe],
                          stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert "TOO BIG" not in out and out.strip().endswith("done"), out
