"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (lu_oracle.c) and of
the reference-derived checkers under oracle/_ref/ (built by build_ref.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (matrixinversion_b200/) never does.

Also holds a small pure-numpy twin (`numpy_invert_one`) of the same algorithm, used to
cross-check the C restatement on small cases.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFDIR = os.path.join(_HERE, "_ref")

MODE_NONE, MODE_SERIAL, MODE_PARALLEL = 0, 1, 2
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile lu_oracle.c (gcc) if the shared object is missing or stale."""
    so = os.path.join(_HERE, "liblu_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("lu_oracle.c", "lu_oracle_impl.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(
            ["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
             srcs[0], "-o", so, "-lm"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            p = ctypes.POINTER(ct)
            f = getattr(L, "oracle_lu_batched_" + suf)
            f.restype = ctypes.c_int
            f.argtypes = [p, _c_i32p, _c_i32p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                          ctypes.c_int, ctypes.c_int, ctypes.c_int]
            f = getattr(L, "oracle_verify_inv_" + suf)
            f.restype = None
            f.argtypes = [p, p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, _c_i64p, _c_i64p,
                          ctypes.POINTER(ctypes.c_double)]
            f = getattr(L, "oracle_pivotedA_" + suf)
            f.restype = None
            f.argtypes = [p, p, _c_i32p, ctypes.c_int]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _suf(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(dtype)


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _ct(dtype):
    return ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def lu_batched(A: np.ndarray, mode: int, *, use_fma: bool = True, lu_only: bool = False, tpm: int = 0,
               threads: int = 0, want_steps: bool = False):
    """Oracle inverse of A[batch, n, n] (a copy is inverted; A is untouched).

    Returns (X, perm[, steps]) -- X the inverses (or packed LU), perm int32[batch, n]
    with reference semantics (perm[i] = original row sitting in row i), steps the pivot
    row picked at each elimination step.
    """
    A = np.ascontiguousarray(A)
    assert A.ndim == 3 and A.shape[1] == A.shape[2]
    b, n, _ = A.shape
    X = A.copy()
    perm = np.empty((b, n), dtype=np.int32)
    steps = np.empty((b, n), dtype=np.int32)
    ct = _ct(A.dtype)
    rc = getattr(lib(), "oracle_lu_batched_" + _suf(A.dtype))(
        _ptr(X, ct), _ptr(perm, ctypes.c_int32), _ptr(steps, ctypes.c_int32), n, b, mode, tpm,
        int(use_fma), int(lu_only), threads)
    if rc < 0:
        raise ValueError("oracle rejected the arguments")
    return (X, perm, steps) if want_steps else (X, perm)


def lu_batched_inplace_timed(X: np.ndarray, mode: int, threads: int = 0) -> int:
    """In-place variant for the CPU-baseline timing leg; returns threads used."""
    b, n, _ = X.shape
    ct = _ct(X.dtype)
    return int(getattr(lib(), "oracle_lu_batched_" + _suf(X.dtype))(
        _ptr(X, ct), None, None, n, b, mode, 0, 2, 0, threads))  # use_fma=2: plain `sum + a*b`


def verify_inv(A: np.ndarray, X: np.ndarray, thr: float = 1e-3):
    """verifyInv restated: (n_correct, n_incorrect, max |r - delta|)."""
    A = np.ascontiguousarray(A)
    X = np.ascontiguousarray(X, dtype=A.dtype)
    b, n, _ = A.shape
    ok, bad, dev = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
    ct = _ct(A.dtype)
    getattr(lib(), "oracle_verify_inv_" + _suf(A.dtype))(
        _ptr(A, ct), _ptr(X, ct), n, b, thr, ctypes.byref(ok), ctypes.byref(bad), ctypes.byref(dev))
    return ok.value, bad.value, dev.value


def pivotedA(A: np.ndarray):
    """pivotedA restated for ONE matrix: (P*A, pivots)."""
    A = np.ascontiguousarray(A)
    n = A.shape[0]
    PA = np.empty_like(A)
    piv = np.empty(n, dtype=np.int32)
    ct = _ct(A.dtype)
    getattr(lib(), "oracle_pivotedA_" + _suf(A.dtype))(_ptr(A, ct), _ptr(PA, ct), _ptr(piv, ctypes.c_int32), n)
    return PA, piv


# ----------------------------------------------------------------------------------------
# reference-derived checkers (oracle/_ref, built from /root/reference by build_ref.sh)
# ----------------------------------------------------------------------------------------

def have_ref(name: str = "ref_verify") -> bool:
    return os.path.exists(os.path.join(_REFDIR, "lib%s.so" % name))


_ref_cache = {}


def _ref(name):
    if name not in _ref_cache:
        _ref_cache[name] = ctypes.CDLL(os.path.join(_REFDIR, "lib%s.so" % name))
    return _ref_cache[name]


def ref_pivotedA(A: np.ndarray):
    """The reference's own pivotedA (parallel_pivot/verify.hpp:106-155), one matrix."""
    A = np.ascontiguousarray(A)
    n = A.shape[0]
    PA = np.empty_like(A)
    piv = np.empty(n, dtype=np.int32)
    ct = _ct(A.dtype)
    f = getattr(_ref("ref_verify"), "ref_pivotedA_" + _suf(A.dtype))
    f.restype = None
    f.argtypes = [ctypes.POINTER(ct), ctypes.POINTER(ct), _c_i32p, ctypes.c_int]
    f(_ptr(A, ct), _ptr(PA, ct), _ptr(piv, ctypes.c_int32), n)
    return PA, piv


def ref_verify_inv(A: np.ndarray, X: np.ndarray):
    """The reference's own verifyInv (parallel_pivot/verify.hpp:50-103): (correct, incorrect)."""
    A = np.ascontiguousarray(A)
    X = np.ascontiguousarray(X, dtype=A.dtype)
    b, n, _ = A.shape
    ct = _ct(A.dtype)
    f = getattr(_ref("ref_verify"), "ref_verify_inv_" + _suf(A.dtype))
    f.restype = None
    f.argtypes = [ctypes.POINTER(ct), ctypes.POINTER(ct), ctypes.c_int, ctypes.c_int,
                  ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    ok, bad = ctypes.c_longlong(), ctypes.c_longlong()
    f(_ptr(A, ct), _ptr(X, ct), n, b, ctypes.byref(ok), ctypes.byref(bad))
    return ok.value, bad.value


def ref_verify_lu_piv(PA: np.ndarray, LU: np.ndarray):
    """The reference's own verifyLUwithPivoting (parallel_pivot/verify.hpp:157-242): PA = ONE permuted
    n x n matrix (what main() would pass after pivotedA), LU = [batch, n, n] factors -> (correct, incorrect)."""
    PA = np.ascontiguousarray(PA)
    LU = np.ascontiguousarray(LU, dtype=PA.dtype)
    b, n, _ = LU.shape
    ct = _ct(PA.dtype)
    f = getattr(_ref("ref_verify"), "ref_verify_lu_piv_" + _suf(PA.dtype))
    f.restype = None
    f.argtypes = [ctypes.POINTER(ct), ctypes.POINTER(ct), ctypes.c_int, ctypes.c_int,
                  ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    ok, bad = ctypes.c_longlong(), ctypes.c_longlong()
    f(_ptr(PA, ct), _ptr(LU, ct), n, b, ctypes.byref(ok), ctypes.byref(bad))
    return ok.value, bad.value


def ref_calc_cond_num(A: np.ndarray) -> float:
    A = np.ascontiguousarray(A)
    ct = _ct(A.dtype)
    f = getattr(_ref("ref_verify"), "ref_calc_cond_num_" + _suf(A.dtype))
    f.restype = ctypes.c_double
    f.argtypes = [ctypes.POINTER(ct), ctypes.c_int]
    return float(f(_ptr(A, ct), A.shape[0]))


_REF_KERNEL = {MODE_NONE: "none", MODE_SERIAL: "serial", MODE_PARALLEL: "parallel"}


def ref_gpu_invert(A: np.ndarray, mode: int, want_piv: bool = False):
    """Run the reference's own CUDA kernel (rebuilt for sm_100) on A[batch, n, n].

    Needs a GPU.  want_piv uses build_ref.sh's pivot-exporting patch of the same kernel
    (SURVEY.md H3).  Returns (X, piv or None, kernel_ms).
    """
    A = np.ascontiguousarray(A)
    b, n, _ = A.shape
    suf = _suf(A.dtype)
    name = "ref_%s%s_%s" % (_REF_KERNEL[mode], "_piv" if want_piv else "", suf)
    ct = _ct(A.dtype)
    f = getattr(_ref(name), name)
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.POINTER(ct), ctypes.POINTER(ct), _c_i32p, ctypes.c_int, ctypes.c_longlong,
                  ctypes.POINTER(ctypes.c_float)]
    X = np.empty_like(A)
    piv = np.full((b, n), -1, dtype=np.int32) if want_piv else None
    ms = ctypes.c_float()
    rc = f(_ptr(A, ct), _ptr(X, ct), _ptr(piv, ctypes.c_int32) if want_piv else None, n, b, ctypes.byref(ms))
    if rc != 0:
        raise RuntimeError("reference kernel %s failed rc=%d" % (name, rc))
    return X, piv, ms.value


def ref_gpu_time(dptr: int, n: int, batch: int, mode: int, dtype, reps: int = 5):
    """Time the reference's own CUDA kernel (rebuilt for sm_100) on a DEVICE buffer of `batch` n x n matrices
    (raw pointer, e.g. tensor.data_ptr()); the buffer is inverted in place `reps` times.  Returns
    (ms_cold, ms_warm_best, matrices_processed) -- cold = first launch, the reference's own convention
    (templated/luBatchedInplace.cu:71-82).  Checker-side only: never on the product path."""
    suf = _suf(np.dtype(dtype))
    name = "ref_%s_%s" % (_REF_KERNEL[mode], suf)
    f = getattr(_ref(name), name + "_device")
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int,
                  ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_longlong)]
    c, w, p = ctypes.c_float(), ctypes.c_float(), ctypes.c_longlong()
    rc = f(ctypes.c_void_p(dptr), None, n, batch, reps, ctypes.byref(c), ctypes.byref(w), ctypes.byref(p))
    if rc != 0:
        raise RuntimeError("reference kernel %s_device failed rc=%d" % (name, rc))
    return c.value, w.value, p.value


def lapack_getrf(A: np.ndarray):
    """LAPACK's own getrf (through scipy) on every matrix of A[b, n, n], in A's precision: the external truth for
    pivot_mode 3.  Returns (LU, ipiv 1-based int32 [b, n], info int32 [b])."""
    from scipy.linalg import get_lapack_funcs
    A = np.ascontiguousarray(A)
    b, n, _ = A.shape
    (getrf,) = get_lapack_funcs(("getrf",), (A[0],))
    LU = np.empty_like(A)
    ipiv = np.empty((b, n), np.int32)
    info = np.empty(b, np.int32)
    for i in range(b):
        lu, piv, inf = getrf(A[i])
        LU[i], ipiv[i], info[i] = lu, piv + 1, inf
    return LU, ipiv, info


def sweep_compare_hook(dptr, n, batch, mode, dtype_name):
    """`--compare oracle.oracle:sweep_compare_hook` of the sweep drivers: the reference kernel's times on the sweep's buffer."""
    if mode == MODE_PARALLEL and dtype_name == "float64" and n % 2:
        return {"reference_gpu_error": "reference bug: misaligned shared-memory carve for odd N in fp64 (parallel_pivot/luBatchedInplace.cuh:142)"}
    cold, warm, done = ref_gpu_time(dptr, n, batch, mode, np.dtype(dtype_name), reps=3)
    return {"reference_gpu_ms_cold": cold, "reference_gpu_ms_warm": warm, "reference_gpu_matrices": done}


# ----------------------------------------------------------------------------------------
# numpy twin (small cases only): same step order, same pivot rules, non-FMA arithmetic
# ----------------------------------------------------------------------------------------

def numpy_invert_one(A: np.ndarray, mode: int):
    """Pure-Python/numpy restatement for ONE small matrix; returns (X, perm, steps).

    Arithmetic is done in A.dtype with separately rounded products (== oracle use_fma=False).
    """
    T = A.dtype.type
    A = A.copy()
    n = A.shape[0]
    perm = list(range(n))
    steps = []
    for k in range(n):
        p = k
        if mode == MODE_SERIAL:
            m = abs(A[k, k])
            for i in range(k + 1, n):
                if abs(A[i, k]) > m:
                    m, p = abs(A[i, k]), i
        elif mode == MODE_PARALLEL:
            tpm = n
            vals, idx = [], []
            for t in range(tpm):
                m, q = abs(A[k, k]), k
                for i in range(k + 1 + t, n, tpm):
                    if abs(A[i, k]) > m:
                        m, q = abs(A[i, k]), i
                vals.append(m)
                idx.append(q)
            stride = tpm // 2
            while stride > 0:
                for t in range(stride):
                    if vals[t] < vals[t + stride]:
                        vals[t], idx[t] = vals[t + stride], idx[t + stride]
                stride >>= 1
            p = idx[0]
        steps.append(p)
        if p != k:
            perm[k], perm[p] = perm[p], perm[k]
            A[[k, p], :] = A[[p, k], :]
        for j in range(k, n):
            s = T(0)
            for l in range(k):
                s = T(s + T(A[k, l] * A[l, j]))
            A[k, j] = T(A[k, j] - s)
        for i in range(k + 1, n):
            s = T(0)
            for l in range(k):
                s = T(s + T(A[i, l] * A[l, k]))
            A[i, k] = T(T(A[i, k] - s) / A[k, k])
    X = np.zeros_like(A)
    for c in range(n):
        y = np.zeros(n, dtype=A.dtype)
        x = np.zeros(n, dtype=A.dtype)
        for i in range(n):
            s = T(0)
            for j in range(i):
                s = T(s + T(A[i, j] * y[j]))
            y[i] = T(T(1 if perm[i] == c else 0) - s)
        for i in range(n - 1, -1, -1):
            s = T(0)
            for j in range(i + 1, n):
                s = T(s + T(A[i, j] * x[j]))
            x[i] = T(T(y[i] - s) / A[i, i])
        X[:, c] = x
    return X, np.array(perm, dtype=np.int32), np.array(steps, dtype=np.int32)
