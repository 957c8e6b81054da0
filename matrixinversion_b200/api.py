"""Host-side mirror of the reference's operator boundary for the luBatchedInplace path.

The reference has no importable API: each variant directory holds a `main()`
(`/root/reference/<variant>/luBatchedInplace.cu`) that reads a template matrix, replicates
it, copies it to the device, launches `batched_lu_subwarp` once, times it with CUDA events,
copies back and runs `verifyInv`.  The functions below are those steps, one per function,
with the same names / argument meaning where the reference has a name, all thin wrappers
over the C ABI in include/lubatched.h (no arithmetic happens in Python and there is no
CPU fallback: the CUDA library must load and a GPU must be present for every compute call).

    variant dir        pivot_mode
    templated/         "none"      (0)
    serial_pivot/      "serial"    (1)
    parallel_pivot/    "parallel"  (2)
    (extension)        "lapack"    (3)   true partial pivoting, LAPACK ipiv + info
"""
from __future__ import annotations

import ctypes
import os
import sys
import time
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import (DTYPE_F32, DTYPE_F64, LAYOUT_BATCH_INTERLEAVED, LAYOUT_MATRIX_MAJOR, PIVOT_LAPACK, PIVOT_NONE, PIVOT_PARALLEL,
                   PIVOT_SERIAL, LubError, check)

PIVOT_MODES = {"none": PIVOT_NONE, "serial": PIVOT_SERIAL, "parallel": PIVOT_PARALLEL,
               "templated": PIVOT_NONE, "serial_pivot": PIVOT_SERIAL, "parallel_pivot": PIVOT_PARALLEL,
               "lapack": PIVOT_LAPACK, "partial": PIVOT_LAPACK}
LAYOUTS = {"matrix": LAYOUT_MATRIX_MAJOR, "matrix_major": LAYOUT_MATRIX_MAJOR,
           "interleaved": LAYOUT_BATCH_INTERLEAVED, "batch_interleaved": LAYOUT_BATCH_INTERLEAVED}


def _mode(pivot_mode) -> int:
    if isinstance(pivot_mode, str):
        try:
            return PIVOT_MODES[pivot_mode]
        except KeyError:
            raise LubError(-2, "unknown pivot_mode %r" % (pivot_mode,)) from None
    return int(pivot_mode)


def _dtype_code(dtype) -> int:
    try:
        import torch
        if isinstance(dtype, torch.dtype):
            return {torch.float32: DTYPE_F32, torch.float64: DTYPE_F64}[dtype]
    except KeyError:
        raise LubError(-3, "dtype must be float32 or float64") from None
    except ImportError:
        pass
    dt = np.dtype(dtype)
    if dt == np.float32:
        return DTYPE_F32
    if dt == np.float64:
        return DTYPE_F64
    raise LubError(-3, "dtype must be float32 or float64")


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _check_batch(shape):
    if len(shape) != 3 or shape[1] != shape[2]:
        raise LubError(-4, "expected a [batch, n, n] array, got shape %s" % (tuple(shape),))
    return int(shape[0]), int(shape[1])


# ------------------------------------------------------------------------------------------
# the hot path
# ------------------------------------------------------------------------------------------

def lu_batched_inplace(A, piv=None, pivot_mode="parallel", stream=None, info=None, layout="matrix"):
    """Invert every matrix of A[batch, n, n] in place (the `batched_lu_subwarp` launch,
    parallel_pivot/luBatchedInplace.cu:127).

    info: optional CUDA int32 [batch] tensor (lu_batched_inplace_ex): with pivot_mode "lapack" 0 or the 1-based
       index of the first exactly-zero pivot; the reference's modes have no status and write 0.
    layout: "matrix" (the reference's A[batch][n][n]) or "interleaved" -- then A is a CUDA tensor of shape
       [n, n, batch] (element (i, j) of consecutive matrices contiguous), n <= 8.

    A: contiguous CUDA torch tensor (float32/float64) -> asynchronous launch on `stream`
       (default: torch's current stream); or a C-contiguous numpy array -> the chunked
       host pipeline (`lu_batched_inplace_host`), synchronous.
    piv: optional int32 [batch, n] buffer of the same kind; receives the reference's
       permutation vector (`pivots[]`, parallel_pivot/luBatchedInplace.cuh:140,161-168).
    Returns A.
    """
    L = _lib.lib()
    mode = _mode(pivot_mode)
    try:
        lay = LAYOUTS[layout] if isinstance(layout, str) else int(layout)
    except KeyError:
        raise LubError(-4, "unknown layout %r" % (layout,)) from None
    if lay == LAYOUT_BATCH_INTERLEAVED:
        if len(A.shape) != 3 or A.shape[0] != A.shape[1]:
            raise LubError(-4, "the interleaved layout expects a [n, n, batch] tensor, got shape %s" % (tuple(A.shape),))
        n, batch = int(A.shape[0]), int(A.shape[2])
    else:
        batch, n = _check_batch(A.shape)
    if _is_torch(A):
        import torch
        if not A.is_cuda:
            raise LubError(-4, "torch tensors must live on a CUDA device (pass numpy for host data)")
        if not A.is_contiguous():
            raise LubError(-4, "A must be contiguous")
        dt = _dtype_code(A.dtype)
        iptr = None
        if info is not None:
            if not (info.is_cuda and info.is_contiguous() and info.dtype == torch.int32 and tuple(info.shape) == (batch,)
                    and info.device == A.device):
                raise LubError(-4, "info must be a contiguous CUDA int32 [batch] tensor on A's device")
            iptr = info.data_ptr()
        pptr = None
        if piv is not None:
            if not (piv.is_cuda and piv.is_contiguous() and piv.dtype == torch.int32 and tuple(piv.shape) == (batch, n)):
                raise LubError(-4, "piv must be a contiguous CUDA int32 [batch, n] tensor")
            if piv.device != A.device:
                raise LubError(-4, "piv must live on A's device")
            pptr = piv.data_ptr()
        with torch.cuda.device(A.device):
            s = stream if stream is not None else torch.cuda.current_stream(A.device)
            sp = s.cuda_stream if hasattr(s, "cuda_stream") else int(s)
            if iptr is None and lay == LAYOUT_MATRIX_MAJOR:
                check(L.lu_batched_inplace_stream(A.data_ptr(), pptr, n, batch, mode, dt, sp))
            else:
                check(L.lu_batched_inplace_ex(A.data_ptr(), pptr, iptr, n, batch, mode, dt, lay, sp))
        return A
    if info is not None or lay != LAYOUT_MATRIX_MAJOR:
        raise LubError(-4, "info / layout are device-side options: pass CUDA tensors")
    if not isinstance(A, np.ndarray) or not A.flags.c_contiguous or not A.flags.writeable:
        raise LubError(-4, "A must be a CUDA torch tensor or a writable C-contiguous numpy array")
    dt = _dtype_code(A.dtype)
    pptr = None
    if piv is not None:
        if not (isinstance(piv, np.ndarray) and piv.dtype == np.int32 and piv.shape == (batch, n) and piv.flags.c_contiguous):
            raise LubError(-4, "piv must be a C-contiguous int32 [batch, n] numpy array")
        pptr = piv.ctypes.data
    check(L.lu_batched_inplace_host(A.ctypes.data, pptr, n, batch, mode, dt))
    return A


def lu_batched_inplace_host_multi(A: np.ndarray, piv=None, pivot_mode="parallel", n_devices: int = 0, register: bool = False,
                                  bind_threads: bool = True) -> np.ndarray:
    """lu_batched_inplace_host_multi: the host-buffer call over several GPUs of this process (contiguous shards, one
    worker thread + pipeline per device).  A: writable C-contiguous numpy [batch, n, n]; n_devices 0 =
    all visible.  register: page-lock a pageable buffer for the call; bind_threads: workers run next to their GPU."""
    L = _lib.lib()
    batch, n = _check_batch(A.shape)
    if not isinstance(A, np.ndarray) or not A.flags.c_contiguous or not A.flags.writeable:
        raise LubError(-4, "A must be a writable C-contiguous numpy array")
    pptr = None
    if piv is not None:
        if not (isinstance(piv, np.ndarray) and piv.dtype == np.int32 and piv.shape == (batch, n) and piv.flags.c_contiguous):
            raise LubError(-4, "piv must be a C-contiguous int32 [batch, n] numpy array")
        pptr = piv.ctypes.data
    flags = (_lib.HOST_REGISTER if register else 0) | (_lib.HOST_BIND_THREADS if bind_threads else 0)
    check(L.lu_batched_inplace_host_multi(A.ctypes.data, pptr, n, batch, _mode(pivot_mode), _dtype_code(A.dtype), int(n_devices), flags))
    return A


def set_option(option, value: int) -> None:
    """lu_batched_set_option: run-time ablation knobs ("staging": 1 = LSU staging instead of TMA; "fp64_tensor": 1 = DFMA
    rank-1 updates instead of DMMA for fp64 N = 32); 0 restores the library's choice."""
    code = {"staging": _lib.OPT_STAGING, "fp64_tensor": _lib.OPT_FP64_TENSOR}.get(option, option)
    check(_lib.lib().lu_batched_set_option(int(code), int(value)))


def bind_thread_near_device(device: int) -> bool:
    """Bind the calling thread to the CPUs next to a GPU (lu_batched_bind_thread_near_device), e.g. before allocating the
    host buffers that will feed it (first touch places them on that NUMA node).  False when the topology is not exposed."""
    return _lib.lib().lu_batched_bind_thread_near_device(int(device)) == 0


def ipiv_to_perm(ipiv: np.ndarray) -> np.ndarray:
    """LAPACK swap lists (pivot_mode "lapack": 1-based ipiv[batch, n]) -> the permutation vectors the other modes
    write to piv (row i of P A is row perm[i] of A), e.g. for verify_lu."""
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
    perm = np.empty_like(ipiv)
    check(_lib.lib().lu_batched_ipiv_to_perm(ipiv.ctypes.data, perm.ctypes.data, ipiv.shape[1], ipiv.shape[0]))
    return perm


def lu_batched_factor_inplace(A, piv=None, pivot_mode="parallel", stream=None, info=None):
    """LU factors only, in place: the state of the reference's shared-memory matrix
    after its k-loop (parallel_pivot/luBatchedInplace.cuh:156-186), which upstream's disabled
    `verifyLU` / `verifyLUwithPivoting` are written for.  A: contiguous CUDA torch tensor
    [batch, n, n]; piv: optional CUDA int32 [batch, n].  Unit-lower L below the diagonal, U on and
    above it, rows in pivoted order.  Returns A."""
    import torch
    L = _lib.lib()
    mode = _mode(pivot_mode)
    batch, n = _check_batch(A.shape)
    if not (_is_torch(A) and A.is_cuda and A.is_contiguous()):
        raise LubError(-4, "A must be a contiguous CUDA torch tensor")
    pptr = None
    if piv is not None:
        if not (piv.is_cuda and piv.is_contiguous() and piv.dtype == torch.int32 and tuple(piv.shape) == (batch, n)):
            raise LubError(-4, "piv must be a contiguous CUDA int32 [batch, n] tensor")
        if piv.device != A.device:
            raise LubError(-4, "piv must live on A's device")
        pptr = piv.data_ptr()
    with torch.cuda.device(A.device):
        s = stream if stream is not None else torch.cuda.current_stream(A.device)
        sp = s.cuda_stream if hasattr(s, "cuda_stream") else int(s)
        if info is None:
            check(L.lu_batched_factor_inplace_stream(A.data_ptr(), pptr, n, batch, mode, _dtype_code(A.dtype), sp))
        else:
            if not (info.is_cuda and info.is_contiguous() and info.dtype == torch.int32 and tuple(info.shape) == (batch,)
                    and info.device == A.device):
                raise LubError(-4, "info must be a contiguous CUDA int32 [batch] tensor on A's device")
            check(L.lu_batched_factor_inplace_ex(A.data_ptr(), pptr, info.data_ptr(), n, batch, mode, _dtype_code(A.dtype), sp))
    return A


def verify_lu(A, LU, piv=None, thr: float = 1e-3):
    """verifyLU (templated/verify.hpp:105-186) / verifyLUwithPivoting (parallel_pivot/verify.hpp:157-242)
    on host arrays: returns (correct, incorrect, max |PA - L U|).  piv: int32 [batch, n] permutation
    vectors (None = no pivoting)."""
    L = _lib.lib()
    A = np.ascontiguousarray(A)
    LU = np.ascontiguousarray(LU, dtype=A.dtype)
    batch, n = _check_batch(A.shape)
    pptr = None
    if piv is not None:
        piv = np.ascontiguousarray(piv, dtype=np.int32)
        if piv.shape != (batch, n):
            raise LubError(-4, "piv must be int32 [batch, n]")
        pptr = piv.ctypes.data
    ok, bad, dev = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
    check(L.lu_batched_verify_lu(A.ctypes.data, LU.ctypes.data, pptr, n, batch, _dtype_code(A.dtype), thr,
                                 ctypes.byref(ok), ctypes.byref(bad), ctypes.byref(dev)))
    return ok.value, bad.value, dev.value


def lu_batched_inplace_ptr(ptr: int, piv_ptr, n: int, batch: int, pivot_mode, dtype, stream_ptr: int = 0) -> None:
    """Raw-pointer form: exactly the C ABI call (device pointers as integers)."""
    check(_lib.lib().lu_batched_inplace_stream(ptr, piv_ptr, n, batch, _mode(pivot_mode), _dtype_code(dtype), stream_ptr))


def set_num_threads(numthreads: int) -> None:
    """NUMTHREADS knob (templated/luBatchedInplace.cu:6); 0 = library default."""
    check(_lib.lib().lu_batched_set_threads(int(numthreads)))


@dataclass
class Geometry:
    threads_per_block: int
    threads_per_matrix: int
    matrices_per_block: int
    num_blocks: int
    dyn_smem_bytes: int


def geometry(n: int, batch: int, pivot_mode="parallel", dtype=np.float32) -> Geometry:
    """Launch geometry -- the numbers main() prints (parallel_pivot/luBatchedInplace.cu:13-18)."""
    tpb, tpm, mpb, smem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    nb = ctypes.c_int64()
    check(_lib.lib().lu_batched_geometry(n, batch, _mode(pivot_mode), _dtype_code(dtype), ctypes.byref(tpb),
                                         ctypes.byref(tpm), ctypes.byref(mpb), ctypes.byref(nb), ctypes.byref(smem)))
    return Geometry(tpb.value, tpm.value, mpb.value, nb.value, smem.value)


def kernel_name(n: int, pivot_mode="parallel", dtype=np.float32) -> str:
    """Kernel family launched for this configuration (profilers, bench reports)."""
    s = _lib.lib().lu_batched_kernel_name(n, _mode(pivot_mode), _dtype_code(dtype))
    if s is None:
        raise _lib.LubError(-1, "bad configuration")
    return s.decode()


def enable_timing(on: bool = True) -> None:
    check(_lib.lib().lu_batched_enable_timing(int(on)))


def last_kernel_ms() -> float:
    """The reference's "Kernel execution time" (templated/luBatchedInplace.cu:71-82)."""
    return float(_lib.lib().lu_batched_last_kernel_ms())


def device_info() -> dict:
    """deviceProps.cu:4-23."""
    v = [ctypes.c_int() for _ in range(5)]
    check(_lib.lib().lu_batched_device_info(*[ctypes.byref(x) for x in v]))
    return dict(zip(("sm_count", "max_smem_optin", "clock_khz", "cc_major", "cc_minor"), (x.value for x in v)))


# ------------------------------------------------------------------------------------------
# verify.hpp-compatible check
# ------------------------------------------------------------------------------------------

def verify_inv(A, A_inv, thr: float = 1e-3):
    """verifyInv (templated/verify.hpp:50-103): returns (correct, incorrect, max |r - delta|).

    numpy arrays are checked by the library's host code, CUDA tensors on the device.
    """
    L = _lib.lib()
    batch, n = _check_batch(A.shape)
    ok, bad, dev = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
    if _is_torch(A):
        if not (A.is_cuda and A_inv.is_cuda and A.is_contiguous() and A_inv.is_contiguous() and A.dtype == A_inv.dtype):
            raise LubError(-4, "A and A_inv must be contiguous CUDA tensors of one dtype")
        import torch
        if A_inv.device != A.device:
            raise LubError(-4, "A and A_inv must live on the same device")
        with torch.cuda.device(A.device):
            # explicit-stream entry point: the caller's lu_batched_set_stream setting stays untouched
            check(L.lu_batched_verify_inv_device_stream(A.data_ptr(), A_inv.data_ptr(), n, batch, _dtype_code(A.dtype), thr,
                                                        ctypes.byref(ok), ctypes.byref(bad), ctypes.byref(dev),
                                                        torch.cuda.current_stream(A.device).cuda_stream))
    else:
        A = np.ascontiguousarray(A)
        A_inv = np.ascontiguousarray(A_inv, dtype=A.dtype)
        check(L.lu_batched_verify_inv(A.ctypes.data, A_inv.ctypes.data, n, batch, _dtype_code(A.dtype), thr,
                                      ctypes.byref(ok), ctypes.byref(bad), ctypes.byref(dev)))
    return ok.value, bad.value, dev.value


# ------------------------------------------------------------------------------------------
# inputs: mtrand*.txt / matrix.txt
# ------------------------------------------------------------------------------------------

def read_template(path: str, n: int, dtype=np.float32) -> np.ndarray:
    """The template matrix as main() reads it (templated/luBatchedInplace.cu:27-34): the
    first n*n whitespace-separated tokens of the file reshaped to [n, n] -- a prefix of the
    token stream, NOT the top-left block of the 32-wide matrix."""
    out = np.empty((n, n), dtype=dtype)
    check(_lib.lib().lu_batched_read_tokens(str(path).encode(), out.ctypes.data, n * n, _dtype_code(dtype)))
    return out


def replicate(template: np.ndarray, batch: int) -> np.ndarray:
    """main()'s fill loop (templated/luBatchedInplace.cu:46-57): batch copies of the template."""
    template = np.ascontiguousarray(template)
    n = template.shape[0]
    out = np.empty((batch, n, n), dtype=template.dtype)
    check(_lib.lib().lu_batched_replicate(template.ctypes.data, out.ctypes.data, n, batch, _dtype_code(template.dtype)))
    return out


def write_to_file(A, path: str) -> None:
    """writeToFile (templated/verify.hpp:31-48): the first matrix of A[batch, n, n] (or A[n, n]) as text."""
    A = np.ascontiguousarray(A)
    first = A[0] if A.ndim == 3 else A
    first = np.ascontiguousarray(first)
    check(_lib.lib().lu_batched_write_matrix(first.ctypes.data, os.fsencode(path), first.shape[0], _dtype_code(first.dtype)))


def print_matrices(A) -> None:
    """printMatrices (templated/verify.hpp:12-29): the first matrix of A on stdout, then an empty line."""
    A = np.ascontiguousarray(A)
    first = np.ascontiguousarray(A[0] if A.ndim == 3 else A)
    check(_lib.lib().lu_batched_write_matrix(first.ctypes.data, None, first.shape[0], _dtype_code(first.dtype)))


def l1_norm(template: np.ndarray) -> float:
    """What the reference prints as "Condition number of the matrix is:": its calc_cond_num
    (templated/verify.hpp:188-281) reduces A to the identity without an augmented block and
    therefore returns ||A||_1 * 1.  Kept so the stdout contract is unchanged."""
    return float(np.abs(template).sum(axis=0).max().astype(template.dtype))


# ------------------------------------------------------------------------------------------
# main(): one run with the reference's stdout contract
# ------------------------------------------------------------------------------------------

def default_num_threads(n: int) -> int:
    """The sweep's NUMTHREADS table (templated/run.py:201-223): largest multiple of n <= 32."""
    return (32 // n) * n


def run_main(matrix_size: int, num_matrices: int, num_threads: int = 0, pivot_mode="parallel",
             input_file: str = "mtrand32_new1.txt", dtype=np.float32, template: np.ndarray | None = None,
             out=sys.stdout, verify: bool = True, warm: bool = False) -> dict:
    """One run of the reference executable (`./custom`), same stdout lines in the same order
    (parallel_pivot/luBatchedInplace.cu:13-18,38,82,104-105,115,133,136,158-162), on this
    library.  Returns the parsed numbers.  `num_threads` here is the library's threads per
    block (0 = default); the reference's per-warp packing is chosen by the library.
    """
    import torch

    n, batch = int(matrix_size), int(num_matrices)
    mode = _mode(pivot_mode)
    if num_threads:
        set_num_threads(num_threads)
    geo = geometry(n, batch, mode, dtype)
    p = lambda *a: print(*a, file=out)
    p("Matrix size:", n)
    p("Number of matrices:", batch)
    p("Number of threads per block:", geo.threads_per_block)
    p("Threads per matrix:", geo.threads_per_matrix)
    p("Matrices per block:", geo.matrices_per_block)
    p("Number of blocks:", geo.num_blocks)
    p("Reading data from file.")
    if template is None:
        template = read_template(input_file, n, dtype)
    p("Condition number of the matrix is:", l1_norm(template))
    t0 = time.perf_counter()
    A = replicate(template, batch)
    p("Time taken to read data: %g seconds" % (time.perf_counter() - t0))
    p("Data read from file.")
    dA = torch.from_numpy(A).cuda()
    p("Data copied to device.")
    if warm:
        lu_batched_inplace(dA.clone(), None, mode)
    enable_timing(True)
    try:
        lu_batched_inplace(dA, None, mode)
        ms = last_kernel_ms()
    finally:
        enable_timing(False)
        if num_threads:
            set_num_threads(0)
    p("Kernel execution time: %g milliseconds" % ms)
    res = {"matrix_size": n, "num_matrices": batch, "num_threads": geo.threads_per_block, "kernel_ms": ms}
    if verify:
        t0 = time.perf_counter()
        dOrig = torch.from_numpy(A).cuda()
        ok, bad, dev = verify_inv(dOrig, dA)
        p("Data copied back to host.")
        p("Correct inversions:", ok)
        p("Incorrect inversions:", bad)
        p("Time taken to verify inverse: %g seconds" % (time.perf_counter() - t0))
        res.update(correct=ok, incorrect=bad, max_abs_dev=dev)
    else:
        p("Data copied back to host.")
    res["A_inv"] = dA
    return res
