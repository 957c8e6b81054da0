"""ctypes binding of liblubatched.so (the C ABI in include/lubatched.h).

This is the binding a maintainer of the reference's sweep drivers would add (see
INTEGRATION.md).  It fails loudly: there is no Python/CPU fallback for the hot path -- if
the shared object is missing it is built with `make` (nvcc), and if that fails, or no
CUDA device is present when a compute entry point is called, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "liblubatched.so")
CUBLAS_LIB_PATH = os.path.join(_HERE, "liblubatched_cublas.so")

PIVOT_NONE, PIVOT_SERIAL, PIVOT_PARALLEL, PIVOT_LAPACK = 0, 1, 2, 3
LAYOUT_MATRIX_MAJOR, LAYOUT_BATCH_INTERLEAVED = 0, 1
HOST_REGISTER, HOST_BIND_THREADS = 1, 2
OPT_STAGING, OPT_FP64_TENSOR = 1, 2
DTYPE_F32, DTYPE_F64 = 0, 1

ERRORS = {
    -1: "LUB_ERR_BAD_N", -2: "LUB_ERR_BAD_MODE", -3: "LUB_ERR_BAD_DTYPE", -4: "LUB_ERR_BAD_ARG",
    -5: "LUB_ERR_CUDA", -6: "LUB_ERR_NO_DEVICE", -7: "LUB_ERR_IO",
}


class LubError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "LUB_ERR"), code, msg))
        self.code = code


def build(jobs: int | None = None) -> None:
    """Compile every CUDA translation unit for sm_100a (in-tree, via csrc/Makefile)."""
    jobs = jobs or max(1, os.cpu_count() or 1)
    subprocess.check_call(["make", "-C", _CSRC, "-j%d" % jobs], stdout=subprocess.DEVNULL)


_P = ctypes.c_void_p
_I32P = ctypes.POINTER(ctypes.c_int32)
_I64P = ctypes.POINTER(ctypes.c_int64)
_IP = ctypes.POINTER(ctypes.c_int)
_DP = ctypes.POINTER(ctypes.c_double)
_FP = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); mirrors include/lubatched.h line by line
SIGNATURES = {
    "lu_batched_inplace": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "lu_batched_inplace_stream": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _P]),
    "lu_batched_inplace_ex": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
    "lu_batched_factor_inplace_ex": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _P]),
    "lu_batched_ipiv_to_perm": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64]),
    "lu_batched_factor_inplace": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "lu_batched_factor_inplace_stream": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _P]),
    "lu_batched_set_stream": (ctypes.c_int, [_P]),
    "lu_batched_inplace_host": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "lu_batched_inplace_host_multi": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "lu_batched_bind_thread_near_device": (ctypes.c_int, [ctypes.c_int]),
    "lu_batched_set_option": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "lu_batched_get_option": (ctypes.c_int, [ctypes.c_int]),
    "lu_batched_set_threads": (ctypes.c_int, [ctypes.c_int]),
    "lu_batched_get_threads": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "lu_batched_kernel_name": (ctypes.c_char_p, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "lu_batched_geometry": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _IP, _IP, _IP, _I64P, _IP]),
    "lu_batched_enable_timing": (ctypes.c_int, [ctypes.c_int]),
    "lu_batched_last_kernel_ms": (ctypes.c_float, []),
    "lu_batched_verify_inv": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_double, _I64P, _I64P, _DP]),
    "lu_batched_verify_lu": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_double, _I64P, _I64P, _DP]),
    "lu_batched_verify_inv_device": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_double, _I64P, _I64P, _DP]),
    "lu_batched_verify_inv_device_stream": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_double, _I64P, _I64P, _DP, _P]),
    "lu_batched_read_tokens": (ctypes.c_int, [ctypes.c_char_p, _P, ctypes.c_int64, ctypes.c_int]),
    "lu_batched_write_matrix": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]),
    "lu_batched_replicate": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int]),
    "lu_batched_device_info": (ctypes.c_int, [_IP, _IP, _IP, _IP, _IP]),
    "lu_batched_last_error": (ctypes.c_char_p, []),
    "lu_batched_version": (ctypes.c_char_p, []),
}
CUBLAS_SIGNATURES = {
    "lu_batched_cublas_baseline": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _FP, _FP]),
}

_lib = None
_cublas = None


def _bind(lib, sigs):
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: loud by design
        fn.restype = res
        fn.argtypes = args
    return lib


def lib():
    """The loaded product library; builds it first if the .so is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = _bind(ctypes.CDLL(LIB_PATH), SIGNATURES)
    return _lib


def cublas_lib():
    global _cublas
    if _cublas is None:
        if not os.path.exists(CUBLAS_LIB_PATH):
            build()
        _cublas = _bind(ctypes.CDLL(CUBLAS_LIB_PATH), CUBLAS_SIGNATURES)
    return _cublas


def check(rc: int) -> None:
    if rc != 0:
        raise LubError(rc, lib().lu_batched_last_error().decode())
