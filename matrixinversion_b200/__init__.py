"""matrixinversion_b200 -- B200-native batched small-matrix inversion (N <= 32, fp32/fp64,
no / serial / parallel pivoting) behind the C ABI of include/lubatched.h.

Scope: the `luBatchedInplace` hot path of sumukhashridhar/matrixInversion and the host
steps around it.  The compute lives in csrc/ (hand-written CUDA for
sm_100a); this package is the ctypes host side.
"""
from ._lib import (DTYPE_F32, DTYPE_F64, LAYOUT_BATCH_INTERLEAVED, LAYOUT_MATRIX_MAJOR, PIVOT_LAPACK, PIVOT_NONE, PIVOT_PARALLEL,
                   PIVOT_SERIAL, LubError, build)
from .api import (Geometry, bind_thread_near_device, default_num_threads, ipiv_to_perm, lu_batched_inplace_host_multi, device_info, enable_timing, geometry, kernel_name, l1_norm,
                  last_kernel_ms, lu_batched_factor_inplace, lu_batched_inplace, lu_batched_inplace_ptr, print_matrices,
                  read_template, replicate, run_main, set_num_threads, set_option, verify_inv, verify_lu, write_to_file)
from .sharding import shard_range

__all__ = [
    "DTYPE_F32", "DTYPE_F64", "LAYOUT_BATCH_INTERLEAVED", "LAYOUT_MATRIX_MAJOR", "PIVOT_LAPACK", "ipiv_to_perm", "lu_batched_inplace_host_multi", "bind_thread_near_device", "PIVOT_NONE", "PIVOT_PARALLEL", "PIVOT_SERIAL", "LubError", "build",
    "Geometry", "default_num_threads", "device_info", "enable_timing", "geometry", "kernel_name", "l1_norm",
    "last_kernel_ms", "lu_batched_factor_inplace", "lu_batched_inplace", "lu_batched_inplace_ptr", "read_template",
    "replicate", "run_main", "set_num_threads", "set_option", "verify_inv", "verify_lu", "shard_range", "print_matrices", "write_to_file",
]
