/*
 * lubatched.h -- C ABI of the B200-native batched small-matrix inversion library
 * (liblubatched.so).  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference (sumukhashridhar/matrixInversion) has no library API: its operator
 * boundary is the templated kernel launch
 *     batched_lu_subwarp<FpType, N, TPM, MPB, B><<<numBlocks, T, shMem>>>(d_A)
 * inside a per-configuration executable (templated/luBatchedInplace.cu:76,
 * serial_pivot/luBatchedInplace.cu:105, parallel_pivot/luBatchedInplace.cu:127) whose
 * knobs are the compile-time macros MATRIXSIZE / NUMMATRICES / NUMTHREADS
 * (templated/luBatchedInplace.cu:4-6) and whose dtype is `using FpType`
 * (templated/verify.hpp:9-10).  Every entry point below names the piece of that
 * executable it replaces.  All knobs are runtime arguments here.
 *
 * Data layout (same as the reference, templated/luBatchedInplace.cuh:89-97):
 *     T A[batch][n][n], row-major, contiguous, matrix b at element offset b*n*n,
 * inverted IN PLACE.  n in [1, 32].
 */
#ifndef LUBATCHED_H_
#define LUBATCHED_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pivot_mode: which reference variant's semantics to reproduce */
#define LUB_PIVOT_NONE 0     /* templated/luBatchedInplace.cuh:78-126 */
#define LUB_PIVOT_SERIAL 1   /* serial_pivot/luBatchedInplace.cuh:104-167 (find_pivot :22-36) */
#define LUB_PIVOT_PARALLEL 2 /* parallel_pivot/luBatchedInplace.cuh:127-199 (find_pivot_parallel :12-44,
                                including the candidates its tree drops for non-power-of-two N) */
/* Extension: TRUE partial pivoting with LAPACK getrf semantics -- the arg-max is taken
 * over the UPDATED column (the reference searches un-eliminated entries, parallel_pivot/luBatchedInplace.cuh:159),
 * first maximum wins.  `piv` then receives LAPACK's ipiv (1-based: at step k the rows at positions k and
 * ipiv[k]-1 were interchanged; identical to cublas<t>getrfBatched's PivotArray), and lu_batched_inplace_ex
 * reports exactly-zero pivots in `info`.  This is the factorisation the reference's disabled check
 * verifyLUwithPivoting (parallel_pivot/verify.hpp:157-242) is written for. */
#define LUB_PIVOT_LAPACK 3
/* layout of the batch in memory */
#define LUB_LAYOUT_MATRIX_MAJOR 0      /* T A[batch][n][n]: the reference's (templated/luBatchedInplace.cuh:89-97) */
#define LUB_LAYOUT_BATCH_INTERLEAVED 1 /* T A[n][n][batch]: element (i, j) of consecutive matrices is contiguous */
/* dtype: the reference's FpType switch (templated/verify.hpp:9-10) */
#define LUB_DTYPE_F32 0
#define LUB_DTYPE_F64 1

/* error codes (the reference's CUDA_CHECK prints and exit(1)s, templated/luBatchedInplace.cuh:9;
 * this library never exits) */
#define LUB_OK 0
#define LUB_ERR_BAD_N (-1)
#define LUB_ERR_BAD_MODE (-2)
#define LUB_ERR_BAD_DTYPE (-3)
#define LUB_ERR_BAD_ARG (-4)
#define LUB_ERR_CUDA (-5)
#define LUB_ERR_NO_DEVICE (-6)
#define LUB_ERR_IO (-7)

/* ---------------------------------------------------------------------------------------
 * The hot path.  Replaces the kernel launch + its launch-geometry block
 * (parallel_pivot/luBatchedInplace.cu:8-11,118-129).
 *
 *   ptr    DEVICE pointer to T[batch][n][n]; overwritten by the inverses.
 *   piv    DEVICE pointer to int32[batch][n] or NULL.  When non-NULL receives the
 *          reference's permutation vector (shared-memory `pivots[]`,
 *          parallel_pivot/luBatchedInplace.cuh:140,146-148,161-168, never exported
 *          upstream): piv[b][i] = original row index sitting in row i after all swaps.
 *          pivot_mode NONE writes the identity.
 *   n, batch, pivot_mode, dtype   as above; batch may be 0.
 *
 * Asynchronous on the library's current stream for the current device (see
 * lu_batched_set_stream); returns LUB_OK or a negative error code.  No CPU fallback
 * exists: without a CUDA device the call fails with LUB_ERR_NO_DEVICE / LUB_ERR_CUDA.
 */
int lu_batched_inplace(void* ptr, int32_t* piv, int n, int64_t batch, int pivot_mode, int dtype);

/* Same, on an explicit stream (a cudaStream_t passed as void*; NULL = legacy default). */
int lu_batched_inplace_stream(void* ptr, int32_t* piv, int n, int64_t batch, int pivot_mode, int dtype,
                              void* stream);

/* Extended form of the hot path.
 *   info    DEVICE int32[batch] or NULL: numerical status per matrix.  pivot_mode LAPACK: 0, or k (1-based) for the
 *           first pivot U(k,k) that is exactly zero (the matrix is singular; its "inverse" is inf / NaN).  The
 *           reference's variants have no status: with modes 0-2 the array is set to 0.
 *   layout  LUB_LAYOUT_MATRIX_MAJOR, or LUB_LAYOUT_BATCH_INTERLEAVED (n <= 8, pivot modes NONE / LAPACK; piv and
 *           info keep their [batch][n] / [batch] layouts): one lane per matrix, every access a fully coalesced
 *           vector access across the batch.
 *   stream  a cudaStream_t passed as void*. */
int lu_batched_inplace_ex(void* ptr, int32_t* piv, int32_t* info, int n, int64_t batch, int pivot_mode, int dtype,
                          int layout, void* stream);
/* lu_batched_factor_inplace with `info` (matrix-major layout).  pivot_mode LAPACK: P A = L U exactly as getrf
 * stores it, piv = ipiv. */
int lu_batched_factor_inplace_ex(void* ptr, int32_t* piv, int32_t* info, int n, int64_t batch, int pivot_mode,
                                 int dtype, void* stream);
/* HOST helper: LAPACK swap lists ipiv[batch][n] (1-based) -> permutation vectors perm[batch][n] with the
 * semantics of the other modes' piv (row i of P A is row perm[i] of A), e.g. for lu_batched_verify_lu. */
int lu_batched_ipiv_to_perm(const int32_t* ipiv, int32_t* perm, int n, int64_t batch);

/* LU factors only: stops where the reference's k-loop ends
 * (parallel_pivot/luBatchedInplace.cuh:156-186, serial_pivot/...cuh:130-160,
 * templated/...cuh:100-110), i.e. before comp_inv / inversion -- the state upstream can only
 * look at by commenting the inversion out, and the one its disabled checks verifyLU
 * (templated/verify.hpp:105-186) and verifyLUwithPivoting (parallel_pivot/verify.hpp:157-242)
 * are written for.  In place: unit-lower L strictly below the diagonal, U on and above it,
 * rows in pivoted order (output row i factorises input row piv[i]); piv as above.  Values
 * equal the reference's left-looking Doolittle factors up to rounding.  Not a tuned path. */
int lu_batched_factor_inplace(void* ptr, int32_t* piv, int n, int64_t batch, int pivot_mode, int dtype);
int lu_batched_factor_inplace_stream(void* ptr, int32_t* piv, int n, int64_t batch, int pivot_mode, int dtype,
                                     void* stream);

/* Stream used by lu_batched_inplace and the helpers below on the calling thread
 * (the reference compiles with --default-stream per-thread, templated/run.py:48). */
int lu_batched_set_stream(void* stream);

/* HOST-pointer convenience replacing main()'s cudaMalloc + H2D + launch + D2H sequence
 * (parallel_pivot/luBatchedInplace.cu:112-135): `host_ptr`/`host_piv` are host buffers
 * (pinned or pageable); the batch is cut into chunks that are copied in, inverted and
 * copied out on rotating streams so transfers overlap the kernels.  Synchronous. */
int lu_batched_inplace_host(void* host_ptr, int32_t* host_piv, int n, int64_t batch, int pivot_mode,
                            int dtype);

/* The same over SEVERAL devices of this process (contiguous shards, no exchange): device d of
 * n_devices (0 = all visible) takes matrices [d * ceil(batch / n_devices), ...), one host thread and one pipeline per
 * device.  This is what a C caller replacing main()'s single-GPU sequence (parallel_pivot/luBatchedInplace.cu:112-135)
 * calls to use the whole box.  flags:
 *   LUB_HOST_REGISTER      page-lock the caller's (pageable) buffers in place for the duration of the call, so the
 *                          copies run as DMA at full speed (a pinned buffer is left alone);
 *   LUB_HOST_BIND_THREADS  bind every worker to the CPUs next to its GPU (/sys/bus/pci/devices/<id>/local_cpulist).
 * Where the host buffer physically lives decides the ceiling (a buffer on one NUMA node feeds every GPU through that
 * node's memory controllers): lu_batched_bind_thread_near_device binds the CALLING thread next to a device, for
 * callers that allocate (first touch) their buffers per device. */
#define LUB_HOST_REGISTER 1
#define LUB_HOST_BIND_THREADS 2
int lu_batched_inplace_host_multi(void* host_ptr, int32_t* host_piv, int n, int64_t batch, int pivot_mode,
                                  int dtype, int n_devices, int flags);
int lu_batched_bind_thread_near_device(int device);

/* NUMTHREADS knob (templated/luBatchedInplace.cu:6; table templated/run.py:201-223):
 * threads per block for subsequent launches of the calling thread, a multiple of 32 in [32, 256];
 * 0 restores the library's per-(n, dtype, pivot_mode) default.  lu_batched_get_threads returns the
 * value in force: the knob if set, else the library default for (n, dtype) with parallel pivoting
 * (the reference's headline variant; other modes: lu_batched_geometry), -1 without a device. */
int lu_batched_set_threads(int numthreads);
int lu_batched_get_threads(int n, int dtype);

/* Ablation knobs of the calling thread (the thesis keeps its variants as sibling directories --
 * swzl/luBatchedInplace.cuh:12-20 staging layout, shfl/luBatchedInplace.cuh:12-53 exchange, no_templ/ runtime N; here the
 * variants that exist as kernels in the library are selectable at run time; same pivots, values equal to rounding):
 *   LUB_OPT_STAGING      0 = library choice (TMA bulk tensor copies into a swizzled image where rows are 128 / 256 bytes or
 *                            padded to a line, 1-D cp.async.bulk span copies into a dense image elsewhere from N = 5 on),
 *                        1 = LSU staging (128-bit loads / cp.async into a padded or dense image) for every size, and the
 *                            one-phase lane = row kernel for pivot_mode 3 (library choice from n = 9 on: two phases, an LU
 *                            factorisation for the permutation, then the Gauss-Jordan of modes 1 / 2) and the generic
 *                            kernel for lu_batched_factor_inplace;
 *   LUB_OPT_FP64_TENSOR  fp64 N = 32, blocked elimination with DMMA rank-4 updates (csrc/lub_dmma.cuh):
 *                        0 = library choice: without pivoting only (as accurate as the unblocked elimination there);
 *                        1 = never (DFMA rank-1 updates with shuffle exchange);
 *                        2 = always: also in the pivot modes, 12 % faster, but the explicit 4 x 4 block inverses amplify
 *                            the conditioning of the diagonal blocks the reference's pivot rule leaves.
 * lu_batched_geometry / lu_batched_kernel_name report what the current setting launches. */
#define LUB_OPT_STAGING 1
#define LUB_OPT_FP64_TENSOR 2
int lu_batched_set_option(int option, int value);
int lu_batched_get_option(int option);

/* Launch geometry the library would use -- the numbers main() prints
 * ("Threads per matrix", "Matrices per block", "Number of blocks",
 * parallel_pivot/luBatchedInplace.cu:13-18). */
int lu_batched_geometry(int n, int64_t batch, int pivot_mode, int dtype, int* threads_per_block,
                        int* threads_per_matrix, int* matrices_per_block, int64_t* num_blocks,
                        int* dyn_smem_bytes);

/* Name of the CUDA kernel family the library launches for this configuration on a 16-byte
 * aligned batch ("lub_tma_kernel", "lub_v3_kernel", "lub_v4_kernel"); for profilers and bench
 * reports (the reference has one kernel, batched_lu_subwarp, parallel_pivot/luBatchedInplace.cuh:70).
 * Static string; NULL on bad arguments. */
const char* lu_batched_kernel_name(int n, int pivot_mode, int dtype);

/* Event-timed kernel time of the most recent lu_batched_inplace* call on this thread, in
 * milliseconds, kernel only (the reference's "Kernel execution time",
 * templated/luBatchedInplace.cu:71-82).  Timing is recorded only after
 * lu_batched_enable_timing(1); the query synchronises on the stop event. */
int lu_batched_enable_timing(int on);
float lu_batched_last_kernel_ms(void);

/* verify.hpp-compatible residual check (verifyInv, templated/verify.hpp:50-103) on HOST
 * buffers: r(i,j) = sum_l A[j][l]*Ainv[l][i] accumulated in T; a matrix is correct iff
 * every |r - delta_ij| < thr (the reference hard-codes thr = 1e-3).  Outputs may be NULL. */
int lu_batched_verify_inv(const void* A, const void* Ainv, int n, int64_t batch, int dtype, double thr,
                          int64_t* n_correct, int64_t* n_incorrect, double* max_abs_dev);

/* verifyLU (templated/verify.hpp:105-186) and verifyLUwithPivoting
 * (parallel_pivot/verify.hpp:157-242) on HOST buffers: L = unit-lower part of LU, U = its upper
 * part; a factorisation is correct iff every |(P A)(i,j) - sum_l L(i,l) U(l,j)| < thr (sum in T).
 * piv = the permutation vectors written by lu_batched_factor_inplace (row i of P A is row piv[i]
 * of A), or NULL for no pivoting (= verifyLU).  The reference permutes one template with pivotedA
 * because all its matrices are equal; here every matrix has its own vector. */
int lu_batched_verify_lu(const void* A, const void* LU, const int32_t* piv, int n, int64_t batch, int dtype,
                         double thr, int64_t* n_correct, int64_t* n_incorrect, double* max_abs_dev);

/* Same predicate evaluated on the DEVICE (buffers are device pointers); used by the sweep
 * driver so that a 1M-matrix check does not need two 4 GB host copies
 * (parallel_pivot/luBatchedInplace.cu:158 passes its vectors by value). */
int lu_batched_verify_inv_device(const void* dA, const void* dAinv, int n, int64_t batch, int dtype,
                                 double thr, int64_t* n_correct, int64_t* n_incorrect,
                                 double* max_abs_dev);

/* Same on an explicit stream (a cudaStream_t passed as void*): does not touch the calling thread's
 * lu_batched_set_stream setting. */
int lu_batched_verify_inv_device_stream(const void* dA, const void* dAinv, int n, int64_t batch, int dtype,
                                        double thr, int64_t* n_correct, int64_t* n_incorrect,
                                        double* max_abs_dev, void* stream);

/* Text input with the reference's semantics (templated/luBatchedInplace.cu:27-34): the
 * first `count` whitespace-separated tokens of `path`, parsed as T -- a PREFIX of the
 * token stream, not a sub-block.  Returns LUB_OK, or LUB_ERR_IO if the file is missing
 * or holds fewer tokens. */
int lu_batched_read_tokens(const char* path, void* out, int64_t count, int dtype);

/* main()'s replicate loop (templated/luBatchedInplace.cu:46-57): fill dst[batch][n][n]
 * (HOST) with copies of tmpl[n][n]. */
int lu_batched_replicate(const void* tmpl, void* dst, int n, int64_t batch, int dtype);

/* Device facts (deviceProps.cu:4-23): SM count, max dynamic smem per block, clock kHz. */
/* printMatrices / writeToFile (templated/verify.hpp:12-48): writes the FIRST matrix of the host buffer A
 * (the reference loops break after k = 0) as text, default ostream formatting, one row per line, to
 * `path`, or to stdout followed by an empty line when path is NULL. */
int lu_batched_write_matrix(const void* A, const char* path, int n, int dtype);

int lu_batched_device_info(int* sm_count, int* max_smem_optin, int* clock_khz, int* cc_major, int* cc_minor);

const char* lu_batched_last_error(void);
const char* lu_batched_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LUBATCHED_H_ */
