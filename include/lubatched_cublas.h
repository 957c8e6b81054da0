/*
 * lubatched_cublas.h -- C ABI of liblubatched_cublas.so, the cuBLAS comparison baseline.
 * Kept in its own shared object so that the product library (liblubatched.so) has no
 * cuBLAS dependency and no library GEMM/LU call anywhere near its hot path.
 */
#ifndef LUBATCHED_CUBLAS_H_
#define LUBATCHED_CUBLAS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cuBLAS getrfBatched + getriBatched baseline on DEVICE buffers (the comparison the
 * reference's README quotes but whose benchmark.cu is absent upstream, README.md:63-69).
 * dA is T[batch][n][n] (overwritten by its LU), dAinv receives the inverses, pivoting != 0
 * uses a pivot array.  ms_getrf / ms_getri (may be NULL) receive event timings. */
int lu_batched_cublas_baseline(void* dA, void* dAinv, int n, int64_t batch, int dtype, int pivoting,
                               float* ms_getrf, float* ms_getri);

#ifdef __cplusplus
}
#endif
#endif /* LUBATCHED_CUBLAS_H_ */
