// lubatched_cublas.cu -- cuBLAS getrfBatched + getriBatched comparison baseline
// (include/lubatched_cublas.h).  The reference quotes its speed-ups against exactly this
// pair of calls (README.md:14-16,36-40) but its benchmark.cu is absent upstream
// (README.md:63-69), so the harness is written new.  Not part of the product path.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/lubatched_cublas.h"

namespace {
template <typename T>
__global__ void fill_ptrs(T** pa, T** pc, T* a, T* c, long long stride, long long batch) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) { pa[i] = a + i * stride; pc[i] = c + i * stride; }
}
}  // namespace

// Row-major A handed to column-major cuBLAS is A^T; (A^T)^-1 = (A^-1)^T, which read back
// row-major is A^-1 -- same bytes as our kernel's output.
extern "C" int lu_batched_cublas_baseline(void* dA, void* dAinv, int n, int64_t batch, int dtype, int pivoting,
                                          float* ms_getrf, float* ms_getri) {
    if (n < 1 || batch < 1 || !dA || !dAinv || (dtype != 0 && dtype != 1)) return -4;
    static cublasHandle_t handle = nullptr;
    if (!handle && cublasCreate(&handle) != CUBLAS_STATUS_SUCCESS) return -5;
    void **pa = nullptr, **pc = nullptr;
    int *piv = nullptr, *info = nullptr;
    cudaEvent_t e0, e1, e2;
    int rc = 0;
    if (cudaMalloc(&pa, batch * sizeof(void*)) != cudaSuccess || cudaMalloc(&pc, batch * sizeof(void*)) != cudaSuccess ||
        cudaMalloc(&info, batch * sizeof(int)) != cudaSuccess ||
        (pivoting && cudaMalloc(&piv, batch * n * sizeof(int)) != cudaSuccess)) {
        rc = -5;
    } else {
        const unsigned blocks = (unsigned)((batch + 255) / 256);
        if (dtype == 0) fill_ptrs<float><<<blocks, 256>>>((float**)pa, (float**)pc, (float*)dA, (float*)dAinv, (long long)n * n, batch);
        else fill_ptrs<double><<<blocks, 256>>>((double**)pa, (double**)pc, (double*)dA, (double*)dAinv, (long long)n * n, batch);
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
        cublasStatus_t s1, s2;
        cudaEventRecord(e0, 0);
        if (dtype == 0) s1 = cublasSgetrfBatched(handle, n, (float**)pa, n, piv, info, (int)batch);
        else s1 = cublasDgetrfBatched(handle, n, (double**)pa, n, piv, info, (int)batch);
        cudaEventRecord(e1, 0);
        if (dtype == 0) s2 = cublasSgetriBatched(handle, n, (const float**)pa, n, piv, (float**)pc, n, info, (int)batch);
        else s2 = cublasDgetriBatched(handle, n, (const double**)pa, n, piv, (double**)pc, n, info, (int)batch);
        cudaEventRecord(e2, 0);
        cudaError_t ce = cudaEventSynchronize(e2);
        if (s1 != CUBLAS_STATUS_SUCCESS || s2 != CUBLAS_STATUS_SUCCESS || ce != cudaSuccess) rc = -5;
        float t = 0.f;
        if (ms_getrf) { cudaEventElapsedTime(&t, e0, e1); *ms_getrf = t; }
        if (ms_getri) { cudaEventElapsedTime(&t, e1, e2); *ms_getri = t; }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    }
    if (pa) cudaFree(pa);
    if (pc) cudaFree(pc);
    if (info) cudaFree(info);
    if (piv) cudaFree(piv);
    return rc;
}
