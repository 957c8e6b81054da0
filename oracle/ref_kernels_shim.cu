// TEST INFRASTRUCTURE ONLY.  Launch shim for the reference's own CUDA kernels, compiled
// for sm_100 from the headers where they lie (-I <reference>/<variant>), one translation
// unit per (variant, dtype) because every variant header defines the same symbol names
// and `extern __shared__ T shmem[]` cannot be instantiated for two T in one TU
// (SURVEY.md H9).  Build-time macros:
//   REF_T        float | double
//   REF_SYM      exported symbol name
//   REF_PIVOUT   defined when the header is build_ref.sh's pivot-exporting patch
//   REF_SMEM_KIND  0 | 1 | 2: which main()'s dynamic shared memory formula applies
//                (templated/luBatchedInplace.cu:68, serial_pivot/...cu:97,
//                parallel_pivot/...cu:118)
// Launch geometry follows the reference: TPM = N, NUMTHREADS from the sweep's table
// (templated/run.py:201-223: the largest multiple of N that is <= 32), MPB = T / N, one
// block per MPB matrices, dynamic smem opted in with cudaFuncSetAttribute.  The kernel's
// compile-time numMatrices only guards the tail; we instantiate it with INT_MAX and pad
// the device buffer to a whole number of blocks instead.
#include <climits>
#include <cstdint>
#include <cstdio>
#include "luBatchedInplace.cuh"

#if REF_SMEM_KIND == 0
#define REF_SMEM(N, TPM) ((N) * (N))
#elif REF_SMEM_KIND == 1
#define REF_SMEM(N, TPM) ((N) * (N) + (N))
#else
#define REF_SMEM(N, TPM) ((N) * (N) + (N) + 2 * (TPM))
#endif

namespace {
template <int N>
int run_one(const REF_T* hA, REF_T* hOut, int32_t* hPiv, long long batch, float* ms) {
    constexpr int TPM = N;
    constexpr int T = (32 / N) * N;
    constexpr int MPB = T / TPM;
    const long long blocks = (batch + MPB - 1) / MPB;
    const long long padded = blocks * MPB;
    const size_t elems = (size_t)N * N;
    REF_T* dA = nullptr;
    int* dP = nullptr;
    if (cudaMalloc(&dA, padded * elems * sizeof(REF_T)) != cudaSuccess) return -2;
    // pad with copies of matrix 0 so the tail lanes factor something finite
    cudaMemcpy(dA, hA, batch * elems * sizeof(REF_T), cudaMemcpyHostToDevice);
    for (long long b = batch; b < padded; ++b)
        cudaMemcpy(dA + b * elems, hA, elems * sizeof(REF_T), cudaMemcpyHostToDevice);
    constexpr int shmem = MPB * (REF_SMEM(N, TPM)) * (int)sizeof(REF_T);
    auto kern = batched_lu_subwarp<REF_T, N, TPM, MPB, INT_MAX>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, shmem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
#ifdef REF_PIVOUT
    if (cudaMalloc(&dP, padded * N * sizeof(int)) != cudaSuccess) return -2;
    cudaEventRecord(e0, 0);
    kern<<<(unsigned)blocks, T, shmem>>>(dA, dP);
#else
    cudaEventRecord(e0, 0);
    kern<<<(unsigned)blocks, T, shmem>>>(dA);
#endif
    cudaEventRecord(e1, 0);
    cudaError_t err = cudaEventSynchronize(e1);
    if (err == cudaSuccess) err = cudaGetLastError();
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms) *ms = t;
    cudaMemcpy(hOut, dA, batch * elems * sizeof(REF_T), cudaMemcpyDeviceToHost);
    if (dP && hPiv) cudaMemcpy(hPiv, dP, batch * N * sizeof(int), cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA);
    if (dP) cudaFree(dP);
    if (err != cudaSuccess) { fprintf(stderr, "ref kernel: %s\n", cudaGetErrorString(err)); return -3; }
    return 0;
}
}  // namespace

// Timing entry (round 2): the reference kernel on a DEVICE buffer the caller owns, launched `reps` times in place
// (A -> A^-1 -> A ...), each launch bracketed by CUDA events exactly as the reference's main() does
// (templated/luBatchedInplace.cu:71-82).  Only whole blocks are launched: *processed = floor(batch / MPB) * MPB.
// ms_cold = the first launch (the reference's convention: one cold launch per process), ms_warm = best of the rest.
namespace {
template <int N>
int time_one(REF_T* dA, int* dP, long long batch, int reps, float* ms_cold, float* ms_warm, long long* processed) {
    constexpr int TPM = N;
    constexpr int T = (32 / N) * N;
    constexpr int MPB = T / TPM;
    const long long blocks = batch / MPB;
    if (processed) *processed = blocks * MPB;
    if (blocks == 0) return -1;
    constexpr int shmem = MPB * (REF_SMEM(N, TPM)) * (int)sizeof(REF_T);
    auto kern = batched_lu_subwarp<REF_T, N, TPM, MPB, INT_MAX>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, shmem) != cudaSuccess) return -2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float cold = 0.f, warm = 1e30f;
    cudaError_t err = cudaSuccess;
    for (int r = 0; r < reps && err == cudaSuccess; ++r) {
        cudaEventRecord(e0, 0);
#ifdef REF_PIVOUT
        kern<<<(unsigned)blocks, T, shmem>>>(dA, dP);
#else
        (void)dP;
        kern<<<(unsigned)blocks, T, shmem>>>(dA);
#endif
        cudaEventRecord(e1, 0);
        err = cudaEventSynchronize(e1);
        if (err == cudaSuccess) err = cudaGetLastError();
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        if (r == 0) cold = t; else if (t < warm) warm = t;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess) { fprintf(stderr, "ref kernel: %s\n", cudaGetErrorString(err)); return -3; }
    if (ms_cold) *ms_cold = cold;
    if (ms_warm) *ms_warm = (reps > 1) ? warm : cold;
    return 0;
}
}  // namespace

#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
extern "C" int REF_CAT(REF_SYM, _device)(REF_T* dA, int* dP, int n, long long batch, int reps, float* ms_cold, float* ms_warm,
                                          long long* processed) {
    switch (n) {
#define C(N) case N: return time_one<N>(dA, dP, batch, reps, ms_cold, ms_warm, processed);
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16)
        C(17) C(18) C(19) C(20) C(21) C(22) C(23) C(24) C(25) C(26) C(27) C(28) C(29) C(30) C(31) C(32)
#undef C
    }
    return -1;
}

// Host buffers in, host buffers out.  Returns 0, or <0 on a CUDA error / bad n.
extern "C" int REF_SYM(const REF_T* hA, REF_T* hOut, int32_t* hPiv, int n, long long batch, float* kernel_ms) {
    switch (n) {
#define C(N) case N: return run_one<N>(hA, hOut, hPiv, batch, kernel_ms);
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16)
        C(17) C(18) C(19) C(20) C(21) C(22) C(23) C(24) C(25) C(26) C(27) C(28) C(29) C(30) C(31) C(32)
#undef C
    }
    return -1;
}
