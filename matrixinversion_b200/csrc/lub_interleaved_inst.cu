// lub_interleaved_inst.cu -- instantiations and launcher of the batch-interleaved layout (lub_interleaved.cuh),
// one translation unit per dtype.  Compile with -DLUB_T=float|double -DLUB_TN=f32|f64.
#include "lub_launch.cuh"
#include "lub_interleaved.cuh"

namespace lub {

template <typename T, int N, int MODE, int VEC>
static cudaError_t launch_il(T* A, int32_t* piv, int32_t* info, long long batch, cudaStream_t stream, cudaEvent_t ev0) {
    auto kern = lub_interleaved_kernel<T, N, MODE, VEC>;
    int dev = 0, sms = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    const int threads = 128;
    const long long groups = batch / VEC;
    long long blocks = (groups + threads - 1) / threads;
    const long long cap = (long long)sms * 16;   // grid-stride loop: a few resident blocks per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return cudaSuccess;
    if (ev0) {
        err = cudaEventRecord(ev0, stream);
        if (err != cudaSuccess) return err;
    }
    kern<<<(unsigned)blocks, threads, 0, stream>>>(A, piv, info, batch);
    return cudaGetLastError();
}

template <typename T, int N, int MODE>
static cudaError_t launch_il_vec(T* A, int32_t* piv, int32_t* info, long long batch, cudaStream_t stream, cudaEvent_t ev0) {
    constexpr int VEC = InterleavedCfg<T, N>::VEC;
    const bool vec_ok = (batch % VEC == 0) && (reinterpret_cast<uintptr_t>(A) % (VEC * sizeof(T)) == 0);
    if (VEC > 1 && vec_ok) return launch_il<T, N, MODE, VEC>(A, piv, info, batch, stream, ev0);
    return launch_il<T, N, MODE, 1>(A, piv, info, batch, stream, ev0);
}

template <typename T, int N>
static cudaError_t launch_il_mode(T* A, int32_t* piv, int32_t* info, long long batch, int mode, cudaStream_t stream, cudaEvent_t ev0) {
    switch (mode) {
        case 0: return launch_il_vec<T, N, 0>(A, piv, info, batch, stream, ev0);
        case 1: return launch_il_vec<T, N, 1>(A, piv, info, batch, stream, ev0);
        case 2: return launch_il_vec<T, N, 2>(A, piv, info, batch, stream, ev0);
        case 3: return launch_il_vec<T, N, 3>(A, piv, info, batch, stream, ev0);
    }
    return cudaErrorInvalidValue;
}

#define LUB_IL_CAT_(a) launch_interleaved_##a
#define LUB_IL_CAT(a) LUB_IL_CAT_(a)
cudaError_t LUB_IL_CAT(LUB_TN)(void* A, int32_t* piv, int32_t* info, int n, long long batch, int mode, cudaStream_t stream, cudaEvent_t ev0) {
    LUB_T* At = static_cast<LUB_T*>(A);
    switch (n) {
#define C(N) case N: return launch_il_mode<LUB_T, N>(At, piv, info, batch, mode, stream, ev0);
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8)
#undef C
    }
    return cudaErrorInvalidValue;
}

}  // namespace lub
