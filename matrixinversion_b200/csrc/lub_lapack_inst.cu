// lub_lapack_inst.cu -- instantiations and launcher of pivot_mode 3 (lub_lapack.cuh; the two-phase kernels are
// lub_bulk_kernel / lub_tma_kernel with MODE = kModeLapack), one translation unit per dtype.  Compile with -DLUB_T=float|double -DLUB_TN=f32|f64.
#include "lub_launch.cuh"
#include "lub_lapack.cuh"

namespace lub {

#ifndef LUB_LAPACK_TWO_PHASE_MIN_N
#define LUB_LAPACK_TWO_PHASE_MIN_N 9
#endif
constexpr int kLapackTwoPhaseMinN = LUB_LAPACK_TWO_PHASE_MIN_N;

template <typename T, int N, bool LUONLY>
static cudaError_t launch_lapack_n(void* A, int32_t* ipiv, int32_t* info, long long batch, int threads_req, cudaStream_t stream,
                                   LaunchInfo* li, int flags, cudaEvent_t ev0) {
    using L = LapackLayout<T, N>;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    LaunchCtx x{batch, threads_req > 0 ? threads_req : 256, stream, li, (flags & kLaunchDryRun) != 0, ev0, dev};
    // N >= kLapackTwoPhaseMinN: the two-phase kernel -- getrf's permutation from an LU factorisation in the lane = row layout
    // (prepass_getrf, N (N - 1) / 2 shuffles), then the permuted-load register Gauss-Jordan of modes 1 / 2
    // (lub_bulk_kernel<..., kModeLapack>); LU only: the first phase alone, factors stored from its registers.
    // LUB_OPT_STAGING = 1 keeps the one-phase kernels below.  Measured: profiles/r02_mode3_twophase.md.
    // ... on the swizzled TMA image where the rows are whole 128-byte lines (fp32 N = 32, fp64 N = 16, 32): the lane = row
    // accesses of the first phase are 8-way bank conflicts on a dense image whose rows are a multiple of 32 banks
    if constexpr (N >= kLapackTwoPhaseMinN && (N * sizeof(T)) % 128 == 0 && TmaCfg<T, N, kModeParallel>::ON) {
        using TC = TmaCfg<T, N, kModeParallel>;
        using VC = V3Cfg<T, N, kModeParallel>;
        using TL = TmaLayout<T, N, TC::GR, TC::GC, kModeLapack>;
        if ((x.dry_run || reinterpret_cast<uintptr_t>(A) % 16 == 0) && batch <= 0x7fffff00ll && !(flags & kLaunchNoTma)) {
            // factors only, fp32: one image per warp at 24 warps per SM (lu_hi_occupancy, lub_launch.cuh)
            constexpr bool HI = LUONLY && sizeof(T) == 4;
            constexpr int TOPT = HI ? (TC::OPT & ~kTmaDB) : TC::OPT, TMAXT = HI ? 768 : TC::MAXT;
            constexpr int NIMG = (TOPT & kTmaDB) ? 2 : 1;
            static KernelCache cache4[kMaxDevices];
            auto kern4 = lub_tma_kernel<T, N, TC::GR, TC::GC, kModeLapack, (TMAXT > kMaxThreads ? 1 : VC::MINB), TC::BSYNC, false, false, false,
                                        TOPT | (LUONLY ? kTmaLuOnly : 0), TMAXT>;
            if (threads_req <= 0) x.threads = HI ? 768 : TC::THREADS;
            return run_kernel(kern4, cache4[dev], x, TMAXT, [](int w) { return TL::smem_bytes(w, NIMG); }, TL::MPW, TL::G,
                              LUONLY ? "lub_tma_kernel<getrf, LUONLY>" : "lub_tma_kernel<getrf>", [&](unsigned blocks, int smem) {
                                  const CUtensorMap* map = nullptr;
                                  cudaError_t e = cached_batch_tmap<T>(&map, A, N, batch, TL::MPW, dev);
                                  if (e != cudaSuccess) return e;
                                  e = ev0 ? cudaEventRecord(ev0, stream) : cudaSuccess;
                                  if (e != cudaSuccess) return e;
                                  kern4<<<blocks, x.threads, smem, stream>>>(*map, static_cast<T*>(A), ipiv, batch, info);
                                  return cudaGetLastError();
                              });
        }
    }
    // N <= 8 where one lane holds a whole matrix (fp32 N <= 8, fp64 N <= 6): the same kernel, every lane inverts / factorises
    // its matrix in its own registers (N = 5 fp32: 0.22 -> 0.04 ms)
    constexpr bool kOneLane = N >= 2 && N < kLapackTwoPhaseMinN && BulkCfg<T, N, kModeLapack>::GR * BulkCfg<T, N, kModeLapack>::GC == 1;
    if constexpr (N >= kLapackTwoPhaseMinN || kOneLane) {
        using BC = BulkCfg<T, N, kModeLapack>;
        using BL = BulkLayout<T, N, BC::GR, BC::GC, kModeLapack>;
        const bool ok = x.dry_run || reinterpret_cast<uintptr_t>(A) % 16 == 0 || (BL::CH == 1 && reinterpret_cast<uintptr_t>(A) % sizeof(T) == 0);
        if (ok && !(flags & kLaunchNoTma)) {
            static KernelCache cache3[kMaxDevices];
            constexpr bool HI = LUONLY && sizeof(T) == 4 && lu_hi_occupancy(N, 4);  // factors only, fp32: 24 warps, one image each
            constexpr int BMAXT = HI ? 768 : BC::MAXT, BMINB = HI ? 1 : BC::MINB, BNIMG = (HI || (BC::OPT & kBulkSingle)) ? 1 : 2;
            auto kern3 = lub_bulk_kernel<T, N, BC::GR, BC::GC, kModeLapack, BMINB, false,
                                         BC::OPT | (LUONLY ? kBulkLuOnly : 0) | (HI ? kBulkSingle : 0), BMAXT>;
            if (threads_req <= 0) x.threads = HI ? 768 : BC::THREADS;
            return run_kernel(kern3, cache3[dev], x, BMAXT, [](int w) { return BL::smem_bytes(w, BNIMG); }, BL::MPW, BL::G,
                              LUONLY ? "lub_bulk_kernel<getrf, LUONLY>" : "lub_bulk_kernel<getrf>", [&](unsigned blocks, int smem) {
                                  cudaError_t e = ev0 ? cudaEventRecord(ev0, stream) : cudaSuccess;
                                  if (e != cudaSuccess) return e;
                                  kern3<<<blocks, x.threads, smem, stream>>>(static_cast<T*>(A), ipiv, batch, info);
                                  return cudaGetLastError();
                              });
        }
    }
    // one phase, lane = row: N <= 8, unaligned batches, LUB_OPT_STAGING = 1.  (Round 2's one-phase kernel on the 2-D lane grid,
    // scripts/tune/lub_lapack2.cuh, is superseded by the two-phase kernels at every size it served: fp64 N = 32 21.2 -> 11.8 ms.)
    static KernelCache cache[kMaxDevices];
    auto kern = lub_lapack_kernel<T, N, LUONLY>;
    return run_kernel(kern, cache[dev], x, kMaxThreads, [](int w) { return L::smem_bytes(w); }, L::MPW, L::G,
                      LUONLY ? "lub_lapack_kernel<LUONLY>" : "lub_lapack_kernel", [&](unsigned blocks, int smem) {
                          cudaError_t e = ev0 ? cudaEventRecord(ev0, stream) : cudaSuccess;
                          if (e != cudaSuccess) return e;
                          kern<<<blocks, x.threads, smem, stream>>>(static_cast<T*>(A), ipiv, info, batch);
                          return cudaGetLastError();
                      });
}

#define LUB_LAPACK_CAT_(a) launch_lapack_##a
#define LUB_LAPACK_CAT(a) LUB_LAPACK_CAT_(a)
cudaError_t LUB_LAPACK_CAT(LUB_TN)(void* A, int32_t* ipiv, int32_t* info, int n, long long batch, int threads, cudaStream_t stream,
                                   LaunchInfo* li, int flags, cudaEvent_t ev0) {
    const bool lu = (flags & kLaunchLuOnly) != 0;
    switch (n) {
#define C(N) case N: return lu ? launch_lapack_n<LUB_T, N, true>(A, ipiv, info, batch, threads, stream, li, flags, ev0) \
                               : launch_lapack_n<LUB_T, N, false>(A, ipiv, info, batch, threads, stream, li, flags, ev0);
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16)
        C(17) C(18) C(19) C(20) C(21) C(22) C(23) C(24) C(25) C(26) C(27) C(28) C(29) C(30) C(31) C(32)
#undef C
    }
    return cudaErrorInvalidValue;
}

}  // namespace lub
