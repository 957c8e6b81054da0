// lub_launch.cuh -- per-(T, N, MODE) launch wrapper: picks the lane layout, sizes the grid
// from the SM count and the occupancy of the instantiation, opts in to dynamic shared
// memory.  Replaces the launch-geometry arithmetic of the reference's main()
// (parallel_pivot/luBatchedInplace.cu:8-11,118-127) and the NUMTHREADS table of its sweep
// driver (templated/run.py:201-223).
#pragma once
#include <type_traits>
#include "lub_kernel.cuh"
#include "lub_fast.cuh"
#include "lub_v3.cuh"
#include "lub_v4.cuh"
#include "lub_tma.cuh"

namespace lub {

struct LaunchInfo {
    int threads_per_block;
    int threads_per_matrix;
    int matrices_per_block;
    long long num_blocks;
    int dyn_smem_bytes;
    int regs_per_thread;
    int blocks_per_sm;
    const char* kernel;
};

// int launcher(A, piv, batch, threads (0 = default), stream, info (may be NULL), flags)
// flags: bit 0 = dry run (fill `info`, launch nothing), bit 1 = LU factors only (no inversion)
constexpr int kLaunchDryRun = 1, kLaunchLuOnly = 2;
using LaunchFn = cudaError_t (*)(void*, int32_t*, long long, int, cudaStream_t, LaunchInfo*, int);

// Lane layout choice.  A matrix is spread over G = GR x GC lanes, each holding an LR x LC
// block in registers.  Smallest G whose block fits the per-lane element budget wins (fewer
// lanes per matrix = fewer shuffles per flop); among the splits of that G the one with the
// fewest per-step shuffles + selects wins.
struct Cfg { int gr, gc; };

constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

constexpr Cfg pick_cfg(int n, int elem_budget) {
    for (int g = 1; g <= 32; g *= 2) {
        int best_cost = 1 << 30, best_gr = 0;
        for (int gr = g; gr >= 1; gr /= 2) {
            const int gc = g / gr;
            const int lr = cdiv(n, gr), lc = cdiv(n, gc);
            if (lr * lc > elem_budget) continue;
            const int cost = 100 * ((gc > 1 ? lr : 0) + (gr > 1 ? lc : 0) + lr) + lr * lc;
            if (cost < best_cost) { best_cost = cost; best_gr = gr; }
        }
        if (best_gr) return Cfg{best_gr, g / best_gr};
    }
    return Cfg{32, 1};
}

template <typename T> constexpr int elem_budget() { return sizeof(T) == 4 ? 64 : 40; }

template <typename T, int N, int MODE>
struct AutoCfg {
    static constexpr Cfg c = pick_cfg(N, elem_budget<T>());
    static constexpr int GR = c.gr, GC = c.gc;
};

// Layout choice for the v3 kernel.  Cost model measured on B200 (profiles/): the shared-memory
// crossbar (shuffles included) moves one word per lane per cycle, so the per-step exchange
// costs LR + LC shuffles per tile and a matrix should sit on as few lanes as its register
// block allows.  Smallest G whose LR x LC block fits the element budget; among its splits the
// one with the least exchange, then the fewest rows (per-row selects).
constexpr Cfg pick_v3_cfg(int n, int es, bool pivoting) {
    const int epv = 16 / es;
    const int chv = (n % epv == 0) ? epv : ((epv == 4 && n % 2 == 0) ? 2 : 1);
    const int ch = pivoting ? 1 : chv;
    const int cpr = n / ch;
    const int budget = (es == 4) ? 64 : 36;  // elements per lane (fp64: 6 x 6 keeps N <= 24 on 16 lanes)
    for (int g = 1; g <= 32; g *= 2) {
        if (g == 2) continue;  // two-lane groups never won a sweep
        int best_cost = 1 << 30;
        Cfg best{0, 0};
        for (int gr = 1; gr <= g; gr *= 2) {
            const int gc = g / gr;
            const int lr = cdiv(n, gr), lc = cdiv(cpr, gc) * ch;
            if (lr * lc > budget) continue;
            const int cost = 64 * ((gc > 1 ? lr : 0) + (gr > 1 ? lc : 0)) + lr;
            if (cost < best_cost) { best_cost = cost; best = Cfg{gr, gc}; }
        }
        if (best.gr) return best;
    }
    return Cfg{4, 8};
}

// Small fp32 matrices need few registers per lane and their tiles are latency-bound (staging wait, pivot
// search): more resident warps win 10-20 % there (profiles/r01_tune_late.jsonl, "minb" sweep).  From N = 13
// on the register cap of three blocks per SM spills.
constexpr int pick_minb(int n, int es, bool pivoting) {
    if (es != 4) return (n == 7 || n == 8) ? 3 : 2;  // fp64: 10-15 % at N = 7, 8; spills from N = 10 on
    if (n <= 6) return 4;
    if (n == 7 || (n >= 9 && n <= 11)) return 3;
    if (n == 12 && pivoting) return 3;
    return 2;
}

// Per-tile block barrier of lub_v3_kernel (keeps a block's warps on the same stretch of straight-line code):
// pays from the sizes on whose unrolled code outgrows the instruction caches -- N >= 24, and N >= 18 with
// the position-wise pivot search of the parallel mode; below that it costs 4-6 % (13 % at N = 23 without
// pivoting), profiles/r01_tune_late.jsonl "bsync" sweep.
constexpr bool pick_bsync(int n, int mode) { return n >= 24 || (mode == kModeParallel && n >= 18); }

// Double-buffered cp.async prefetch of the dense image (lub_v3_kernel PFD): where it was measured to win
// (profiles/r01_tune_late.jsonl "pfd": without pivoting 10-29 % for N = 9..23, 27, 29; serial 3-12 %; parallel
// 1-16 % up to N = 17).  At N = 15, 21, 30, 31 with a pivot search the second image costs a resident block;
// multiples of 4 have their own 16-byte-image / TMA paths.
constexpr bool pick_pfd(int n, int es, int mode) {
    if (es != 4) return false;
    const bool small = n == 9 || n == 10 || n == 11 || n == 13 || n == 14 || n == 17;
    if (mode == kModeParallel) return small;
    if (mode == kModeSerial) return small || n == 18 || n == 19 || n == 23 || n == 27 || n == 29;
    return small || n == 15 || n == 18 || n == 19 || n == 21 || n == 22 || n == 23 || n == 25 || n == 26 || n == 27 || n == 29;
}

// The same prefetch in the fp64 kernel (lub_v4_kernel PFD), odd N (the sizes whose image is dense in every
// mode): without pivoting -10..-26 % for N = 7..21, pivot modes -7..-13 % for N = 7, 9, 13, 17, 19; N = 11, 15, 23
// with a pivot search lose a resident block to the second image (profiles/r01_tune_late.jsonl "pfd64").
constexpr bool pick_pfd64(int n, int mode) {
    if (pick_dense_even(n, mode)) return true;  // fp64 N = 10, 14, 18, 20 without pivoting: -20..-27 % (dense + prefetch)
    if (n % 2 == 0 || n < 7 || n > 21) return false;
    if (mode == kModeNone) return true;
    return n == 7 || n == 9 || n == 13 || n == 17 || n == 19 || (n == 21 && mode == kModeSerial);
}

template <typename T, int N, int MODE>
struct V3Cfg {
#if defined(LUB_FORCE_GR) && defined(LUB_FORCE_GC)
    static constexpr int GR = LUB_FORCE_GR, GC = LUB_FORCE_GC;
#else
    static constexpr Cfg c = pick_v3_cfg(N, (int)sizeof(T), MODE != kModeNone);
    static constexpr int GR = c.gr, GC = c.gc;
#endif
    // resident 256-thread blocks per SM the kernel is compiled for (register cap 128 / 80 / 64 per thread)
    static constexpr int MINB = pick_minb(N, (int)sizeof(T), MODE != kModeNone);
    static constexpr bool BSYNC = pick_bsync(N, MODE);
};

#ifndef LUB_USE_TMA
#define LUB_USE_TMA 1
#endif
constexpr bool kUseTma = LUB_USE_TMA != 0;

// Which configurations run the TMA-staged kernel (lub_tma.cuh), on which lane grid and with or without
// the per-tile block barrier.  Measured choices (profiles/r01_tune_tma.jsonl, r01_tune_late.jsonl):
//   * rows of one 128-byte line (N = 32 fp32, N = 16 fp64): every mode, the v3 lane grid;
//   * rows of two lines (N = 32 fp64): pivot modes only -- without pivoting the rolled-step kernel
//     (lub_v4.cuh) is still ahead;
//   * fp32 N = 20, 24, 28 (rows zero-padded to one line by the TMA unit): no pivoting and serial pivoting,
//     plus parallel pivoting at N = 24; elsewhere the position-wise pivot search on the swizzled image
//     (4-way bank conflicts on its column walk) loses to the odd-stride image of lub_v3.cuh.  Smaller N
//     would need a padded image too large for two blocks per SM.
struct TmaChoice { bool on; int gr, gc; bool bsync; };
constexpr TmaChoice pick_tma(int n, int es, int mode, Cfg v3) {
    const int rowb = n * es;
    const bool piv = mode != kModeNone;
    if (rowb == 128) return TmaChoice{true, v3.gr, v3.gc, piv};
    if (rowb == 256) return TmaChoice{piv, v3.gr, v3.gc, true};
    if (es == 4 && n == 20) return TmaChoice{mode != kModeParallel, 4, 2, piv};
    if (es == 4 && n == 24) return TmaChoice{true, 8, 2, mode == kModeParallel};
    if (es == 4 && n == 28) return TmaChoice{mode != kModeParallel, 4, 4, false};
    return TmaChoice{false, v3.gr, v3.gc, false};
}
template <typename T, int N, int MODE>
struct TmaCfg {
    static constexpr TmaChoice c = pick_tma(N, (int)sizeof(T), MODE, Cfg{V3Cfg<T, N, MODE>::GR, V3Cfg<T, N, MODE>::GC});
    static constexpr bool ON = kUseTma && c.on;
    static constexpr int GR = c.gr, GC = c.gc;
    static constexpr bool BSYNC = c.bsync;
};

constexpr int kMaxDevices = 64;

struct KernelCache { int ready_threads; int blocks_per_sm; int sms; int regs; };

template <typename K>
cudaError_t prepare(K kern, KernelCache& c, int dev, int threads, int smem) {
    if (c.ready_threads == threads) return cudaSuccess;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    int sms = 0;
    err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    cudaFuncAttributes fa;
    err = cudaFuncGetAttributes(&fa, kern);
    if (err != cudaSuccess) return err;
    c.blocks_per_sm = occ; c.sms = sms; c.regs = fa.numRegs; c.ready_threads = threads;
    return cudaSuccess;
}

template <typename T, int N, int MODE>
cudaError_t launch(void* A, int32_t* piv, long long batch, int threads, cudaStream_t stream,
                   LaunchInfo* info, int flags) {
    const int dry_run = flags & kLaunchDryRun;
    if (threads <= 0) threads = 256;
    const int warps = threads / 32;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;

    if (flags & kLaunchLuOnly) {  // factors only: the generic kernel's LU variant (not a tuned path)
        using AC = AutoCfg<T, N, MODE>;
        using GLU = Layout<T, N, AC::GR, AC::GC, MODE>;
        static KernelCache cache_lu[kMaxDevices] = {};
        auto kern = lub_invert_kernel<T, N, AC::GR, AC::GC, MODE, true>;
        const int smem_lu = GLU::HEADER_BYTES + warps * GLU::WARP_BYTES;
        KernelCache& cl = cache_lu[dev];
        err = prepare(kern, cl, dev, threads, smem_lu);
        if (err != cudaSuccess) return err;
        const long long ntiles_lu = (batch + GLU::MPW - 1) / GLU::MPW;
        long long blocks_lu = (ntiles_lu + warps - 1) / warps;
        if (blocks_lu > (long long)cl.sms * cl.blocks_per_sm) blocks_lu = (long long)cl.sms * cl.blocks_per_sm;
        if (info) {
            info->threads_per_block = threads; info->threads_per_matrix = GLU::G; info->matrices_per_block = warps * GLU::MPW;
            info->num_blocks = blocks_lu; info->dyn_smem_bytes = smem_lu; info->regs_per_thread = cl.regs;
            info->blocks_per_sm = cl.blocks_per_sm; info->kernel = "lub_invert_kernel<LUONLY>";
        }
        if (dry_run || batch == 0) return cudaSuccess;
        kern<<<(unsigned)blocks_lu, threads, smem_lu, stream>>>(static_cast<T*>(A), piv, batch);
        return cudaGetLastError();
    }

    using VC = V3Cfg<T, N, MODE>;
    // fp64 with a multi-lane layout runs the rolled-step kernel (lub_v4.cuh): 5-18 % faster there
    // (profiles/r01_tune_f64.jsonl); fp32 and one-lane-per-matrix layouts stay on lub_v3.cuh
    constexpr bool USE_V4 = (sizeof(T) == 8) && (VC::GR * VC::GC > 1);
    using FL = typename std::conditional<USE_V4, V4Layout<T, N, VC::GR, VC::GC, MODE>, V3Layout<T, N, VC::GR, VC::GC, MODE>>::type;
    using GL = Layout<T, N, AutoCfg<T, N, MODE>::GR, AutoCfg<T, N, MODE>::GC, MODE>;
    // the fast kernel's vector accesses need a 16-byte aligned batch (cudaMalloc gives 256);
    // anything else (a view starting mid-buffer at an odd element) takes the generic kernel
    const bool fast = dry_run || (reinterpret_cast<uintptr_t>(A) % 16 == 0);

    static KernelCache cache_fast[kMaxDevices] = {}, cache_gen[kMaxDevices] = {};
    int smem, mpw, g;
    KernelCache* c;
    using TC = TmaCfg<T, N, MODE>;
    constexpr bool USE_TMA = TC::ON;
    if constexpr (USE_TMA) {
        if (fast) {
            using TL = TmaLayout<T, N, TC::GR, TC::GC, MODE>;
            // no pivoting: results leave through the image and a bulk store (5-18 % faster than register
            // stores + in-place prefetch)
            auto kern = lub_tma_kernel<T, N, TC::GR, TC::GC, MODE, VC::MINB, TC::BSYNC, false, MODE == kModeNone>;
            smem = TL::smem_bytes(warps); mpw = TL::MPW; g = TL::G; c = &cache_fast[dev];
            err = prepare(kern, *c, dev, threads, smem);
            if (err != cudaSuccess) return err;
            const long long ntiles = (batch + mpw - 1) / mpw;
            long long blocks = (ntiles + warps - 1) / warps;
            const long long resident = (long long)c->sms * c->blocks_per_sm;
            if (blocks > resident) blocks = resident;
            if (info) {
                info->threads_per_block = threads; info->threads_per_matrix = g; info->matrices_per_block = warps * mpw;
                info->num_blocks = blocks; info->dyn_smem_bytes = smem; info->regs_per_thread = c->regs;
                info->blocks_per_sm = c->blocks_per_sm; info->kernel = "lub_tma_kernel";
            }
            if (dry_run || batch == 0) return cudaSuccess;
            CUtensorMap map;
            err = make_batch_tmap<T>(&map, A, N, batch, mpw);
            if (err != cudaSuccess) return err;
            kern<<<(unsigned)blocks, threads, smem, stream>>>(map, static_cast<T*>(A), piv, batch);
            return cudaGetLastError();
        }
    }
    // no pivoting on the 16-byte image: the next tile is prefetched with cp.async while this one is
    // eliminated and the results leave straight from the registers (12-22 % faster, N = 16..24,
    // profiles/r01_tune_prefetch.jsonl); no per-tile block barrier there
    // (below N = 12 the tiles are so small that the plain path wins: N = 8 0.129 -> 0.100 ms)
    constexpr bool V3_PF = !USE_V4 && (MODE == kModeNone) && V3Layout<T, N, VC::GR, VC::GC, MODE>::ROWVEC && N >= 12;
    constexpr bool V3_PFD = !USE_V4 && pick_pfd(N, (int)sizeof(T), MODE);
    constexpr bool V4_PFD = USE_V4 && pick_pfd64(N, MODE);
    if (fast) {
        mpw = FL::MPW; g = FL::G; c = &cache_fast[dev];
        if constexpr (V3_PFD) {
            smem = FL::HEADER_BYTES + warps * FL::WARP_BYTES_PFD;
            err = prepare(lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, VC::BSYNC, 0, false, true>, *c, dev, threads, smem);
        } else if constexpr (USE_V4 && V4_PFD) {
            smem = FL::HEADER_BYTES + warps * FL::WARP_BYTES_PFD;
            err = prepare(lub_v4_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false, 0, true>, *c, dev, threads, smem);
        } else if constexpr (USE_V4) {
            smem = FL::HEADER_BYTES + warps * FL::WARP_BYTES;
            err = prepare(lub_v4_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false, 0>, *c, dev, threads, smem);
        } else if constexpr (V3_PF) {
            smem = FL::HEADER_BYTES + warps * FL::WARP_BYTES_PF;
            err = prepare(lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false, 0, true>, *c, dev, threads, smem);
        } else {
            smem = FL::HEADER_BYTES + warps * FL::WARP_BYTES;
            err = prepare(lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, VC::BSYNC>, *c, dev, threads, smem);
        }
    } else {
        smem = GL::HEADER_BYTES + warps * GL::WARP_BYTES; mpw = GL::MPW; g = GL::G; c = &cache_gen[dev];
        err = prepare(lub_invert_kernel<T, N, AutoCfg<T, N, MODE>::GR, AutoCfg<T, N, MODE>::GC, MODE>, *c, dev, threads, smem);
    }
    if (err != cudaSuccess) return err;

    const long long ntiles = (batch + mpw - 1) / mpw;
    long long blocks = (ntiles + warps - 1) / warps;
    const long long resident = (long long)c->sms * c->blocks_per_sm;
    if (blocks > resident) blocks = resident;  // persistent: every warp strides over tiles
    if (info) {
        info->threads_per_block = threads;
        info->threads_per_matrix = g;
        info->matrices_per_block = warps * mpw;
        info->num_blocks = blocks;
        info->dyn_smem_bytes = smem;
        info->regs_per_thread = c->regs;
        info->blocks_per_sm = c->blocks_per_sm;
        info->kernel = !fast ? "lub_invert_kernel" : (USE_V4 ? "lub_v4_kernel" : "lub_v3_kernel");
    }
    if (dry_run || batch == 0) return cudaSuccess;
    if (fast) {
        if constexpr (V3_PFD)
            lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, VC::BSYNC, 0, false, true>
                <<<(unsigned)blocks, threads, smem, stream>>>(static_cast<T*>(A), piv, batch);
        else if constexpr (USE_V4 && V4_PFD)
            lub_v4_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false, 0, true>
                <<<(unsigned)blocks, threads, smem, stream>>>(static_cast<T*>(A), piv, batch);
        else if constexpr (USE_V4)
            lub_v4_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false, 0>
                <<<(unsigned)blocks, threads, smem, stream>>>(static_cast<T*>(A), piv, batch);
        else if constexpr (V3_PF)
            lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false, 0, true>
                <<<(unsigned)blocks, threads, smem, stream>>>(static_cast<T*>(A), piv, batch);
        else
            lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, VC::BSYNC>
                <<<(unsigned)blocks, threads, smem, stream>>>(static_cast<T*>(A), piv, batch);
    } else
        lub_invert_kernel<T, N, AutoCfg<T, N, MODE>::GR, AutoCfg<T, N, MODE>::GC, MODE>
            <<<(unsigned)blocks, threads, smem, stream>>>(static_cast<T*>(A), piv, batch);
    return cudaGetLastError();
}

// Each instantiation TU exports one of these for its (dtype, mode, N-range).
#define LUB_DEFINE_GETTER(NAME, T, MODE, N0, N1, N2, N3, N4, N5, N6, N7)                         \
    namespace lub {                                                                               \
    LaunchFn NAME(int n) {                                                                        \
        switch (n) {                                                                              \
            case N0: return &launch<T, N0, MODE>;                                                 \
            case N1: return &launch<T, N1, MODE>;                                                 \
            case N2: return &launch<T, N2, MODE>;                                                 \
            case N3: return &launch<T, N3, MODE>;                                                 \
            case N4: return &launch<T, N4, MODE>;                                                 \
            case N5: return &launch<T, N5, MODE>;                                                 \
            case N6: return &launch<T, N6, MODE>;                                                 \
            case N7: return &launch<T, N7, MODE>;                                                 \
        }                                                                                         \
        return nullptr;                                                                           \
    }                                                                                             \
    }

}  // namespace lub
