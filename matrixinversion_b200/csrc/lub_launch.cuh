// lub_launch.cuh -- per-(T, N, MODE) launch wrapper: picks the lane layout, sizes the grid
// from the SM count and the occupancy of the instantiation, opts in to dynamic shared
// memory.  Replaces the launch-geometry arithmetic of the reference's main()
// (parallel_pivot/luBatchedInplace.cu:8-11,118-127) and the NUMTHREADS table of its sweep
// driver (templated/run.py:201-223).
#pragma once
#include <atomic>
#include <mutex>
#include <type_traits>
#include "lub_kernel.cuh"
#include "lub_fast.cuh"
#include "lub_v3.cuh"
#include "lub_v4.cuh"
#include "lub_tma.cuh"
#include "lub_dmma.cuh"
#include "lub_bulk.cuh"

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
constexpr int kLaunchDryRun = 1, kLaunchLuOnly = 2, kLaunchNoTma = 4, kLaunchNoDmma = 8, kLaunchForceDmma = 16;  // 4, 8, 16: lu_batched_set_option
// ev0 (may be NULL): recorded on the stream immediately before the kernel launch, i.e. after the one-time
// preparation and the tensor-map encode, so that the ABI's "kernel execution time" is the kernel's
using LaunchFn = cudaError_t (*)(void*, int32_t*, long long, int, cudaStream_t, LaunchInfo*, int, cudaEvent_t);

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

// Which configurations run the bulk-copy staged kernel (lub_bulk.cuh): the sizes whose rows are not 16-byte
// multiples, i.e. the ones a tensor map cannot describe.  Lane grid as pick_v3_cfg, with the vector width the
// dense image allows; two images per warp, so the block shape follows from shared memory: two (three, four for
// small N) 256-thread blocks per SM where they fit, else one 384-thread block.
constexpr Cfg pick_bulk_cfg(int n, int es, int mode = kModeParallel) {
    // fp64 without pivoting, N = 25..30: 16 lanes per matrix (7 x 7 / 8 x 8 doubles per lane, 255 registers, one 256-thread
    // block per SM) halve the shuffles per matrix: -5..-12 % against the 32-lane grid (profiles/r02_tune_bulk.md)
    if (es == 8 && mode == kModeNone && n >= 25 && n <= 30) return Cfg{4, 4};
    const int epv = 16 / es;
    const int ch = (n % epv == 0) ? epv : ((epv == 4 && n % 2 == 0) ? 2 : 1);
    const int cpr = n / ch;
    const int budget = (es == 4) ? 64 : 36;
    for (int g = 1; g <= 32; g *= 2) {
        if (g == 2) continue;
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
struct BulkChoice { bool on; int gr, gc, minb, maxt, threads, opt; };
#ifndef LUB_BULK_MIN_N
#define LUB_BULK_MIN_N 5
#endif
#ifndef LUB_BULK_F64_EXC
#define LUB_BULK_F64_EXC 1
#endif
#ifndef LUB_BULK_SMALL4
#define LUB_BULK_SMALL4 1
#endif
#ifndef LUB_BULK_PAR4
#define LUB_BULK_PAR4 1
#endif
// `lapack`: the two-phase pivot_mode 3 kernel (MODE = kModeLapack; call with mode = kModeParallel for the block shape): every
// N, two pivot vectors per matrix in shared memory.
// `force`: every N (the factors-only kernels of modes 0 - 2).
constexpr BulkChoice pick_bulk(int n, int es, int mode, bool lapack = false, bool force = false) {
    const Cfg c = pick_bulk_cfg(n, es, mode);
    // fp32: the sizes without 16-byte rows; fp64: every size but the two TMA / DMMA ones (N = 16, 32) -- there the
    // bulk-staged block layout also beats the rolled-step kernel of lub_v4.cuh (N = 31: 7.7 -> 5.8 ms without pivoting)
    // (fp64 exceptions, measured slower: profiles/r02_bulk_ab_f64_*.json)
    const bool f64_off = n == 16 || n == 32 || (LUB_BULK_F64_EXC && (n == 8 || (mode == kModeNone && (n == 13 || n == 18 || n == 20)) ||
                         (mode == kModeSerial && n == 20)));
    // fp32 parallel pivoting at N = 20, 24, 28 (other modes: TMA): the position-aware row-wise search on the dense image
    const bool f32_par4 = LUB_BULK_PAR4 && mode == kModeParallel && (n == 20 || n == 24 || n == 28);
    // fp32 N = 12, 16 in every mode (N = 12 serial: 0.43 -> 0.27 ms, N = 16: -3..-5 %; N = 8 loses: 0.10 -> 0.14 without pivoting)
    const bool f32_small4 = LUB_BULK_SMALL4 && (n == 12 || n == 16);
    // parallel pivoting with one lane per matrix, N = 6..8 (fp64: 6): the whole inversion in the lane's registers
    // (invert_in_registers: search, conditional row swaps, the same arithmetic -- bitwise equal results): N = 6 0.066 -> 0.054 ms,
    // N = 7 0.078 -> 0.076, N = 8 0.199 (lub_v3_kernel) -> 0.182, fp64 N = 6 0.102 -> 0.097 (profiles/r02_tune_lane_small_n.jsonl)
    // fp64 N = 7, serial and parallel, on the forced one-lane shape below: 0.244 / 0.235 -> 0.163 / 0.142 ms (N = 8: slower, not taken)
    const bool lane12 = !lapack && !force && ((mode == kModeParallel && ((es == 4 && n >= 6 && n <= 8) || (es == 8 && n == 6))) ||
                                              (es == 8 && n == 7 && (mode == kModeParallel || mode == kModeSerial)) ||
                                              (es == 4 && n == 8 && mode == kModeSerial));  // 0.167 -> 0.156 ms with the tournament argmax
    const bool on = lapack || force || lane12 || (n >= LUB_BULK_MIN_N && ((es == 4) ? (n % 4 != 0 || f32_par4 || f32_small4) : !f64_off));
    if (!on) return BulkChoice{false, c.gr, c.gc, 1, kMaxThreads, 256, 0};
    // fp64 N = 7, 8, pivot_mode 3 and the factors-only kernels: one lane per matrix as well (49 / 64 doubles per lane: one
    // 256-thread block per SM with 255 registers, ONE image of 32 matrices per warp) -- mode 3 N = 8: 0.67 -> 0.25 ms
    if ((lapack || force || lane12) && es == 8 && (n == 7 || n == 8))
        return BulkChoice{true, 1, 1, 1, kMaxThreads, 256, kBulkGroupSearch | kBulkSingle | (lane12 ? kBulkLane : 0)};
    const int mpw = 32 / (c.gr * c.gc);
    const int img = (mpw * n * n * es + 15) / 16 * 16 + 16;
    const int perm = ((mode != kModeNone) ? (mpw * n * 4 + 15) / 16 * 16 : 0) * (lapack ? 2 : 1);
    const int wb = 2 * img + perm + 16;
    int minb = pick_minb(n, es, mode != kModeNone);
    while (minb > 1 && minb * (64 + 8 * wb + 1024) > 233472) --minb;
    // measured (profiles/r02_tune_bulk.md): the lean elimination step pays from N = 25 on (+-1 % below, -8 % at N = 18
    // without pivoting); one lane per matrix (N <= 8) searches its own matrix, larger groups search warp-wide
    // (N = 15 serial: 0.52 vs 0.87 ms); from N = 25 on (two matrices per warp, 8 x 8 blocks) one 384-thread block per SM
    // with 168 registers beats two 256-thread blocks with 128 (N = 27: 1.53 vs 1.67 ms without pivoting, 2.36 vs 2.49
    // parallel); fp64 needs the 168 registers from N = 21 on (6 x 6 doubles per lane: 2.11 vs 2.82 ms)
    const int opt = ((es == 4 && n >= 25) ? kBulkLean : 0) | (n <= 8 ? kBulkGroupSearch : 0) | (lane12 ? kBulkLane : 0);
    const bool big = (es == 4) ? (n >= 25) : (n >= 21);
    if (es == 8 && mode == kModeNone && n >= 25 && n <= 30) return BulkChoice{true, c.gr, c.gc, 1, kMaxThreads, 256, opt};
    // (a 384-thread block must fit the 227 KB of an SM: pivot_mode 3 with 32 small matrices per warp -- two pivot vectors
    // each -- does not at fp64 N = 6)
    if ((minb == 1 || big) && 64 + 12 * wb <= 232448) return BulkChoice{true, c.gr, c.gc, 1, 384, 384, opt};
    if (minb == 1 || big) return BulkChoice{true, c.gr, c.gc, 1, kMaxThreads, 256, opt};
    return BulkChoice{true, c.gr, c.gc, minb, kMaxThreads, 256, opt};
}
template <typename T, int N, int MODE>
struct BulkCfg {
    static constexpr BulkChoice c = (MODE == kModeLapack) ? pick_bulk(N, (int)sizeof(T), kModeParallel, true) : pick_bulk(N, (int)sizeof(T), MODE);
    static constexpr bool ON = kUseTma && c.on;
    static constexpr int GR = c.gr, GC = c.gc, MINB = c.minb, MAXT = c.maxt, THREADS = c.threads, OPT = c.opt;
};

// factors only (lu_batched_factor_inplace), modes 0 - 2: the bulk-copy staged kernel at every N it is asked for
// The lane = row factorisation needs ~80 registers, not the 128-168 of the Gauss-Jordan phase, and it is bound by the latency of
// its shuffle chain: 24 warps per SM with ONE image per warp (no prefetch; the other warps cover the load) beat 12 warps
// with two (profiles/r02_tune_lu_only_occupancy.jsonl: fp32 N = 31 3.39 -> 2.62 ms, N = 24 2.01 -> 1.78; fp64 N = 31 5.89 -> 5.35
// as three 256-thread blocks).  HI_F32 also serves the factors-only form of pivot_mode 3 (fp64 there: spills, slower).
constexpr bool lu_hi_occupancy(int n, int es) { return es == 4 ? n >= 17 : n >= 21; }
template <typename T, int N, int MODE>
struct BulkLuCfg {
    static constexpr BulkChoice c = pick_bulk(N, (int)sizeof(T), MODE == kModeNone ? kModeNone : kModeParallel, false, true);
    static constexpr bool HI = lu_hi_occupancy(N, (int)sizeof(T));
    static constexpr int GR = c.gr, GC = c.gc, MINB = HI ? 1 : c.minb, MAXT = HI ? 768 : c.maxt;
    static constexpr int THREADS = HI ? (sizeof(T) == 4 ? 768 : 256) : c.threads, NIMG = (HI || (c.opt & kBulkSingle)) ? 1 : 2;
    // largest block the launcher accepts (and sizes the shared-memory opt-in for): fp64 is compiled for 768 threads only to get
    // the 80-register budget that lets three 256-thread blocks share an SM
    static constexpr int CAP = HI ? THREADS : c.maxt;
    static constexpr int OPT = c.opt | (HI ? kBulkSingle : 0);
};

// Which configurations run the TMA-staged kernel (lub_tma.cuh), on which lane grid and with or without
// the per-tile block barrier.  Measured choices (profiles/r01_tune_tma.jsonl, r01_tune_late.jsonl):
//   * rows of one 128-byte line (N = 32 fp32, N = 16 fp64): every mode, the v3 lane grid;
//   * rows of two lines (N = 32 fp64): pivot modes only -- without pivoting the rolled-step kernel
//     (lub_v4.cuh) is still ahead;
//   * fp32 N = 20, 24, 28 (rows zero-padded to one line by the TMA unit): no pivoting and serial pivoting,
//     plus parallel pivoting at N = 24; elsewhere the position-wise pivot search on the swizzled image
//     (4-way bank conflicts on its column walk) loses to the odd-stride image of lub_v3.cuh.  Smaller N
//     would need a padded image too large for two blocks per SM.
// `opt`: kTmaLean | kTmaDB (lub_tma.cuh); `maxt`: block size the kernel is compiled for; `threads`: the library's
// default block size for the configuration.  Round-2 measurements (profiles/r02_tune_headline.jsonl):
//   * N = 32 fp32, pivot modes: lean step + two images per warp, one 384-thread block per SM, no per-tile barrier:
//     2.62 -> 2.33 ms (parallel), 2.71 -> 2.51 ms (serial);
//   * N = 32 fp32 without pivoting: the same with the image output path, 256 threads: 2.35 -> 2.07 ms;
//   * N = 16 fp64, pivot modes: two images per warp, two 192-thread blocks per SM: 1.34 -> 1.23 ms;
//   * N = 24 / 28 fp32 and N = 32 fp64: no gain from either (second image costs resident warps), unchanged.
struct TmaChoice { bool on; int gr, gc; bool bsync; int opt; int maxt; int threads; };
constexpr TmaChoice pick_tma(int n, int es, int mode, Cfg v3) {
    const int rowb = n * es;
    const bool piv = mode != kModeNone;
    if (es == 4 && n == 32) return TmaChoice{true, v3.gr, v3.gc, false, kTmaLean | kTmaDB, 384, piv ? 384 : 256};
    if (es == 8 && n == 16 && piv) return TmaChoice{true, v3.gr, v3.gc, true, kTmaDB, 384, 192};
    if (rowb == 128) return TmaChoice{true, v3.gr, v3.gc, piv, 0, kMaxThreads, 256};
    if (rowb == 256) return TmaChoice{piv, v3.gr, v3.gc, true, 0, kMaxThreads, 256};
    if (es == 4 && n == 20) return TmaChoice{mode != kModeParallel, 4, 2, piv, 0, kMaxThreads, 256};
    if (es == 4 && n == 24) return TmaChoice{true, 8, 2, mode == kModeParallel, 0, kMaxThreads, 256};
    if (es == 4 && n == 28) return TmaChoice{mode != kModeParallel, 4, 4, false, 0, kMaxThreads, 256};
    return TmaChoice{false, v3.gr, v3.gc, false, 0, kMaxThreads, 256};
}
template <typename T, int N, int MODE>
struct TmaCfg {
    static constexpr TmaChoice c = pick_tma(N, (int)sizeof(T), MODE, Cfg{V3Cfg<T, N, MODE>::GR, V3Cfg<T, N, MODE>::GC});
    static constexpr bool ON = kUseTma && c.on && !pick_bulk(N, (int)sizeof(T), MODE).on;
    static constexpr int GR = c.gr, GC = c.gc;
    static constexpr bool BSYNC = c.bsync;
    static constexpr int OPT = c.opt, MAXT = c.maxt, THREADS = c.threads;
};

constexpr int kMaxDevices = 64;
constexpr int kMaxWarps = 32;

// Per-(kernel instantiation, device) launch facts, filled once and read by any number of host threads
// (thread-safe for distinct streams / devices).  The dynamic shared memory opt-in is set ONCE,
// to the size of the largest block the kernel is compiled for, so a thread that lowered the NUMTHREADS knob
// cannot shrink it under another thread's launch; occupancy is cached per block size and published with a
// release store after the entry is complete.
struct KernelCache {
    std::mutex mu;
    bool attr_set = false;
    int sms = 0, regs = 0;
    std::atomic<int> occ[kMaxWarps + 1];
    KernelCache() { for (auto& o : occ) o.store(0, std::memory_order_relaxed); }
};

template <typename K>
cudaError_t prepare(K kern, KernelCache& c, int dev, int threads, int smem, int smem_max, int* occ_out) {
    const int w = threads / 32;
    int occ = c.occ[w].load(std::memory_order_acquire);
    if (occ > 0) { *occ_out = occ; return cudaSuccess; }
    std::lock_guard<std::mutex> lk(c.mu);
    cudaError_t err;
    if (!c.attr_set) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (err != cudaSuccess) return err;
        err = cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) return err;
        cudaFuncAttributes fa;
        err = cudaFuncGetAttributes(&fa, kern);
        if (err != cudaSuccess) return err;
        c.regs = fa.numRegs;
        c.attr_set = true;
    }
    occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    c.occ[w].store(occ, std::memory_order_release);
    *occ_out = occ;
    return cudaSuccess;
}

// The tensor map of the last TMA launch of this host thread: re-encoding costs a driver call per launch, and
// sweeps / benches launch the same (pointer, batch) over and over.
struct TmapCache {
    void* A = nullptr; long long batch = -1; int n = 0, es = 0, mpw = 0, dev = -1;
    CUtensorMap map;
};
template <typename T>
inline cudaError_t cached_batch_tmap(const CUtensorMap** out, void* A, int n, long long batch, int mpw, int dev) {
    static thread_local TmapCache c;
    if (!(c.A == A && c.batch == batch && c.n == n && c.es == (int)sizeof(T) && c.mpw == mpw && c.dev == dev)) {
        c.A = nullptr;
        cudaError_t err = make_batch_tmap<T>(&c.map, A, n, batch, mpw);
        if (err != cudaSuccess) return err;
        c.A = A; c.batch = batch; c.n = n; c.es = (int)sizeof(T); c.mpw = mpw; c.dev = dev;
    }
    *out = &c.map;
    return cudaSuccess;
}

// One launch of `kern` (persistent: at most one resident grid) with the geometry report the ABI exposes.
// smem_of(warps) = dynamic shared memory for a block of that many warps; max_threads = what kern is compiled for.
// `pre` runs after the one-time preparation and before the start event (tensor-map encode), `go` launches.
struct LaunchCtx { long long batch; int threads; cudaStream_t stream; LaunchInfo* info; bool dry_run; cudaEvent_t ev0; int dev; };

template <typename K, typename SmemOf, typename Go>
cudaError_t run_kernel(K kern, KernelCache& c, const LaunchCtx& x, int max_threads, SmemOf smem_of, int mpw, int g,
                       const char* name, Go go) {
    if (x.threads > max_threads || x.threads < 32 || (x.threads % 32)) return cudaErrorInvalidConfiguration;
    const int warps = x.threads / 32;
    const int smem = smem_of(warps);
    int occ = 0;
    cudaError_t err = prepare(kern, c, x.dev, x.threads, smem, smem_of(max_threads / 32), &occ);
    if (err != cudaSuccess) return err;
    const long long ntiles = (x.batch + mpw - 1) / mpw;
    long long blocks = (ntiles + warps - 1) / warps;
    const long long resident = (long long)c.sms * occ;
    if (blocks > resident) blocks = resident;  // persistent: every warp strides over tiles
    if (x.info) {
        x.info->threads_per_block = x.threads; x.info->threads_per_matrix = g; x.info->matrices_per_block = warps * mpw;
        x.info->num_blocks = blocks; x.info->dyn_smem_bytes = smem; x.info->regs_per_thread = c.regs;
        x.info->blocks_per_sm = occ; x.info->kernel = name;
    }
    if (x.dry_run || x.batch == 0) return cudaSuccess;
    return go((unsigned)blocks, smem);
}

template <typename T, int N, int MODE>
cudaError_t launch(void* A, int32_t* piv, long long batch, int threads_req, cudaStream_t stream,
                   LaunchInfo* info, int flags, cudaEvent_t ev0) {
    const bool dry_run = (flags & kLaunchDryRun) != 0;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    T* At = static_cast<T*>(A);
    auto start = [&]() -> cudaError_t { return ev0 ? cudaEventRecord(ev0, stream) : cudaSuccess; };
    LaunchCtx x{batch, threads_req > 0 ? threads_req : 256, stream, info, dry_run, ev0, dev};

    if (flags & kLaunchLuOnly) {
        // Factors only.  N >= 9: the permutation is known before any arithmetic (pre-pass of the inverse kernels), so the
        // factors are an LU factorisation without a search in the lane = row-position layout (lu_core_static, lub_lapack.cuh)
        // on the staged image -- N = 32 fp32 parallel: 9.7 -> 3.0 ms, against 2.35 ms for the inverse
        // (profiles/r02_lu_only.md).  Rows of whole 128-byte lines on the swizzled TMA image (no bank conflicts in
        // the lane = row accesses), everything else on the bulk-copy image.  Smaller N, unaligned batches and
        // LUB_OPT_STAGING = 1: the generic kernel's LU variant.
        // N <= 8 where one lane holds a whole matrix (fp32 N <= 8, fp64 N <= 6): the same kernel, the lane factorises its matrix
        // in its own registers (N = 8 fp32: 0.50 -> 0.1 ms).
        constexpr bool kLuOneLane = BulkLuCfg<T, N, MODE>::GR * BulkLuCfg<T, N, MODE>::GC == 1 && N >= 2;
        if constexpr ((N >= 9 || kLuOneLane) && kUseTma) {
            const bool aligned = dry_run || reinterpret_cast<uintptr_t>(A) % 16 == 0;
            if (aligned && !(flags & kLaunchNoTma)) {
                using TCp = TmaCfg<T, N, kModeParallel>;
                if constexpr ((N * sizeof(T)) % 128 == 0 && TCp::ON) {
                    if (batch <= 0x7fffff00ll) {
                        using VCp = V3Cfg<T, N, kModeParallel>;
                        using TL = TmaLayout<T, N, TCp::GR, TCp::GC, MODE>;
                        constexpr bool HI = sizeof(T) == 4;  // one image per warp, 24 warps per SM (see BulkLuCfg)
                        constexpr int TOPT = HI ? (TCp::OPT & ~kTmaDB) : TCp::OPT, TMAXT = HI ? 768 : TCp::MAXT;
                        constexpr int NIMG = (TOPT & kTmaDB) ? 2 : 1;
                        static KernelCache cache_tlu[kMaxDevices];
                        auto kern = lub_tma_kernel<T, N, TCp::GR, TCp::GC, MODE, (TMAXT > kMaxThreads ? 1 : VCp::MINB), TCp::BSYNC, false,
                                                   MODE == kModeNone, false, TOPT | kTmaLuOnly, TMAXT>;
                        if (threads_req <= 0) x.threads = HI ? 768 : TCp::THREADS;
                        return run_kernel(kern, cache_tlu[dev], x, TMAXT, [](int w) { return TL::smem_bytes(w, NIMG); }, TL::MPW, TL::G,
                                          "lub_tma_kernel<LUONLY>", [&](unsigned blocks, int smem) {
                                              const CUtensorMap* map = nullptr;
                                              cudaError_t e = cached_batch_tmap<T>(&map, A, N, batch, TL::MPW, dev);
                                              if (e != cudaSuccess) return e;
                                              e = start();
                                              if (e != cudaSuccess) return e;
                                              kern<<<blocks, x.threads, smem, stream>>>(*map, At, piv, batch, nullptr);
                                              return cudaGetLastError();
                                          });
                    }
                } else {
                    using BC = BulkLuCfg<T, N, MODE>;
                    using BL = BulkLayout<T, N, BC::GR, BC::GC, MODE>;
                    static KernelCache cache_blu[kMaxDevices];
                    auto kern = lub_bulk_kernel<T, N, BC::GR, BC::GC, MODE, BC::MINB, false, BC::OPT | kBulkLuOnly, BC::MAXT>;
                    if (threads_req <= 0) x.threads = BC::THREADS;
                    return run_kernel(kern, cache_blu[dev], x, BC::CAP, [](int w) { return BL::smem_bytes(w, BC::NIMG); }, BL::MPW, BL::G,
                                      "lub_bulk_kernel<LUONLY>", [&](unsigned blocks, int smem) {
                                          cudaError_t e = start();
                                          if (e != cudaSuccess) return e;
                                          kern<<<blocks, x.threads, smem, stream>>>(At, piv, batch, nullptr);
                                          return cudaGetLastError();
                                      });
                }
            }
        }
        using AC = AutoCfg<T, N, MODE>;
        using GLU = Layout<T, N, AC::GR, AC::GC, MODE>;
        static KernelCache cache_lu[kMaxDevices];
        auto kern = lub_invert_kernel<T, N, AC::GR, AC::GC, MODE, true>;
        return run_kernel(kern, cache_lu[dev], x, kMaxThreads, [](int w) { return GLU::HEADER_BYTES + w * GLU::WARP_BYTES; }, GLU::MPW, GLU::G,
                          "lub_invert_kernel<LUONLY>", [&](unsigned blocks, int smem) {
                              cudaError_t e = start();
                              if (e != cudaSuccess) return e;
                              kern<<<blocks, x.threads, smem, stream>>>(At, piv, batch);
                              return cudaGetLastError();
                          });
    }

    using VC = V3Cfg<T, N, MODE>;
    // fp64 with a multi-lane layout runs the rolled-step kernel (lub_v4.cuh): 5-18 % faster there
    // (profiles/r01_tune_f64.jsonl); fp32 and one-lane-per-matrix layouts stay on lub_v3.cuh
    constexpr bool USE_V4 = (sizeof(T) == 8) && (VC::GR * VC::GC > 1);
    using FL = typename std::conditional<USE_V4, V4Layout<T, N, VC::GR, VC::GC, MODE>, V3Layout<T, N, VC::GR, VC::GC, MODE>>::type;
    // the fast kernels' vector accesses need a 16-byte aligned batch (cudaMalloc gives 256);
    // anything else (a view starting mid-buffer at an odd element) takes the generic kernel
    const bool fast = dry_run || (reinterpret_cast<uintptr_t>(A) % 16 == 0);
    static KernelCache cache_fast[kMaxDevices], cache_gen[kMaxDevices];

    using TC = TmaCfg<T, N, MODE>;
    // fp64 N = 32: blocked Gauss-Jordan on the FP64 tensor cores (lub_dmma.cuh), four 128-thread blocks per SM: 6.45 -> 5.29 ms
    // without pivoting, 6.31 -> 5.54 ms with.  Library default: WITHOUT pivoting only -- the block step multiplies by an
    // explicit 4 x 4 inverse, which is as accurate as the unblocked elimination on the diagonally dominant matrices the
    // no-pivot mode is for, but amplifies the conditioning of the diagonal blocks under the reference's pivot rule (27 %
    // of uniform(0,1) matrices get a >= 10x larger residual, profiles/r02_dmma.md); LUB_OPT_FP64_TENSOR = 2 forces it.
    const bool no_tma = (flags & kLaunchNoTma) != 0;
    if constexpr (sizeof(T) == 8 && N == 32 && kUseTma) {
        const bool want = (MODE == kModeNone || (flags & kLaunchForceDmma)) && !(flags & kLaunchNoDmma);
        if (fast && batch <= 0x7fffff00ll && !no_tma && want) {
            using TL = TmaLayout<T, N, 8, 4, MODE>;
            auto kern = lub_dmma_kernel<MODE, 2, true>;
            if (threads_req <= 0) x.threads = 128;
            static KernelCache cache_dmma[kMaxDevices];
            return run_kernel(kern, cache_dmma[dev], x, kMaxThreads, [](int w) { return dmma_smem_bytes(w, TL::PERM_BYTES); }, TL::MPW, TL::G,
                              "lub_dmma_kernel", [&](unsigned blocks, int smem) {
                                  const CUtensorMap* map = nullptr;
                                  cudaError_t e = cached_batch_tmap<T>(&map, A, N, batch, TL::MPW, dev);
                                  if (e != cudaSuccess) return e;
                                  e = start();
                                  if (e != cudaSuccess) return e;
                                  kern<<<blocks, x.threads, smem, stream>>>(*map, At, piv, batch);
                                  return cudaGetLastError();
                              });
        }
    }
    // TMA tile coordinates are 32-bit: larger batches take the v3 / v4 kernels (64-bit index arithmetic)
    if constexpr (TC::ON) {
        if (fast && batch <= 0x7fffff00ll && !no_tma) {
            using TL = TmaLayout<T, N, TC::GR, TC::GC, MODE>;
            constexpr int NIMG = (TC::OPT & kTmaDB) ? 2 : 1;
            // no pivoting: results leave through the image and a bulk store (5-18 % faster than register
            // stores + in-place prefetch)
            auto kern = lub_tma_kernel<T, N, TC::GR, TC::GC, MODE, (TC::MAXT > kMaxThreads ? 1 : VC::MINB), TC::BSYNC, false, MODE == kModeNone,
                                       false, TC::OPT, TC::MAXT>;
            if (threads_req <= 0) x.threads = TC::THREADS;
            return run_kernel(kern, cache_fast[dev], x, TC::MAXT, [](int w) { return TL::smem_bytes(w, NIMG); }, TL::MPW, TL::G, "lub_tma_kernel",
                              [&](unsigned blocks, int smem) {
                                  const CUtensorMap* map = nullptr;
                                  cudaError_t e = cached_batch_tmap<T>(&map, A, N, batch, TL::MPW, dev);
                                  if (e != cudaSuccess) return e;
                                  e = start();
                                  if (e != cudaSuccess) return e;
                                  kern<<<blocks, x.threads, smem, stream>>>(*map, At, piv, batch, nullptr);
                                  return cudaGetLastError();
                              });
        }
    }
    // rows that are not 16-byte multiples: 1-D bulk copies instead of a tensor map (lub_bulk.cuh); any batch size
    using BC = BulkCfg<T, N, MODE>;
    if constexpr (BC::ON) {
        using BL = BulkLayout<T, N, BC::GR, BC::GC, MODE>;
        // scalar-row layouts (odd N) take any element-aligned pointer: a view starting at an odd matrix index stays on this path
        if ((fast || (BL::CH == 1 && reinterpret_cast<uintptr_t>(A) % sizeof(T) == 0)) && !no_tma) {
            auto kern = lub_bulk_kernel<T, N, BC::GR, BC::GC, MODE, BC::MINB, false, BC::OPT, BC::MAXT>;
            if (threads_req <= 0) x.threads = BC::THREADS;
            return run_kernel(kern, cache_fast[dev], x, BC::MAXT, [](int w) { return BL::smem_bytes(w, (BC::OPT & kBulkSingle) ? 1 : 2); }, BL::MPW, BL::G, "lub_bulk_kernel",
                              [&](unsigned blocks, int smem) {
                                  cudaError_t e = start();
                                  if (e != cudaSuccess) return e;
                                  kern<<<blocks, x.threads, smem, stream>>>(At, piv, batch, nullptr);
                                  return cudaGetLastError();
                              });
        }
    }
    if (!fast) {
        using GL = Layout<T, N, AutoCfg<T, N, MODE>::GR, AutoCfg<T, N, MODE>::GC, MODE>;
        auto kern = lub_invert_kernel<T, N, AutoCfg<T, N, MODE>::GR, AutoCfg<T, N, MODE>::GC, MODE>;
        return run_kernel(kern, cache_gen[dev], x, kMaxThreads, [](int w) { return GL::HEADER_BYTES + w * GL::WARP_BYTES; }, GL::MPW, GL::G,
                          "lub_invert_kernel", [&](unsigned blocks, int smem) {
                              cudaError_t e = start();
                              if (e != cudaSuccess) return e;
                              kern<<<blocks, x.threads, smem, stream>>>(At, piv, batch);
                              return cudaGetLastError();
                          });
    }
    // When the TMA path is compiled in but not taken (batch beyond 32-bit tile coordinates, LUB_OPT_STAGING) the fast cache entry
    // would be shared by two kernels: keep a second one.
    static KernelCache cache_fast2[kMaxDevices];
    KernelCache& cf = (TC::ON || BC::ON) ? cache_fast2[dev] : cache_fast[dev];
    // no pivoting on the 16-byte image: the next tile is prefetched with cp.async while this one is
    // eliminated and the results leave straight from the registers (12-22 % faster, N = 16..24,
    // profiles/r01_tune_prefetch.jsonl); no per-tile block barrier there
    // (below N = 12 the tiles are so small that the plain path wins: N = 8 0.129 -> 0.100 ms)
    constexpr bool V3_PF = !USE_V4 && (MODE == kModeNone) && V3Layout<T, N, VC::GR, VC::GC, MODE>::ROWVEC && N >= 12;
    constexpr bool V3_PFD = !USE_V4 && pick_pfd(N, (int)sizeof(T), MODE);
    constexpr bool V4_PFD = USE_V4 && pick_pfd64(N, MODE);
    // round 2 (profiles/r02_tune_v3.jsonl): the lean elimination step pays in lub_v3_kernel from N = 25 on (N = 31: -7..-9 % in
    // every mode; N <= 28: +-0); without pivoting N = 30, 31 get the double-buffered prefetch they could not afford at two
    // 256-thread blocks per SM (second image) by running one 384-thread block (N = 31: 3.14 -> 2.66 ms)
    constexpr bool V3_LEAN = !USE_V4 && sizeof(T) == 4 && N >= 25;
    constexpr bool V3_BIG = !USE_V4 && sizeof(T) == 4 && MODE == kModeNone && (N == 30 || N == 31);
    auto plain = [&](auto kern, int warp_bytes, const char* name, int max_threads = kMaxThreads) {
        return run_kernel(kern, cf, x, max_threads, [warp_bytes](int w) { return FL::HEADER_BYTES + w * warp_bytes; }, FL::MPW, FL::G, name,
                          [&](unsigned blocks, int smem) {
                              cudaError_t e = start();
                              if (e != cudaSuccess) return e;
                              kern<<<blocks, x.threads, smem, stream>>>(At, piv, batch);
                              return cudaGetLastError();
                          });
    };
    if constexpr (V3_BIG) {
        if (threads_req <= 0) x.threads = 384;
        return plain(lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, 1, false, false, true, true, 384>, FL::WARP_BYTES_PFD, "lub_v3_kernel", 384);
    } else if constexpr (V3_PFD)
        return plain(lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, VC::BSYNC, false, true, V3_LEAN>, FL::WARP_BYTES_PFD, "lub_v3_kernel");
    else if constexpr (USE_V4 && V4_PFD)
        return plain(lub_v4_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false, true>, FL::WARP_BYTES_PFD, "lub_v4_kernel");
    else if constexpr (USE_V4)
        return plain(lub_v4_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false>, FL::WARP_BYTES, "lub_v4_kernel");
    else if constexpr (V3_PF)
        return plain(lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, false, true, false, V3_LEAN>, FL::WARP_BYTES_PF, "lub_v3_kernel");
    else
        return plain(lub_v3_kernel<T, N, VC::GR, VC::GC, MODE, VC::MINB, VC::BSYNC, false, false, V3_LEAN>, FL::WARP_BYTES, "lub_v3_kernel");
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
