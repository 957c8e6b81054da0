// lub_launch.cuh -- per-(T, N, MODE) launch wrapper: picks the lane layout, sizes the grid
// from the SM count and the occupancy of the instantiation, opts in to dynamic shared
// memory.  Replaces the launch-geometry arithmetic of the reference's main()
// (parallel_pivot/luBatchedInplace.cu:8-11,118-127) and the NUMTHREADS table of its sweep
// driver (templated/run.py:201-223).
#pragma once
#include "lub_kernel.cuh"

namespace lub {

struct LaunchInfo {
    int threads_per_block;
    int threads_per_matrix;
    int matrices_per_block;
    long long num_blocks;
    int dyn_smem_bytes;
    int regs_per_thread;
    int blocks_per_sm;
};

// int launcher(A, piv, batch, threads (0 = default), stream, info (may be NULL), dry_run)
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

constexpr int kMaxDevices = 64;

template <typename T, int N, int MODE>
cudaError_t launch(void* A, int32_t* piv, long long batch, int threads, cudaStream_t stream,
                   LaunchInfo* info, int dry_run) {
    constexpr int GR = AutoCfg<T, N, MODE>::GR, GC = AutoCfg<T, N, MODE>::GC;
    using L = Layout<T, N, GR, GC, MODE>;
    auto kern = lub_invert_kernel<T, N, GR, GC, MODE>;

    if (threads <= 0) threads = 128;
    const int warps = threads / 32;
    const int smem = L::HEADER_BYTES + warps * L::WARP_BYTES;

    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;

    // per-device, per-thread-count cache of (attribute set, occupancy, SM count)
    struct Cache { int ready_threads; int blocks_per_sm; int sms; int regs; };
    static Cache cache[kMaxDevices] = {};
    Cache& c = cache[dev];
    if (c.ready_threads != threads) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
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
    }

    const long long ntiles = (batch + L::MPW - 1) / L::MPW;
    long long blocks = (ntiles + warps - 1) / warps;
    const long long resident = (long long)c.sms * c.blocks_per_sm;
    if (blocks > resident) blocks = resident;  // persistent: every warp strides over tiles
    if (info) {
        info->threads_per_block = threads;
        info->threads_per_matrix = L::G;
        info->matrices_per_block = warps * L::MPW;
        info->num_blocks = blocks;
        info->dyn_smem_bytes = smem;
        info->regs_per_thread = c.regs;
        info->blocks_per_sm = c.blocks_per_sm;
    }
    if (dry_run || batch == 0) return cudaSuccess;
    kern<<<(unsigned)blocks, threads, smem, stream>>>(static_cast<T*>(A), piv, batch);
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
