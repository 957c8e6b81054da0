// lub_tma.cuh -- TMA-staged variant of the in-register Gauss-Jordan kernel (lub_v3.cuh) for the
// sizes whose rows are whole 128-byte lines: N = 32 fp32 (the headline configuration of
// parallel_pivot/luBatchedInplace.cu) and N = 16 fp64 (one line per row), N = 32 fp64 (two lines
// per row, BASELINE config 5).
//
// Why: lub_v3_kernel is bound by the LSU pipe (shared-memory wavefronts + shuffles + LDG/STG,
// profiles/r01_prof_headline_*.md: ~616 wavefronts per matrix, 66 % of the pipe's peak).  About a
// quarter of those wavefronts only move the tile between HBM and shared memory.  Here that part
// is done by the TMA unit instead: one lane issues cp.async.bulk.tensor (3-D box: N x N x MPW
// matrices) into a 128-byte-swizzled image and, for the pivoting modes, one bulk tensor store
// writes the finished tile back.  The swizzle (16-byte chunk index ^ row % 8) makes both the
// row-wise pivot search (lane = row, LDS.128 along the row) and the permuted register load free
// of systematic bank conflicts without any padding, which is what TMA needs (a dense box).
// Without pivoting the results leave straight from the registers: every lane owns whole
// 32-byte sectors of its rows.
#pragma once
#include <cuda.h>
#include "lub_v3.cuh"
#include "lub_lapack.cuh"

namespace lub {

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LUB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LUB_DONE_%=;\n"
        "bra LUB_WAIT_%=;\n"
        "LUB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// one warp tile (MPW whole matrices starting at matrix `first`); LPR = 128-byte lines per row
template <int LPR>
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, void* bar, int first) {
    if (LPR == 1) tma_load_3d(dst, map, bar, 0, 0, first);
    else tma_load_4d(dst, map, bar, 0, 0, 0, first);
}
template <int LPR>
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, const void* src, int first) {
    if (LPR == 1) tma_store_3d(map, src, 0, 0, first);
    else tma_store_4d(map, src, 0, 0, 0, first);
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- layout -----------------------------------------------------------------------------------
template <typename T, int N, int GR, int GC, int MODE>
struct TmaLayout {
    static constexpr int ES = sizeof(T);
    static constexpr int EPV = 16 / ES;
    static constexpr int ROWB = N * ES;                    // row bytes in global memory
    static constexpr int RB = (ROWB + 127) / 128 * 128;    // row pitch of the image: whole 128-byte swizzle lines
    static constexpr bool PADDED = RB != ROWB;             // N < 32 fp32: the TMA unit zero-fills the rest of the line
    static_assert(ROWB % 16 == 0, "TMA needs 16-byte multiples as global strides");
    static_assert(!PADDED || RB == 128, "padding is only done up to one line");
    static constexpr int LPR = RB / 128;  // lines per row
    static constexpr int G = GR * GC;
    static_assert(G >= 1 && G <= 32 && (32 % G) == 0, "G must divide 32");
    static constexpr int MPW = 32 / G;
    static constexpr int CH = EPV;
    static constexpr int CPR = N / CH;  // 16-byte chunks per row that hold data
    static constexpr int CPL = (CPR + GC - 1) / GC;  // chunks per lane; a lane may also hold zero-filled padding chunks
    static_assert(GC * CPL * 16 <= RB, "lanes must stay inside the image row");
    static constexpr int LC = CPL * CH;
    static constexpr int LR = (N + GR - 1) / GR;
    static constexpr int MAT_BYTES = N * RB;
    static constexpr int IMG_BYTES = MPW * MAT_BYTES;
    static_assert(IMG_BYTES % 1024 == 0, "swizzle atoms are 1 KB");
    // pivot_mode 3 keeps two vectors per matrix: the permutation the load applies, and LAPACK's ipiv for the caller
    static constexpr int PERM1_BYTES = (MODE != kModeNone) ? MPW * N * 4 : 0;
    static constexpr int PERM_BYTES = (MODE == kModeLapack) ? 2 * PERM1_BYTES : PERM1_BYTES;
    static constexpr int HEADER_BYTES = 64;  // slot ranks of the reference tree (exact tie-break path)
    static constexpr int smem_bytes(int warps, int nimg = 1) {  // + 1 KB slack to align the images by hand
        return 1024 + warps * (nimg * IMG_BYTES + PERM_BYTES) + warps * 16 + HEADER_BYTES;
    }
};

// byte offset, inside the swizzled image of a TILE, of byte b (< RB) of tile row `row` (= matrix in the
// tile * N + row in the matrix): the 16-byte chunk index within a 128-byte line is XORed with the line
// index mod 8 (tile images are 1 KB aligned; a matrix inside a tile need not be)
template <int RB>
__device__ __forceinline__ int swz_byte(int row, int b) {
    if (RB == 128) return row * 128 + (b ^ ((row & 7) << 4));
    const int o = row * RB + b;
    return o ^ (((o >> 7) & 7) << 4);
}
// byte offset of element (row, col)
template <int RB, int ES>
__device__ __forceinline__ int swz_off(int row, int col) { return swz_byte<RB>(row, col * ES); }

// Exact warp-wide pivot search on the swizzled image (explicit tree priorities): the rare path for
// matrices with equal |values| in one column.  Same search as prepass_group (lub_fast.cuh).
// `img` is the tile image, `row0` the tile row at which this matrix starts.
template <typename T, int N, int MODE>
__device__ __noinline__ void prepass_exact_swz(const unsigned char* img, int row0, int* perm, const int8_t* slot_rank, int lane) {
    using U = typename FpBits<T>::U;
    constexpr int ES = sizeof(T), RB = (N * ES + 127) / 128 * 128;
    for (int i = lane; i < N; i += 32) perm[i] = i;
    __syncwarp();
    for (int k = 0; k < N - 1; ++k) {
        U best_v = FpBits<T>::absbits(*reinterpret_cast<const T*>(img + swz_off<RB, ES>(row0 + perm[k], k)));
        unsigned best_p = 0;
        for (int t = lane; t < N - 1 - k; t += 32) {
            int pr;
            if (MODE == kModeParallel) {
                pr = slot_rank[t];
                if (pr < 0) continue;
            } else {
                pr = t;
            }
            const U v = FpBits<T>::absbits(*reinterpret_cast<const T*>(img + swz_off<RB, ES>(row0 + perm[k + 1 + t], k)));
            const unsigned p = ((unsigned)(pr + 1) << 8) | (unsigned)(t + 1);
            if (v > best_v || (v == best_v && p < best_p)) { best_v = v; best_p = p; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const U ov = __shfl_xor_sync(0xffffffffu, best_v, off);
            const unsigned op = __shfl_xor_sync(0xffffffffu, best_p, off);
            if (ov > best_v || (ov == best_v && op < best_p)) { best_v = ov; best_p = op; }
        }
        __syncwarp();
        if (lane == 0 && best_p != 0) {
            const int p = k + (int)(best_p & 0xffu);
            const int tmp = perm[k];
            perm[k] = perm[p];
            perm[p] = tmp;
        }
        __syncwarp();
    }
}

// warp-wide max |v| of a float in ONE instruction (CREDUX.MAXABS.F32, sm_100a); NaNs are ignored
__device__ __forceinline__ float warp_max_abs(float v) {
    float r;
    asm volatile("redux.sync.max.abs.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

#ifndef LUB_PREPASS_F32
#define LUB_PREPASS_F32 1
#endif

// Row-wise pivot search (see prepass_rowwise in lub_fast.cuh) on the swizzled image: lane = original
// row, one LDS.128 per 16-byte chunk of the row, conflict-free because of the swizzle.
//
// fp32 runs the search in floating point on the FMA pipe: on B200 every 16-lane ALU-pipe instruction
// (LOP3, SHF, ISETP, SEL) holds the scheduler's dispatch port for two cycles and an FFMA/FMUL for one
// (profiles/r01_tune_v6.md), and the integer form is five ALU instructions per step.  Here a step is
//   v = x * alive (FMUL)  ->  m = max |v| over the warp (CREDUX.MAXABS.F32)  ->  hit = (|v| == m) as 1.0/0.0 (FSET)
//   ->  when += hit * k (FFMA)  ->  alive -= hit * alive (FFMA)
// with alive in {1.0, 0.0}; every operation is exact.  The comparisons are the reference's
// (fabs(a) > fabs(b) on the un-eliminated entries, serial_pivot/luBatchedInplace.cuh:24-41).  A step whose
// maximum is shared (equal |values|, or an all-zero remaining column, where retired lanes "hit" as well)
// retires several lanes at once, the survivors then do not add up to one and the matrix is redone
// by the exact search -- the same rule as the integer form.
// `img` is the tile image, `row0` the tile row of the first of the MI matrices searched in lock step.
template <typename T, int N, int MODE, int MI>
__device__ __forceinline__ void prepass_rowwise_swz(const unsigned char* img, int row0, int* perm0, const int8_t* slot_rank, int lane) {
    using U = typename FpBits<T>::U;
    constexpr int ES = sizeof(T), EPV = 16 / ES, RB = (N * ES + 127) / 128 * 128;
    const int row = (lane < N) ? lane : 0;
    T x[MI][EPV];
    if constexpr (sizeof(T) == 4 && (LUB_PREPASS_F32 != 0)) {
        float alive[MI], when[MI];
#pragma unroll
        for (int m = 0; m < MI; ++m) { alive[m] = (lane < N) ? 1.0f : 0.0f; when[m] = 0.0f; }
#pragma unroll
        for (int k = 0; k < N - 1; ++k) {
            if ((k % EPV) == 0) {
#pragma unroll
                for (int m = 0; m < MI; ++m)
                    ld_vec<T, EPV>(reinterpret_cast<const T*>(img + swz_byte<RB>(row0 + m * N + row, (k / EPV) << 4)), x[m]);
            }
            float v[MI], mx[MI];
#pragma unroll
            for (int m = 0; m < MI; ++m) v[m] = x[m][k % EPV] * alive[m];
#pragma unroll
            for (int m = 0; m < MI; ++m) mx[m] = warp_max_abs(v[m]);
#pragma unroll
            for (int m = 0; m < MI; ++m) {
                const float hit = (fabsf(v[m]) == mx[m]) ? 1.0f : 0.0f;
                when[m] = fmaf(hit, (float)k, when[m]);
                alive[m] = fmaf(-hit, alive[m], alive[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            const bool ok = __popc(__ballot_sync(0xffffffffu, alive[m] != 0.0f)) == 1;  // warp-uniform
            // (the index is clamped: with NaN inputs a retired row can "hit" again on an all-zero step and run past N)
            const int pos = (alive[m] != 0.0f) ? (N - 1) : (int)fminf(when[m], (float)(N - 1));
            if (ok) {
                if (lane < N) perm0[m * N + pos] = lane;
            } else {
                prepass_exact_swz<T, N, MODE>(img, row0 + m * N, perm0 + m * N, slot_rank, lane);
            }
        }
    } else {
        // integer keys: the upper word of |x| (fp64: FpBits<double>::hi31 -- one REDUX per step; equal upper words are a tie)
        uint32_t alive[MI];
        int when[MI];
#pragma unroll
        for (int m = 0; m < MI; ++m) { alive[m] = (lane < N) ? ~0u : 0u; when[m] = N - 1; }
#pragma unroll
        for (int k = 0; k < N - 1; ++k) {
            if ((k % EPV) == 0) {
#pragma unroll
                for (int m = 0; m < MI; ++m)
                    ld_vec<T, EPV>(reinterpret_cast<const T*>(img + swz_byte<RB>(row0 + m * N + row, (k / EPV) << 4)), x[m]);
            }
            uint32_t key[MI], mx[MI];
#pragma unroll
            for (int m = 0; m < MI; ++m) key[m] = ((FpBits<T>::hi31(x[m][k % EPV]) << 1) | 1u) & alive[m];
#pragma unroll
            for (int m = 0; m < MI; ++m) mx[m] = __reduce_max_sync(0xffffffffu, key[m]);
#pragma unroll
            for (int m = 0; m < MI; ++m) {
                const bool hit = key[m] == mx[m];
                when[m] = hit ? k : when[m];
                alive[m] = hit ? 0u : alive[m];
            }
        }
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            const bool ok = __popc(__ballot_sync(0xffffffffu, alive[m] != 0u)) == 1;  // warp-uniform
            if (ok) {
                if (lane < N) perm0[m * N + when[m]] = lane;
            } else {
                prepass_exact_swz<T, N, MODE>(img, row0 + m * N, perm0 + m * N, slot_rank, lane);
            }
        }
    }
}

// Position-wise pivot search (lane = row POSITION) on the swizzled image, for parallel pivoting with N not a
// power of two: the reference tree (parallel_pivot/luBatchedInplace.cuh:34-42) then never merges some
// slots into slot 0, and which rows are candidates at step k depends on where they sit.  Same search as
// the position-wise branch of prepass_warp_ptrs (lub_fast.cuh); only the image addressing differs.
template <typename T, int N, int MODE, int MI>
__device__ __forceinline__ void prepass_poswise_swz(const unsigned char* img, int row0, int* perm0, const int8_t* slot_rank, int lane) {
    using U = typename FpBits<T>::U;
    constexpr int ES = sizeof(T), RB = (N * ES + 127) / 128 * 128;
    constexpr unsigned ALL = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u);
    constexpr unsigned REACH = ReachMask<N>::value;
    int prow[MI];  // original row sitting at position `lane`
    unsigned multi[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) { prow[m] = (lane < N) ? lane : 0; multi[m] = 0u; }
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
        const unsigned vmask = ((MODE == kModeParallel) ? ((REACH << (k + 1)) | (1u << k)) : (ALL << k)) & ALL;
        const bool valid = (vmask == ((ALL << k) & ALL)) ? (lane >= k && lane < N) : (((vmask >> lane) & 1u) != 0u);
        U v[MI], mx[MI];
        unsigned bal[MI];
#pragma unroll
        for (int m = 0; m < MI; ++m)
            v[m] = FpBits<T>::absbits(*reinterpret_cast<const T*>(img + swz_off<RB, ES>(row0 + m * N + prow[m], k)));
#pragma unroll
        for (int m = 0; m < MI; ++m) mx[m] = warp_max_bits(valid ? v[m] : U(0));
#pragma unroll
        for (int m = 0; m < MI; ++m) bal[m] = __ballot_sync(0xffffffffu, valid && v[m] == mx[m]);
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            const int wl = __ffs(bal[m]) - 1;  // lowest position among the maxima; ties are redone exactly below
            if (MODE == kModeParallel) multi[m] |= bal[m] & (bal[m] - 1u);
            const int other = __shfl_sync(0xffffffffu, prow[m], lane == k ? wl : k);
            prow[m] = (lane == k || lane == wl) ? other : prow[m];
        }
    }
#pragma unroll
    for (int m = 0; m < MI; ++m)
        if (lane < N) perm0[m * N + lane] = prow[m];
    if (MODE == kModeParallel) {
#pragma unroll
        for (int m = 0; m < MI; ++m)
            if (multi[m] != 0u) {  // warp-uniform, rare (needs two equal |values| in one column)
                __syncwarp();
                prepass_exact_swz<T, N, MODE>(img, row0 + m * N, perm0 + m * N, slot_rank, lane);
            }
    }
}

// pivot_mode 3 on the swizzled image: getrf's permutation from the LU factorisation of prepass_getrf (lub_lapack.cuh), lane =
// original row.  The swizzle makes the lane = row accesses (16-byte chunk q of 32 different rows) conflict-free, which the dense
// image of lub_bulk.cuh cannot be when the row is a multiple of 32 banks (N = 32: 5.1 vs 3.2 ms at N = 31 for the factors).
template <typename T, int N, bool LU>
__device__ __forceinline__ int prepass_getrf_swz(unsigned char* __restrict__ img, int row0, int* __restrict__ perm, int* __restrict__ ipiv_s, int lane) {
    constexpr int ES = sizeof(T), EPV = 16 / ES, RB = (N * ES + 127) / 128 * 128;
    static_assert(N % EPV == 0, "whole 16-byte chunks per row");
    const bool mine = lane < N;
    const int row = mine ? lane : 0;
    T a[N];
#pragma unroll
    for (int q = 0; q < N / EPV; ++q) ld_vec<T, EPV>(reinterpret_cast<const T*>(img + swz_byte<RB>(row0 + row, q << 4)), &a[q * EPV]);
    int pos;
    const int first_zero = getrf_core<T, N, LU>(a, pos, perm, ipiv_s, lane);
    if constexpr (LU) {
        __syncwarp();  // every lane has long read its row; now the rows change places
        if (mine) {
#pragma unroll
            for (int q = 0; q < N / EPV; ++q) st_vec<T, EPV>(reinterpret_cast<T*>(img + swz_byte<RB>(row0 + pos, q << 4)), &a[q * EPV]);
        }
    }
    return first_zero;
}

// factors only, modes 0 - 2, on the swizzled image (lu_core_static, lub_lapack.cuh): lane = row position
template <typename T, int N>
__device__ __forceinline__ void lu_rows_swz(unsigned char* __restrict__ img, int row0, const int* __restrict__ perm, int lane) {
    constexpr int ES = sizeof(T), EPV = 16 / ES, RB = (N * ES + 127) / 128 * 128;
    static_assert(N % EPV == 0, "whole 16-byte chunks per row");
    const bool mine = lane < N;
    const int row = mine ? (perm != nullptr ? perm[lane] : lane) : 0;
    T a[N];
#pragma unroll
    for (int q = 0; q < N / EPV; ++q) ld_vec<T, EPV>(reinterpret_cast<const T*>(img + swz_byte<RB>(row0 + row, q << 4)), &a[q * EPV]);
    lu_core_static<T, N>(a, lane);
    __syncwarp();  // every lane has long read its row; now the rows change places
    if (mine) {
#pragma unroll
        for (int q = 0; q < N / EPV; ++q) st_vec<T, EPV>(reinterpret_cast<T*>(img + swz_byte<RB>(row0 + lane, q << 4)), &a[q * EPV]);
    }
}

// 32-byte global store (STG.256, sm_100): one full sector per lane and instruction
__device__ __forceinline__ void st_global_256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st_global_256(double* p, const double* v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

// One warp = one tile of MPW matrices; persistent over tiles.  BSYNC as in lub_v3_kernel.
// PF (no-pivot mode only): the image is free once the registers are loaded, so the next tile is
// requested right then, in place, and lands while this one is eliminated (+2.5 %).  The same idea
// for the pivot modes -- results leaving through two alternating 2 KB output slices -- measured
// 8 % SLOWER on N = 32 fp32 (profiles/r01_tune_tma.jsonl, "bs3"): the kernel is bound by issue slots
// and the LSU pipe, not by the wait for its input, so that variant is not kept.
// OUTIMG (no-pivot mode only, excludes PF): the results leave like in the pivot modes -- 16-byte vector
// stores into the swizzled image, then one bulk tensor store -- instead of straight from the registers.
// ST256 (register-store path only): 32-byte stores; needs a 32-byte aligned batch and an even number of
// chunks per lane.
// OPT (round 2): bit 0 = lean elimination step (gj_eliminate_lean); bit 1 = DB, two images per warp: the next
// tile is requested right after the register load, into the image the previous tile left through (its bulk
// store has long been read by then), so no warp waits on HBM for its input and the output image stays
// untouched until its own bulk store has drained -- the pivot modes, whose column scatter needs the image
// to the very end, get the prefetch that PF gives the no-pivot path.  MAXT = threads the kernel is compiled
// for: 16 KB per warp means 12 warps per SM (one 384-thread block, up to 168 registers per thread).
constexpr int kTmaLean = 1, kTmaDB = 2, kTmaFused = 4;
constexpr int kTmaLuOnly = 8;  // factors only: pivot_mode 3 stops after prepass_getrf; modes 0 - 2 run lu_rows_swz under the known permutation
template <typename T, int N, int GR, int GC, int MODE, int MINB = 2, bool BSYNC = true, bool PF = false, bool OUTIMG = false,
          bool ST256 = false, int OPT = 0, int MAXT = kMaxThreads>
__global__ void __launch_bounds__(MAXT, MINB)
lub_tma_kernel(const __grid_constant__ CUtensorMap tmap, T* __restrict__ A, int32_t* __restrict__ piv, long long batch,
               int32_t* __restrict__ info = nullptr) {
    static_assert(!PF || MODE == kModeNone, "in-place prefetch: the pivot modes need the image for the column scatter");
    static_assert(!OUTIMG || (MODE == kModeNone && !PF), "OUTIMG is the no-pivot output path without in-place prefetch");
    constexpr bool VIA_IMG = (MODE != kModeNone) || OUTIMG;  // results go through the image and a bulk store
    constexpr bool LEAN = (OPT & kTmaLean) != 0, DB = (OPT & kTmaDB) != 0, LUONLY = (OPT & kTmaLuOnly) != 0;
    static_assert(!DB || (VIA_IMG && !PF), "DB is the double-buffered form of the image-output path");
    constexpr int NIMG = DB ? 2 : 1;
    using L = TmaLayout<T, N, GR, GC, MODE>;
    constexpr int G = L::G, MPW = L::MPW, LR = L::LR, LC = L::LC, CH = L::CH, CPL = L::CPL;
    constexpr int RB = L::RB, ES = L::ES;
    extern __shared__ unsigned char smem_dyn[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    // carve: [images, 1 KB aligned][perm][mbarriers][slot ranks]
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* img0 = base + (size_t)warp * (NIMG * L::IMG_BYTES);
    unsigned char* img = img0;
    unsigned char* after = base + (size_t)nwarps * (NIMG * L::IMG_BYTES);
    int* perm_all = reinterpret_cast<int*>(after + (size_t)warp * L::PERM_BYTES);
    int* ipiv_all = perm_all + L::PERM1_BYTES / 4;  // MODE == kModeLapack only
    unsigned long long* bar0 = reinterpret_cast<unsigned long long*>(after + (size_t)nwarps * L::PERM_BYTES) + 2 * warp;
    unsigned long long* bar = bar0;
    int8_t* slot_rank = reinterpret_cast<int8_t*>(after + (size_t)nwarps * L::PERM_BYTES + (size_t)nwarps * 16);

    if (MODE == kModeParallel && threadIdx.x < N) slot_rank[threadIdx.x] = (int8_t)tree_slot_rank(threadIdx.x, N);
    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 1, 1); }
    // every entry of perm[] is a row index from the start: a search that NaN inputs derail may skip entries, never invent one
    if (MODE != kModeNone)
        for (int x = lane; x < MPW * N; x += 32) perm_all[x] = x % N;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int g = lane % G;
    const int ml = lane / G;
    const int gr = g / GC;
    const int gc = g % GC;
    const int grp_base = ml * G;
    unsigned parity = 0;
    unsigned iter = 0;  // DB: tiles this warp has started; image / barrier iter & 1, phase (iter >> 1) & 1

    const long long ntiles = (batch + MPW - 1) / MPW;
    const long long tstride = (long long)gridDim.x * nwarps;
    if ((PF || DB) && lane == 0) {
        const long long t0 = (long long)blockIdx.x * nwarps + warp;
        if (t0 < ntiles) {
            mbar_expect_tx(bar, (unsigned)L::IMG_BYTES);
            tma_load_tile<L::LPR>(img, &tmap, bar, (int)(t0 * MPW));
        }
    }
#pragma unroll 1
    for (long long tbase = (long long)blockIdx.x * nwarps; tbase < ntiles; tbase += (long long)gridDim.x * nwarps) {
        if (BSYNC) __syncthreads();
        const long long tile = tbase + warp;
        if (tile >= ntiles) continue;
        const long long first = tile * MPW;
        const int nm = (batch - first < MPW) ? (int)(batch - first) : MPW;
        T* gspan = A + first * (long long)(N * N);

        // ---- HBM -> swizzled image, by the TMA unit (matrices past the batch end read as zero) ----
        if (DB) {
            img = img0 + (iter & 1u) * L::IMG_BYTES;
            bar = bar0 + (iter & 1u);
            parity = (iter >> 1) & 1u;
            ++iter;
        }
        if (!PF && !DB && lane == 0) {
            if (VIA_IMG) tma_store_wait_read();  // last round's tile has left the image
            mbar_expect_tx(bar, (unsigned)L::IMG_BYTES);
            tma_load_tile<L::LPR>(img, &tmap, bar, (int)first);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;

        const int trow0 = ml * N;  // first tile row of this lane's matrix
        int* perm = perm_all + ml * N;
        if constexpr (MODE == kModeLapack) {
#pragma unroll 1
            for (int m = 0; m < MPW; ++m) {
                const int fz = prepass_getrf_swz<T, N, LUONLY>(img, m * N, perm_all + m * N, ipiv_all + m * N, lane);
                if (lane == 0 && m < nm && info != nullptr) info[first + m] = fz;
            }
            __syncwarp();
        } else if (MODE != kModeNone) {
            constexpr int MI = (MPW < 2) ? MPW : 2;
#pragma unroll 1
            for (int m = 0; m < MPW; m += MI) {
                if constexpr (RowwiseOk<N, MODE>::value)
                    prepass_rowwise_swz<T, N, MODE, MI>(img, m * N, perm_all + m * N, slot_rank, lane);
                else
                    prepass_poswise_swz<T, N, MODE, MI>(img, m * N, perm_all + m * N, slot_rank, lane);
            }
            __syncwarp();
        }

        if constexpr (LUONLY && MODE != kModeLapack) {
#pragma unroll 1
            for (int m = 0; m < MPW; ++m) lu_rows_swz<T, N>(img, m * N, (MODE != kModeNone) ? perm_all + m * N : nullptr, lane);
        }
        if constexpr (LUONLY) {  // factors only: the image already holds the result; fetch the next tile and store this one
            if (DB) {
                const long long nxt = tile + tstride;
                if (lane == 0 && nxt < ntiles) {
                    tma_store_wait_read();
                    mbar_expect_tx(bar0 + (iter & 1u), (unsigned)L::IMG_BYTES);
                    tma_load_tile<L::LPR>(img0 + (iter & 1u) * L::IMG_BYTES, &tmap, bar0 + (iter & 1u), (int)(nxt * MPW));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_tile<L::LPR>(&tmap, img, (int)first);
                tma_store_commit();
            }
        } else {
        // ---- registers <- image: rows permuted, LR x LC block per lane ---------------------
        T a[LR][LC];
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
            int prow = rok ? i : 0;
            if (MODE != kModeNone) prow = rok ? perm[i] : 0;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                if (rok) {  // chunks past the data of a padded row read the zeros the TMA unit filled in
                    ld_vec<T, CH>(reinterpret_cast<const T*>(img + swz_byte<RB>(trow0 + prow, (gc * CPL + q) << 4)), &a[li][q * CH]);
                } else {
#pragma unroll
                    for (int w = 0; w < CH; ++w) a[li][q * CH + w] = T(0);
                }
            }
        }

        if (PF) {
            __syncwarp();  // every lane has its block: the image can take the next tile
            const long long nxt = tile + tstride;
            if (lane == 0 && nxt < ntiles) {
                mbar_expect_tx(bar, (unsigned)L::IMG_BYTES);
                tma_load_tile<L::LPR>(img, &tmap, bar, (int)(nxt * MPW));
            }
        }
        if (DB) {  // the other image: last round's tile left it through a bulk store issued a whole search ago
            const long long nxt = tile + tstride;
            if (lane == 0 && nxt < ntiles) {
                tma_store_wait_read();
                mbar_expect_tx(bar0 + (iter & 1u), (unsigned)L::IMG_BYTES);
                tma_load_tile<L::LPR>(img0 + (iter & 1u) * L::IMG_BYTES, &tmap, bar0 + (iter & 1u), (int)(nxt * MPW));
            }
        }

        T dinv[LR];
#pragma unroll
        for (int li = 0; li < LR; ++li) dinv[li] = T(0);
        if (LEAN) gj_eliminate_lean<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, gr, gc, grp_base);
        else gj_eliminate<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, gr, gc, grp_base);

        // ---- scale by 1/pivot; undo the row permutation as a column scatter -------------------
#pragma unroll
        for (int li = 0; li < LR; ++li) {
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) a[li][lj] *= dinv[li];
        }
        if (OUTIMG) {
            __syncwarp();  // every lane has its block: the image may be overwritten
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
                for (int q = 0; q < CPL; ++q)
                    if (rok) st_vec<T, CH>(reinterpret_cast<T*>(img + swz_byte<RB>(trow0 + i, (gc * CPL + q) << 4)), &a[li][q * CH]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_tile<L::LPR>(&tmap, img, (int)first);
                tma_store_commit();
            }
        } else if (MODE == kModeNone) {
            T* gm = gspan + (size_t)ml * (N * N) + gc * LC;
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (ml < nm) && ((li * GR + GR - 1 < N) || (i < N));
                if constexpr (ST256) {
                    static_assert(!ST256 || (CPL % 2 == 0 && GC * CPL <= L::CPR), "32-byte stores: whole chunk pairs of real data");
#pragma unroll
                    for (int q = 0; q < CPL; q += 2)
                        if (rok) st_global_256(gm + i * N + q * CH, &a[li][q * CH]);
                } else {
#pragma unroll
                    for (int q = 0; q < CPL; ++q)
                        if (rok && ((GC * CPL <= L::CPR) || (gc * CPL + q < L::CPR))) st_vec<T, CH>(gm + i * N + q * CH, &a[li][q * CH]);
                }
            }
        } else {
            int pcb[LC];  // byte offset of the destination column inside a row
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) {
                const int j = gc * LC + lj;
                pcb[lj] = ((GC * LC <= N) || (j < N)) ? perm[j] * ES : -1;  // padding columns are not written
            }
            __syncwarp();  // all lanes hold their blocks and columns: the image may be overwritten
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
                for (int lj = 0; lj < LC; ++lj)
                    if (rok && ((GC * LC <= N) || (pcb[lj] >= 0))) *reinterpret_cast<T*>(img + swz_byte<RB>(trow0 + i, pcb[lj])) = a[li][lj];
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the TMA unit
            __syncwarp();
            if (lane == 0) {
                tma_store_tile<L::LPR>(&tmap, img, (int)first);  // matrices past the batch end are clipped
                tma_store_commit();
            }
        }
        }  // !LUONLY
        int32_t* pivp = piv;
        asm volatile("" : "+l"(pivp));  // opaque: no second copy of the tile loop for piv == NULL
        if (pivp != nullptr) {
            int32_t* pdst = pivp + first * N;
            for (int e = lane; e < nm * N; e += 32)
                pdst[e] = (MODE == kModeLapack) ? ipiv_all[e] : ((MODE != kModeNone) ? perm_all[e] : (e % N));
        }
        __syncwarp();
    }
    if (VIA_IMG && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete
}

// ---- host: tensor map over the batch viewed as [batch][N][N], box = one warp tile ----------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
        if (qres != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

template <typename T>
inline cudaError_t make_batch_tmap(CUtensorMap* map, void* A, int n, long long batch, int mpw) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    const CUtensorMapDataType dt = (sizeof(T) == 4) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const int row_bytes = n * (int)sizeof(T);
    CUresult r;
    if (row_bytes <= 128) {  // [batch][n][n], one swizzle line per row; a shorter row is zero-filled up to the line
        const cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)n * row_bytes};
        const cuuint32_t box[3] = {(cuuint32_t)(128 / sizeof(T)), (cuuint32_t)n, (cuuint32_t)mpw};
        r = enc(map, dt, 3, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {  // rows of several lines: [batch][n][lines][128 bytes]; the box is still whole matrices, rows contiguous
        const int epl = 128 / (int)sizeof(T), lpr = row_bytes / 128;
        const cuuint64_t dims[4] = {(cuuint64_t)epl, (cuuint64_t)lpr, (cuuint64_t)n, (cuuint64_t)batch};
        const cuuint64_t strides[3] = {128, (cuuint64_t)row_bytes, (cuuint64_t)n * row_bytes};
        const cuuint32_t box[4] = {(cuuint32_t)epl, (cuuint32_t)lpr, (cuuint32_t)n, (cuuint32_t)mpw};
        r = enc(map, dt, 4, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace lub
