// lub_bulk.cuh -- the in-register Gauss-Jordan kernel (lub_v3.cuh) staged by 1-D bulk copies, for the sizes
// whose rows are NOT whole 16-byte multiples (every N that is not a multiple of 4 in fp32, odd N in fp64):
// a tensor map cannot describe them (global strides must be 16-byte multiples), so lub_v3_kernel moved
// their tiles through the LSU (LDG/LDGSTS + STS, LDS + STG: ~170 issued instructions per fp32 N = 31 matrix
// and an un-prefetched wait, profiles/r02_prof_n31_f32_parallel.md).  But a warp tile -- MPW whole
// matrices, contiguous in `T A[batch][n][n]` (templated/luBatchedInplace.cuh:89-97) -- is a contiguous
// byte span, and cp.async.bulk (UBLKCP) moves such a span with ONE instruction each way:
//
//   * load: the 16-byte aligned part of the span, [s & ~15, e & ~15), by one cp.async.bulk that completes on
//     the image's mbarrier; the image starts at buf + (s & 15) so that global and shared 16-byte chunks line
//     up.  Up to three words of a ragged end travel by 4-byte cp.async.  Bytes of the neighbouring tile that
//     the round-down drags along are never used.
//   * two images per warp: the next tile is requested right after the register load, so no warp waits
//     on HBM for its input (as the TMA kernel's DB option, lub_tma.cuh);
//   * store: results go into the image (pivot modes: the column scatter that undoes the row permutation),
//     the aligned interior leaves by one cp.async.bulk, up to three words at either end by plain stores --
//     neighbouring tiles never write each other's bytes.
//
// The image is dense (row stride N): odd N makes every column walk conflict-free, N = 2 mod 4 two-way.
//
// Pivot search for parallel pivoting with N not a power of two (prepass_rowpos below): the reference tree
// (parallel_pivot/luBatchedInplace.cuh:12-44) only merges some of its slots into slot 0, so whether a row is
// a candidate at step k depends on WHERE it sits.  lub_fast.cuh searches with lane = position and fetches
// the value through a run-time row address (16 issued instructions per step).  Here lane = ORIGINAL row,
// as in the row-wise search: its column entries come from static addresses, and the lane carries its
// row's current position as a one-hot word.
#pragma once
#include "lub_tma.cuh"
#include "lub_lapack.cuh"
#include "lub_interleaved.cuh"

namespace lub {

__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

template <typename T, int N, int GR, int GC, int MODE>
struct BulkLayout {
    static constexpr int ES = sizeof(T);
    static constexpr int EPV = 16 / ES;
    // widest vector the dense rows allow: the image inherits the alignment of global memory
    static constexpr int CH = (N % EPV == 0) ? EPV : ((EPV == 4 && N % 2 == 0) ? 2 : 1);
    static constexpr int G = GR * GC;
    static_assert(G >= 1 && G <= 32 && (32 % G) == 0, "G must divide 32");
    static constexpr int MPW = 32 / G;
    static constexpr int CPR = N / CH;
    static constexpr int CPL = cdiv_(CPR, GC);
    static constexpr int LC = CPL * CH;
    static constexpr int LR = cdiv_(N, GR);
    static constexpr int P = N, MS = N * N;
    static constexpr int SPAN_BYTES = MPW * MS * ES;
    static constexpr bool ALIGNED = (SPAN_BYTES % 16) == 0;  // every full tile starts on 16 bytes
    static constexpr int IMG_BYTES = roundup_(SPAN_BYTES, 16) + 16;
    // pivot_mode 3 keeps two vectors per matrix: the permutation the load applies, and LAPACK's ipiv for the caller
    static constexpr int PERM1_BYTES = (MODE != kModeNone) ? roundup_(MPW * N * 4, 16) : 0;
    static constexpr int PERM_BYTES = (MODE == kModeLapack) ? 2 * PERM1_BYTES : PERM1_BYTES;
    static constexpr int HEADER_BYTES = 64;
    static constexpr int warp_bytes(int nimg) { return nimg * IMG_BYTES + PERM_BYTES + 16; }
    static constexpr int smem_bytes(int warps, int nimg) { return HEADER_BYTES + warps * warp_bytes(nimg); }
};

// Row-wise pivot search in floating point (see prepass_rowwise_swz, lub_tma.cuh) on a plain image with row
// stride P: serial pivoting, and parallel pivoting when the reference tree reaches all of its slots.
template <int N, int MODE, int P, int MS, int MI>
__device__ __forceinline__ void prepass_rowwise_f32(const float* __restrict__ img0, int* __restrict__ perm0,
                                                    const int8_t* __restrict__ slot_rank, int lane) {
    const int roff = ((lane < N) ? lane : 0) * P;
    float alive[MI], when[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) { alive[m] = (lane < N) ? 1.0f : 0.0f; when[m] = 0.0f; }
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
        float v[MI], mx[MI];
#pragma unroll
        for (int m = 0; m < MI; ++m) v[m] = img0[m * MS + roff + k] * alive[m];
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
            prepass_exact<float, N, MODE, P>(img0 + m * MS, perm0 + m * N, slot_rank, lane);
        }
    }
}

// Position-aware row-wise search (parallel pivoting, N not a power of two).  Lane r owns original row r and
// carries pb = 1 << (the position row r sits at).  Step k:
//   candidates  = rows whose position is k (the seed every tree slot starts from) or k + 1 + t for a slot t
//                 the tree merges into slot 0: (pb & VM_k) != 0, VM_k a compile-time mask;
//   pick        = the candidate with the largest |A[r][k]| (un-eliminated entries): one CREDUX;
//   swap        = the picked row takes position k for good, the row that sat at k takes the picked row's
//                 position: a REDUX.MAX broadcasts it (one lane contributes), both lanes XOR their word with
//                 hp ^ (1 << k).  (REDUX.OR / .ADD / .XOR issue every 8-10 cycles on B200, REDUX.MAX / .MIN and
//                 CREDUX every 2.2: profiles/r02_search_costs.jsonl.)
// From step KH on the tree reaches every remaining slot (small remaining counts: e.g. N = 18 from step 9), positions
// no longer matter and the search continues as the plain row-wise one on the FMA pipe (fp32).
// Ties: a step with several equal maxima (or an all-zero column, where every lane "hits") needs the tree's own
// order.  Every lane records its hits; unless they add up to one per step and the final positions cover
// 0..N-1 once each, the matrix is redone by the exact search (same rule as the other fast searches).
// (A variant that hands the displaced row its next validity bit with a VOTE.ANY, so that the exchange leaves the
// critical path of a step, issues 4 more instructions per step and measured 9-40 % slower, N = 18..31, before and
// after the exchange became a CREDUX.MAX -- profiles/r02_tune_bulk.md.)
constexpr int rowpos_tail_start(int n, int mode) {  // first step from which every later step sees all rows below it
    if (mode != kModeParallel) return 0;
    const unsigned all = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
    const unsigned reach = reach_mask_of(n);
    int k0 = 0;
    for (int k = 0; k < n - 1; ++k)
        if ((((reach << (k + 1)) | (1u << k)) & all) != ((all << k) & all)) k0 = k + 1;
    return k0;
}
template <int N, int MODE> struct RowposTail { static constexpr int value = rowpos_tail_start(N, MODE); };

// The picked row and the row at position k (bitk) swap positions.  `acc` collects the position word of this lane at
// the steps it was picked at -- a plain data dependency per step (a predicated counter makes ptxas defer the
// additions and rebuild them from predicates it has to spill into a register: four instructions per step).
__device__ __forceinline__ void rowpos_swap(unsigned& pb, unsigned& acc, bool hit, unsigned bitk) {
    const unsigned mine = hit ? pb : 0u;
    const unsigned hp = __reduce_max_sync(0xffffffffu, mine);
    acc |= mine;
    // if (hit || pb == bit k) pb ^= hp ^ bit k;
    asm("{ .reg .pred q, r;\n"
        "setp.ne.s32 q, %1, 0;\n"
        "setp.eq.or.u32 r, %0, %2, q;\n"
        "@r lop3.b32 %0, %0, %3, %2, 0x96;\n}"
        : "+r"(pb) : "r"((int)hit), "r"(bitk), "r"(hp));
}

template <typename T, int N, int MODE, int P, int MS, int MI>
__device__ __forceinline__ void prepass_rowpos(const T* __restrict__ img0, int* __restrict__ perm0,
                                               const int8_t* __restrict__ slot_rank, int lane) {
    constexpr unsigned ALL = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u);
    constexpr unsigned REACH = ReachMask<N>::value;
    constexpr int K0 = RowposTail<N, MODE>::value;
    // the row-wise tail pays for its switch-over when it covers a few steps; fp64 keeps one code path
    constexpr int KH = (sizeof(T) == 4 && K0 + 4 <= N - 1) ? K0 : (N - 1);
    const int roff = ((lane < N) ? lane : 0) * P;
    unsigned pb[MI], nh[MI];  // nh: the positions this lane was picked from (one bit per hit)
#pragma unroll
    for (int m = 0; m < MI; ++m) { pb[m] = (lane < N) ? (1u << lane) : 0u; nh[m] = 0u; }
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        const unsigned vmask = ((MODE == kModeParallel) ? ((REACH << (k + 1)) | (1u << k)) : (ALL << k)) & ALL;
        // (hit is consumed where it is produced: a predicate kept across the MI chains gets spilled into a register)
        if constexpr (sizeof(T) == 4) {
            float v[MI], mx[MI];
#pragma unroll
            for (int m = 0; m < MI; ++m) v[m] = sel_t((pb[m] & vmask) != 0u, img0[m * MS + roff + k], 0.0f);
#pragma unroll
            for (int m = 0; m < MI; ++m) mx[m] = warp_max_abs(v[m]);
#pragma unroll
            for (int m = 0; m < MI; ++m) rowpos_swap(pb[m], nh[m], fabsf(v[m]) == mx[m], 1u << k);
        } else {
            uint32_t v[MI], mx[MI];  // fp64: the upper word of |x| (FpBits<double>::hi31); equal upper words are a tie
#pragma unroll
            for (int m = 0; m < MI; ++m) v[m] = ((pb[m] & vmask) != 0u) ? FpBits<T>::hi31(img0[m * MS + roff + k]) : 0u;
#pragma unroll
            for (int m = 0; m < MI; ++m) mx[m] = __reduce_max_sync(0xffffffffu, v[m]);
#pragma unroll
            for (int m = 0; m < MI; ++m) rowpos_swap(pb[m], nh[m], v[m] == mx[m], 1u << k);
        }
    }
    float alive[MI], when[MI];
    if constexpr (KH < N - 1) {
#pragma unroll
        for (int m = 0; m < MI; ++m) { alive[m] = ((pb[m] & ((ALL << KH) & ALL)) != 0u) ? 1.0f : 0.0f; when[m] = 0.0f; }
#pragma unroll
        for (int k = KH; k < N - 1; ++k) {
            float v[MI], mx[MI];
#pragma unroll
            for (int m = 0; m < MI; ++m) v[m] = (float)img0[m * MS + roff + k] * alive[m];
#pragma unroll
            for (int m = 0; m < MI; ++m) mx[m] = warp_max_abs(v[m]);
#pragma unroll
            for (int m = 0; m < MI; ++m) {
                const float h = (fabsf(v[m]) == mx[m]) ? 1.0f : 0.0f;
                when[m] = fmaf(h, (float)k, when[m]);
                alive[m] = fmaf(-h, alive[m], alive[m]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MI; ++m) {
        int pos = __ffs((int)pb[m]) - 1;  // rows picked by the position-aware steps (and, without a tail, the last row)
        if constexpr (KH < N - 1) {
            const bool late = (pb[m] & ((ALL << KH) & ALL)) != 0u;
            if (late) pos = (alive[m] != 0.0f) ? (N - 1) : (int)fminf(when[m], 31.0f);
        }
        const unsigned bit = (lane < N && pos >= 0) ? (1u << pos) : 0u;
        const unsigned cover = __reduce_or_sync(0xffffffffu, bit);
        const int hits = (int)__reduce_add_sync(0xffffffffu, (unsigned)__popc(nh[m]));
        const unsigned odd = __ballot_sync(0xffffffffu, (lane < N) && __popc(pb[m]) != 1);
        if (cover == ALL && hits == KH && odd == 0u) {  // warp-uniform
            if (lane < N) perm0[m * N + pos] = lane;
        } else {
            prepass_exact<T, N, MODE, P>(img0 + m * MS, perm0 + m * N, slot_rank, lane);
        }
    }
}

// OPT bits
constexpr int kBulkLean = 1;      // gj_eliminate_lean
constexpr int kBulkOldSearch = 2; // the position-wise search of lub_fast.cuh (for comparison)
constexpr int kBulkSingle = 4;    // one image per warp (no prefetch)
constexpr int kBulkLuOnly = 64;   // factors only: pivot_mode 3 stops after prepass_getrf; modes 0 - 2 run lu_rows_dense under the known permutation
constexpr int kBulkGetrfSingle = 128; // pivot_mode 3, fp32: one matrix at a time in the search phase (for comparison)
constexpr int kBulkLane = 256;    // modes 1 / 2 with one lane per matrix: the whole inversion in the lane's registers (invert_in_registers)
constexpr int kBulkGroupSearch = 16; // N <= 16: every lane group searches its own matrix (prepass_group) instead of warp-wide searches

template <typename T, int N, int GR, int GC, int MODE, int MINB = 1, bool BSYNC = false, int OPT = kBulkLean, int MAXT = kMaxThreads>
__global__ void __launch_bounds__(MAXT, MINB)
lub_bulk_kernel(T* __restrict__ A, int32_t* __restrict__ piv, long long batch, int32_t* __restrict__ info = nullptr) {
    using L = BulkLayout<T, N, GR, GC, MODE>;
    constexpr bool LEAN = (OPT & kBulkLean) != 0, OLDS = (OPT & kBulkOldSearch) != 0;
    constexpr int NIMG = (OPT & kBulkSingle) ? 1 : 2;
    constexpr bool LUONLY = (OPT & kBulkLuOnly) != 0;
    constexpr int G = L::G, MPW = L::MPW, LR = L::LR, LC = L::LC, CH = L::CH, CPL = L::CPL, CPR = L::CPR;
    constexpr int P = L::P, MS = L::MS, ES = L::ES;
    constexpr bool LANE3 = MODE == kModeLapack && G == 1 && !LUONLY;  // pivot_mode 3 with one lane per matrix
    // serial / parallel pivoting with one lane per matrix, the whole inversion in the lane's registers (OPT & kBulkLane)
    constexpr bool LANE12 = (OPT & kBulkLane) != 0 && G == 1 && (MODE == kModeSerial || MODE == kModeParallel) && !LUONLY;
    constexpr bool IMG_DONE = LUONLY || LANE3 || LANE12;              // the result is in the image before the Gauss-Jordan phase
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    int8_t* slot_rank = reinterpret_cast<int8_t*>(smem_raw);
    constexpr int WARP_BYTES = NIMG * L::IMG_BYTES + L::PERM_BYTES + 16;
    unsigned char* wbase = smem_raw + L::HEADER_BYTES + (size_t)warp * WARP_BYTES;
    int* perm_all = reinterpret_cast<int*>(wbase + NIMG * L::IMG_BYTES);
    int* ipiv_all = perm_all + L::PERM1_BYTES / 4;  // MODE == kModeLapack only
    unsigned long long* bar0 = reinterpret_cast<unsigned long long*>(wbase + NIMG * L::IMG_BYTES + L::PERM_BYTES);

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

    const long long ntiles = (batch + MPW - 1) / MPW;
    const long long tstride = (long long)gridDim.x * nwarps;
    // Byte offsets are taken from the 16-byte boundary at or below A: layouts with scalar rows (CH == 1, odd N) then serve any
    // element-aligned batch pointer -- a view that starts at an odd matrix index is 4, 8 or 12 bytes off -- through the same
    // head / tail handling; the layouts with vector rows need (and the launcher guarantees) a0 == 0.
    const long long a0 = (CH == 1) ? (long long)(reinterpret_cast<uintptr_t>(A) & 15) : 0;
    const unsigned char* Ab = reinterpret_cast<const unsigned char*>(A) - a0;
    const long long batch_bytes = a0 + batch * (long long)(MS * ES);

    // lane 0: request the span of `tile` into image buffer `buf`, completion on `bar` (+ its own cp.async group)
    auto request = [&](long long tile, unsigned char* buf, unsigned long long* bar) {
        const long long s = a0 + tile * (long long)L::SPAN_BYTES;
        long long e = s + L::SPAN_BYTES;
        if (e > batch_bytes) e = batch_bytes;
        const long long s16 = s & ~15ll, e16 = e & ~15ll;
        const unsigned bytes = (unsigned)(e16 - s16);
        mbar_expect_tx(bar, bytes);
        if (bytes) bulk_load(buf, Ab + s16, bytes, bar);
        for (long long b = e16; b < e; b += 4) cp_async4(buf + (b - s16), Ab + b);  // ragged end: at most three words
        cp_async_commit();
    };

    // The matrices past the end of a partial last tile are never loaded, but their pivot search runs along with the others (its
    // result is discarded).  On uninitialised shared memory two lanes can then claim one position of the scratch permutation -- a
    // benign write-write race that compute-sanitizer reports (profiles/r02_sanitizer.md).  The images are therefore zeroed ONCE,
    // here: a later partial tile finds the words of an earlier, complete tile.  (Zeroing inside the tile loop instead cost 3-7 %
    // at fp32 N = 13..24: profiles/r02_tune_bulk.md section 7.)
    for (int x = lane; x < NIMG * L::IMG_BYTES / 4; x += 32) reinterpret_cast<unsigned*>(wbase)[x] = 0u;
    fence_proxy_async();
    __syncwarp();
    unsigned iter = 0;
    if (NIMG == 2 && lane == 0) {
        const long long t0 = (long long)blockIdx.x * nwarps + warp;
        if (t0 < ntiles) request(t0, wbase, bar0);
    }
#pragma unroll 1
    for (long long tbase = (long long)blockIdx.x * nwarps; tbase < ntiles; tbase += tstride) {
        if (BSYNC) __syncthreads();
        const long long tile = tbase + warp;
        if (tile >= ntiles) continue;
        const long long first = tile * MPW;
        const int nm = (batch - first < MPW) ? (int)(batch - first) : MPW;
        const long long s = a0 + tile * (long long)L::SPAN_BYTES;
        const long long e = s + (long long)nm * (MS * ES);
        const unsigned mis = (L::ALIGNED && CH != 1) ? 0u : (unsigned)(s & 15);

        const unsigned cur = (NIMG == 2) ? (iter & 1u) : 0u;
        unsigned char* buf = wbase + cur * L::IMG_BYTES;
        unsigned long long* bar = bar0 + cur;
        const unsigned parity = (NIMG == 2) ? ((iter >> 1) & 1u) : (iter & 1u);
        ++iter;
        if (NIMG == 1 && lane == 0) {
            tma_store_wait_read();  // last round's tile has left the image
            request(tile, buf, bar);
        }
        if (lane == 0) cp_async_wait<0>();
        mbar_wait(bar, parity);
        __syncwarp();

        T* img = reinterpret_cast<T*>(buf + mis);
        T* mimg = img + ml * MS;
        int* perm = perm_all + ml * N;
        if constexpr (LANE12) {
            // invert_in_registers (lub_interleaved.cuh): search on the un-eliminated column, row interchanges by conditional swaps,
            // the arithmetic of gj_eliminate operation for operation -- bitwise the results of the other matrix-major kernels
            T a[N][N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int q = 0; q < CPR; ++q) ld_vec<T, CH>(mimg + i * P + q * CH, &a[i][q * CH]);
            }
            int pv[N];
            invert_in_registers<T, N, MODE, (MODE == kModeSerial && N == 8)>(a, pv);  // tournament argmax where measured to win
#pragma unroll
            for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int q = 0; q < CPR; ++q) st_vec<T, CH>(mimg + i * P + q * CH, &a[i][q * CH]);
            }
#pragma unroll
            for (int k = 0; k < N; ++k) perm_all[ml * N + k] = pv[k];
            __syncwarp();
        } else if constexpr (LANE3) {
            // pivot_mode 3, one lane per matrix (N <= 8 fp32, N <= 6 fp64): the whole inversion in the lane's own registers --
            // invert_in_registers (lub_interleaved.cuh), the arithmetic of lub_lapack_kernel operation for operation, so the
            // results are bitwise those of the lane = row kernel and of the batch-interleaved layout
            T a[N][N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int q = 0; q < CPR; ++q) ld_vec<T, CH>(mimg + i * P + q * CH, &a[i][q * CH]);
            }
            int pv[N];
            const int fz = invert_in_registers<T, N, kModeLapack>(a, pv);
#pragma unroll
            for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int q = 0; q < CPR; ++q) st_vec<T, CH>(mimg + i * P + q * CH, &a[i][q * CH]);
            }
#pragma unroll
            for (int k = 0; k < N; ++k) ipiv_all[ml * N + k] = pv[k];
            if (ml < nm && info != nullptr) info[first + ml] = fz;
            __syncwarp();
        } else if constexpr (MODE == kModeLapack && G == 1) {
            // pivot_mode 3, factors only, one lane per matrix: getf2 in the lane's own registers -- isamax (first maximum), row
            // interchange by conditional swaps, reciprocal scaling, fma(-l, r, a): the recurrence of lub_lapack_kernel<LUONLY>
            using UB = typename FpBits<T>::U;
            T a[N][N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int q = 0; q < CPR; ++q) ld_vec<T, CH>(mimg + i * P + q * CH, &a[i][q * CH]);
            }
            int fz = 0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                UB best;
                const int p = argmax_first<T, N, false>(a, k, best);
                if (best == UB(0) && fz == 0) fz = k + 1;
                ipiv_all[ml * N + k] = p + 1;
#pragma unroll
                for (int i = k + 1; i < N; ++i) {
                    const bool c = (p == i);
#pragma unroll
                    for (int j = 0; j < N; ++j) cswap(c, a[k][j], a[i][j]);
                }
                const T rinv = T(1) / a[k][k];
#pragma unroll
                for (int i = k + 1; i < N; ++i) {
                    const T l = a[i][k] * rinv;
                    a[i][k] = l;
#pragma unroll
                    for (int j = k + 1; j < N; ++j) a[i][j] = fma(-l, a[k][j], a[i][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int q = 0; q < CPR; ++q) st_vec<T, CH>(mimg + i * P + q * CH, &a[i][q * CH]);
            }
            if (ml < nm && info != nullptr) info[first + ml] = fz;
            __syncwarp();
        } else if constexpr (MODE == kModeLapack) {
            // true partial pivoting: the permutation comes out of an LU factorisation of the staged matrix (prepass_getrf)
            if constexpr (sizeof(T) == 4 && MPW >= 2 && N <= 20 && (OPT & kBulkGetrfSingle) == 0) {
                // fp32, N <= 20: two matrices of the tile at a time, their steps interleaved (getrf_core_x2): -5..-10 %; from
                // N = 24 on the phase is bound by throughput, not by its dependent chain, and the doubled code costs 3 %
                // (profiles/r02_mode3_twophase.md)
#pragma unroll 1
                for (int m = 0; m < MPW; m += 2) {
                    int fzA, fzB;
                    prepass_getrf_x2<N, P, MS, LUONLY>(img + m * MS, perm_all + m * N, ipiv_all + m * N, lane, fzA, fzB);
                    if (lane == 0 && info != nullptr) {
                        if (m < nm) info[first + m] = fzA;
                        if (m + 1 < nm) info[first + m + 1] = fzB;
                    }
                }
            } else {
#pragma unroll 1
                for (int m = 0; m < MPW; ++m) {
                    const int fz = prepass_getrf<T, N, P, LUONLY>(img + m * MS, perm_all + m * N, ipiv_all + m * N, lane);
                    if (lane == 0 && m < nm && info != nullptr) info[first + m] = fz;
                }
            }
            __syncwarp();
        } else if (MODE != kModeNone && N <= 16 && (OPT & kBulkGroupSearch) != 0) {
            prepass_group<T, N, G, MODE, P>(mimg, perm, slot_rank, g);
            __syncwarp();
        } else if (MODE != kModeNone) {
            constexpr int MI = (MPW < 4) ? MPW : 4;
#pragma unroll 1
            for (int m = 0; m < MPW; m += MI) {
                if constexpr (OLDS || (sizeof(T) == 8 && RowwiseOk<N, MODE>::value))
                    prepass_warp<T, N, MODE, P, MS, MI, false, false>(img + m * MS, perm_all + m * N, slot_rank, lane);
                else if constexpr (RowwiseOk<N, MODE>::value)
                    prepass_rowwise_f32<N, MODE, P, MS, MI>(img + m * MS, perm_all + m * N, slot_rank, lane);
                else
                    prepass_rowpos<T, N, MODE, P, MS, MI>(img + m * MS, perm_all + m * N, slot_rank, lane);
            }
            __syncwarp();
        }

        if constexpr (LUONLY && MODE != kModeLapack && G == 1) {
            // factors only, one lane per matrix (N <= 8 fp32, N <= 6 fp64): the lane loads its matrix with the rows permuted,
            // factorises it in its own registers -- no exchange at all -- and writes it back in place
            static_assert(LR == N && LC == N, "one lane holds the whole matrix");
            T a[N][N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const T* rowp = mimg + ((MODE != kModeNone) ? perm[i] : i) * P;
#pragma unroll
                for (int q = 0; q < CPR; ++q) ld_vec<T, CH>(rowp + q * CH, &a[i][q * CH]);
            }
#pragma unroll
            for (int k = 0; k < N - 1; ++k) {
                const T rinv = T(1) / a[k][k];
#pragma unroll
                for (int i = k + 1; i < N; ++i) {
                    const T l = a[i][k] * rinv;
                    a[i][k] = l;
#pragma unroll
                    for (int j = k + 1; j < N; ++j) a[i][j] = fma(-l, a[k][j], a[i][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int q = 0; q < CPR; ++q) st_vec<T, CH>(mimg + i * P + q * CH, &a[i][q * CH]);
            }
        } else if constexpr (LUONLY && MODE != kModeLapack) {
            // factors only, modes 0 - 2: the permutation is known; LU without a search, lane = row position (lu_rows_dense)
#pragma unroll 1
            for (int m = 0; m < MPW; ++m) lu_rows_dense<T, N, P>(img + m * MS, (MODE != kModeNone) ? perm_all + m * N : nullptr, lane);
        }
        if constexpr (IMG_DONE && NIMG == 2) {  // the image already holds the result; fetch the next tile
            __syncwarp();
            const long long nxt = tile + tstride;
            if (lane == 0 && nxt < ntiles) {
                tma_store_wait_read();
                request(nxt, wbase + (cur ^ 1u) * L::IMG_BYTES, bar0 + (cur ^ 1u));
            }
        }
        if constexpr (!IMG_DONE) {
        // ---- registers <- image: rows permuted, LR x LC block per lane ---------------------
        T a[LR][LC];
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
            int prow = rok ? i : 0;
            if (MODE != kModeNone) prow = rok ? perm[i] : 0;
            const T* rowp = mimg + prow * P;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int cq = gc * CPL + q;
                const bool ok = rok && ((GC * CPL <= CPR) || (cq < CPR));
                if (ok) {
                    ld_vec<T, CH>(rowp + cq * CH, &a[li][q * CH]);
                } else {
#pragma unroll
                    for (int w = 0; w < CH; ++w) a[li][q * CH + w] = T(0);
                }
            }
        }

        if (NIMG == 2) {  // the other image: last round's tile left it through a bulk store issued a whole search ago
            const long long nxt = tile + tstride;
            if (lane == 0 && nxt < ntiles) {
                tma_store_wait_read();
                request(nxt, wbase + (cur ^ 1u) * L::IMG_BYTES, bar0 + (cur ^ 1u));
            }
        }

        T dinv[LR];
#pragma unroll
        for (int li = 0; li < LR; ++li) dinv[li] = T(0);
        if (LEAN) gj_eliminate_lean<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, gr, gc, grp_base);
        else gj_eliminate<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, gr, gc, grp_base);

        // ---- scale by 1/pivot; results into the image (pivot modes: column scatter = un-permute) ----
#pragma unroll
        for (int li = 0; li < LR; ++li) {
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) a[li][lj] *= dinv[li];
        }
        if (MODE == kModeNone) {
            __syncwarp();  // every lane has its block: the image may be overwritten
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int cq = gc * CPL + q;
                    if (rok && ((GC * CPL <= CPR) || (cq < CPR))) st_vec<T, CH>(mimg + i * P + cq * CH, &a[li][q * CH]);
                }
            }
        } else {
            int pcol[LC];
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) {
                const int j = gc * LC + lj;
                pcol[lj] = ((GC * LC <= N) || (j < N)) ? perm[j] : -1;
            }
            __syncwarp();  // all lanes hold their blocks and columns: the image may be overwritten
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
                for (int lj = 0; lj < LC; ++lj)
                    if (rok && ((GC * LC <= N) || (pcol[lj] >= 0))) mimg[i * P + pcol[lj]] = a[li][lj];
            }
        }
        }  // !IMG_DONE
        fence_proxy_async();  // generic-proxy writes -> visible to the bulk-copy unit
        __syncwarp();
        {
            // aligned interior by one bulk store; up to three words at either end by plain stores
            const long long s16u = (s + 15) & ~15ll, e16 = e & ~15ll;
            unsigned char* gdst = reinterpret_cast<unsigned char*>(A) - a0;
            if (lane == 0) {
                if (e16 > s16u) bulk_store(gdst + s16u, buf + mis + (s16u - s), (unsigned)(e16 - s16u));
                tma_store_commit();
            }
            const int hw = (int)(s16u - s) >> 2, tw = (int)(e - e16) >> 2;  // head / tail words (0..3)
            if ((!L::ALIGNED || CH == 1) && lane < hw)
                *reinterpret_cast<unsigned*>(gdst + s + 4 * lane) = *reinterpret_cast<const unsigned*>(buf + mis + 4 * lane);
            if (lane >= 4 && lane < 4 + tw)
                *reinterpret_cast<unsigned*>(gdst + e16 + 4 * (lane - 4)) = *reinterpret_cast<const unsigned*>(buf + mis + (e16 - s) + 4 * (lane - 4));
        }
        int32_t* pivp = piv;
        asm volatile("" : "+l"(pivp));  // opaque: no second copy of the tile loop for piv == NULL
        if (pivp != nullptr) {
            int32_t* pdst = pivp + first * N;
            for (int x = lane; x < nm * N; x += 32)
                pdst[x] = (MODE == kModeLapack) ? ipiv_all[x] : ((MODE != kModeNone) ? perm_all[x] : (x % N));
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete
}

}  // namespace lub
