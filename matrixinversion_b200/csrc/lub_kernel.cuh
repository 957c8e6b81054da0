// lub_kernel.cuh -- the hot path: batched in-place inversion of N x N matrices, N <= 32.
//
// Replaces the reference's kernel family `batched_lu_subwarp`
//   templated/luBatchedInplace.cuh:78-126       (pivot_mode none)
//   serial_pivot/luBatchedInplace.cuh:104-167   (pivot_mode serial)
//   parallel_pivot/luBatchedInplace.cuh:127-199 (pivot_mode parallel)
// with a different algorithm that produces the same result (same pivot permutation
// bit-exactly, same inverse to rounding):
//
//   1. A warp stages a contiguous span of MPW = 32/G matrices global -> shared with
//      128-bit coalesced loads ("the image"; verbatim copy of global memory).
//   2. Pivot pre-pass (serial / parallel modes).  The reference factorises left-looking,
//      so when it searches column k that column still holds ORIGINAL entries (SURVEY.md
//      Q1): the whole row permutation is a function of the input only.  It is computed
//      here from the image before any arithmetic, reproducing find_pivot's "lowest row
//      wins ties" and find_pivot_parallel's shared-memory tree -- including the slots the
//      tree never merges for non-power-of-two N (Q2) -- through a static per-slot
//      (reachable, priority) table instead of log2(N) block barriers per step.
//   3. The G = GR x GC lanes of a matrix load it into registers in a 2-D cyclic layout,
//      rows already permuted (row p <- image row perm[p]), so elimination needs no swaps.
//   4. Register-resident Gauss-Jordan with deferred row scaling: N rank-1 FMA updates on
//      the lane's LR x LC block (packed FFMA2 for fp32), pivot row / column pieces moved by
//      warp shuffles.  2N^3 flops, no shared-memory traffic, no block barriers.
//   5. The block is scaled by 1/pivot, written back to the image with the column
//      permutation that undoes step 3 (A^-1 = (PA)^-1 P), and the span is stored
//      image -> global with 128-bit coalesced stores; perm goes to `piv` if requested.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lub {

constexpr int kModeNone = 0, kModeSerial = 1, kModeParallel = 2;
constexpr int kModeLapack = 3;  // true partial pivoting (getrf): lub_lapack*.cuh, lub_bulk.cuh, lub_interleaved.cuh
constexpr int kMaxThreads = 256;  // upper bound of the NUMTHREADS knob (keeps 255 registers available)

// ---- small helpers ---------------------------------------------------------------------

template <typename T> struct FpBits;
template <> struct FpBits<float> {
    using U = uint32_t;
    static __device__ __forceinline__ U absbits(float x) { return __float_as_uint(x) & 0x7fffffffu; }
    static __device__ __forceinline__ uint32_t hi31(float x) { return absbits(x); }
};
template <> struct FpBits<double> {
    using U = unsigned long long;
    static __device__ __forceinline__ U absbits(double x) {
        return (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffull;
    }
    // the upper word of |x|: orders like |x| wherever two values differ in their upper 32 bits.  The fast pivot
    // searches compare these (ONE warp reduction per step instead of two); two candidates that agree in the upper
    // word look like a tie to them, and every tie is redone by the exact 64-bit search.
    static __device__ __forceinline__ uint32_t hi31(double x) { return (uint32_t)__double2hiint(x) & 0x7fffffffu; }
};

// |v| compared through its bit pattern: identical to the reference's `fabs(a) > fabs(b)`
// for every non-NaN input (non-negative IEEE values order like unsigned integers).

__device__ __forceinline__ float rcp_t(float x) {
    // MUFU.RCP + one Newton step: <= 1 ulp, 3 instructions (the reference's sweep builds
    // with --use_fast_math, i.e. a plain approximate division, templated/run.py:47).
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    float e = fmaf(-x, r, 1.0f);
    return fmaf(r, e, r);
}
__device__ __forceinline__ double rcp_t(double x) { return 1.0 / x; }

template <typename T>
__device__ __forceinline__ T shfl_t(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    return __ldcs(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) { __stcs(reinterpret_cast<uint4*>(p), v); }

// Priority of slot t in find_pivot_parallel's tree (parallel_pivot/luBatchedInplace.cuh:34-42)
// for `tpm` slots: -1 if the slot never reaches slot 0; otherwise a rank such that among
// slots holding the same (maximal) value the LOWEST rank is the one the tree returns.
// Bit l of the rank = "moved at tree level l" (level 0 = widest stride); a slot that stays
// put wins the tie at that level, and later levels dominate earlier ones.
__host__ __device__ constexpr int tree_slot_rank(int t, int tpm) {
    int loc = t, rank = 0, lvl = 0;
    for (int s = tpm / 2; s > 0; s >>= 1, ++lvl) {
        if (loc >= 2 * s) return -1;
        if (loc >= s) { loc -= s; rank |= 1 << lvl; }
    }
    return rank;
}

// ---- layout ----------------------------------------------------------------------------

template <typename T, int N, int GR, int GC, int MODE>
struct Layout {
    static constexpr int G = GR * GC;   // lanes cooperating on one matrix
    static_assert(G >= 1 && G <= 32 && (32 % G) == 0, "G must divide 32");
    static constexpr int MPW = 32 / G;  // matrices per warp tile
    static constexpr int LR = (N + GR - 1) / GR;  // rows held by a lane (cyclic over GR)
    static constexpr int LC = (N + GC - 1) / GC;  // cols held by a lane (cyclic over GC)
    static constexpr int EPV = 16 / (int)sizeof(T);
    static constexpr int IMG_ELEMS = MPW * N * N;
    // +16: the image is shifted by (global address & 15) so that 16-byte chunks line up
    static constexpr int IMG_BYTES = ((IMG_ELEMS * (int)sizeof(T) + 15) / 16) * 16 + 16;
    static constexpr int PERM_BYTES = (MODE != kModeNone) ? ((MPW * N * 4 + 15) / 16) * 16 : 0;
    static constexpr int WARP_BYTES = IMG_BYTES + PERM_BYTES;
    static constexpr int HEADER_BYTES = 64;  // slot-rank table (parallel mode), int8[N] padded
};

// ---- span copy global <-> image ----------------------------------------------------------

template <typename T>
__device__ __forceinline__ void copy_in(T* __restrict__ img, const T* __restrict__ src, int total,
                                        int capacity, int lane) {
    constexpr int EPV = 16 / (int)sizeof(T);
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
    int nhead = mis ? (int)((16u - mis) / sizeof(T)) : 0;
    if (nhead > total) nhead = total;
    const int nvec = (total - nhead) / EPV;
    if (lane < nhead) img[lane] = src[lane];
    const T* s = src + nhead;
    T* d = img + nhead;
    int c = lane;
    for (; c + 96 < nvec; c += 128) {  // 4 independent 16-byte loads in flight per lane
        uint4 v0 = ld_stream16(s + (size_t)c * EPV);
        uint4 v1 = ld_stream16(s + (size_t)(c + 32) * EPV);
        uint4 v2 = ld_stream16(s + (size_t)(c + 64) * EPV);
        uint4 v3 = ld_stream16(s + (size_t)(c + 96) * EPV);
        *reinterpret_cast<uint4*>(d + c * EPV) = v0;
        *reinterpret_cast<uint4*>(d + (c + 32) * EPV) = v1;
        *reinterpret_cast<uint4*>(d + (c + 64) * EPV) = v2;
        *reinterpret_cast<uint4*>(d + (c + 96) * EPV) = v3;
    }
    for (; c < nvec; c += 32) *reinterpret_cast<uint4*>(d + c * EPV) = ld_stream16(s + (size_t)c * EPV);
    const int tb = nhead + nvec * EPV;
    if (lane < total - tb) img[tb + lane] = src[tb + lane];
    // matrices beyond the batch tail: zeros (their lanes run the same code, results dropped)
    for (int e = total + lane; e < capacity; e += 32) img[e] = T(0);
}

template <typename T>
__device__ __forceinline__ void copy_out(T* __restrict__ dst, const T* __restrict__ img, int total, int lane) {
    constexpr int EPV = 16 / (int)sizeof(T);
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u);
    int nhead = mis ? (int)((16u - mis) / sizeof(T)) : 0;
    if (nhead > total) nhead = total;
    const int nvec = (total - nhead) / EPV;
    if (lane < nhead) dst[lane] = img[lane];
    T* d = dst + nhead;
    const T* s = img + nhead;
    for (int c = lane; c < nvec; c += 32)
        st_stream16(d + (size_t)c * EPV, *reinterpret_cast<const uint4*>(s + c * EPV));
    const int tb = nhead + nvec * EPV;
    if (lane < total - tb) dst[tb + lane] = img[tb + lane];
}

// ---- pivot pre-pass -----------------------------------------------------------------------
//
// Generic version: the G lanes of a matrix share the scan of column k.  Slot t (the
// reference's thread t, holding row k+1+t) is examined by lane t % G; candidates are
// ordered by (|value| descending, priority ascending) where priority 0 is the seed
// (|A[k][k]|, k) that every reference slot starts from, so a strictly larger value is
// required to move away from row k -- exactly `val > thread_max_val` / `vals[t] < vals[t+s]`.
//   serial   : priority(t) = t + 1           (lowest row wins, find_pivot :22-36)
//   parallel : priority(t) = tree rank + 1, unreachable slots skipped (find_pivot_parallel :12-44)

template <typename T, int N, int G, int MODE>
__device__ __forceinline__ void pivot_prepass(const T* __restrict__ mimg, int* __restrict__ perm,
                                              const int8_t* __restrict__ slot_rank, int g) {
    using U = typename FpBits<T>::U;
    for (int i = g; i < N; i += G) perm[i] = i;
    __syncwarp();
    for (int k = 0; k < N - 1; ++k) {
        U best_v = FpBits<T>::absbits(mimg[perm[k] * N + k]);
        unsigned best_p = 0;  // (priority << 8) | (t + 1); 0 = stay on row k
        for (int t = g; t < N - 1 - k; t += G) {
            int pr;
            if (MODE == kModeParallel) {
                pr = slot_rank[t];
                if (pr < 0) continue;
            } else {
                pr = t;
            }
            const U v = FpBits<T>::absbits(mimg[perm[k + 1 + t] * N + k]);
            const unsigned p = ((unsigned)(pr + 1) << 8) | (unsigned)(t + 1);
            if (v > best_v || (v == best_v && p < best_p)) { best_v = v; best_p = p; }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            const U ov = __shfl_xor_sync(0xffffffffu, best_v, off);
            const unsigned op = __shfl_xor_sync(0xffffffffu, best_p, off);
            if (ov > best_v || (ov == best_v && op < best_p)) { best_v = ov; best_p = op; }
        }
        __syncwarp();  // every lane has read perm[] for this step
        if (g == 0 && best_p != 0) {
            const int p = k + (int)(best_p & 0xffu);
            const int tmp = perm[k];
            perm[k] = perm[p];
            perm[p] = tmp;
        }
        __syncwarp();
    }
}

// ---- rank-1 update of one local row -------------------------------------------------------

template <int LC>
__device__ __forceinline__ void row_update(float (&a)[LC], const float (&r)[LC], float nf) {
    // a[j] += nf * r[j]; packed FFMA2 (two fp32 FMAs per issue slot on sm_100)
    const float2 nf2 = make_float2(nf, nf);
#pragma unroll
    for (int j = 0; j + 1 < LC; j += 2) {
        const float2 d = __ffma2_rn(nf2, make_float2(r[j], r[j + 1]), make_float2(a[j], a[j + 1]));
        a[j] = d.x;
        a[j + 1] = d.y;
    }
    if (LC & 1) a[LC - 1] = fmaf(nf, r[LC - 1], a[LC - 1]);
}
template <int LC>
__device__ __forceinline__ void row_update(double (&a)[LC], const double (&r)[LC], double nf) {
#pragma unroll
    for (int j = 0; j < LC; ++j) a[j] = fma(nf, r[j], a[j]);
}

// ---- the kernel ---------------------------------------------------------------------------

// LUONLY: stop after the factorisation and write the factors instead of the inverse -- the state the
// reference's shared-memory matrix is in after its k-loop (parallel_pivot/luBatchedInplace.cuh:156-186,
// before comp_inv): unit-lower L strictly below the diagonal, U on and above it, rows in pivoted order
// (row i of the output factorises input row piv[i]).  This is what verifyLU / verifyLUwithPivoting
// (templated/verify.hpp:105-186, parallel_pivot/verify.hpp:157-242) check.  Right-looking elimination
// on the same lane grid; same values as the reference's left-looking Doolittle up to rounding.
template <typename T, int N, int GR, int GC, int MODE, bool LUONLY = false>
__global__ void __launch_bounds__(kMaxThreads)
lub_invert_kernel(T* __restrict__ A, int32_t* __restrict__ piv, long long batch) {
    using L = Layout<T, N, GR, GC, MODE>;
    constexpr int G = L::G, MPW = L::MPW, LR = L::LR, LC = L::LC;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    int8_t* slot_rank = reinterpret_cast<int8_t*>(smem_raw);
    unsigned char* wbase = smem_raw + L::HEADER_BYTES + (size_t)warp * L::WARP_BYTES;
    int* perm_all = reinterpret_cast<int*>(wbase + L::IMG_BYTES);

    if (MODE == kModeParallel) {
        if (threadIdx.x < N) slot_rank[threadIdx.x] = (int8_t)tree_slot_rank(threadIdx.x, N);
        __syncthreads();
    }

    const int g = lane % G;    // lane within the matrix group
    const int ml = lane / G;   // matrix within the warp tile
    const int gr = g / GC;
    const int gc = g % GC;
    const int grp_base = ml * G;

    const long long ntiles = (batch + MPW - 1) / MPW;
    for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < ntiles;
         tile += (long long)gridDim.x * nwarps) {
        const long long first = tile * MPW;
        const int nm = (batch - first < MPW) ? (int)(batch - first) : MPW;
        const int total = nm * N * N;
        T* gspan = A + first * (long long)(N * N);
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(gspan) & 15u);
        T* img = reinterpret_cast<T*>(wbase + mis);

        copy_in<T>(img, gspan, total, L::IMG_ELEMS, lane);
        __syncwarp();

        T* mimg = img + ml * (N * N);
        int* perm = perm_all + ml * N;
        if (MODE != kModeNone) pivot_prepass<T, N, G, MODE>(mimg, perm, slot_rank, g);

        // ---- registers <- image, rows permuted -------------------------------------------
        T a[LR][LC];
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
            int prow = i;
            if (MODE != kModeNone) prow = rok ? perm[i] : 0;
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) {
                const int j = lj * GC + gc;
                const bool ok = rok && ((lj * GC + GC - 1 < N) || (j < N));
                a[li][lj] = ok ? mimg[prow * N + j] : T(0);
            }
        }

        if constexpr (LUONLY) {
            // ---- right-looking LU: row k is final at step k, column k below it takes the multipliers
#pragma unroll
            for (int k = 0; k < N - 1; ++k) {
                const int gro = k % GR, lk = k / GR;
                const int gco = k % GC, ck = k / GC;
                const bool own_col = (GC == 1) || (gc == gco);
                T r[LC], c[LR];
#pragma unroll
                for (int lj = 0; lj < LC; ++lj)
                    r[lj] = (GR > 1) ? shfl_t(a[lk][lj], grp_base + gro * GC + gc) : a[lk][lj];
#pragma unroll
                for (int li = 0; li < LR; ++li)
                    c[li] = (GC > 1) ? shfl_t(a[li][ck], grp_base + gr * GC + gco) : a[li][ck];
                const T pv = (GC > 1) ? shfl_t(r[ck], grp_base + gr * GC + gco) : r[ck];
                const T rinv = rcp_t(pv);
#pragma unroll
                for (int lj = 0; lj < LC; ++lj) r[lj] = (lj * GC + gc > k) ? r[lj] : T(0);  // only columns right of k change
#pragma unroll
                for (int li = 0; li < LR; ++li) {
                    const T l = (li * GR + gr > k) ? c[li] * rinv : T(0);                   // only rows below k change
                    row_update<LC>(a[li], r, -l);
                    a[li][ck] = (own_col && (li * GR + gr > k)) ? l : a[li][ck];
                }
            }
            __syncwarp();  // all lanes finished reading the image
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
                for (int lj = 0; lj < LC; ++lj) {
                    const int j = lj * GC + gc;
                    const bool ok = rok && ((lj * GC + GC - 1 < N) || (j < N));
                    if (ok) mimg[i * N + j] = a[li][lj];
                }
            }
        } else {
        // ---- Gauss-Jordan, deferred scaling ----------------------------------------------
        T dinv[LR];
#pragma unroll
        for (int li = 0; li < LR; ++li) dinv[li] = T(0);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int gro = k % GR, lk = k / GR;  // owner lane-row / local row of row k
            const int gco = k % GC, ck = k / GC;  // owner lane-col / local col of col k
            const bool own_row = (GR == 1) || (gr == gro);
            const bool own_col = (GC == 1) || (gc == gco);
            T r[LC], c[LR];
#pragma unroll
            for (int lj = 0; lj < LC; ++lj)
                r[lj] = (GR > 1) ? shfl_t(a[lk][lj], grp_base + gro * GC + gc) : a[lk][lj];
#pragma unroll
            for (int li = 0; li < LR; ++li)
                c[li] = (GC > 1) ? shfl_t(a[li][ck], grp_base + gr * GC + gco) : a[li][ck];
            const T pv = (GC > 1) ? shfl_t(r[ck], grp_base + gr * GC + gco) : r[ck];
            const T rinv = rcp_t(pv);
            // column k of the augmented identity takes over slot k: broadcast row has a 1 there
            r[ck] = own_col ? T(1) : r[ck];
            T nf[LR];
#pragma unroll
            for (int li = 0; li < LR; ++li) nf[li] = -(c[li] * rinv);
            nf[lk] = own_row ? T(0) : nf[lk];
            const T diag = own_row ? T(1) : T(0);
#pragma unroll
            for (int li = 0; li < LR; ++li) a[li][ck] = own_col ? ((li == lk) ? diag : T(0)) : a[li][ck];
#pragma unroll
            for (int li = 0; li < LR; ++li) row_update<LC>(a[li], r, nf[li]);
            dinv[lk] = own_row ? rinv : dinv[lk];
        }

        // ---- scale, undo the permutation on the way back to the image ---------------------
        __syncwarp();  // all lanes finished reading the image
        int pcol[LC];
#pragma unroll
        for (int lj = 0; lj < LC; ++lj) {
            const int j = lj * GC + gc;
            const bool ok = (lj * GC + GC - 1 < N) || (j < N);
            pcol[lj] = (MODE != kModeNone) ? (ok ? perm[j] : 0) : j;
        }
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) {
                const int j = lj * GC + gc;
                const bool ok = rok && ((lj * GC + GC - 1 < N) || (j < N));
                if (ok) mimg[i * N + pcol[lj]] = a[li][lj] * dinv[li];
            }
        }
        }
        __syncwarp();
        copy_out<T>(gspan, img, total, lane);
        if (piv != nullptr) {
            int32_t* pdst = piv + first * N;
            for (int e = lane; e < nm * N; e += 32)
                pdst[e] = (MODE != kModeNone) ? perm_all[e] : (e % N);
        }
        __syncwarp();  // image and perm are reused by the next tile
    }
}

}  // namespace lub
