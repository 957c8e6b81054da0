// lub_lapack.cuh -- pivot_mode 3: TRUE partial pivoting (LAPACK getrf semantics) with `ipiv` and `info`.
//
// the reference's two pivoting variants search column k BEFORE it has been eliminated
// (parallel_pivot/luBatchedInplace.cuh:159, serial_pivot/...cuh:133), which does not bound element growth -- on
// uniform(0,1) 32 x 32 fp32 matrices its own 1e-3 check (templated/verify.hpp:50-103) fails for 5-8 % of the inputs,
// against 0.2 % with LAPACK's rule.  The check the reference wrote for a real partial-pivoting factorisation,
// verifyLUwithPivoting (parallel_pivot/verify.hpp:157-242), is what this mode is built to pass; zero pivots are
// reported per matrix instead of silently producing inf / NaN.
//
// Because the pivot row is only known once column k has been updated, the permutation cannot be computed ahead of
// the arithmetic as in modes 1 / 2, and on a 2-D lane grid the pivot row would sit at a run-time REGISTER index.
// This kernel therefore uses the layout in which the pivot row is a run-time LANE index instead: lane = row, the N
// entries of the row in registers (column index = register index, static after unrolling), G = 2^ceil(log2 N)
// lanes per matrix.  Rows never move: every lane tracks the POSITION its row has in LAPACK's swapped order, which
// is all that ipiv (a list of position swaps) and the first-maximum tie rule (isamax) need.  Per step: the
// candidates (position >= k) reduce |a[k]| with REDUX.MAX, the lowest position among the maxima wins
// (REDUX.MIN), the winner scales its row and broadcasts it with N shuffles, everybody else eliminates.
// The exchange is N words per lane and step (the 2-D grid of the other modes needs (N/4 + N/4)), so this mode
// is bound by the shuffle crossbar at about 3x the time of mode 2 -- and about 6x faster than cuBLAS
// getrfBatched + getriBatched; it is the numerically safe mode, not the headline.
//
//   LUONLY = false: in-place inverse.  Gauss-Jordan with the pivots of getrf; with rho(k) = the row that was pivot at
//                   step k, the in-place array W ends with A^-1[k][rho(k')] = W[rho(k)][k'] (row k of the inverse
//                   sits in the lane that was pivot at step k, its register k' belongs to column rho(k')).
//   LUONLY = true : the getrf output itself: P A = L U, unit-lower L below the diagonal, U on and above it, rows
//                   in final (swapped) order.
//   ipiv[b][k] = 1-based position the row at position k was swapped with at step k (LAPACK / cuBLAS PivotArray).
//   info[b]    = 0, or k + 1 for the first exactly-zero pivot U(k,k) (then the inverse of that matrix is not
//                meaningful: inf / NaN, as LAPACK's getri refuses it).
#pragma once
#include "lub_fast.cuh"

namespace lub {


constexpr int pow2_ceil(int n) { int g = 1; while (g < n) g *= 2; return g; }

// warp-wide (sub-warp-wide) maximum of |v| as an ordered bit pattern, over the lanes named in mask
__device__ __forceinline__ uint32_t group_max_bits(unsigned mask, uint32_t v) { return __reduce_max_sync(mask, v); }
__device__ __forceinline__ unsigned long long group_max_bits(unsigned mask, unsigned long long v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_max_sync(mask, hi);
    const uint32_t ml = __reduce_max_sync(mask, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

// Round 2, two-phase form of pivot_mode 3 (used by lub_bulk_kernel<..., MODE = kModeLapack>): the permutation of getrf
// is a function of the UPDATED columns, so it is found by running the LU factorisation itself (n^3 / 3 FMAs, a third
// of the inversion) in the lane = row layout on a staged image -- the pivot row is a run-time lane, its trailing
// part travels by N - 1 - k shuffles -- and then the inverse is computed by the same permuted-load / register
// Gauss-Jordan / column-scatter machinery as modes 1 / 2, whose diagonal pivots under that permutation ARE getrf's.
// N (N - 1) / 2 shuffles for the search phase instead of N^2 for a lane = row Gauss-Jordan.
//   mimg: the matrix (row stride P) in shared memory; perm[i] <- original row that getrf moves to position i;
//   ipiv_s[k] <- LAPACK's 1-based swap position of step k; returns info (0, or k + 1 for the first exactly-zero pivot).
// The arithmetic is the getf2 recurrence of lub_lapack_kernel<LUONLY> (reciprocal scaling, fma(-l, r, a)), so both
// mode-3 kernels pick the same pivots.  Every lane of the warp must call this; lanes >= N idle along.
// isamax of one step: the first maximum (lowest position) of |col| over the rows at positions >= k
template <typename T>
__device__ __forceinline__ void getrf_search(T col, int pos, bool mine, int k, int lane, int& p, int& pl, bool& zero) {
    using U = typename FpBits<T>::U;
    const bool cand = mine && (pos >= k);
    const U key = cand ? FpBits<T>::absbits(col) : U(0);
    const U mx = group_max_bits(0xffffffffu, key);
    const unsigned sel = (cand && key == mx) ? (((unsigned)pos << 8) | (unsigned)lane) : 0xffffu;
    const unsigned win = __reduce_min_sync(0xffffffffu, sel);
    p = (int)(win >> 8);
    pl = (int)(win & 0xffu);
    zero = (mx == U(0));
}

// LU = true: the factors themselves are the result (lu_batched_factor_inplace, pivot_mode 3): the multipliers are kept in
// column k and every lane writes its row back into the image at the row's final position -- P A = L U, rows in swapped
// order, as lub_lapack_kernel<LUONLY> leaves them.
// getrf_core works on the row every lane has loaded (a[], lane = original row) and leaves the row's final position in pos;
// the wrappers below load / store the rows from a dense image (here) or a 128-byte-swizzled one (lub_tma.cuh).
template <typename T, int N, bool LU>
__device__ __forceinline__ int getrf_core(T (&a)[N], int& pos, int* __restrict__ perm, int* __restrict__ ipiv_s, int lane) {
    const bool mine = lane < N;
    pos = lane;          // position of this lane's row in LAPACK's row order (rows never move)
    int first_zero = 0;  // info
    int p, pl;
    bool zero;
    if constexpr (sizeof(T) == 4) {
        // Software pipeline (fp32): a step's critical path is search (two warp reductions) -> broadcast of the pivot row ->
        // update, and one warp runs one matrix, so the order of the code is the order of execution.  Column k + 1 is
        // therefore updated FIRST, the search of step k + 1 is issued right behind it and the rest of the pivot row travels
        // while those reductions are in flight; every lane also takes the reciprocal of its own candidate ahead of time,
        // so the pivot's comes with one shuffle instead of a division on the critical path (N = 32: 7.19 -> 6.78 ms).
        T rown = T(1) / a[0];
        getrf_search<T>(a[0], pos, mine, 0, lane, p, pl, zero);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (zero && first_zero == 0) first_zero = k + 1;
            if (pos == k) pos = p;
            if (lane == pl) pos = k;
            if (lane == 0) ipiv_s[k] = p + 1;
            if (k < N - 1) {
                const int plk = pl;
                const T rinv = __shfl_sync(0xffffffffu, rown, plk);
                const bool below = mine && pos > k;
                const T l = below ? a[k] * rinv : T(0);
                if (LU && below) a[k] = l;
                {
                    const T r = __shfl_sync(0xffffffffu, a[k + 1], plk);
                    a[k + 1] = fma(-l, r, a[k + 1]);
                }
                rown = T(1) / a[k + 1];
                getrf_search<T>(a[k + 1], pos, mine, k + 1, lane, p, pl, zero);
#pragma unroll
                for (int j = k + 2; j < N; ++j) {
                    const T r = __shfl_sync(0xffffffffu, a[j], plk);
                    a[j] = fma(-l, r, a[j]);
                }
            }
        }
    } else {
        // fp64: the plain order -- the pipelined form holds more doubles live than 168 registers take (N = 27: 9.1 vs 10.2 ms)
#pragma unroll
        for (int k = 0; k < N; ++k) {
            getrf_search<T>(a[k], pos, mine, k, lane, p, pl, zero);
            if (zero && first_zero == 0) first_zero = k + 1;
            if (pos == k) pos = p;
            if (lane == pl) pos = k;
            if (lane == 0) ipiv_s[k] = p + 1;
            if (k < N - 1) {
                const T pv = __shfl_sync(0xffffffffu, a[k], pl);
                const T rinv = T(1) / pv;
                const bool below = mine && pos > k;
                const T l = below ? a[k] * rinv : T(0);
                if (LU && below) a[k] = l;
#pragma unroll
                for (int j = k + 1; j < N; ++j) {
                    const T r = __shfl_sync(0xffffffffu, a[j], pl);
                    a[j] = fma(-l, r, a[j]);
                }
            }
        }
    }
    if (mine) perm[pos] = lane;
    return first_zero;
}

// fp32: TWO matrices per call, both in the lane = row layout, their steps interleaved by hand.  One matrix is one dependent
// chain (search -> broadcast -> update -> search ...) and a block has 12 warps, so the phase runs at about half of the
// shuffle throughput that bounds it; warp-wide operations keep their program order, hence the interleaving is done in the
// source.  Same recurrence per matrix as getrf_core.
template <int N, bool LU>
__device__ __forceinline__ void getrf_core_x2(float (&a)[N], float (&b)[N], int& posA, int& posB, int* __restrict__ ipivA,
                                              int* __restrict__ ipivB, int lane, int& fzA, int& fzB) {
    const bool mine = lane < N;
    posA = posB = lane;
    fzA = fzB = 0;
    int pA, plA, pB, plB;
    bool zA, zB;
    float rownA = 1.0f / a[0], rownB = 1.0f / b[0];
    getrf_search<float>(a[0], posA, mine, 0, lane, pA, plA, zA);
    getrf_search<float>(b[0], posB, mine, 0, lane, pB, plB, zB);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (zA && fzA == 0) fzA = k + 1;
        if (zB && fzB == 0) fzB = k + 1;
        if (posA == k) posA = pA;
        if (posB == k) posB = pB;
        if (lane == plA) posA = k;
        if (lane == plB) posB = k;
        if (lane == 0) { ipivA[k] = pA + 1; ipivB[k] = pB + 1; }
        if (k < N - 1) {
            const int sA = plA, sB = plB;
            const float rinvA = __shfl_sync(0xffffffffu, rownA, sA);
            const float rinvB = __shfl_sync(0xffffffffu, rownB, sB);
            const bool belowA = mine && posA > k, belowB = mine && posB > k;
            const float lA = belowA ? a[k] * rinvA : 0.0f, lB = belowB ? b[k] * rinvB : 0.0f;
            if (LU && belowA) a[k] = lA;
            if (LU && belowB) b[k] = lB;
            {
                const float rA = __shfl_sync(0xffffffffu, a[k + 1], sA);
                const float rB = __shfl_sync(0xffffffffu, b[k + 1], sB);
                a[k + 1] = fmaf(-lA, rA, a[k + 1]);
                b[k + 1] = fmaf(-lB, rB, b[k + 1]);
            }
            rownA = 1.0f / a[k + 1];
            rownB = 1.0f / b[k + 1];
            getrf_search<float>(a[k + 1], posA, mine, k + 1, lane, pA, plA, zA);
            getrf_search<float>(b[k + 1], posB, mine, k + 1, lane, pB, plB, zB);
#pragma unroll
            for (int j = 2; j < N; ++j) {  // (constant trip count: a bound that depends on k can be left partly rolled)
                if (j >= k + 2) {
                    const float rA = __shfl_sync(0xffffffffu, a[j], sA);
                    const float rB = __shfl_sync(0xffffffffu, b[j], sB);
                    a[j] = fmaf(-lA, rA, a[j]);
                    b[j] = fmaf(-lB, rB, b[j]);
                }
            }
        }
    }
}

// dense-image wrapper of getrf_core_x2: matrices m and m + 1 of a tile (mimg, mimg + MS)
template <int N, int P, int MS, bool LU>
__device__ __forceinline__ void prepass_getrf_x2(float* __restrict__ mimg, int* __restrict__ perm, int* __restrict__ ipiv_s, int lane,
                                                 int& fzA, int& fzB) {
    const bool mine = lane < N;
    const float* rowA = mimg + (mine ? lane : 0) * P;
    const float* rowB = rowA + MS;
    float a[N], b[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { a[j] = rowA[j]; b[j] = rowB[j]; }
    int posA, posB;
    getrf_core_x2<N, LU>(a, b, posA, posB, ipiv_s, ipiv_s + N, lane, fzA, fzB);
    if (mine) { perm[posA] = lane; perm[N + posB] = lane; }
    if constexpr (LU) {
        __syncwarp();  // every lane has long read its rows; now the rows change places
        if (mine) {
            float* dA = mimg + posA * P;
            float* dB = mimg + MS + posB * P;
#pragma unroll
            for (int j = 0; j < N; ++j) { dA[j] = a[j]; dB[j] = b[j]; }
        }
    }
}

template <typename T, int N, int P, bool LU>
__device__ __forceinline__ int prepass_getrf(T* __restrict__ mimg, int* __restrict__ perm, int* __restrict__ ipiv_s, int lane) {
    const bool mine = lane < N;
    const T* rowp = mimg + (mine ? lane : 0) * P;
    T a[N];
#pragma unroll
    for (int j = 0; j < N; ++j) a[j] = rowp[j];
    int pos;
    const int first_zero = getrf_core<T, N, LU>(a, pos, perm, ipiv_s, lane);
    if constexpr (LU) {
        __syncwarp();  // every lane has long read its row; now the rows change places
        if (mine) {
            T* dst = mimg + pos * P;
#pragma unroll
            for (int j = 0; j < N; ++j) dst[j] = a[j];
        }
    }
    return first_zero;
}

// ---- factors only, pivot modes 0 - 2 (lu_batched_factor_inplace) ------------------------------------------------------------
// The reference's variants know their permutation before any arithmetic, so the factors "after the k-loop"
// (templated/luBatchedInplace.cuh:99-113; rows in pivoted order, L below the diagonal with a unit diagonal implied, U on and
// above it -- what verifyLU / verifyLUwithPivoting read, templated/verify.hpp:105-186) are an LU factorisation WITHOUT a search:
// lane = row POSITION (the lane loads the row the permutation puts there), so the pivot row of step k sits in lane k, a
// compile-time lane; its trailing part travels by N - 1 - k shuffles; right-looking update with the reciprocal of the pivot.
// Same arithmetic per entry as the Gauss-Jordan kernels' elimination (fma(-l, r, a), l = a * (1 / pivot)).
template <typename T, int N>
__device__ __forceinline__ void lu_core_static(T (&a)[N], int lane) {
    const bool mine = lane < N;
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
        const T pv = __shfl_sync(0xffffffffu, a[k], k);
        const T rinv = T(1) / pv;
        const bool below = mine && lane > k;
        const T l = below ? a[k] * rinv : T(0);
        if (below) a[k] = l;
#pragma unroll
        for (int j = k + 1; j < N; ++j) {
            const T r = __shfl_sync(0xffffffffu, a[j], k);
            a[j] = fma(-l, r, a[j]);
        }
    }
}
// dense image (row stride P); perm == nullptr: no pivoting
template <typename T, int N, int P>
__device__ __forceinline__ void lu_rows_dense(T* __restrict__ mimg, const int* __restrict__ perm, int lane) {
    const bool mine = lane < N;
    const int row = mine ? (perm != nullptr ? perm[lane] : lane) : 0;
    const T* rowp = mimg + row * P;
    T a[N];
#pragma unroll
    for (int j = 0; j < N; ++j) a[j] = rowp[j];
    lu_core_static<T, N>(a, lane);
    __syncwarp();  // every lane has long read its row; now the rows change places
    if (mine) {
        T* dst = mimg + lane * P;
#pragma unroll
        for (int j = 0; j < N; ++j) dst[j] = a[j];
    }
}

template <typename T, int N>
struct LapackLayout {
    static constexpr int G = pow2_ceil(N), MPW = 32 / G, P = N | 1;
    static constexpr int smem_bytes(int warps) { return warps * (MPW * N * P * (int)sizeof(T) + MPW * N * 4) + 16; }
};

template <typename T, int N, bool LUONLY>
__global__ void __launch_bounds__(256)
lub_lapack_kernel(T* __restrict__ A, int32_t* __restrict__ ipiv, int32_t* __restrict__ info, long long batch) {
    using U = typename FpBits<T>::U;
    constexpr int G = LapackLayout<T, N>::G, MPW = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane % G, ml = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (ml * G));
    // per warp: an output image of MPW matrices with an odd row stride (conflict-free row scatter) + rho[]
    constexpr int P = N | 1;
    T* img = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (MPW * N * P);
    int* rho_all = reinterpret_cast<int*>(reinterpret_cast<T*>(smem_raw) + (size_t)nwarps * (MPW * N * P)) + warp * (MPW * N);
    T* mimg = img + ml * (N * P);
    int* rho = rho_all + ml * N;

    const long long ntiles = (batch + MPW - 1) / MPW;
#pragma unroll 1
    for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < ntiles; tile += (long long)gridDim.x * nwarps) {
        const long long b = tile * MPW + ml;
        const bool live = (b < batch) && (g < N);   // lanes past the matrix / past the batch run along on zeros
        T* gm = A + (b < batch ? b : 0) * (long long)(N * N);
        T a[N];
#pragma unroll
        for (int j = 0; j < N; ++j) a[j] = live ? gm[g * N + j] : T(g == j ? 1 : 0);
        int pos = g;          // position of this lane's row in LAPACK's row order (rows never move here)
        int mystep = g;       // step at which this lane's row was the pivot
        int first_zero = 0;   // info
#pragma unroll
        for (int k = 0; k < N; ++k) {
            // ---- isamax over the rows at positions >= k of the UPDATED column k: first maximum wins ----
            const bool cand = (pos >= k) && (g < N);
            const U key = cand ? FpBits<T>::absbits(a[k]) : U(0);
            const U mx = group_max_bits(gmask, key);
            const unsigned sel = (cand && key == mx) ? (((unsigned)pos << 8) | (unsigned)g) : 0xffffu;
            const unsigned win = __reduce_min_sync(gmask, sel);
            const int p = (int)(win >> 8), pl = (int)(win & 0xffu);   // position and lane of the pivot row
            if (mx == U(0) && first_zero == 0) first_zero = k + 1;
            if (pos == k) pos = p;           // the row that sat at position k moves to p ...
            const bool is_piv = (g == pl);
            if (is_piv) { pos = k; mystep = k; }   // ... and the pivot row to k
            if (g == 0 && live && ipiv != nullptr) ipiv[b * N + k] = p + 1;
            if (!LUONLY && g == 0) rho[k] = pl;
            // ---- eliminate ----
            const T pv = __shfl_sync(0xffffffffu, a[k], pl, G);
            const T rinv = T(1) / pv;
            if (LUONLY) {
                // rows below position k: multiplier l = a[k] / pivot, trailing update of columns > k
                const bool below = pos > k;
                const T l = below ? a[k] * rinv : T(0);
                if (below) a[k] = l;
#pragma unroll
                for (int j = k + 1; j < N; ++j) {
                    const T r = __shfl_sync(0xffffffffu, a[j], pl, G);
                    a[j] = fma(-l, r, a[j]);
                }
            } else {
                // Gauss-Jordan, in place: the pivot row is scaled by 1/pivot and gets 1/pivot in column k; every other
                // row t subtracts a[t][k] times the scaled pivot row and keeps -a[t][k]/pivot in column k
                const T t = is_piv ? T(0) : a[k];
                a[k] = is_piv ? rinv : T(0);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    if (j == k) continue;
                    if (is_piv) a[j] *= rinv;
                    const T r = __shfl_sync(0xffffffffu, a[j], pl, G);
                    a[j] = fma(-t, r, a[j]);
                }
                a[k] = fma(-t, rinv, a[k]);
            }
        }
        // ---- results -> image (row order restored), image -> global, coalesced ----
        __syncwarp();
        if (g < N) {
            if (LUONLY) {
#pragma unroll
                for (int j = 0; j < N; ++j) mimg[pos * P + j] = a[j];
            } else {
#pragma unroll
                for (int j = 0; j < N; ++j) mimg[mystep * P + rho[j]] = a[j];
            }
        }
        __syncwarp();
        const long long first = tile * MPW;
        const int nm = (batch - first < MPW) ? (int)(batch - first) : MPW;
        T* gspan = A + first * (long long)(N * N);
        for (int e = lane; e < nm * N * N; e += 32) {
            const int m = e / (N * N), rem = e % (N * N);
            gspan[e] = img[m * (N * P) + (rem / N) * P + (rem % N)];
        }
        if (g == 0 && live && info != nullptr) info[b] = first_zero;
        __syncwarp();
    }
}

}  // namespace lub
