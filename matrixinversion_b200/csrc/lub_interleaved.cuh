// lub_interleaved.cuh -- the batch-interleaved layout option of the north star (N <= 8):
//     T A[n][n][batch]   element (i, j) of matrix b at offset (i * n + j) * batch + b
// instead of the reference's matrix-major T A[batch][n][n] (templated/luBatchedInplace.cuh:89-97).  With the
// batch index innermost ONE LANE OWNS A WHOLE MATRIX (up to 64 registers) and every global access of a warp is one
// fully coalesced vector access: lane l reads VEC consecutive matrices' element (i, j) as one 4 / 8 / 16-byte word,
// the warp 128 / 256 / 512 contiguous bytes.  No shared memory, no shuffles, no staging: the kernel is a straight
// stream of loads, FMAs and stores, which is why this layout is the natural one for tiny matrices (a matrix-major
// 4 x 4 fp32 matrix is 64 bytes: half a cache line, shared by two lanes' worth of work).
//
// All four pivot modes are supported; pivoting is data-dependent SELECTS on statically indexed registers (each
// lane has its own matrix, so "row p" is a per-lane value, never a register index):
//   mode 1 / 2  the reference's rules (serial_pivot/luBatchedInplace.cuh:22-36; parallel_pivot/...cuh:12-44 with its
//               tree emulated literally, dropped slots included), searching the UN-eliminated column:
//               a pre-pass applies the swaps to the matrix, then the no-pivot elimination runs;
//   mode 3      LAPACK getrf semantics: search the updated column inside the elimination, ipiv + info.
// In-place Gauss-Jordan, then the column interchanges in reverse order (A^-1 = (P A)^-1 P, what LAPACK's getri does).
// piv keeps its [batch][n] layout: the permutation vector of modes 1 / 2 (identity for mode 0) or ipiv (mode 3).
#pragma once
#include "lub_lapack.cuh"

namespace lub {

template <typename T>
__device__ __forceinline__ void cswap(bool c, T& x, T& y) {
    const T t = c ? y : x;
    y = c ? x : y;
    x = t;
}

// argmax |a[i][k]|, i = k .. N - 1, FIRST maximum on ties (find_pivot, serial_pivot/luBatchedInplace.cuh:22-36; LAPACK's isamax):
// an adjacent-pair tournament -- the left entry of a comparison always covers lower rows than the right one, and a strict "<" keeps
// the left entry on a tie -- instead of one chain of N - k dependent compare-selects.  TOURN = false: the chain.  Measured
// (profiles/r02_tune_lane_small_n.jsonl, r02_interleaved_vs_matrix_major.jsonl): the tournament wins where the lane is fed from a
// staged image (matrix-major N = 8 fp32 serial: 0.167 -> 0.156 ms) and loses in the batch-interleaved kernel (its v[] / ix[]
// arrays cost registers there: N = 7 fp32 mode 3 0.097 -> 0.147 ms), so it is a template choice.
template <typename T, int N, bool TOURN>
__device__ __forceinline__ int argmax_first(const T (&a)[N][N], const int k, typename FpBits<T>::U& best) {
    using U = typename FpBits<T>::U;
    if constexpr (!TOURN) {
        int p = k;
        best = FpBits<T>::absbits(a[k][k]);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const U v = FpBits<T>::absbits(a[i][k]);
            if (v > best) { best = v; p = i; }   // strict: the first maximum wins
        }
        return p;
    }
    U v[N];
    int ix[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { v[i] = (i >= k) ? FpBits<T>::absbits(a[i][k]) : U(0); ix[i] = i; }
#pragma unroll
    for (int w = 1; w < N; w *= 2) {
#pragma unroll
        for (int t = 0; t < N; ++t) {
            if (t >= k && ((t - k) % (2 * w)) == 0 && t + w < N) {
                if (v[t] < v[t + w]) { v[t] = v[t + w]; ix[t] = ix[t + w]; }
            }
        }
    }
    best = v[k];
    return ix[k];
}

// One matrix, entirely in the registers of one lane.  piv_out[k]: see above.  Returns info (mode 3) or 0.
template <typename T, int N, int MODE, bool TOURN = false>
__device__ __forceinline__ int invert_in_registers(T (&a)[N][N], int (&piv_out)[N]) {
    using U = typename FpBits<T>::U;
    int sw[N];  // row interchanged with row k at step k (0-based; sw[k] == k: none)
    int info = 0;
    T dinv[N];
    if (MODE == kModeSerial || MODE == kModeParallel) {
        int perm[N];
#pragma unroll
        for (int i = 0; i < N; ++i) perm[i] = i;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            int p = k;
            // The tree is emulated literally for every N >= 3.  (Until late in round 2 a power-of-two N took the serial search
            // here, on the argument that the tree then reaches every slot and a tie stays with the lower slot.  The slots
            // are right, the tie rule is not: stride 2 moves slot 2's row into slot 0 before stride 1 compares it with slot
            // 1, so of two equal maxima in slots 1 and 2 the HIGHER row wins -- found by the tie-heavy integer matrices of
            // tests/test_gpu_parity.py once the matrix-major N = 8 kernel used this function.)
            constexpr bool TREE = (MODE == kModeParallel) && N >= 3;
            if (!TREE) {
                U best;
                p = argmax_first<T, N, TOURN>(a, k, best);   // lowest row wins ties
            } else {
                // find_pivot_parallel, literally: TPM = N slots, slot t seeded with row k and offered row k + 1 + t,
                // then the halving tree with integer-divided strides (slots it never merges are dropped, Q2)
                U v[N];
                int ix[N];
                const U seed = FpBits<T>::absbits(a[k][k]);
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    v[t] = seed; ix[t] = k;
                    if (k + 1 + t < N) {
                        const U w = FpBits<T>::absbits(a[k + 1 + t][k]);
                        if (w > seed) { v[t] = w; ix[t] = k + 1 + t; }
                    }
                }
#pragma unroll
                for (int s = N / 2; s > 0; s >>= 1) {
#pragma unroll
                    for (int t = 0; t < s; ++t)
                        if (v[t] < v[t + s]) { v[t] = v[t + s]; ix[t] = ix[t + s]; }
                }
                p = ix[0];
            }
            sw[k] = p;
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool c = (p == i);
#pragma unroll
                for (int j = 0; j < N; ++j) cswap(c, a[k][j], a[i][j]);
                cswap(c, perm[k], perm[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) piv_out[i] = perm[i];
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (MODE == kModeLapack) {
            U best;
            const int p = argmax_first<T, N, TOURN>(a, k, best);   // isamax: first maximum
            if (best == U(0) && info == 0) info = k + 1;
            sw[k] = p;
            piv_out[k] = p + 1;
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const bool c = (p == i);
#pragma unroll
                for (int j = 0; j < N; ++j) cswap(c, a[k][j], a[i][j]);
            }
        }
        if (MODE == kModeLapack) {
            // the arithmetic of lub_lapack_kernel, operation for operation (results are bitwise equal): the pivot row
            // is scaled by 1/pivot, every other row t subtracts a[t][k] times it and keeps -a[t][k]/pivot in column k
            const T rinv = T(1) / a[k][k];
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (j != k) a[k][j] *= rinv;
            a[k][k] = rinv;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (i == k) continue;
                const T t = a[i][k];
#pragma unroll
                for (int j = 0; j < N; ++j)
                    if (j != k) a[i][j] = fma(-t, a[k][j], a[i][j]);
                a[i][k] = fma(-t, rinv, T(0));
            }
        } else {
            // the arithmetic of gj_eliminate (lub_v3.cuh), operation for operation (results are bitwise equal to the
            // matrix-major kernels): rows are scaled by 1/pivot only at the end, column k becomes the multipliers
            const T rinv = rcp_t(a[k][k]);
            T r[N];
#pragma unroll
            for (int j = 0; j < N; ++j) r[j] = a[k][j];
            r[k] = T(1);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (i == k) continue;
                const T nf = -(a[i][k] * rinv);
                a[i][k] = T(0);
#pragma unroll
                for (int j = 0; j < N; ++j) a[i][j] = fma(nf, r[j], a[i][j]);
            }
            a[k][k] = T(1);
            dinv[k] = rinv;
        }
    }
    if (MODE != kModeLapack) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) a[i][j] *= dinv[i];
    }
    if (MODE != kModeNone) {  // undo the row interchanges on the columns, last one first
#pragma unroll
        for (int k = N - 1; k >= 0; --k) {
#pragma unroll
            for (int j = k + 1; j < N; ++j) {
                const bool c = (sw[k] == j);
#pragma unroll
                for (int i = 0; i < N; ++i) cswap(c, a[i][k], a[i][j]);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) piv_out[i] = i;
    }
    return info;
}

template <typename T, int VEC> struct VecOf;
template <> struct VecOf<float, 1> { using V = float; };
template <> struct VecOf<float, 2> { using V = float2; };
template <> struct VecOf<float, 4> { using V = float4; };
template <> struct VecOf<double, 1> { using V = double; };
template <> struct VecOf<double, 2> { using V = double2; };

// VEC consecutive matrices per lane (one 4 / 8 / 16-byte word per element): needs batch % VEC == 0 and a base pointer
// aligned to VEC * sizeof(T), otherwise the VEC = 1 instantiation is launched.
template <typename T, int N, int MODE, int VEC>
__global__ void __launch_bounds__(128)
lub_interleaved_kernel(T* __restrict__ A, int32_t* __restrict__ piv, int32_t* __restrict__ info, long long batch) {
    using V = typename VecOf<T, VEC>::V;
    const long long groups = batch / VEC;
    for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += (long long)gridDim.x * blockDim.x) {
        T a[VEC][N][N];
        V* base = reinterpret_cast<V*>(A) + gidx;
#pragma unroll
        for (int e = 0; e < N * N; ++e) {
            const V w = base[(long long)e * groups];
            const T* c = reinterpret_cast<const T*>(&w);
#pragma unroll
            for (int m = 0; m < VEC; ++m) a[m][e / N][e % N] = c[m];
        }
#pragma unroll
        for (int m = 0; m < VEC; ++m) {
            int pv[N];
            const int st = invert_in_registers<T, N, MODE>(a[m], pv);
            const long long b = gidx * VEC + m;
            if (piv != nullptr) {
#pragma unroll
                for (int k = 0; k < N; ++k) piv[b * N + k] = pv[k];
            }
            if (info != nullptr) info[b] = st;
        }
#pragma unroll
        for (int e = 0; e < N * N; ++e) {
            V w;
            T* c = reinterpret_cast<T*>(&w);
#pragma unroll
            for (int m = 0; m < VEC; ++m) c[m] = a[m][e / N][e % N];
            base[(long long)e * groups] = w;
        }
    }
}

// matrices per lane: as many as fit ~64 data registers and one 16-byte word
template <typename T, int N>
struct InterleavedCfg {
    static constexpr int MAXV = 16 / (int)sizeof(T);
    static constexpr int REGS = N * N * ((int)sizeof(T) / 4);
    static constexpr int VEC = (MAXV >= 4 && REGS * 4 <= 64) ? 4 : ((REGS * 2 <= 72) ? 2 : 1);
};

}  // namespace lub
