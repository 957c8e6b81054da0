// lub_fast.cuh -- helpers shared by the v3 / v4 / v5 kernels: padded 128-bit image layout and
// copies (pivot_mode none), vector load/store helpers, lane-dependent select / predicated move
// wrappers, and the pivot pre-pass (warp-wide REDUX search + exact group search).
//
// History: this file used to hold the second-generation kernel (shared-memory mailbox exchange).
// Nsight Compute showed a 128-bit LDS/STS costs four wavefronts even as a broadcast and that
// predicated owner stores pay per quarter-warp, i.e. a mailbox costs twice the wavefronts of
// shuffles (profiles/r01_prof_v2_n32.md); the kernel was replaced by lub_v3.cuh and removed.
#pragma once
#include "lub_kernel.cuh"

namespace lub {

constexpr int cdiv_(int a, int b) { return (a + b - 1) / b; }
constexpr int gcd_(int a, int b) { return b == 0 ? a : gcd_(b, a % b); }
constexpr int roundup_(int a, int b) { return cdiv_(a, b) * b; }

// conflict degree of `rows` lanes walking one column of an image whose rows are `pw`
// 32-bit words apart, elements `ew` words wide
constexpr int column_conflict(int rows, int pw, int ew) {
    const int slots = 32 / ew;
    const int distinct = slots / gcd_((pw / ew) % slots == 0 ? slots : (pw / ew) % slots, slots);
    return cdiv_(rows, distinct);
}

template <typename T, int N>
constexpr int pick_row_pad() {  // elements; multiple of 16 bytes
    constexpr int ES = sizeof(T), EPV = 16 / ES, EW = ES / 4;
    int best = 0, best_deg = 1 << 30;
    for (int pad = 0; pad <= 3 * EPV; pad += EPV) {
        const int deg = column_conflict(N, (N + pad) * EW, EW);
        if (deg < best_deg) { best_deg = deg; best = pad; }
    }
    return best;
}

template <typename T, int N, int P, int MPW>
constexpr int pick_mat_pad() {  // elements; multiple of 16 bytes; spreads the tile's matrices over banks
    constexpr int ES = sizeof(T), EPV = 16 / ES, EW = ES / 4;
    if (MPW == 1) return 0;
    for (int pad = 0; pad < 32 / EW; pad += EPV) {
        const int msw = (N * P + pad) * EW;
        if ((msw % 8) == 4) return pad;  // 4 * odd words: eight matrices land on eight different bank groups
    }
    return 0;
}


// Lane-dependent selects are written as PTX selp so that the compiler keeps them as one
// FSEL each: given a C++ ?: it if-converts the whole rank-1 update into two divergent copies
// (one per value of "this lane owns column k"), doubling the FMA work of every warp.
__device__ __forceinline__ float sel_t(bool p, float x, float y) {
    float d;
    asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f32 %0, %1, %2, q; }" : "=f"(d) : "f"(x), "f"(y), "r"((int)p));
    return d;
}
__device__ __forceinline__ double sel_t(bool p, double x, double y) {
    double d;
    asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(d) : "d"(x), "d"(y), "r"((int)p));
    return d;
}

// In-place predicated overwrite: `if (p) x = v` as ONE predicated move on x's own register.
// (A select creates a new value; when x is half of an FFMA2 register pair the compiler then has
// to rebuild the pair with an extra MOV per row per step.)
__device__ __forceinline__ void set_if(bool p, float& x, float v) {
    asm("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q mov.f32 %0, %2; }" : "+f"(x) : "r"((int)p), "f"(v));
}
__device__ __forceinline__ void set_if(bool p, double& x, double v) {
    asm("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q mov.f64 %0, %2; }" : "+d"(x) : "r"((int)p), "d"(v));
}

// ---- vector helpers -------------------------------------------------------------------------

template <typename T, int CH> struct Vec;
template <> struct Vec<float, 4> { using V = float4; };
template <> struct Vec<float, 2> { using V = float2; };
template <> struct Vec<float, 1> { using V = float; };
template <> struct Vec<double, 2> { using V = double2; };
template <> struct Vec<double, 1> { using V = double; };

template <typename T, int CH>
__device__ __forceinline__ void ld_vec(const T* __restrict__ p, T* __restrict__ dst) {
    using V = typename Vec<T, CH>::V;
    const V v = *reinterpret_cast<const V*>(p);
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < CH; ++i) dst[i] = e[i];
}
template <typename T, int CH>
__device__ __forceinline__ void st_vec(T* __restrict__ p, const T* __restrict__ src) {
    using V = typename Vec<T, CH>::V;
    V v;
    T* e = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < CH; ++i) e[i] = src[i];
    *reinterpret_cast<V*>(p) = v;
}

// ---- padded copy global <-> image (ROWVEC) ---------------------------------------------------

template <typename T, typename L, int N>
__device__ __forceinline__ void copy_in_padded(unsigned char* __restrict__ img, const T* __restrict__ src,
                                               int nchunks, int lane) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(img);
    auto dst_of = [](int q) {
        const int rowidx = q / L::CPR16;
        int o = q + rowidx * L::RPAD16;
        if (L::MPAD16 != 0) o += (rowidx / N) * L::MPAD16;
        return o;
    };
    int q = lane;
    for (; q + 96 < nchunks; q += 128) {
        const uint4 v0 = ld_stream16(s + q), v1 = ld_stream16(s + q + 32);
        const uint4 v2 = ld_stream16(s + q + 64), v3 = ld_stream16(s + q + 96);
        d[dst_of(q)] = v0;
        d[dst_of(q + 32)] = v1;
        d[dst_of(q + 64)] = v2;
        d[dst_of(q + 96)] = v3;
    }
    for (; q < nchunks; q += 32) d[dst_of(q)] = ld_stream16(s + q);
}

// same mapping, asynchronous (LDGSTS): no registers, the warp does not wait
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory"); }

template <typename T, typename L, int N>
__device__ __forceinline__ void copy_in_padded_async(unsigned char* __restrict__ img, const T* __restrict__ src,
                                                     int nchunks, int lane) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(img);
    for (int q = lane; q < nchunks; q += 32) {
        const int rowidx = q / L::CPR16;
        int o = q + rowidx * L::RPAD16;
        if (L::MPAD16 != 0) o += (rowidx / N) * L::MPAD16;
        cp_async16(d + o, s + q);
    }
}

template <typename T, typename L, int N>
__device__ __forceinline__ void copy_out_padded(T* __restrict__ dst, const unsigned char* __restrict__ img,
                                                int nchunks, int lane) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    const uint4* s = reinterpret_cast<const uint4*>(img);
    for (int q = lane; q < nchunks; q += 32) {
        const int rowidx = q / L::CPR16;
        int o = q + rowidx * L::RPAD16;
        if (L::MPAD16 != 0) o += (rowidx / N) * L::MPAD16;
        st_stream16(d + q, s[o]);
    }
}

// ---- warp-wide pivot pre-pass (one matrix, lane = row position) -----------------------------

constexpr unsigned reach_mask_of(int n) {  // bit t set <=> slot t of an n-slot tree reaches slot 0
    unsigned m = 0;
    for (int t = 0; t < n && t < 32; ++t)
        if (tree_slot_rank(t, n) >= 0) m |= 1u << t;
    return m;
}
template <int N> struct ReachMask { static constexpr unsigned value = reach_mask_of(N); };

__device__ __forceinline__ uint32_t warp_max_bits(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned long long warp_max_bits(unsigned long long v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

template <typename T, int N, int G, int MODE, int P>
__device__ __forceinline__ void prepass_group(const T* __restrict__ mimg, int* __restrict__ perm,
                                              const int8_t* __restrict__ slot_rank, int g);

// Exact but slow warp-wide search (explicit priorities, butterfly reductions); out of line so
// that its code stays off the hot path.
template <typename T, int N, int MODE, int P>
__device__ __noinline__ void prepass_exact(const T* mimg, int* perm, const int8_t* slot_rank, int lane) {
    prepass_group<T, N, 32, MODE, P>(mimg, perm, slot_rank, lane);
}

#ifndef LUB_ROWWISE_PREPASS
#define LUB_ROWWISE_PREPASS 1
#endif
constexpr bool kRowwisePrepass = LUB_ROWWISE_PREPASS != 0;
constexpr bool rowwise_prepass_ok(int n, int mode) {
    const unsigned all = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
    return kRowwisePrepass && (mode == kModeSerial || (reach_mask_of(n) & (all >> 1)) == (all >> 1));
}

// Row-wise search (lane = ORIGINAL row).  Usable whenever the candidate set of step k is simply
// "every row not picked yet" -- serial mode, and parallel mode when the reference tree reaches all
// of its slots (N a power of two): find_pivot looks at un-eliminated entries only
// (serial_pivot/luBatchedInplace.cuh:24-41, parallel_pivot/...cuh:25-68), so a row's key for column k
// never changes and row positions matter for tie-breaks alone.  Each step is LDS (static address)
// -> key -> REDUX.MAX -> compare -> two selects: no ballot, no shuffle, and a dependency chain a
// third as long as the position-wise search below.  The row picked at step k ends at position k for
// good, so lane r just remembers the step at which it was picked (= its final position).  A step with
// two equal maxima retires two lanes at once; then the survivors do not add up to one at the end
// and the matrix is redone by the exact search (rare: needs equal |values| in one column).
template <int N, int MODE> struct RowwiseOk { static constexpr bool value = rowwise_prepass_ok(N, MODE); };

template <typename T, int N, int MODE, int P, int MI, bool INLINE_EXACT, bool VEC = false>
__device__ __forceinline__ void prepass_rowwise(const T* const (&img)[MI], int* const (&perm)[MI],
                                                const int8_t* __restrict__ slot_rank, int lane) {
    using U = uint32_t;  // keys: the upper word of |x| (fp64: see FpBits<double>::hi31 -- one REDUX per step)
    const int roff = ((lane < N) ? lane : 0) * P;
    U alive[MI];
    int when[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) { alive[m] = (lane < N) ? ~U(0) : U(0); when[m] = N - 1; }
    constexpr int EPV = 16 / (int)sizeof(T);
    T x[MI][EPV];
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
        U key[MI], mx[MI];
        if (VEC && (k % EPV) == 0) {  // 16-byte image: one vector load feeds EPV steps
#pragma unroll
            for (int m = 0; m < MI; ++m) ld_vec<T, EPV>(img[m] + roff + k, x[m]);
        }
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            const T xv = VEC ? x[m][k % EPV] : img[m][roff + k];
            key[m] = ((FpBits<T>::hi31(xv) << 1) | U(1)) & alive[m];
        }
#pragma unroll
        for (int m = 0; m < MI; ++m) mx[m] = __reduce_max_sync(0xffffffffu, key[m]);
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            const bool hit = key[m] == mx[m];
            when[m] = hit ? k : when[m];
            alive[m] = hit ? U(0) : alive[m];
        }
    }
#pragma unroll
    for (int m = 0; m < MI; ++m) {
        const bool ok = __popc(__ballot_sync(0xffffffffu, alive[m] != U(0))) == 1;  // warp-uniform
        if (ok) {
            if (lane < N) perm[m][when[m]] = lane;
        } else {
            if (INLINE_EXACT) prepass_group<T, N, 32, MODE, P>(img[m], perm[m], slot_rank, lane);
            else prepass_exact<T, N, MODE, P>(img[m], perm[m], slot_rank, lane);
        }
    }
}

// MI matrices are searched in lock step: their dependency chains (LDS -> REDUX -> VOTE -> SHFL)
// are independent, so the scheduler overlaps them.
// INLINE_EXACT: inline the exact fallback instead of calling it (a call needs the callee's
// register budget, which a warp that has shrunk its allocation with setmaxnreg does not have).
// The MI matrices may live anywhere (img[m], perm[m]): a producer warp searches two tiles at once.
template <typename T, int N, int MODE, int P, int MI, bool INLINE_EXACT = false, bool VEC = false>
__device__ __forceinline__ void prepass_warp_ptrs(const T* const (&img)[MI], int* const (&perm)[MI],
                                                  const int8_t* __restrict__ slot_rank, int lane) {
    using U = typename FpBits<T>::U;
    constexpr unsigned ALL = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u);
    constexpr unsigned REACH = ReachMask<N>::value;
    if constexpr (RowwiseOk<N, MODE>::value) {
        prepass_rowwise<T, N, MODE, P, MI, INLINE_EXACT, VEC>(img, perm, slot_rank, lane);
    } else {
    int prow[MI];  // original row sitting at position `lane`
    unsigned multi[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) { prow[m] = (lane < N) ? lane : 0; multi[m] = 0u; }
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
        // rows this step may pick: serial = every row below k; parallel = the slots the
        // reference tree merges into slot 0 (slot t <-> row k+1+t), plus the seed row k
        const unsigned vmask = ((MODE == kModeParallel) ? ((REACH << (k + 1)) | (1u << k)) : (ALL << k)) & ALL;
        // (when every slot is reachable the mask is just "position >= k")
        const bool valid = (vmask == ((ALL << k) & ALL)) ? (lane >= k && lane < N) : (((vmask >> lane) & 1u) != 0u);
        U v[MI], mx[MI];
        unsigned bal[MI];
#pragma unroll
        for (int m = 0; m < MI; ++m) v[m] = FpBits<T>::absbits(img[m][prow[m] * P + k]);
#pragma unroll
        for (int m = 0; m < MI; ++m) mx[m] = warp_max_bits(valid ? v[m] : U(0));
#pragma unroll
        for (int m = 0; m < MI; ++m) bal[m] = __ballot_sync(0xffffffffu, valid && v[m] == mx[m]);
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            // Lowest position among the maxima.  Serial mode: that IS find_pivot's answer (strict
            // '>' from the seed row k upwards).  Parallel mode: identical whenever the maximum is
            // unique or row k itself (the seed every tree slot starts from) is among the maxima.
            // Equal maxima in several OTHER slots need the tree's rank order; any step with more
            // than one maximum is only flagged here (no branch in this loop: a branch per step
            // costs the scheduler the overlap between the MI chains) and such a matrix is redone
            // with the exact search afterwards.
            const int wl = __ffs(bal[m]) - 1;
            if (MODE == kModeParallel) multi[m] |= bal[m] & (bal[m] - 1u);  // non-zero <=> two or more maxima
            // swap the row ids of positions k and wl (a no-op when wl == k)
            const int other = __shfl_sync(0xffffffffu, prow[m], lane == k ? wl : k);
            prow[m] = (lane == k || lane == wl) ? other : prow[m];
        }
    }
#pragma unroll
    for (int m = 0; m < MI; ++m)
        if (lane < N) perm[m][lane] = prow[m];
    if (MODE == kModeParallel) {
#pragma unroll
        for (int m = 0; m < MI; ++m)
            if (multi[m] != 0u) {  // warp-uniform, rare (needs two equal |values| in one column)
                __syncwarp();
                if (INLINE_EXACT) prepass_group<T, N, 32, MODE, P>(img[m], perm[m], slot_rank, lane);
                else prepass_exact<T, N, MODE, P>(img[m], perm[m], slot_rank, lane);
            }
    }
    }
}

// MI matrices of one tile: img0 + m * MS, perm0 + m * N
template <typename T, int N, int MODE, int P, int MS, int MI, bool INLINE_EXACT = false, bool VEC = false>
__device__ __forceinline__ void prepass_warp(const T* __restrict__ img0, int* __restrict__ perm0,
                                             const int8_t* __restrict__ slot_rank, int lane) {
    const T* img[MI];
    int* perm[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) { img[m] = img0 + m * MS; perm[m] = perm0 + m * N; }
    prepass_warp_ptrs<T, N, MODE, P, MI, INLINE_EXACT, VEC>(img, perm, slot_rank, lane);
}

// generic sub-warp pre-pass on a strided image (N <= 16): see pivot_prepass in lub_kernel.cuh
template <typename T, int N, int G, int MODE, int P>
__device__ __forceinline__ void prepass_group(const T* __restrict__ mimg, int* __restrict__ perm,
                                              const int8_t* __restrict__ slot_rank, int g) {
    using U = typename FpBits<T>::U;
    for (int i = g; i < N; i += G) perm[i] = i;
    __syncwarp();
    for (int k = 0; k < N - 1; ++k) {
        U best_v = FpBits<T>::absbits(mimg[perm[k] * P + k]);
        unsigned best_p = 0;
        for (int t = g; t < N - 1 - k; t += G) {
            int pr;
            if (MODE == kModeParallel) {
                pr = slot_rank[t];
                if (pr < 0) continue;
            } else {
                pr = t;
            }
            const U v = FpBits<T>::absbits(mimg[perm[k + 1 + t] * P + k]);
            const unsigned p = ((unsigned)(pr + 1) << 8) | (unsigned)(t + 1);
            if (v > best_v || (v == best_v && p < best_p)) { best_v = v; best_p = p; }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            const U ov = __shfl_xor_sync(0xffffffffu, best_v, off);
            const unsigned op = __shfl_xor_sync(0xffffffffu, best_p, off);
            if (ov > best_v || (ov == best_v && op < best_p)) { best_v = ov; best_p = op; }
        }
        __syncwarp();
        if (g == 0 && best_p != 0) {
            const int p = k + (int)(best_p & 0xffu);
            const int tmp = perm[k];
            perm[k] = perm[p];
            perm[p] = tmp;
        }
        __syncwarp();
    }
}

}  // namespace lub
