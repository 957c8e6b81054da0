// lub_v3.cuh -- third-generation hot-path kernel (the one the library launches for fp32).  Same
// algorithm and results as lub_kernel.cuh; the data movement is re-planned around the cost model that
// Nsight Compute measurements on B200 established (profiles/r01_*.md):
//
//   * the shared-memory crossbar delivers one 32-bit word per lane per cycle per SM, whatever
//     the instruction: a 128-bit LDS/STS is four wavefronts (one per quarter-warp) even when
//     every lane reads the same address, and a predicated store still pays for every quarter
//     that has an active lane.  A shuffle is one wavefront per word too, but needs no store.
//     => the per-step pivot row / column exchange uses shuffles (LR + LC + 1 wavefronts per
//        tile step) instead of a shared-memory mailbox (twice that), and matrices are spread
//        over as FEW lanes as the register file allows (exchange per matrix ~ GR + GC).
//   * pivoting modes use an element-granular image with an ODD row stride, so the pivot
//     search's column walk and the permuted row gather are bank-conflict free; the price is
//     scalar instead of 128-bit shared-memory instructions, which cost the same wavefronts.
//   * pivot_mode none keeps a 128-bit padded image (helpers in lub_fast.cuh).
#pragma once
#include "lub_fast.cuh"

namespace lub {

// worst bank multiplicity of one warp-wide scalar access where lane (ml, gr, gc) touches element
// offset ml*ms + gr*p + gc*cm; ew = words per element (64-bit elements are served per half-warp)
constexpr int lane_conflict(int p, int ms, int gr_n, int gc_n, int cm, int ew) {
    const int g = gr_n * gc_n, slots = 32 / ew;
    int worst = 0;
    for (int half = 0; half < ew; ++half) {
        int cnt[32] = {};
        for (int lane = half * (32 / ew); lane < (half + 1) * (32 / ew); ++lane) {
            const int ml = lane / g, gg = lane % g, gr = gg / gc_n, gc = gg % gc_n;
            const int s = (ml * ms + gr * p + gc * cm) % slots;
            if (++cnt[s] > worst) worst = cnt[s];
        }
    }
    return worst;
}
struct ScStrides { int p, pad; };
constexpr ScStrides pick_sc_strides(int n, int gr, int gc, int cm, int ew) {
    const int slots = 32 / ew;
    int best = 1 << 30;
    ScStrides s{n | 1, 0};
    for (int p = n | 1; p <= (n | 1) + 8; p += 2) {
        for (int pad = 0; pad < slots; ++pad) {
            const int c = lane_conflict(p, n * p + pad, gr, gc, cm, ew) * 4096 + (p - n) * 64 + pad;
            if (c < best) { best = c; s = ScStrides{p, pad}; }
        }
    }
    return s;
}

#ifndef LUB_V3_VECPIV
#define LUB_V3_VECPIV 1
#endif
#ifndef LUB_V3_VECPIV_MIN_N
#define LUB_V3_VECPIV_MIN_N 16
#endif
#ifndef LUB_V3_DENSE_EVEN
#define LUB_V3_DENSE_EVEN 1
#endif
#ifndef LUB_V3_DENSE_MIN_N
#define LUB_V3_DENSE_MIN_N 9
#endif

template <typename T, int N, int GR, int GC, int MODE>
struct V3Layout {
    static constexpr int ES = sizeof(T);
    static constexpr int EW = ES / 4;
    static constexpr int EPV = 16 / ES;
    static constexpr int CHV = (N % EPV == 0) ? EPV : ((EPV == 4 && N % 2 == 0) ? 2 : 1);
    // The row-wise pivot search reads whole rows at static addresses, so it can use the 16-byte
    // image of the no-pivot path (vector staging: 4x fewer LDS/STS instructions); the position-wise
    // search walks columns through a dynamic row and needs the element-granular odd-stride image.
    // (only when the 16-byte chunks split evenly over the lane columns: otherwise the chunk
    // granularity pads LC and the extra FMA work costs more than the staging saves -- N=24: +12 %)
    // (N = 16 included: 0.87 -> 0.66 ms in the pivot modes; N = 8 is mixed and keeps the group search)
    static constexpr bool VECPIV = (LUB_V3_VECPIV != 0) && MODE != kModeNone && N >= LUB_V3_VECPIV_MIN_N && CHV == EPV &&
                                   ((N / EPV) % GC) == 0 && rowwise_prepass_ok(N, MODE);
    // Odd N: the dense image (row stride N) already has an odd stride, so the pivot modes stage it with the
    // plain 128-bit span copy instead of the element scatter (which costs ~700 instructions per matrix at
    // N = 31 and 3-way conflicts on its stores, profiles/r01_tune_v6.md section 5).  N = 2 mod 4: the
    // dense stride gives 2-way conflicts on the column walks of the pivot search, still cheaper than the
    // scatter (N = 30 parallel 4.32 -> 3.19 ms); N = 0 mod 4 keeps the padded image (4-way and worse).
    static constexpr bool DENSE = (MODE != kModeNone) && !VECPIV && N >= LUB_V3_DENSE_MIN_N &&
                                  ((N % 2 == 1) || (LUB_V3_DENSE_EVEN >= 1 && N % 4 == 2) || (LUB_V3_DENSE_EVEN >= 2));
    static constexpr bool SC = (MODE != kModeNone) && !VECPIV && !DENSE;  // element-granular image, odd row stride
    static constexpr int CH = SC ? 1 : CHV;
    static constexpr int G = GR * GC;
    static_assert(G >= 1 && G <= 32 && (32 % G) == 0, "G must divide 32");
    static constexpr int MPW = 32 / G;
    static constexpr int CPR = N / CH;
    static constexpr int CPL = cdiv_(CPR, GC);  // chunks per lane: lane-col gc holds chunks [gc*CPL, (gc+1)*CPL)
    static constexpr int LC = CPL * CH;
    static constexpr int LR = cdiv_(N, GR);     // rows per lane, cyclic over GR
    static constexpr bool ROWVEC = !SC && (CH == EPV);
    static constexpr int SLOTS = 32 / EW;
    // element-granular image: odd row stride P (conflict-free column walks for the pivot search)
    // and matrix padding, searched together so that the register load -- lane (ml, gr, gc) reads
    // element ml*MS + gr*P + gc*LC -- spreads the warp over as many banks as possible
    // (2-way conflicts there cost ~11 % of the LSU time, profiles/r01_prof_headline_*.md)
    static constexpr ScStrides SCS = pick_sc_strides(N, GR, GC, LC, EW);
    static constexpr int P = SC ? SCS.p : (ROWVEC ? N + pick_row_pad<T, N>() : N);
    static constexpr int MPAD = SC ? SCS.pad : (ROWVEC ? pick_mat_pad<T, N, P, MPW>() : 0);
    static constexpr int MS = N * P + MPAD;
    static constexpr bool ALIGNED = ((MPW * N * N * ES) % 16) == 0;  // every tile span starts on 16 bytes
    static constexpr int IMG_BYTES = roundup_(MPW * MS * ES, 16) + 16;
    static constexpr int PERM_BYTES = (MODE != kModeNone) ? roundup_(MPW * N * 4, 16) : 0;
    static constexpr int WARP_BYTES = IMG_BYTES + PERM_BYTES;
    // prefetching kernel: + a one-matrix output buffer (pivot modes only)
    static constexpr int OUT_BYTES = (MODE != kModeNone) ? roundup_(MS * ES, 16) : 0;
    static constexpr int WARP_BYTES_PF = IMG_BYTES + OUT_BYTES + PERM_BYTES;
    static constexpr int WARP_BYTES_PFD = 2 * IMG_BYTES + PERM_BYTES;  // dense image, double-buffered (PFD)
    static constexpr int HEADER_BYTES = 64;
    static constexpr int CPR16 = N * ES / 16;
    static constexpr int RPAD16 = (P - N) * ES / 16, MPAD16 = MPAD * ES / 16;
};

// image offset (elements) of flat element f of a tile span, element-granular image
template <typename L, int N>
__device__ __forceinline__ int sc_off(int f) {
    const int rg = f / N;  // row index within the tile (constant divisor)
    int o = f + rg * (L::P - N);
    if (L::MPAD != 0) o += (rg / N) * L::MPAD;
    return o;
}

template <typename T, typename L, int N>
__device__ __forceinline__ void copy_in_scatter(T* __restrict__ img, const T* __restrict__ src, int total, int lane) {
    constexpr int EPV = L::EPV;
    int nhead = 0;
    if (!L::ALIGNED) {
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        nhead = mis ? (int)((16u - mis) / sizeof(T)) : 0;
        if (nhead > total) nhead = total;
        if (lane < nhead) img[sc_off<L, N>(lane)] = src[lane];
    }
    const int nvec = (total - nhead) / EPV;
    const T* s = src + nhead;
    auto put = [&](int q, uint4 v) {
        const T* e = reinterpret_cast<const T*>(&v);
        const int f0 = nhead + q * EPV;
        if (L::ALIGNED && (N % EPV) == 0) {
            const int o = sc_off<L, N>(f0);  // a 16-byte chunk never straddles a row
#pragma unroll
            for (int w = 0; w < EPV; ++w) img[o + w] = e[w];
        } else {
#pragma unroll
            for (int w = 0; w < EPV; ++w) img[sc_off<L, N>(f0 + w)] = e[w];
        }
    };
    int q = lane;
    for (; q + 96 < nvec; q += 128) {
        const uint4 v0 = ld_stream16(s + (size_t)q * EPV), v1 = ld_stream16(s + (size_t)(q + 32) * EPV);
        const uint4 v2 = ld_stream16(s + (size_t)(q + 64) * EPV), v3 = ld_stream16(s + (size_t)(q + 96) * EPV);
        put(q, v0); put(q + 32, v1); put(q + 64, v2); put(q + 96, v3);
    }
    for (; q < nvec; q += 32) put(q, ld_stream16(s + (size_t)q * EPV));
    const int tb = nhead + nvec * EPV;
    if (lane < total - tb) img[sc_off<L, N>(tb + lane)] = src[tb + lane];
}

template <typename T, typename L, int N>
__device__ __forceinline__ void copy_out_gather(T* __restrict__ dst, const T* __restrict__ img, int total, int lane) {
    constexpr int EPV = L::EPV;
    int nhead = 0;
    if (!L::ALIGNED) {
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u);
        nhead = mis ? (int)((16u - mis) / sizeof(T)) : 0;
        if (nhead > total) nhead = total;
        if (lane < nhead) dst[lane] = img[sc_off<L, N>(lane)];
    }
    const int nvec = (total - nhead) / EPV;
    T* d = dst + nhead;
    for (int q = lane; q < nvec; q += 32) {
        uint4 v;
        T* e = reinterpret_cast<T*>(&v);
        const int f0 = nhead + q * EPV;
        if (L::ALIGNED && (N % EPV) == 0) {
            const int o = sc_off<L, N>(f0);
#pragma unroll
            for (int w = 0; w < EPV; ++w) e[w] = img[o + w];
        } else {
#pragma unroll
            for (int w = 0; w < EPV; ++w) e[w] = img[sc_off<L, N>(f0 + w)];
        }
        st_stream16(d + (size_t)q * EPV, v);
    }
    const int tb = nhead + nvec * EPV;
    if (lane < total - tb) dst[tb + lane] = img[sc_off<L, N>(tb + lane)];
}

// Dense span copy global -> image with cp.async (no registers, the warp does not wait): 16-byte chunks, and
// 4-byte pieces for a head / tail that is not a chunk multiple (odd-N tiles start 8 bytes off every other
// tile; a partial last tile).  sizeof(T) is a multiple of 4, the batch pointer of sizeof(T).
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
template <typename T>
__device__ __forceinline__ void copy_in_dense_async(unsigned char* __restrict__ buf, const T* __restrict__ src, int total, int lane) {
    // the image starts at buf + (src & 15), so that global and shared 16-byte chunks line up (as copy_in does)
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
    const int bytes = total * (int)sizeof(T);
    int head = mis ? (int)(16u - mis) : 0;
    if (head > bytes) head = bytes;
    const unsigned char* s = reinterpret_cast<const unsigned char*>(src);
    unsigned char* img = buf + mis;
    if (lane * 4 < head) cp_async4(img + lane * 4, s + lane * 4);
    const int nvec = (bytes - head) / 16;
    for (int q = lane; q < nvec; q += 32) cp_async16(img + head + q * 16, s + head + q * 16);
    for (int b = head + nvec * 16 + lane * 4; b < bytes; b += 128) cp_async4(img + b, s + b);
}

// BSYNC: one block barrier per tile (see the loop).
// In-register Gauss-Jordan on the LR x LC block of every lane: row k and column k travel by
// shuffles, rows are scaled by 1/pivot only at the end (dinv), column k is overwritten with the
// multipliers as it is cleared (in-place inverse).  Shared by lub_v3_kernel and lub_tma_kernel.
// (Tried and dropped: doing the lane-dependent fix-ups with 0/1 masks on the FMA pipe instead of selects --
// bit-identical results, but the compiler rebuilds the FFMA2 register pairs with extra MOVs and the
// kernel gets 4 % slower, profiles/r01_tune_v6.md.)
template <typename T, int N, int GR, int GC, int CH, int CPL, int LR, int LC>
__device__ __forceinline__ void gj_eliminate(T (&a)[LR][LC], T (&dinv)[LR], int gr, int gc, int grp_base) {
    constexpr int G = GR * GC;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int gro = k % GR, lk = k / GR;
        const int cj = k / CH, gco = cj / CPL, ck = (cj % CPL) * CH + (k % CH);
        const bool own_row = (GR == 1) || (gr == gro);
        const bool own_col = (GC == 1) || (gc == gco);
        T r[LC], c[LR];
#pragma unroll
        for (int lj = 0; lj < LC; ++lj)
            r[lj] = (GR > 1) ? shfl_t(a[lk][lj], grp_base + gro * GC + gc) : a[lk][lj];
#pragma unroll
        for (int li = 0; li < LR; ++li)
            c[li] = (GC > 1) ? shfl_t(a[li][ck], grp_base + gr * GC + gco) : a[li][ck];
        // straight from the owner (not via r[ck]): all shuffles of a step leave in one batch
        const T pv = (G > 1) ? shfl_t(a[lk][ck], grp_base + gro * GC + gco) : a[lk][ck];
        const T rinv = rcp_t(pv);
        set_if(own_col, r[ck], T(1));
        T nf[LR];
#pragma unroll
        for (int li = 0; li < LR; ++li) nf[li] = -(c[li] * rinv);
        set_if(own_row, nf[lk], T(0));
        const T diag = sel_t(own_row, T(1), T(0));
#pragma unroll
        for (int li = 0; li < LR; ++li) set_if(own_col, a[li][ck], (li == lk) ? diag : T(0));
#pragma unroll
        for (int li = 0; li < LR; ++li) row_update<LC>(a[li], r, nf[li]);
        set_if(own_row, dinv[lk], rinv);
    }
}

// ---- lean step (round 2) ----------------------------------------------------------------------
// Same arithmetic as gj_eliminate, bit for bit, with the lane-dependent fix-ups re-planned around what
// the B200 sub-partition actually charges (profiles/r02_issue_costs.md): an FSEL / MOV / predicated MOV
// runs on the 16-lane ALU pipe (2 dispatch cycles) while a predicated FMUL runs on the FMA pipe (1 cycle);
// and clearing column k BEFORE the rank-1 update overwrites a register the pending column shuffle still
// reads, so the compiler renamed the pair and paid a MOV for the partner (8-9 MOVs per step).  Here
//   * column k is not cleared at all: the owners of the column let the FFMA2 run over it and take their
//     multiplier afterwards (a[li][ck] = nf[li]), an in-place write whose inputs have long arrived;
//   * that copy, and the copy of 1/pivot into dinv, is a predicated multiplication by an opaque 1.0
//     (a __constant__ the assembler cannot fold): FMA pipe, exact for every value;
//   * r[ck] = 1 on the column owners is no longer needed (their column ck is overwritten anyway).
// Per step and warp: 17 SHFL + 8 FMUL + 9 predicated FMUL + 2 predicated FFMA + 3 (reciprocal) besides the
// 32 FFMA2, against 17 + 8 + ~11 FSEL + ~9 MOV + 3 before.  (The same fix-ups as in-place FSELs on the ALU pipe:
// 2.52 ms against 2.33 ms on the headline, profiles/r02_tune_headline.jsonl -- the MOVs come back.)
static __constant__ float kLubOne = 1.0f;
static __constant__ float kLubZero = 0.0f;

__device__ __forceinline__ void pset_fma(bool p, float& x, float v, float one) {
    asm("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q mul.rn.f32 %0, %2, %3; }" : "+f"(x) : "r"((int)p), "f"(v), "f"(one));
}
__device__ __forceinline__ void pset_fma(bool p, double& x, double v, double) { set_if(p, x, v); }
// x = x * zero + c on the lanes where p holds (c = 0 or 1): a constant written by an FFMA that depends on x,
// so the assembler cannot hoist it out of the step and turn the write back into an FSEL.  Exact for every
// finite x; a non-finite x (zero pivot: the matrix is singular and the result inf/NaN as in the reference,
// SURVEY Q7) stays non-finite.
__device__ __forceinline__ void pconst_fma(bool p, float& x, float zero, float c) {
    asm("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q fma.rn.f32 %0, %0, %2, %3; }" : "+f"(x) : "r"((int)p), "f"(zero), "f"(c));
}
__device__ __forceinline__ void pconst_fma(bool p, double& x, double, double c) { set_if(p, x, c); }

// One elimination step (pivot k); k is a compile-time constant after unrolling.
template <typename T, int N, int GR, int GC, int CH, int CPL, int LR, int LC>
__device__ __forceinline__ void gj_step_lean(T (&a)[LR][LC], T (&dinv)[LR], const int k, int gr, int gc, int grp_base, T one, T zero) {
    constexpr int G = GR * GC;
    const int gro = k % GR, lk = k / GR;
    const int cj = k / CH, gco = cj / CPL, ck = (cj % CPL) * CH + (k % CH);
    const bool own_row = (GR == 1) || (gr == gro);
    const bool own_col = (GC == 1) || (gc == gco);
    T r[LC], c[LR];
#pragma unroll
    for (int lj = 0; lj < LC; ++lj) r[lj] = (GR > 1) ? shfl_t(a[lk][lj], grp_base + gro * GC + gc) : a[lk][lj];
#pragma unroll
    for (int li = 0; li < LR; ++li) c[li] = (GC > 1) ? shfl_t(a[li][ck], grp_base + gr * GC + gco) : a[li][ck];
    const T pv = (G > 1) ? shfl_t(a[lk][ck], grp_base + gro * GC + gco) : a[lk][ck];
    const T rinv = rcp_t(pv);
    T nf[LR];
#pragma unroll
    for (int li = 0; li < LR; ++li) nf[li] = -(c[li] * rinv);
    if (GR == 1) nf[lk] = T(0);
    else pconst_fma(own_row, nf[lk], zero, zero);
#pragma unroll
    for (int li = 0; li < LR; ++li) row_update<LC>(a[li], r, nf[li]);
#pragma unroll
    for (int li = 0; li < LR; ++li) {
        if (GC == 1) a[li][ck] = nf[li];
        else pset_fma(own_col, a[li][ck], nf[li], one);
    }
    if (G == 1) a[lk][ck] = T(1);
    else pconst_fma(own_row && own_col, a[lk][ck], zero, one);
    if (GR == 1) dinv[lk] = rinv;
    else pset_fma(own_row, dinv[lk], rinv, one);
}

template <typename T, int N, int GR, int GC, int CH, int CPL, int LR, int LC>
__device__ __forceinline__ void gj_eliminate_lean(T (&a)[LR][LC], T (&dinv)[LR], int gr, int gc, int grp_base) {
    const T one = (T)kLubOne, zero = (T)kLubZero;
#pragma unroll
    for (int k = 0; k < N; ++k) gj_step_lean<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, k, gr, gc, grp_base, one, zero);
}

// PF (16-byte image layouts): the tile image is free again as soon as the registers are loaded, so
// the NEXT tile is fetched into it with cp.async while this one is eliminated -- no warp waits on
// HBM for its input.  The results then leave through a separate one-matrix output buffer, one
// matrix of the tile at a time (pivot modes: the column scatter needs shared memory), or straight
// from the registers (no pivoting: every lane owns whole 32-byte sectors of its rows).
// PFD (dense image: V3Layout::DENSE in the pivot modes, any N that is not a multiple of 4 without pivoting;
// any alignment of the tile spans): two images per warp; the next tile is
// fetched with cp.async into the idle one while this tile is searched, eliminated and written back from
// the other -- the pivot modes need their image until the very end (column scatter), so the in-place
// prefetch of PF does not apply.
// LEAN: the round-2 elimination step (gj_eliminate_lean above); MAXT: block size the kernel is compiled for.
template <typename T, int N, int GR, int GC, int MODE, int MINB = 1, bool BSYNC = true, bool PF = false, bool PFD = false, bool LEAN = false,
          int MAXT = kMaxThreads>
__global__ void __launch_bounds__(MAXT, MINB)
lub_v3_kernel(T* __restrict__ A, int32_t* __restrict__ piv, long long batch) {
    using L = V3Layout<T, N, GR, GC, MODE>;
    static_assert(!PF || L::ROWVEC, "prefetch needs the 16-byte image");
    static_assert(!PFD || (!L::SC && !L::ROWVEC && !PF), "double-buffered prefetch: dense image");
    constexpr int G = L::G, MPW = L::MPW, LR = L::LR, LC = L::LC, CH = L::CH, CPL = L::CPL, CPR = L::CPR;
    constexpr int P = L::P, MS = L::MS;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    int8_t* slot_rank = reinterpret_cast<int8_t*>(smem_raw);
    unsigned char* wbase = smem_raw + L::HEADER_BYTES + (size_t)warp * (PF ? L::WARP_BYTES_PF : (PFD ? L::WARP_BYTES_PFD : L::WARP_BYTES));
    unsigned char* obase = wbase + L::IMG_BYTES;  // PF: one-matrix output buffer; PFD: the second image
    int* perm_all = reinterpret_cast<int*>(wbase + L::IMG_BYTES + (PF ? L::OUT_BYTES : (PFD ? L::IMG_BYTES : 0)));
    int cur = 0;  // PFD: which image holds the tile being worked on

    if (MODE == kModeParallel) {
        if (threadIdx.x < N) slot_rank[threadIdx.x] = (int8_t)tree_slot_rank(threadIdx.x, N);
        __syncthreads();
    }
    // every entry of perm[] is a row index from the start: a search that NaN inputs derail may skip entries, never invent one
    if (MODE != kModeNone) {
        for (int x = threadIdx.x & 31; x < MPW * N; x += 32) perm_all[x] = x % N;
        __syncwarp();
    }

    const int g = lane % G;
    const int ml = lane / G;
    const int gr = g / GC;
    const int gc = g % GC;
    const int grp_base = ml * G;

    const long long ntiles = (batch + MPW - 1) / MPW;
    const long long tstride = (long long)gridDim.x * nwarps;
    if constexpr (PF) {
        const long long t0 = (long long)blockIdx.x * nwarps + warp;
        if (t0 < ntiles) {
            const long long first0 = t0 * MPW;
            const int nm0 = (batch - first0 < MPW) ? (int)(batch - first0) : MPW;
            copy_in_padded_async<T, L, N>(wbase, A + first0 * (long long)(N * N), nm0 * N * L::CPR16, lane);
        }
        cp_async_commit();
    }
    if constexpr (PFD) {
        const long long t0 = (long long)blockIdx.x * nwarps + warp;
        if (t0 < ntiles) {
            const long long first0 = t0 * MPW;
            const int nm0 = (batch - first0 < MPW) ? (int)(batch - first0) : MPW;
            copy_in_dense_async<T>(wbase, A + first0 * (long long)(N * N), nm0 * N * N, lane);
        }
        cp_async_commit();
    }
#pragma unroll 1
    for (long long tbase = (long long)blockIdx.x * nwarps; tbase < ntiles; tbase += (long long)gridDim.x * nwarps) {
        // Re-align the block's warps once per tile: they all run the same ~50 KB of straight-line
        // code, and warps that drift apart thrash the instruction caches (no_instruction stalls).
        if (BSYNC) __syncthreads();
        const long long tile = tbase + warp;
        if (tile >= ntiles) continue;
        const long long first = tile * MPW;
        const int nm = (batch - first < MPW) ? (int)(batch - first) : MPW;
        T* gspan = A + first * (long long)(N * N);
        T* img;
        if constexpr (PF) {
            img = reinterpret_cast<T*>(wbase);
            cp_async_wait<0>();  // this tile, requested one round ago
        } else if constexpr (PFD) {
            img = reinterpret_cast<T*>((cur ? obase : wbase) + (unsigned)(reinterpret_cast<uintptr_t>(gspan) & 15u));
            cp_async_wait<0>();  // this tile, requested one round ago
            __syncwarp();        // ... by every lane; and the other image's write-back (last round) is done
            const long long nxt = tile + tstride;
            if (nxt < ntiles) {
                const long long firstn = nxt * MPW;
                const int nmn = (batch - firstn < MPW) ? (int)(batch - firstn) : MPW;
                copy_in_dense_async<T>(cur ? wbase : obase, A + firstn * (long long)(N * N), nmn * N * N, lane);
            }
            cp_async_commit();
            cur ^= 1;
        } else if constexpr (L::SC) {
            img = reinterpret_cast<T*>(wbase);
            copy_in_scatter<T, L, N>(img, gspan, nm * N * N, lane);
        } else if constexpr (L::ROWVEC) {
            img = reinterpret_cast<T*>(wbase);
            copy_in_padded<T, L, N>(wbase, gspan, nm * N * L::CPR16, lane);
        } else {
            const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(gspan) & 15u);
            img = reinterpret_cast<T*>(wbase + mis);
            copy_in<T>(img, gspan, nm * N * N, nm * N * N, lane);
        }
        __syncwarp();

        T* mimg = img + ml * MS;
        int* perm = perm_all + ml * N;
        if (MODE != kModeNone) {
            if (N > 16 || L::VECPIV) {
                constexpr int MI = (MPW < 4) ? MPW : 4;
#pragma unroll 1
                for (int m = 0; m < MPW; m += MI)
                    prepass_warp<T, N, MODE, P, MS, MI, false, L::VECPIV>(img + m * MS, perm_all + m * N, slot_rank, lane);
            } else {
                prepass_group<T, N, G, MODE, P>(mimg, perm, slot_rank, g);
            }
            __syncwarp();
        }

        // ---- registers <- image: rows permuted, LR x LC block per lane ---------------------
        T a[LR][LC];
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
            int prow = i;
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

        if constexpr (PF) {
            __syncwarp();  // every lane has its block: the image can take the next tile
            const long long nxt = tile + tstride;
            if (nxt < ntiles) {
                const long long firstn = nxt * MPW;
                const int nmn = (batch - firstn < MPW) ? (int)(batch - firstn) : MPW;
                copy_in_padded_async<T, L, N>(wbase, A + firstn * (long long)(N * N), nmn * N * L::CPR16, lane);
            }
            cp_async_commit();
        }

        // ---- Gauss-Jordan with deferred row scaling; exchange by shuffles ---------------------
        T dinv[LR];
#pragma unroll
        for (int li = 0; li < LR; ++li) dinv[li] = T(0);
        if (LEAN) gj_eliminate_lean<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, gr, gc, grp_base);
        else gj_eliminate<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, gr, gc, grp_base);

        // ---- scale by 1/pivot, undo the row permutation as a column scatter -------------------
        __syncwarp();
#pragma unroll
        for (int li = 0; li < LR; ++li) {
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) a[li][lj] *= dinv[li];
        }
        if constexpr (PF && MODE == kModeNone) {
            T* gm = gspan + (size_t)ml * (N * N);
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (ml < nm) && ((li * GR + GR - 1 < N) || (i < N));
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int cq = gc * CPL + q;
                    if (rok && ((GC * CPL <= CPR) || (cq < CPR))) st_vec<T, CH>(gm + i * N + cq * CH, &a[li][q * CH]);
                }
            }
        } else if constexpr (PF) {
            int pcol[LC];
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) {
                const int j = gc * LC + lj;
                pcol[lj] = ((GC * LC <= N) || (j < N)) ? perm[j] : -1;
            }
            T* oimg = reinterpret_cast<T*>(obase);
#pragma unroll 1
            for (int m = 0; m < nm; ++m) {
                if (ml == m) {
#pragma unroll
                    for (int li = 0; li < LR; ++li) {
                        const int i = li * GR + gr;
                        const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
                        for (int lj = 0; lj < LC; ++lj)
                            if (rok && pcol[lj] >= 0) oimg[i * P + pcol[lj]] = a[li][lj];
                    }
                }
                __syncwarp();
                copy_out_padded<T, L, N>(gspan + (size_t)m * (N * N), obase, N * L::CPR16, lane);
                __syncwarp();
            }
        } else if constexpr (MODE == kModeNone) {
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
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
                for (int lj = 0; lj < LC; ++lj)
                    if (rok && pcol[lj] >= 0) mimg[i * P + pcol[lj]] = a[li][lj];
            }
        }
        __syncwarp();
        if constexpr (PF) {
        } else if constexpr (L::SC) copy_out_gather<T, L, N>(gspan, img, nm * N * N, lane);
        else if constexpr (L::ROWVEC) copy_out_padded<T, L, N>(gspan, wbase, nm * N * L::CPR16, lane);
        else copy_out<T>(gspan, img, nm * N * N, lane);
        int32_t* pivp = piv;
        asm volatile("" : "+l"(pivp));  // opaque: keeps the compiler from cloning the whole tile loop on piv == NULL
        if (pivp != nullptr) {
            int32_t* pdst = pivp + first * N;
            for (int e = lane; e < nm * N; e += 32)
                pdst[e] = (MODE != kModeNone) ? perm_all[e] : (e % N);
        }
        __syncwarp();
    }
}

}  // namespace lub
