// lub_v4.cuh -- fourth-generation hot-path kernel (the one the library launches for G > 1).
// Same algorithm and results as the earlier generations (see lub_kernel.cuh for the maths and
// lub_v3.cuh for the data-movement cost model); what changes is the instruction stream:
//
//   * rows AND columns are dealt to the lanes cyclically (row i -> lane-row i % GR, column j ->
//     lane-col j % GC).  Elimination step k then touches local row k / GR and local column
//     k / GC, so min(GR, GC) consecutive steps use the SAME registers and differ only in which
//     lanes own the pivot row / column -- a run-time lane id.  Those steps are executed by one
//     rolled inner loop: the unrolled code shrinks by that factor (4x for the 4 x 4 grid used at
//     N = 32), which is what removes the instruction-cache stalls Nsight Compute showed
//     (no_instruction = 19-23 % of the elimination phase's samples in profiles/r01_v3_*.md);
//   * column k of the eliminated matrix is cleared by multiplying the register PAIR that holds
//     it with a per-lane (1, 0) / (0, 1) / (1, 1) mask -- one packed multiply in place instead
//     of a select that breaks the FFMA2 register pair and costs a second move to rebuild it;
//   * the image is element-granular for every mode; row stride P (odd, so column walks are
//     conflict-free) and matrix stride are searched at compile time for the (P, MS) that makes
//     the 2-D register load hit 32 different banks in natural row order.
#pragma once
#include "lub_v3.cuh"

#ifndef LUB_V4_DENSE
#define LUB_V4_DENSE 1
#endif

namespace lub {

// worst bank multiplicity of one warp-wide scalar access where lane (ml, gr, gc) touches
// element offset ml*ms + gr*p + gc; ew = words per element (64-bit elements are served per half-warp)
constexpr int load_conflict(int p, int ms, int gr_n, int gc_n, int ew) {
    const int g = gr_n * gc_n, mpw = 32 / g, slots = 32 / ew;
    int worst = 0;
    for (int half = 0; half < ew; ++half) {
        int cnt[32] = {};
        for (int lane = half * (32 / ew); lane < (half + 1) * (32 / ew); ++lane) {
            const int ml = lane / g, gg = lane % g, gr = gg / gc_n, gc = gg % gc_n;
            if (ml >= mpw) continue;
            const int s = (ml * ms + gr * p + gc) % slots;
            if (++cnt[s] > worst) worst = cnt[s];
        }
    }
    return worst;
}

struct Strides { int p, ms; };

// even N without pivoting for which the dense image (+ prefetch) was measured to win although the conflict
// model prefers a padded layout (filled from profiles/r01_tune_late.jsonl "pfd64 even")
constexpr bool pick_dense_even(int n, int mode) { return mode == kModeNone && (n == 10 || n == 14 || n == 18 || n == 20); }

// odd_only: the pivot search walks columns (row stride must be odd to be conflict-free); without
// pivoting any stride will do and a conflict-free register load usually exists with an even one
constexpr Strides pick_strides(int n, int gr, int gc, int ew, bool odd_only) {
    const int slots = 32 / ew;
    int best = 1 << 30;
    Strides s{n | 1, n * (n | 1)};
    for (int p = odd_only ? (n | 1) : n; p < n + 2 * slots; p += odd_only ? 2 : 1) {
        for (int pad = 0; pad < slots; ++pad) {
            const int ms = n * p + pad;
            const int c = load_conflict(p, ms, gr, gc, ew) * 4096 + (p - n) * 64 + pad;
            if (c < best) { best = c; s = Strides{p, ms}; }
        }
    }
    return s;
}

template <typename T, int N, int GR, int GC, int MODE>
struct V4Layout {
    static constexpr int ES = sizeof(T);
    static constexpr int EW = ES / 4;
    static constexpr int EPV = 16 / ES;
    static constexpr int G = GR * GC;
    static_assert(G >= 1 && G <= 32 && (32 % G) == 0, "G must divide 32");
    static constexpr int MPW = 32 / G;
    static constexpr int LR = cdiv_(N, GR);  // local rows:    i = li * GR + gr
    static constexpr int LC = cdiv_(N, GC);  // local columns: j = lj * GC + gc
    static constexpr int GM = GR < GC ? GR : GC;  // steps that share one code body
    static constexpr Strides S0 = pick_strides(N, GR, GC, EW, MODE != kModeNone);
    // Dense image (row stride N, no padding): staged with the plain 128-bit span copy instead of the
    // element scatter.  Taken when it is as conflict-free for the register load as the best padded
    // layout and -- with a pivot search walking columns -- N is odd (profiles/r01_tune_v6.md, section 5).
    // (odd N: always -- measured 0..-9 % even where the conflict model prefers a padded layout)
    static constexpr bool DENSE = (LUB_V4_DENSE != 0) && (MODE == kModeNone || (N % 2 == 1)) &&
                                  ((N % 2 == 1) || LUB_V4_DENSE >= 2 || pick_dense_even(N, MODE) ||
                                   load_conflict(N, N * N, GR, GC, EW) <= load_conflict(S0.p, S0.ms, GR, GC, EW));
    static constexpr Strides S = DENSE ? Strides{N, N * N} : S0;
    static constexpr int P = S.p, MS = S.ms, MPAD = MS - N * P;
    static constexpr bool ALIGNED = ((MPW * N * N * ES) % 16) == 0;
    static constexpr int IMG_BYTES = roundup_(MPW * MS * ES, 16) + 16;
    static constexpr int PERM_BYTES = (MODE != kModeNone) ? roundup_(MPW * N * 4, 16) : 0;
    static constexpr int WARP_BYTES = IMG_BYTES + PERM_BYTES;
    static constexpr int WARP_BYTES_PFD = 2 * IMG_BYTES + PERM_BYTES;  // dense image, double-buffered (PFD)
    static constexpr int HEADER_BYTES = 64;
};

// a[j] += nf * r[j] with column `cz` first multiplied by zmask (1 keeps it, 0 clears it)
template <int LC>
__device__ __forceinline__ void row_update_masked(float (&a)[LC], const float (&r)[LC], float nf, int cz, float zmask) {
    const float2 nf2 = make_float2(nf, nf);
#pragma unroll
    for (int j = 0; j + 1 < LC; j += 2) {
        float2 acc = make_float2(a[j], a[j + 1]);
        if (j == (cz & ~1)) acc = __fmul2_rn(acc, (cz & 1) ? make_float2(1.0f, zmask) : make_float2(zmask, 1.0f));
        const float2 d = __ffma2_rn(nf2, make_float2(r[j], r[j + 1]), acc);
        a[j] = d.x;
        a[j + 1] = d.y;
    }
    if (LC & 1) {
        float acc = a[LC - 1];
        if (cz == LC - 1) acc *= zmask;
        a[LC - 1] = fmaf(nf, r[LC - 1], acc);
    }
}
template <int LC>
__device__ __forceinline__ void row_update_masked(double (&a)[LC], const double (&r)[LC], double nf, int cz, double zmask) {
#pragma unroll
    for (int j = 0; j < LC; ++j) a[j] = fma(nf, r[j], (j == cz) ? a[j] * zmask : a[j]);
}

// PFD (dense layouts): two images per warp, the next tile is fetched with cp.async into the idle one while
// this tile is worked on (see lub_v3_kernel).
template <typename T, int N, int GR, int GC, int MODE, int MINB = 1, bool BSYNC = true, bool PFD = false>
__global__ void __launch_bounds__(kMaxThreads, MINB)
lub_v4_kernel(T* __restrict__ A, int32_t* __restrict__ piv, long long batch) {
    using L = V4Layout<T, N, GR, GC, MODE>;
    static_assert(!PFD || L::DENSE, "double-buffered prefetch needs the dense image");
    constexpr int G = L::G, MPW = L::MPW, LR = L::LR, LC = L::LC, GM = L::GM, P = L::P, MS = L::MS;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    int8_t* slot_rank = reinterpret_cast<int8_t*>(smem_raw);
    unsigned char* wbase = smem_raw + L::HEADER_BYTES + (size_t)warp * (PFD ? L::WARP_BYTES_PFD : L::WARP_BYTES);
    unsigned char* obase = wbase + L::IMG_BYTES;  // PFD: the second image
    int* perm_all = reinterpret_cast<int*>(wbase + (PFD ? 2 : 1) * L::IMG_BYTES);
    int cur = 0;

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
        // Re-align the block's warps once per tile: they all run the same straight-line code, and
        // warps that drift apart thrash the instruction caches.
        if (BSYNC) __syncthreads();
        const long long tile = tbase + warp;
        if (tile >= ntiles) continue;
        const long long first = tile * MPW;
        const int nm = (batch - first < MPW) ? (int)(batch - first) : MPW;
        T* gspan = A + first * (long long)(N * N);
        T* img = reinterpret_cast<T*>(wbase);
        if constexpr (PFD) {
            img = reinterpret_cast<T*>((cur ? obase : wbase) + (unsigned)(reinterpret_cast<uintptr_t>(gspan) & 15u));
            cp_async_wait<0>();  // this tile, requested one round ago
            __syncwarp();
            const long long nxt = tile + (long long)gridDim.x * nwarps;
            if (nxt < ntiles) {
                const long long firstn = nxt * MPW;
                const int nmn = (batch - firstn < MPW) ? (int)(batch - firstn) : MPW;
                copy_in_dense_async<T>(cur ? wbase : obase, A + firstn * (long long)(N * N), nmn * N * N, lane);
            }
            cp_async_commit();
            cur ^= 1;
        } else if constexpr (L::DENSE) {  // the image is the span itself, shifted so that 16-byte chunks line up
            img = reinterpret_cast<T*>(wbase + (unsigned)(reinterpret_cast<uintptr_t>(gspan) & 15u));
            copy_in<T>(img, gspan, nm * N * N, nm * N * N, lane);
        } else {
            copy_in_scatter<T, L, N>(img, gspan, nm * N * N, lane);
        }
        __syncwarp();

        T* mimg = img + ml * MS;
        int* perm = perm_all + ml * N;
        if (MODE != kModeNone) {
            if (N > 16) {
                constexpr int MI = (MPW < 4) ? MPW : 4;
#pragma unroll 1
                for (int m = 0; m < MPW; m += MI)
                    prepass_warp<T, N, MODE, P, MS, MI>(img + m * MS, perm_all + m * N, slot_rank, lane);
            } else {
                prepass_group<T, N, G, MODE, P>(mimg, perm, slot_rank, g);
            }
            __syncwarp();
        }

        // ---- registers <- image: rows permuted, LR x LC block per lane (cyclic x cyclic) -------
        T a[LR][LC];
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
            int prow = i;
            if (MODE != kModeNone) prow = rok ? perm[i] : 0;
            const T* rowp = mimg + prow * P + gc;
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) {
                const bool ok = rok && ((lj * GC + GC - 1 < N) || (lj * GC + gc < N));
                a[li][lj] = ok ? rowp[lj * GC] : T(0);
            }
        }

        // ---- Gauss-Jordan with deferred row scaling; exchange by shuffles -----------------------
        T dinv[LR];
#pragma unroll
        for (int li = 0; li < LR; ++li) dinv[li] = T(0);
        constexpr int NSTEP = N;
#pragma unroll
        for (int kb = 0; kb < (NSTEP + GM - 1) / GM; ++kb) {
            constexpr int dummy = 0; (void)dummy;
            const int lk = (kb * GM) / GR;          // local row of rows kb*GM .. kb*GM+GM-1
            const int ck = (kb * GM) / GC;          // local column of those columns
            const int gro0 = (kb * GM) % GR, gco0 = (kb * GM) % GC;
#pragma unroll 1
            for (int s = 0; s < GM; ++s) {
                if (kb * GM + s >= NSTEP) break;
                const int gro = gro0 + s, gco = gco0 + s;   // run-time owners of pivot row / column
                const bool own_row = (GR == 1) || (gr == gro);
                const bool own_col = (GC == 1) || (gc == gco);
                const int src_row = grp_base + gro * GC + gc;   // lane holding row k for my columns
                const int src_col = grp_base + gr * GC + gco;   // lane holding column k for my rows
                T r[LC], c[LR];
#pragma unroll
                for (int lj = 0; lj < LC; ++lj) r[lj] = (GR > 1) ? shfl_t(a[lk][lj], src_row) : a[lk][lj];
#pragma unroll
                for (int li = 0; li < LR; ++li) c[li] = (GC > 1) ? shfl_t(a[li][ck], src_col) : a[li][ck];
                // straight from the owner (not via r[ck]): all shuffles of a step leave in one batch
                const T pv = (G > 1) ? shfl_t(a[lk][ck], grp_base + gro * GC + gco) : a[lk][ck];
                const T rinv = rcp_t(pv);
                // slot k now belongs to column k of the augmented identity: the broadcast row has
                // a 1 there, and the column itself is cleared (mask 0) before the update
                r[ck] = sel_t(own_col, T(1), r[ck]);
                const T zmask = sel_t(own_col, T(0), T(1));
                T nf[LR];
#pragma unroll
                for (int li = 0; li < LR; ++li) nf[li] = -(c[li] * rinv);
                nf[lk] = sel_t(own_row, T(0), nf[lk]);
#pragma unroll
                for (int li = 0; li < LR; ++li) row_update_masked<LC>(a[li], r, nf[li], ck, zmask);
                // the pivot row's own slot-k entry is the 1 of the identity column
                a[lk][ck] = sel_t(own_row && own_col, T(1), a[lk][ck]);
                dinv[lk] = sel_t(own_row, rinv, dinv[lk]);
            }
        }

        // ---- scale by 1/pivot, undo the row permutation as a column scatter -------------------
        __syncwarp();
        int pcol[LC];
#pragma unroll
        for (int lj = 0; lj < LC; ++lj) {
            const int j = lj * GC + gc;
            const bool ok = (lj * GC + GC - 1 < N) || (j < N);
            pcol[lj] = ok ? ((MODE != kModeNone) ? perm[j] : j) : -1;
        }
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
            const T sc = dinv[li];
#pragma unroll
            for (int lj = 0; lj < LC; ++lj)
                if (rok && pcol[lj] >= 0) mimg[i * P + pcol[lj]] = a[li][lj] * sc;
        }
        __syncwarp();
        if constexpr (L::DENSE) {
            copy_out<T>(gspan, img, nm * N * N, lane);
        } else {
            copy_out_gather<T, L, N>(gspan, img, nm * N * N, lane);
        }
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
