// lub_tma2.cuh (tuning harness only; measured SLOWER than lub_tma_kernel<OPT = lean | DB>, profiles/r02_headline_floor.md) -- fused form of the TMA-staged kernel for the pivoting modes of the fp32 sizes whose pivot search
// is the row-wise one (serial pivoting; parallel pivoting with N a power of two -- the headline, N = 32).
//
// What round 1's kernel (lub_tma.cuh) left on the table, measured (profiles/r02_headline_floor.md):
//   * a warp ran its phases one after the other -- wait for the tile, search the pivots (a chain of 31 dependent
//     CREDUX steps, ~1400 cycles of latency that issues only ~370 instructions), load registers, eliminate, scatter,
//     store -- and relied on the other 3 warps of its scheduler to fill the holes.  The elimination alone keeps the
//     FMA pipe ~75 % busy, so whenever two warps of a scheduler sat in a latency phase the pipe idled: the kernel
//     took 760 SM-cycles per matrix where the elimination alone needs 500.
// Here
//   * two images per warp (12 warps x 16 KB): the tile after this one is requested as soon as the image it will
//     land in is free, i.e. right after this tile's registers are loaded;
//   * the pivot search of the NEXT tile is issued inside the elimination of THIS one, in the same straight-line
//     block: the search's long dependency chain fills the issue slots the elimination leaves, and no warp ever
//     sits in a search-only phase (except for its very first tile);
//   * the lean elimination step (lub_v3.cuh: fix-ups after the update, on the FMA pipe).
// Same results as lub_tma_kernel bit for bit (same search, same arithmetic in the same order).
#pragma once
#include "../../matrixinversion_b200/csrc/lub_tma.cuh"

namespace lub {

// Row-wise pivot search (prepass_rowwise_swz, fp32 form) cut into resumable pieces: the state of the two matrices
// of a tile lives in registers between the steps.
template <int MI>
struct RowSearch {
    float alive[MI], when[MI], x[MI][4];
};

template <int N, int MI>
__device__ __forceinline__ void rowsearch_init(RowSearch<MI>& s, int lane) {
#pragma unroll
    for (int m = 0; m < MI; ++m) { s.alive[m] = (lane < N) ? 1.0f : 0.0f; s.when[m] = 0.0f; }
}

// step k (compile-time after unrolling) of the search on the MI matrices starting at tile row row0
template <int N, int MI>
__device__ __forceinline__ void rowsearch_step(RowSearch<MI>& s, const int k, const unsigned char* img, int row0, int row) {
    constexpr int RB = (N * 4 + 127) / 128 * 128;
    if ((k % 4) == 0) {
#pragma unroll
        for (int m = 0; m < MI; ++m)
            ld_vec<float, 4>(reinterpret_cast<const float*>(img + swz_byte<RB>(row0 + m * N + row, (k / 4) << 4)), s.x[m]);
    }
    float v[MI], mx[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) v[m] = s.x[m][k % 4] * s.alive[m];
#pragma unroll
    for (int m = 0; m < MI; ++m) mx[m] = warp_max_abs(v[m]);
#pragma unroll
    for (int m = 0; m < MI; ++m) {
        const float hit = (fabsf(v[m]) == mx[m]) ? 1.0f : 0.0f;
        s.when[m] = fmaf(hit, (float)k, s.when[m]);
        s.alive[m] = fmaf(-hit, s.alive[m], s.alive[m]);
    }
}

// write the permutation vectors (or redo a matrix with equal maxima exactly, see prepass_rowwise_swz)
template <int N, int MODE, int MI>
__device__ __forceinline__ void rowsearch_finish(const RowSearch<MI>& s, const unsigned char* img, int row0, int* perm0,
                                                 const int8_t* slot_rank, int lane) {
#pragma unroll
    for (int m = 0; m < MI; ++m) {
        const bool ok = __popc(__ballot_sync(0xffffffffu, s.alive[m] != 0.0f)) == 1;  // warp-uniform
        if (ok) {
            if (lane < N) perm0[m * N + ((s.alive[m] != 0.0f) ? (N - 1) : (int)s.when[m])] = lane;
        } else {
            prepass_exact_swz<float, N, MODE>(img, row0 + m * N, perm0 + m * N, slot_rank, lane);
        }
    }
}

template <int N, int GR, int GC, int MODE>
struct Tma2Layout : TmaLayout<float, N, GR, GC, MODE> {
    using B = TmaLayout<float, N, GR, GC, MODE>;
    static_assert(B::MPW == 2, "the fused search handles the two matrices of a tile in lock step");
    static constexpr int WARP_BYTES = 2 * B::IMG_BYTES + 2 * B::PERM_BYTES + 16;
    static constexpr int smem_bytes(int warps) { return 1024 + warps * WARP_BYTES + B::HEADER_BYTES; }
};

// K1: elimination steps issued before the warp looks for the next tile (the time its TMA load gets to land).
template <int N, int GR, int GC, int MODE, int MAXT = 384, int K1 = 6>
__global__ void __launch_bounds__(MAXT, 1)
lub_tma2_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ A, int32_t* __restrict__ piv, long long batch) {
    static_assert(MODE != kModeNone && RowwiseOk<N, MODE>::value, "row-wise pivot search only");
    using T = float;
    using L = Tma2Layout<N, GR, GC, MODE>;
    constexpr int G = L::G, MPW = L::MPW, LR = L::LR, LC = L::LC, CH = L::CH, CPL = L::CPL;
    constexpr int RB = L::RB, ES = L::ES, IMG = L::IMG_BYTES;
    extern __shared__ unsigned char smem_dyn[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    // carve: [images of all warps, 1 KB aligned][perm x 2][mbarriers x 2] per warp, then the slot ranks
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* img0 = base + (size_t)warp * (2 * IMG);
    unsigned char* after = base + (size_t)nwarps * (2 * IMG);
    int* perm0 = reinterpret_cast<int*>(after + (size_t)warp * (2 * L::PERM_BYTES));
    unsigned long long* bar0 = reinterpret_cast<unsigned long long*>(after + (size_t)nwarps * (2 * L::PERM_BYTES)) + 2 * warp;
    int8_t* slot_rank = reinterpret_cast<int8_t*>(after + (size_t)nwarps * (2 * L::PERM_BYTES + 16));

    if (MODE == kModeParallel && threadIdx.x < N) slot_rank[threadIdx.x] = (int8_t)tree_slot_rank(threadIdx.x, N);
    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 1, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int g = lane % G, ml = lane / G, gr = g / GC, gc = g % GC, grp_base = ml * G;
    const int trow0 = ml * N;
    const int srow = (lane < N) ? lane : 0;
    const long long ntiles = (batch + MPW - 1) / MPW;
    const long long tstride = (long long)gridDim.x * nwarps;
    long long tile = (long long)blockIdx.x * nwarps + warp;
    if (tile >= ntiles) return;
    const T one = kLubOne, zero = kLubZero;

    // ---- first tile of this warp: load, wait, search on its own ----
    if (lane == 0) {
        mbar_expect_tx(bar0, (unsigned)IMG);
        tma_load_tile<L::LPR>(img0, &tmap, bar0, (int)(tile * MPW));
    }
    mbar_wait(bar0, 0u);
    prepass_rowwise_swz<T, N, MODE, 2>(img0, 0, perm0, slot_rank, lane);
    __syncwarp();

    unsigned it = 0;
#pragma unroll 1
    for (;; ++it) {
        const unsigned cur = it & 1u, oth = cur ^ 1u;
        unsigned char* img = img0 + cur * IMG;
        unsigned char* img_next = img0 + oth * IMG;
        int* perm_t = perm0 + cur * (MPW * N);      // this tile's permutation vectors (both matrices)
        int* perm_next = perm0 + oth * (MPW * N);
        const int* perm = perm_t + ml * N;
        const long long first = tile * MPW;
        const int nm = (batch - first < MPW) ? (int)(batch - first) : MPW;
        const long long nxt = tile + tstride;
        const bool has_next = nxt < ntiles;  // warp-uniform

        // ---- registers <- image: rows permuted, LR x LC block per lane ----
        T a[LR][LC];
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
            const int prow = rok ? perm[i] : 0;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                if (rok) {
                    ld_vec<T, CH>(reinterpret_cast<const T*>(img + swz_byte<RB>(trow0 + prow, (gc * CPL + q) << 4)), &a[li][q * CH]);
                } else {
#pragma unroll
                    for (int w = 0; w < CH; ++w) a[li][q * CH + w] = T(0);
                }
            }
        }
        // ---- request the next tile into the other image (the tile before this one left it through a bulk store) ----
        if (lane == 0 && has_next) {
            tma_store_wait_read();
            mbar_expect_tx(bar0 + oth, (unsigned)IMG);
            tma_load_tile<L::LPR>(img_next, &tmap, bar0 + oth, (int)(nxt * MPW));
        }

        T dinv[LR];
#pragma unroll
        for (int li = 0; li < LR; ++li) dinv[li] = T(0);
#pragma unroll
        for (int k = 0; k < K1; ++k) gj_step_lean<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, k, gr, gc, grp_base, one, zero);

        // ---- the rest of the elimination, with the next tile's pivot search issued in between ----
        if (has_next) mbar_wait(bar0 + oth, ((it + 1u) >> 1) & 1u);
        RowSearch<2> rs;
        rowsearch_init<N, 2>(rs, lane);
#pragma unroll
        for (int k = K1; k < N; ++k) {
            gj_step_lean<T, N, GR, GC, CH, CPL, LR, LC>(a, dinv, k, gr, gc, grp_base, one, zero);
            const int s0 = (k - K1) * (N - 1) / (N - K1), s1 = (k - K1 + 1) * (N - 1) / (N - K1);
#pragma unroll
            for (int s = s0; s < s1; ++s) rowsearch_step<N, 2>(rs, s, img_next, 0, srow);
        }
        if (has_next) rowsearch_finish<N, MODE, 2>(rs, img_next, 0, perm_next, slot_rank, lane);

        // ---- scale by 1/pivot; undo the row permutation as a column scatter; bulk store ----
#pragma unroll
        for (int li = 0; li < LR; ++li) {
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) a[li][lj] *= dinv[li];
        }
        int pcb[LC];
#pragma unroll
        for (int lj = 0; lj < LC; ++lj) {
            const int j = gc * LC + lj;
            pcb[lj] = ((GC * LC <= N) || (j < N)) ? perm[j] * ES : -1;
        }
        __syncwarp();
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
#pragma unroll
            for (int lj = 0; lj < LC; ++lj)
                if (rok && ((GC * LC <= N) || (pcb[lj] >= 0))) *reinterpret_cast<T*>(img + swz_byte<RB>(trow0 + i, pcb[lj])) = a[li][lj];
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_tile<L::LPR>(&tmap, img, (int)first);  // matrices past the batch end are clipped
            tma_store_commit();
        }
        int32_t* pivp = piv;
        asm volatile("" : "+l"(pivp));
        if (pivp != nullptr) {
            int32_t* pdst = pivp + first * N;
            for (int e = lane; e < nm * N; e += 32) pdst[e] = perm_t[e];
        }
        __syncwarp();
        if (!has_next) break;
        tile = nxt;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace lub
