// SUPERSEDED (kept as a record, not part of the library): the two-phase pivot_mode 3 kernels (prepass_getrf + lub_bulk_kernel /
// lub_tma_kernel<kModeLapack>) beat this kernel at every size it served -- profiles/r02_mode3_twophase.md.
// lub_lapack2.cuh -- pivot_mode 3 (true partial pivoting, LAPACK getrf semantics, `ipiv` + `info`) on a 2-D lane
// grid, for N = 17..32.  Same contract as lub_lapack.cuh (the check it is built to pass
// is verifyLUwithPivoting, parallel_pivot/verify.hpp:157-242), another data layout:
//
// lub_lapack_kernel keeps lane = row, so the pivot row -- only known once column k has been updated -- is a run-time
// LANE and never a run-time register; the price is N shuffles per step and matrix (the whole pivot row travels to
// every lane).  Here a matrix lives on the 8 x 4 lane grid of the other modes (rows cyclic over 8 lane rows, columns
// blocked over 4 lane columns, LR x LC <= 4 x 8 block per lane): the exchange per step is LC + LR + 1 <= 13 shuffles.
// The pivot row then does sit at a run-time register index li_p -- but that index is WARP-UNIFORM (it comes out of
// a warp reduction), so a uniform switch over the LR <= 4 possible values picks the registers with static indices
// and no divergence.  Rows never move ("implicit pivoting"); every lane tracks, for each of its rows,
//   pos    = the position the row has in LAPACK's swapped order (ipiv is a list of position swaps, and isamax
//            breaks ties by position),
//   mystep = the step at which the row was the pivot (row k of the inverse sits in the row that was pivot at step k).
// Per step: candidates = rows not yet used as a pivot, key = |a[.][k]| of the UPDATED column k; max by CREDUX, the
// lowest position among the maxima by one REDUX.MIN that also carries the owner (lane row, register index).
// Staging: 1-D bulk copies, two images per warp (lub_bulk.cuh).
//
//   LUONLY = false: in-place inverse; Gauss-Jordan with deferred row scaling as in modes 0-2.  With rho(k) = the
//                   row that was pivot at step k, the in-place array W ends with A^-1[k][rho(k')] = W[rho(k)][k'].
//   LUONLY = true : the getrf output: P A = L U, unit-lower L below the diagonal, U on and above it, rows in final order.
#pragma once
#include "lub_bulk.cuh"
#include "lub_lapack.cuh"

namespace lub {

template <typename T, int N>
struct Lapack2Layout {
    static constexpr int GR = 8, GC = 4;
    using B = BulkLayout<T, N, GR, GC, kModeParallel>;
    static constexpr int TAB_BYTES = roundup_(N * 4, 16);  // rho[] and ipiv[] of the matrix in flight
    static constexpr int WARP_BYTES = 2 * B::IMG_BYTES + 2 * TAB_BYTES + 16;
    static constexpr int smem_bytes(int warps) { return warps * WARP_BYTES; }
};

template <typename T, int N, bool LUONLY, int MAXT = kMaxThreads, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB)
lub_lapack2_kernel(T* __restrict__ A, int32_t* __restrict__ ipiv, int32_t* __restrict__ info, long long batch) {
    using L2 = Lapack2Layout<T, N>;
    using L = typename L2::B;
    using U = typename FpBits<T>::U;
    constexpr int GR = L2::GR, GC = L2::GC, LR = L::LR, LC = L::LC, CH = L::CH, CPL = L::CPL, CPR = L::CPR;
    constexpr int P = L::P, MS = L::MS, ES = L::ES;
    static_assert(L::MPW == 1 && LR <= 4, "one matrix per warp, at most four rows per lane");
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    unsigned char* wbase = smem_raw + (size_t)warp * L2::WARP_BYTES;
    int* rho = reinterpret_cast<int*>(wbase + 2 * L::IMG_BYTES);
    int* ipiv_s = reinterpret_cast<int*>(wbase + 2 * L::IMG_BYTES + L2::TAB_BYTES);
    unsigned long long* bar0 = reinterpret_cast<unsigned long long*>(wbase + 2 * L::IMG_BYTES + 2 * L2::TAB_BYTES);

    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 1, 1); }
    for (int x = lane; x < N; x += 32) { rho[x] = x; ipiv_s[x] = x + 1; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int gr = lane / GC;
    const int gc = lane % GC;

    const long long ntiles = batch;
    const long long tstride = (long long)gridDim.x * nwarps;
    const unsigned char* Ab = reinterpret_cast<const unsigned char*>(A);
    const long long batch_bytes = batch * (long long)(MS * ES);

    auto request = [&](long long tile, unsigned char* buf, unsigned long long* bar) {  // lane 0, as in lub_bulk_kernel
        const long long s = tile * (long long)L::SPAN_BYTES;
        long long e = s + L::SPAN_BYTES;
        if (e > batch_bytes) e = batch_bytes;
        const long long s16 = s & ~15ll, e16 = e & ~15ll;
        const unsigned bytes = (unsigned)(e16 - s16);
        mbar_expect_tx(bar, bytes);
        if (bytes) bulk_load(buf, Ab + s16, bytes, bar);
        for (long long b = e16; b < e; b += 4) cp_async4(buf + (b - s16), Ab + b);
        cp_async_commit();
    };

    unsigned iter = 0;
    if (lane == 0) {
        const long long t0 = (long long)blockIdx.x * nwarps + warp;
        if (t0 < ntiles) request(t0, wbase, bar0);
    }
#pragma unroll 1
    for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < ntiles; tile += tstride) {
        const long long s = tile * (long long)L::SPAN_BYTES;
        const long long e = s + L::SPAN_BYTES;
        const unsigned mis = L::ALIGNED ? 0u : (unsigned)(s & 15);
        const unsigned cur = iter & 1u;
        unsigned char* buf = wbase + cur * L::IMG_BYTES;
        const unsigned parity = (iter >> 1) & 1u;
        ++iter;
        if (lane == 0) cp_async_wait<0>();
        mbar_wait(bar0 + cur, parity);
        __syncwarp();
        T* mimg = reinterpret_cast<T*>(buf + mis);

        // ---- registers <- image: LR x LC block per lane, rows where they are ----
        T a[LR][LC];
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            const bool rok = (li * GR + GR - 1 < N) || (i < N);
            const T* rowp = mimg + (rok ? i : 0) * P;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int cq = gc * CPL + q;
                if (rok && ((GC * CPL <= CPR) || (cq < CPR))) {
                    ld_vec<T, CH>(rowp + cq * CH, &a[li][q * CH]);
                } else {
#pragma unroll
                    for (int w = 0; w < CH; ++w) a[li][q * CH + w] = T(0);
                }
            }
        }
        {  // the other image: last round's matrix left it through a bulk store issued a whole elimination ago
            const long long nxt = tile + tstride;
            if (lane == 0 && nxt < ntiles) {
                tma_store_wait_read();
                request(nxt, wbase + (cur ^ 1u) * L::IMG_BYTES, bar0 + (cur ^ 1u));
            }
        }

        int pos[LR], mystep[LR];
        T dinv[LR];
        unsigned act = 0u;  // bit li: row li * GR + gr has not been a pivot yet
#pragma unroll
        for (int li = 0; li < LR; ++li) {
            const int i = li * GR + gr;
            pos[li] = i; mystep[li] = i; dinv[li] = T(0);
            if (i < N) act |= 1u << li;
        }
        int first_zero = 0;

#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int cj = k / CH, gco = cj / CPL, ck = (cj % CPL) * CH + (k % CH);
            const bool own_col = (gc == gco);
            // ---- isamax over the rows not yet used, on the UPDATED column k; first maximum in position order wins ----
            unsigned code = 0xffffu;  // (not a maximum) << 15 | position << 8 | register index << 3 | lane row
            if constexpr (sizeof(T) == 4) {
                float v[LR], lm = 0.0f;
#pragma unroll
                for (int li = 0; li < LR; ++li) {
                    v[li] = sel_t(own_col && ((act >> li) & 1u), a[li][ck], 0.0f);
                    lm = fmaxf(lm, fabsf(v[li]));
                }
                const float mx = warp_max_abs(lm);
#pragma unroll
                for (int li = 0; li < LR; ++li) {
                    const bool cand = own_col && ((act >> li) & 1u);
                    const unsigned c = ((fabsf(v[li]) == mx) ? 0u : 0x8000u) | ((unsigned)pos[li] << 8) | (unsigned)(li << 3) | (unsigned)gr;
                    code = min(code, cand ? c : 0xffffu);
                }
            } else {
                U key[LR], lm = U(0);
#pragma unroll
                for (int li = 0; li < LR; ++li) {
                    key[li] = (own_col && ((act >> li) & 1u)) ? FpBits<T>::absbits(a[li][ck]) : U(0);
                    lm = key[li] > lm ? key[li] : lm;
                }
                const U mx = warp_max_bits(lm);
#pragma unroll
                for (int li = 0; li < LR; ++li) {
                    const bool cand = own_col && ((act >> li) & 1u);
                    const unsigned c = ((key[li] == mx) ? 0u : 0x8000u) | ((unsigned)pos[li] << 8) | (unsigned)(li << 3) | (unsigned)gr;
                    code = min(code, cand ? c : 0xffffu);
                }
            }
            const unsigned win = __reduce_min_sync(0xffffffffu, code);  // warp-uniform
            const int p = (int)((win >> 8) & 31u), li_p = (int)((win >> 3) & 3u), gr_p = (int)(win & 7u);
            const bool piv_lane_row = (gr == gr_p);

            // ---- the pivot row travels to its column owners' lanes; li_p is uniform: static register indices ----
            T r[LC], pv = T(0);
#pragma unroll
            for (int c = 0; c < LR; ++c) {
                if (li_p == c) {
#pragma unroll
                    for (int lj = 0; lj < LC; ++lj) r[lj] = shfl_t(a[c][lj], gr_p * GC + gc);
                    pv = shfl_t(a[c][ck], gr_p * GC + gco);
                    if (piv_lane_row) { pos[c] = -1; mystep[c] = k; }   // pos fixed up below (the row at k first)
                }
            }
            // LAPACK's swap of positions k and p: the row that sat at k moves to p, the pivot row to k
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                if (pos[li] == k) pos[li] = p;
                if (pos[li] < 0) pos[li] = k;
            }
            const unsigned pm = piv_lane_row ? (1u << li_p) : 0u;  // bit li: my row li is this step's pivot row
            act &= ~pm;
            if (pv == T(0) && first_zero == 0) first_zero = k + 1;
            if (lane == 0) { ipiv_s[k] = p + 1; rho[k] = li_p * GR + gr_p; }

            T c[LR];
#pragma unroll
            for (int li = 0; li < LR; ++li) c[li] = shfl_t(a[li][ck], gr * GC + gco);
            const T rinv = T(1) / pv;
            if (LUONLY) {
                // rows still to be used: multiplier l = a[.][k] / pivot kept in column k, trailing update of columns > k
#pragma unroll
                for (int lj = 0; lj < LC; ++lj) {
                    const int j = gc * LC + lj;
                    if (j <= k) r[lj] = T(0);
                }
#pragma unroll
                for (int li = 0; li < LR; ++li) {
                    const bool below = ((act >> li) & 1u) != 0u;
                    const T l = below ? c[li] * rinv : T(0);
                    row_update<LC>(a[li], r, -l);
                    if (own_col && below) a[li][ck] = l;
                }
            } else {
                // Gauss-Jordan in place, rows scaled by 1 / pivot at the end: every row but the pivot row subtracts its multiple
                // of the pivot row and keeps its multiplier in column k; the pivot row keeps a 1 there (-> 1 / pivot)
#pragma unroll
                for (int li = 0; li < LR; ++li) {
                    const bool is_piv = ((pm >> li) & 1u) != 0u;
                    const T nf = is_piv ? T(0) : -(c[li] * rinv);
                    row_update<LC>(a[li], r, nf);
                    if (own_col) a[li][ck] = is_piv ? T(1) : nf;
                    if (is_piv) dinv[li] = rinv;
                }
            }
        }

        // ---- results -> image (rows / columns back in order), image -> global ----
        __syncwarp();  // rho[] complete; every lane has long loaded its block: the image may be overwritten
        if (LUONLY) {
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const bool rok = (li * GR + GR - 1 < N) || (li * GR + gr < N);
#pragma unroll
                for (int lj = 0; lj < LC; ++lj) {
                    const int j = gc * LC + lj;
                    if (rok && ((GC * LC <= N) || (j < N))) mimg[pos[li] * P + j] = a[li][lj];
                }
            }
        } else {
            int pcol[LC];
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) {
                const int j = gc * LC + lj;
                pcol[lj] = ((GC * LC <= N) || (j < N)) ? rho[j] : -1;
            }
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const bool rok = (li * GR + GR - 1 < N) || (li * GR + gr < N);
#pragma unroll
                for (int lj = 0; lj < LC; ++lj)
                    if (rok && ((GC * LC <= N) || (pcol[lj] >= 0))) mimg[mystep[li] * P + pcol[lj]] = a[li][lj] * dinv[li];
            }
        }
        fence_proxy_async();
        __syncwarp();
        {
            const long long s16u = (s + 15) & ~15ll, e16 = e & ~15ll;
            unsigned char* gdst = reinterpret_cast<unsigned char*>(A);
            if (lane == 0) {
                if (e16 > s16u) bulk_store(gdst + s16u, buf + mis + (s16u - s), (unsigned)(e16 - s16u));
                tma_store_commit();
            }
            const int hw = (int)(s16u - s) >> 2, tw = (int)(e - e16) >> 2;  // head / tail words (0..3)
            if (!L::ALIGNED && lane < hw)
                *reinterpret_cast<unsigned*>(gdst + s + 4 * lane) = *reinterpret_cast<const unsigned*>(buf + mis + 4 * lane);
            if (!L::ALIGNED && lane >= 4 && lane < 4 + tw)
                *reinterpret_cast<unsigned*>(gdst + e16 + 4 * (lane - 4)) = *reinterpret_cast<const unsigned*>(buf + mis + (e16 - s) + 4 * (lane - 4));
        }
        if (ipiv != nullptr)
            for (int x = lane; x < N; x += 32) ipiv[tile * N + x] = ipiv_s[x];
        if (lane == 0 && info != nullptr) info[tile] = first_zero;
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace lub
