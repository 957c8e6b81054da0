// lub_dmma.cuh -- fp64 N = 32 on the FP64 tensor-core path: a BLOCKED in-register Gauss-Jordan whose rank-4 trailing
// updates are DMMA (mma.sync.m8n8k4.f64) instead of 32 DFMA per lane and step.  This is the north star's "optional
// DMMA fp64 trailing-update variant kept only if ncu shows a gain" -- evaluated and KEPT, see profiles/r02_dmma.md:
//
//   * measured on B200 (scripts/micro/dmma_rate.cu): a DFMA whose three operands all come fresh from the register
//     file issues every 3.0 cycles (42.6 FMA / clk / SM), a DMMA m8n8k4 every 16.0 (64.0 FMA / clk / SM, the FP64
//     peak) -- and ONE instruction does the work of eight, with the operand broadcast happening inside the tensor
//     core instead of through 26 32-bit shuffles per elimination step;
//   * the fp64 kernel of round 1 (lub_tma_kernel<double, 32, 8, 4>) was bound by exactly those shuffles (LSU data
//     pipe 70 % busy, two SHFL.32 per value) and by the DFMA issue rate.
//
// Layout: one matrix per warp, held as 4 x 4 accumulator tiles of 8 x 8 in the DMMA C-fragment layout -- lane
// (g = lane >> 2, t = lane & 3) owns rows 8I + g and columns 8J + 2t + {0, 1}: 32 doubles per lane.
//
// One block step eliminates the four pivots K = 4kb .. 4kb + 3 at once (same pivots, in the same order, as four
// unblocked steps -- the row permutation was applied on the way in, exactly as in the other kernels):
//     P    = inv(A[K][K])                   4 x 4 Gauss-Jordan on 16 lanes (one element per lane, 12 shuffles)
//     C'   = A[:][K] * P                    4 DMMA   (new columns K are -C', and -C' is the A operand of the update)
//     R'   = P * A[K][:]                    4 DMMA   (new rows K)
//     A   += (-C') * A[K][:]               16 DMMA   (rank-4 update of the whole matrix; rows / columns K are
//                                                     overwritten afterwards with R', -C' and P)
// The only cross-lane traffic is the conversion of the row panel into B fragments and of the column panel (twice)
// into A fragments: 41 64-bit shuffles per block step against 104 32-bit... = 82 against 104 32-bit shuffles for the
// four unblocked steps, and 24 tensor instructions against 128 DFMA.
#pragma once
#include "lub_tma.cuh"

namespace lub {

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// c[I][J][s]: element (8I + g, 8J + 2t + s) of the (row-permuted) matrix; on return the inverse.
__device__ __forceinline__ void gj_eliminate_dmma32(double (&c)[4][4][2], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        const int Ik = kb >> 1, h = kb & 1;        // the tile row / column holding K, and which half of it
        const bool rowK = (g >> 2) == h;           // this lane owns rows of K (in tile row Ik)
        const bool colK = (t >> 1) == h;           // this lane owns columns of K (in tile column Ik)
        const int gi = g & 3;
        // ---- P0 = A[K][K] on 16 lanes: lane (4h + i, t) <- A[4kb + i][4kb + t] ----
        double p;
        {
            const int src = ((4 * h + gi) << 2) + 2 * h + (t >> 1);
            const double x0 = shfl_d(c[Ik][Ik][0], src), x1 = shfl_d(c[Ik][Ik][1], src);
            p = (t & 1) ? x1 : x0;
        }
        // ---- P = inv(P0): Gauss-Jordan on the 4 x 4 block, row gi / column t of every quad-group of 16 lanes ----
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int base = (g & 4) << 2;                       // first lane of this lane's group of 16
            const double pv = shfl_d(p, base + (k << 2) + k);
            const double rk = shfl_d(p, base + (k << 2) + t);     // pivot row, my column
            const double ck = shfl_d(p, base + (gi << 2) + k);    // my row, pivot column
            const double rinv = 1.0 / pv;
            const double rs = (t == k) ? rinv : rk * rinv;       // scaled pivot row (1/pivot in the pivot column)
            const double upd = (t == k) ? -(ck * rinv) : fma(-ck, rs, p);
            p = (gi == k) ? rs : upd;
        }
        // P[i][j] now sits in lane (4h' + i, j) of BOTH halves h' (each half inverted what it fetched; they fetched the same)
        const double aP = rowK ? p : 0.0;                                        // A fragment: rows 4h + i = P[i][:]
        const double pT = shfl_d(p, ((g & 4) << 2) + (t << 2) + gi);             // lane (., i, j) <- P[j][i]
        const double bP = rowK ? pT : 0.0;                                       // B fragment: B[k = t][n = 4h + j] = P[t][j]
        // ---- row panel A[K][:] as B fragments: B_J[k = t][n = g] = A[4kb + t][8J + g] ----
        double bR[4];
        {
            const int src = ((4 * h + t) << 2) + (g >> 1);
#pragma unroll
            for (int J = 0; J < 4; ++J) {
                const double x0 = shfl_d(c[Ik][J][0], src), x1 = shfl_d(c[Ik][J][1], src);
                bR[J] = (g & 1) ? x1 : x0;
            }
        }
        // ---- column panel A[:][K] as A fragments: A_I[m = g][k = t] = A[8I + g][4kb + t] ----
        double aC[4];
        {
            const int src = (g << 2) + 2 * h + (t >> 1);
#pragma unroll
            for (int I = 0; I < 4; ++I) {
                const double y0 = shfl_d(c[I][Ik][0], src), y1 = shfl_d(c[I][Ik][1], src);
                aC[I] = (t & 1) ? y1 : y0;
            }
        }
        // ---- C' = A[:][K] * P (lands in the C-layout positions of columns K), R' = P * A[K][:] (in rows K) ----
        double d1[4][2], d2[4][2];
#pragma unroll
        for (int I = 0; I < 4; ++I) { d1[I][0] = 0.0; d1[I][1] = 0.0; dmma8x8x4(d1[I][0], d1[I][1], aC[I], bP); }
#pragma unroll
        for (int J = 0; J < 4; ++J) { d2[J][0] = 0.0; d2[J][1] = 0.0; dmma8x8x4(d2[J][0], d2[J][1], aP, bR[J]); }
        // ---- -C' as A fragments (rows of K excluded: they are replaced, not updated) ----
        double aN[4];
        {
            const int src = (g << 2) + 2 * h + (t >> 1);
#pragma unroll
            for (int I = 0; I < 4; ++I) {
                const double y0 = shfl_d(d1[I][0], src), y1 = shfl_d(d1[I][1], src);
                aN[I] = -((t & 1) ? y1 : y0);
            }
            aN[Ik] = rowK ? 0.0 : aN[Ik];
        }
        // ---- rank-4 update of the whole matrix ----
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J < 4; ++J) dmma8x8x4(c[I][J][0], c[I][J][1], aN[I], bR[J]);
        // ---- rows K <- R', columns K <- -C', block K x K <- P ----
        double pc[2];  // P in the C layout: lane (4h + i, 2h + j / 2), slot j % 2 <- P[i][j]
        {
            const int base = (g & 4) << 2;
            pc[0] = shfl_d(p, base + (gi << 2) + ((2 * (t & 1)) & 3));
            pc[1] = shfl_d(p, base + (gi << 2) + ((2 * (t & 1) + 1) & 3));
        }
#pragma unroll
        for (int J = 0; J < 4; ++J) {
#pragma unroll
            for (int s = 0; s < 2; ++s) c[Ik][J][s] = rowK ? d2[J][s] : c[Ik][J][s];
        }
#pragma unroll
        for (int I = 0; I < 4; ++I) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const double v = (I == Ik && rowK) ? pc[s] : -d1[I][s];
                c[I][Ik][s] = colK ? v : c[I][Ik][s];
            }
        }
    }
}

// One warp = one matrix (tile); persistent over tiles; TMA staging, pivot pre-pass, permuted register load, column
// scatter and bulk tensor store exactly as in lub_tma_kernel<double, 32, 8, 4> -- only the register layout and the
// elimination differ.  MODE none: results go back through the image too (the OUTIMG path).
template <int MODE, int MINB = 2, bool BSYNC = true>
__global__ void __launch_bounds__(kMaxThreads, MINB)
lub_dmma_kernel(const __grid_constant__ CUtensorMap tmap, double* __restrict__ A, int32_t* __restrict__ piv, long long batch) {
    using T = double;
    constexpr int N = 32;
    using L = TmaLayout<T, N, 8, 4, MODE>;
    constexpr int RB = L::RB, ES = L::ES;
    static_assert(L::MPW == 1 && RB == 256, "one 32 x 32 fp64 matrix per warp");
    extern __shared__ unsigned char smem_dyn[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* img = base + (size_t)warp * L::IMG_BYTES;
    unsigned char* after = base + (size_t)nwarps * L::IMG_BYTES;
    int* perm = reinterpret_cast<int*>(after + (size_t)warp * L::PERM_BYTES);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(after + (size_t)nwarps * L::PERM_BYTES) + 2 * warp;
    int8_t* slot_rank = reinterpret_cast<int8_t*>(after + (size_t)nwarps * L::PERM_BYTES + (size_t)nwarps * 16);

    if (MODE == kModeParallel && threadIdx.x < N) slot_rank[threadIdx.x] = (int8_t)tree_slot_rank(threadIdx.x, N);
    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    unsigned parity = 0;
    const long long ntiles = batch;
#pragma unroll 1
    for (long long tbase = (long long)blockIdx.x * nwarps; tbase < ntiles; tbase += (long long)gridDim.x * nwarps) {
        if (BSYNC) __syncthreads();
        const long long tile = tbase + warp;
        if (tile >= ntiles) continue;
        if (lane == 0) {
            tma_store_wait_read();  // last round's tile has left the image
            mbar_expect_tx(bar, (unsigned)L::IMG_BYTES);
            tma_load_tile<L::LPR>(img, &tmap, bar, (int)tile);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;

        if (MODE != kModeNone) {
            prepass_rowwise_swz<T, N, MODE, 1>(img, 0, perm, slot_rank, lane);
            __syncwarp();
        }
        // ---- registers <- image: rows permuted, C-fragment layout ----
        double c[4][4][2];
#pragma unroll
        for (int I = 0; I < 4; ++I) {
            const int i = 8 * I + g;
            const int prow = (MODE != kModeNone) ? perm[i] : i;
#pragma unroll
            for (int J = 0; J < 4; ++J)
                ld_vec<T, 2>(reinterpret_cast<const T*>(img + swz_byte<RB>(prow, (4 * J + t) << 4)), c[I][J]);
        }
        gj_eliminate_dmma32(c, lane);
        // ---- undo the row permutation as a column scatter (A^-1 = (P A)^-1 P); bulk tensor store ----
        __syncwarp();  // every lane holds its block: the image may be overwritten
        if (MODE == kModeNone) {
#pragma unroll
            for (int I = 0; I < 4; ++I)
#pragma unroll
                for (int J = 0; J < 4; ++J)
                    st_vec<T, 2>(reinterpret_cast<T*>(img + swz_byte<RB>(8 * I + g, (4 * J + t) << 4)), c[I][J]);
        } else {
            int pcb[4][2];
#pragma unroll
            for (int J = 0; J < 4; ++J)
#pragma unroll
                for (int s = 0; s < 2; ++s) pcb[J][s] = perm[8 * J + 2 * t + s] * ES;
            __syncwarp();
#pragma unroll
            for (int I = 0; I < 4; ++I)
#pragma unroll
                for (int J = 0; J < 4; ++J)
#pragma unroll
                    for (int s = 0; s < 2; ++s) *reinterpret_cast<T*>(img + swz_byte<RB>(8 * I + g, pcb[J][s])) = c[I][J][s];
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_tile<L::LPR>(&tmap, img, (int)tile);
            tma_store_commit();
        }
        int32_t* pivp = piv;
        asm volatile("" : "+l"(pivp));
        if (pivp != nullptr) pivp[tile * N + lane] = (MODE != kModeNone) ? perm[lane] : lane;
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace lub
