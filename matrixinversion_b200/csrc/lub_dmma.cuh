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
//     P    = inv(A[K][K])                   4 x 4 Gauss-Jordan on 16 lanes (one element per lane, 12 64-bit shuffles)
//     C'   = A[:][K] * P                    4 DMMA   (new columns K are -C', and -C' is the A operand of the update)
//     A   += [-C' ; P - I] * A[K][:]       16 DMMA   (rank-4 update of the whole matrix; on the rows of K the A operand is
//                                                     P - I, which turns them into P * A[K][:]; columns K are overwritten
//                                                     afterwards with -C' and P)
// The panels travel through a 2.3 KB per-warp scratch in their natural layout (owners store 16-byte pairs, everybody
// loads its one fragment element per tile): 12 STS.128 + 15 LDS per block step, no selects -- the first version,
// which converted accumulator <-> fragment layouts with shuffles (41 64-bit shuffles + as many selects per block
// step), ran 3406 instructions per matrix and was no faster than the DFMA kernel (6.58 vs 6.33 ms).
#pragma once
#include "lub_tma.cuh"

namespace lub {

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// 1 / x without the IEEE division's special-case branches: MUFU.RCP64H seed + two Newton steps (<= 1 ulp for normal x;
// a zero pivot still gives inf / NaN, like the division in the other fp64 kernels -- SURVEY.md Q7)
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

// per-warp scratch for the panels of one block step (doubles): row panel 4 x 32 (row pitch 36: conflict-free fragment
// reads), column panel 32 x 4, the 4 x 4 block
struct DmmaScratch {
    static constexpr int RP = 36;
    static constexpr int kRow = 0, kCol = 4 * RP, kP = 4 * RP + 128, kDoubles = 4 * RP + 128 + 16;
    static constexpr int BYTES = ((kDoubles * 8 + 15) / 16) * 16;
};

// c[I][J][s]: element (8I + g, 8J + 2t + s) of the (row-permuted) matrix; on return the inverse.  `sc`: this warp's scratch.
__device__ __forceinline__ void gj_eliminate_dmma32(double (&c)[4][4][2], double* __restrict__ sc, int lane) {
    const int g = lane >> 2, t = lane & 3, gi = g & 3;
    double* Rp = sc + DmmaScratch::kRow;
    double* Cp = sc + DmmaScratch::kCol;
    double* Ps = sc + DmmaScratch::kP;
    const int base16 = (g & 4) << 2;                                   // first lane of this lane's group of 16
    const double* rp_rd = Rp + t * DmmaScratch::RP + g;                 // B fragment: Rp[t][8J + g]
    const double* cp_rd = Cp + (g << 2) + t;                            // A fragment: Cp[8I + g][t]
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        const int Ik = kb >> 1, h = kb & 1;        // the tile row / column holding K, and which half of it
        const bool rowK = (g >> 2) == h;           // this lane owns rows of K (in tile row Ik)
        const bool colK = (t >> 1) == h;           // this lane owns columns of K (in tile column Ik)
        // ---- the two panels -> scratch, in natural layout (owners only) ----
        __syncwarp();
        if (rowK) {
#pragma unroll
            for (int J = 0; J < 4; ++J) st_vec<double, 2>(Rp + gi * DmmaScratch::RP + 8 * J + 2 * t, c[Ik][J]);
        }
        if (colK) {
#pragma unroll
            for (int I = 0; I < 4; ++I) st_vec<double, 2>(Cp + ((8 * I + g) << 2) + 2 * (t & 1), c[I][Ik]);
        }
        __syncwarp();
        // ---- P = inv(A[K][K]): Gauss-Jordan on the 4 x 4 block, one element per lane in each group of 16 lanes ----
        double p = Rp[gi * DmmaScratch::RP + 4 * kb + t];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double pv = shfl_d(p, base16 + (k << 2) + k);
            const double rk = shfl_d(p, base16 + (k << 2) + t);     // pivot row, my column
            const double ck = shfl_d(p, base16 + (gi << 2) + k);    // my row, pivot column
            const double rinv = rcp_fast(pv);
            const double rs = (t == k) ? rinv : rk * rinv;         // scaled pivot row (1/pivot in the pivot column)
            const double upd = (t == k) ? -(ck * rinv) : fma(-ck, rs, p);
            p = (gi == k) ? rs : upd;
        }
        if (g < 4) Ps[(gi << 2) + t] = p;                           // P[i][j] (both groups of 16 hold the same P)
        // ---- fragments of the raw panels ----
        double bR[4], aC[4];
#pragma unroll
        for (int J = 0; J < 4; ++J) bR[J] = rp_rd[8 * J];
#pragma unroll
        for (int I = 0; I < 4; ++I) aC[I] = cp_rd[32 * I];
        __syncwarp();
        // A fragment of (P - I) on rows K (folded into the update of tile row Ik: rows K become P * A[K][:]);
        // B fragment of P on columns K (C' = A[:][K] * P lands in the C-layout positions of columns K)
        const double pT = Ps[(t << 2) + gi];                        // P[t][gi]
        const double aPI = rowK ? (p - ((gi == t) ? 1.0 : 0.0)) : 0.0;
        const double bP = rowK ? pT : 0.0;
        double pc[2];                                               // P in the C layout: (4h + i, 4h + 2(t & 1) + s)
        ld_vec<double, 2>(Ps + (gi << 2) + 2 * (t & 1), pc);
        double d1[4][2];
#pragma unroll
        for (int I = 0; I < 4; ++I) { d1[I][0] = 0.0; d1[I][1] = 0.0; dmma8x8x4(d1[I][0], d1[I][1], aC[I], bP); }
        // ---- -C' as A fragments, through the column-panel scratch ----
        if (colK) {
#pragma unroll
            for (int I = 0; I < 4; ++I) st_vec<double, 2>(Cp + ((8 * I + g) << 2) + 2 * (t & 1), d1[I]);
        }
        __syncwarp();
        double aN[4];
#pragma unroll
        for (int I = 0; I < 4; ++I) aN[I] = -cp_rd[32 * I];
        aN[Ik] = rowK ? aPI : aN[Ik];               // rows K: += (P - I) * A[K][:] instead of -= C' * A[K][:]
        // ---- rank-4 update of the whole matrix ----
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J < 4; ++J) dmma8x8x4(c[I][J][0], c[I][J][1], aN[I], bR[J]);
        // ---- columns K <- -C', block K x K <- P ----
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

constexpr int dmma_smem_bytes(int warps, int perm_bytes) {
    return 1024 + warps * (32 * 256 + perm_bytes) + warps * 16 + 64 + warps * DmmaScratch::BYTES;
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
    double* scratch = reinterpret_cast<double*>(after + (size_t)nwarps * L::PERM_BYTES + (size_t)nwarps * 16 + L::HEADER_BYTES) +
                      (size_t)warp * (DmmaScratch::BYTES / 8);

    if (MODE == kModeParallel && threadIdx.x < N) slot_rank[threadIdx.x] = (int8_t)tree_slot_rank(threadIdx.x, N);
    if (lane == 0) mbar_init(bar, 1);
    if (MODE != kModeNone) perm[lane] = lane;  // always row indices, whatever a search derailed by NaN inputs leaves unwritten
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
        gj_eliminate_dmma32(c, scratch, lane);
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
