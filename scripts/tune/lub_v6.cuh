// lub_v6.cuh -- warp-specialised, TMA-staged kernel for the pivoting modes at the sizes whose rows
// are one 128-byte line (N = 32 fp32: the headline configuration of
// parallel_pivot/luBatchedInplace.cu; N = 16 fp64).  Same algorithm and bit-identical results as
// lub_tma_kernel (lub_tma.cuh); what changes is which warp does what.
//
// Why: in lub_tma_kernel every warp walks a tile through stage-in wait (17 % of warp time), pivot
// pre-pass (24 %), register load + Gauss-Jordan + scatter (59 %) in sequence
// (profiles/r01_prof_headline_n32_f32_parallel_1M.md).  The register file holds 16 such warps, so on
// average fewer than ten of them are in the FMA/shuffle-heavy part and the issue slots sit at 57 %.
// Here one persistent block per SM splits the roles:
//   * PRODUCER warps (few registers, setmaxnreg.dec) own the input side of a ring of shared-memory
//     slots: wait until a slot is drained, have the TMA unit load the next tile into it
//     (one tile of look-ahead per producer), run the row-wise pivot search, write the permutation
//     (and the exported piv[] rows), publish the slot;
//   * CONSUMER warps (128 registers, setmaxnreg.inc) take a published slot, load their permuted
//     register blocks, run gj_eliminate, scatter the inverse into the slot image and hand it to the
//     TMA unit for the bulk store; the slot is released once the store has read it.
// Tiles are dealt round-robin (tile q -> slot q % NB, producer q % NPW, consumer q % NCW) with NB a
// multiple of both warp counts, so every slot keeps its producer and consumer and the hand-over is a
// pair of mbarriers per slot (hardware-suspended waits, no polling).
#pragma once
#include "../../matrixinversion_b200/csrc/lub_tma.cuh"

namespace lub {

__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(void* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <typename T, int N, int GR, int GC, int MODE, int NB>
struct V6Layout : TmaLayout<T, N, GR, GC, MODE> {
    using B = TmaLayout<T, N, GR, GC, MODE>;
    static constexpr int PERM_SLOT = B::MPW * N * 4;
    // [1 KB slack to align by hand][NB images][NB perms][3 x NB mbarriers: landed, full, empty][slot ranks]
    static constexpr int SMEM_BYTES = 1024 + NB * (B::IMG_BYTES + PERM_SLOT) + 3 * NB * 8 + 64;
};

// Exact warp-wide search, inlined (a producer that gave registers back cannot afford a callee's
// register budget).  Same search as prepass_exact_swz.
template <typename T, int N, int MODE>
__device__ __forceinline__ void prepass_exact_swz_inl(const unsigned char* mimg, int* perm, const int8_t* slot_rank, int lane) {
    using U = typename FpBits<T>::U;
    constexpr int RB = N * (int)sizeof(T), ES = sizeof(T);
    for (int i = lane; i < N; i += 32) perm[i] = i;
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k < N - 1; ++k) {
        U best_v = FpBits<T>::absbits(*reinterpret_cast<const T*>(mimg + swz_off<RB, ES>(perm[k], k)));
        unsigned best_p = 0;
#pragma unroll 1
        for (int t = lane; t < N - 1 - k; t += 32) {
            int pr;
            if (MODE == kModeParallel) {
                pr = slot_rank[t];
                if (pr < 0) continue;
            } else {
                pr = t;
            }
            const U v = FpBits<T>::absbits(*reinterpret_cast<const T*>(mimg + swz_off<RB, ES>(perm[k + 1 + t], k)));
            const unsigned p = ((unsigned)(pr + 1) << 8) | (unsigned)(t + 1);
            if (v > best_v || (v == best_v && p < best_p)) { best_v = v; best_p = p; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const U ov = __shfl_xor_sync(0xffffffffu, best_v, off);
            const unsigned op = __shfl_xor_sync(0xffffffffu, best_p, off);
            if (ov > best_v || (ov == best_v && op < best_p)) { best_v = ov; best_p = op; }
        }
        __syncwarp();
        if (lane == 0 && best_p != 0) {
            const int p = k + (int)(best_p & 0xffu);
            const int tmp = perm[k];
            perm[k] = perm[p];
            perm[p] = tmp;
        }
        __syncwarp();
    }
}

// Row-wise search on the swizzled image (prepass_rowwise_swz of lub_tma.cuh) with the inlined exact
// fallback; MI matrices in lock step.
template <typename T, int N, int MODE, int MI>
__device__ __forceinline__ void prepass_rowwise_swz_inl(const unsigned char* img0, int* perm0, const int8_t* slot_rank, int lane) {
    using U = typename FpBits<T>::U;
    constexpr int ES = sizeof(T), EPV = 16 / ES, RB = N * ES, MAT = N * RB;
    const int row = (lane < N) ? lane : 0;
    const unsigned char* rowp = img0 + row * RB;
    const int xr = (row & 7) << 4;
    U alive[MI];
    int when[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) { alive[m] = (lane < N) ? ~U(0) : U(0); when[m] = N - 1; }
    // one 16-byte chunk of the row per outer iteration, NOT unrolled: unrolled, the compiler hoists
    // every load to the top and the producer's small register budget spills
#pragma unroll 1
    for (int c = 0; c < N / EPV; ++c) {
        T x[MI][EPV];
#pragma unroll
        for (int m = 0; m < MI; ++m)
            ld_vec<T, EPV>(reinterpret_cast<const T*>(rowp + m * MAT + ((c << 4) ^ xr)), x[m]);
#pragma unroll
        for (int j = 0; j < EPV; ++j) {
            const int k = c * EPV + j;
            if (k < N - 1) {
                U key[MI], mx[MI];
#pragma unroll
                for (int m = 0; m < MI; ++m) key[m] = ((FpBits<T>::absbits(x[m][j]) << 1) | U(1)) & alive[m];
#pragma unroll
                for (int m = 0; m < MI; ++m) mx[m] = warp_max_bits(key[m]);
#pragma unroll
                for (int m = 0; m < MI; ++m) {
                    const bool hit = key[m] == mx[m];
                    when[m] = hit ? k : when[m];
                    alive[m] = hit ? U(0) : alive[m];
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MI; ++m) {
        const bool ok = __popc(__ballot_sync(0xffffffffu, alive[m] != U(0))) == 1;  // warp-uniform
        if (ok) {
            if (lane < N) perm0[m * N + when[m]] = lane;
        } else {
            prepass_exact_swz_inl<T, N, MODE>(img0 + m * MAT, perm0 + m * N, slot_rank, lane);
        }
    }
}

// NCW consumer warps (ids 0 .. NCW-1), NPW producer warps (the highest ids: the issue arbiter favours
// them, and a starved producer starves the ring), NB slots.  NB is a multiple of both NCW and NPW, so a
// slot always has the same producer and the same consumer and its three mbarriers (landed: TMA bytes,
// full: producer -> consumer, empty: consumer -> producer) strictly alternate: a one-bit phase is
// enough.  CREG / PREG: registers per thread after the split.  NCG: consumer warps per re-alignment
// group (named barrier once per tile, 0 = none): the elimination is ~40 KB of straight-line code and
// warps that drift apart thrash the instruction caches.  LA: tiles a producer keeps in flight ahead of
// the one it is searching.  DBG (tuning harness only, wrong results): 1 = consumers skip the
// elimination, 2 = producers skip the pivot search (identity permutation).
template <typename T, int N, int GR, int GC, int MODE, int NCW, int NPW, int NB, int CREG, int PREG, int NCG, int LA = 2, int DBG = 0>
__global__ void __launch_bounds__((NCW + NPW) * 32, 1)
lub_v6_kernel(const __grid_constant__ CUtensorMap tmap, T* __restrict__ A, int32_t* __restrict__ piv, long long batch) {
    static_assert(MODE != kModeNone, "the producer/consumer split pays only with a pivot search");
    static_assert(NCW % 4 == 0 && NPW % 4 == 0, "register re-allocation is per 4-warp group");
    static_assert(NCG == 0 || NCW % NCG == 0, "consumer groups must tile the consumer warps");
    static_assert(NB % NCW == 0 && NB % NPW == 0, "a slot must keep its producer and its consumer");
    // a slot is released when its consumer picks up its NEXT tile: the ring must hold two tiles per consumer
    static_assert(NB >= 2 * NCW, "ring too short: hand-over could deadlock");
    using L = V6Layout<T, N, GR, GC, MODE, NB>;
    constexpr int G = L::G, MPW = L::MPW, LR = L::LR, LC = L::LC, CH = L::CH, CPL = L::CPL;
    constexpr int RB = L::RB, ES = L::ES, MAT = L::MAT_BYTES;
    constexpr int R0 = (65536 / ((NCW + NPW) * 32)) / 8 * 8;  // what the launch bound lets ptxas use
    constexpr bool REGSPLIT = (CREG != R0) || (PREG != R0);
    static_assert(!REGSPLIT || (NCW * (CREG - R0) <= NPW * (R0 - PREG)), "consumers may only take what producers release");
    static_assert(CREG % 8 == 0 && PREG % 8 == 0 && PREG >= 24 && CREG <= 255, "setmaxnreg operand");
    extern __shared__ unsigned char smem_dyn[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* imgs = base;
    unsigned char* perms = base + (size_t)NB * L::IMG_BYTES;
    unsigned long long* bar_land = reinterpret_cast<unsigned long long*>(perms + (size_t)NB * L::PERM_SLOT);
    unsigned long long* bar_full = bar_land + NB;
    unsigned long long* bar_empty = bar_full + NB;
    int8_t* slot_rank = reinterpret_cast<int8_t*>(bar_empty + NB);

    if (MODE == kModeParallel && threadIdx.x < N) slot_rank[threadIdx.x] = (int8_t)tree_slot_rank(threadIdx.x, N);
    if (threadIdx.x < 3 * NB) mbar_init(bar_land + threadIdx.x, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const long long ntiles = (batch + MPW - 1) / MPW;
    // tiles of this block: blockIdx.x, blockIdx.x + gridDim.x, ...; q is the block-local sequence number
    const long long Q = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto tile_first = [&](long long q) { return (blockIdx.x + q * (long long)gridDim.x) * (long long)MPW; };

    if (warp >= NCW) {
        // ================================ producer ================================
        if (REGSPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PREG));
        const int pw = warp - NCW;
        long long q_load = pw;  // next tile of this producer whose load has not been issued yet
#pragma unroll 1
        for (long long q = pw; q < Q; q += NPW) {
            // keep up to LA tiles in flight beyond the one searched now; only the load of tile q itself
            // may block on its slot, the others go out if their slot happens to be free
#pragma unroll 1
            while (q_load < Q && q_load <= q + (long long)LA * NPW) {
                const int sl = (int)(q_load % NB);
                const unsigned ul = (unsigned)(q_load / NB);
                if (ul > 0) {  // the previous tenant must have left (its bulk store has read the image)
                    if (q_load == q) {
                        mbar_wait(bar_empty + sl, (ul - 1u) & 1u);
                    } else {
                        const int freed = __shfl_sync(0xffffffffu, (int)mbar_test(bar_empty + sl, (ul - 1u) & 1u), 0);
                        if (!freed) break;
                    }
                }
                if (lane == 0) {
                    mbar_expect_tx(bar_land + sl, (unsigned)L::IMG_BYTES);
                    tma_load_3d(imgs + (size_t)sl * L::IMG_BYTES, &tmap, bar_land + sl, 0, 0, (int)tile_first(q_load));
                }
                q_load += NPW;
            }
            const int s = (int)(q % NB);
            const unsigned use = (unsigned)(q / NB);
            unsigned char* img = imgs + (size_t)s * L::IMG_BYTES;
            int* perm_all = reinterpret_cast<int*>(perms + (size_t)s * L::PERM_SLOT);
            mbar_wait(bar_land + s, use & 1u);  // this tile has landed
            if (DBG & 2) {
                for (int e = lane; e < MPW * N; e += 32) perm_all[e] = e % N;
            } else {
                constexpr int MI = (MPW < 2) ? MPW : 2;
#pragma unroll 1
                for (int m = 0; m < MPW; m += MI)
                    prepass_rowwise_swz_inl<T, N, MODE, MI>(img + m * MAT, perm_all + m * N, slot_rank, lane);
            }
            __syncwarp();
            int32_t* pivp = piv;
            if (pivp != nullptr) {
                const long long first = tile_first(q);
                const int nm = (batch - first < MPW) ? (int)(batch - first) : MPW;
                int32_t* pdst = pivp + first * N;
                for (int e = lane; e < nm * N; e += 32) pdst[e] = perm_all[e];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + s);
        }
    } else {
        // ================================ consumer ================================
        if (REGSPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CREG));
        const int cw = warp;
        const int g = lane % G;
        const int ml = lane / G;
        const int gr = g / GC;
        const int gc = g % GC;
        const int grp_base = ml * G;
        int prev_s = -1;
        const long long rounds = (Q + NCW - 1) / NCW;
#pragma unroll 1
        for (long long rd = 0; rd < rounds; ++rd) {
            if constexpr (NCG > 0) named_bar_sync(1 + cw / NCG, NCG * 32);  // re-align the group (every warp of it runs every round)
            const long long q = rd * NCW + cw;
            if (q >= Q) continue;
            const int s = (int)(q % NB);
            const unsigned use = (unsigned)(q / NB);
            mbar_wait(bar_full + s, use & 1u);
            unsigned char* img = imgs + (size_t)s * L::IMG_BYTES;
            const int* perm = reinterpret_cast<const int*>(perms + (size_t)s * L::PERM_SLOT) + ml * N;
            unsigned char* mimg = img + ml * MAT;

            // ---- registers <- image: rows permuted, LR x LC block per lane ---------------------
            T a[LR][LC];
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (li * GR + GR - 1 < N) || (i < N);
                const int prow = rok ? perm[i] : 0;
                const unsigned char* rowp = mimg + prow * RB;
                const int xr = (prow & 7) << 4;
#pragma unroll
                for (int qq = 0; qq < CPL; ++qq) {
                    if (rok) {
                        ld_vec<T, CH>(reinterpret_cast<const T*>(rowp + (((gc * CPL + qq) << 4) ^ xr)), &a[li][qq * CH]);
                    } else {
#pragma unroll
                        for (int w = 0; w < CH; ++w) a[li][qq * CH + w] = T(0);
                    }
                }
            }
            // the previous tile's bulk store has long since read its slot: release it
            if (prev_s >= 0 && lane == 0) {
                tma_store_wait_read();
                mbar_arrive(bar_empty + prev_s);
            }

            T dinv[LR];
#pragma unroll
            for (int li = 0; li < LR; ++li) dinv[li] = (DBG & 1) ? T(1) : T(0);
            if (!(DBG & 1)) gj_eliminate<T, N, GR, GC, CH, CPL, LR, LC, (DBG & 12)>(a, dinv, gr, gc, grp_base);

            // ---- scale by 1/pivot; undo the row permutation as a column scatter -------------------
#pragma unroll
            for (int li = 0; li < LR; ++li) {
#pragma unroll
                for (int lj = 0; lj < LC; ++lj) a[li][lj] *= dinv[li];
            }
            int pcb[LC];  // byte offset of the destination column inside a row
#pragma unroll
            for (int lj = 0; lj < LC; ++lj) pcb[lj] = perm[gc * LC + lj] * ES;
            __syncwarp();  // all lanes hold their blocks and columns: the image may be overwritten
#pragma unroll
            for (int li = 0; li < LR; ++li) {
                const int i = li * GR + gr;
                const bool rok = (li * GR + GR - 1 < N) || (i < N);
                unsigned char* rowp = mimg + i * RB;
                const int xr = (i & 7) << 4;
#pragma unroll
                for (int lj = 0; lj < LC; ++lj)
                    if (rok) *reinterpret_cast<T*>(rowp + (pcb[lj] ^ xr)) = a[li][lj];
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the TMA unit
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&tmap, img, 0, 0, (int)tile_first(q));  // rows past the batch end are clipped
                tma_store_commit();
            }
            prev_s = s;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete
    }
}

}  // namespace lub
