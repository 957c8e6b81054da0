// Micro-benchmark (dev only, not part of the product): what does ONE Gauss-Jordan step of the headline
// configuration (N = 32 fp32, 4 x 4 lanes per matrix, 8 x 8 register block, two matrices per warp) cost on a
// B200 SM when nothing else runs -- no HBM traffic, no shared-memory staging, no pivot search?  The kernel
// inverts a register-resident matrix over and over (A -> A^-1 -> A ...), so every variant executes exactly the
// instruction stream the library executes between its register load and its column scatter.
//
//   variant 0  gj_eliminate        (round-1 step: FSEL / predicated-MOV fix-ups, column cleared before the update)
//   variant 1  gj_eliminate_lean   (round-2 step: predicated FMA-pipe fix-ups after the update)
//   variant 2  minimal step        (17 SHFL + reciprocal + 8 FMUL + 32 FFMA2, NO fix-ups: wrong results --
//                                   the least any register Gauss-Jordan on this lane grid can issue)
//   variant 3  minimal step, no shuffles (lanes use their own registers: FMA pipe + reciprocal only)
//   variant 4  32 FFMA2 per step only
//
// Output: SM-cycles per matrix (the unit of DESIGN.md's budget: 1.707 ms for 1e6 matrices on 148 SMs at
// 1.965 GHz = 497 SM-cycles per matrix for EVERYTHING), per variant and per number of resident warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o step_floor step_floor.cu && ./step_floor
#include <cstdio>
#include <cuda_runtime.h>
#include "../../matrixinversion_b200/csrc/lub_v3.cuh"

using namespace lub;

constexpr int N = 32, CH = 4;

template <int VARIANT, int GR, int GC, int LR, int LC, int CPL>
__device__ __forceinline__ void step_minimal(float (&a)[LR][LC], float (&dinv)[LR], int gr, int gc, int grp_base) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int gro = k % GR, lk = k / GR;
        const int cj = k / CH, gco = cj / CPL, ck = (cj % CPL) * CH + (k % CH);
        float r[LC], c[LR];
        if (VARIANT == 4) {
#pragma unroll
            for (int li = 0; li < LR; ++li) row_update<LC>(a[li], a[(li + 1) % LR], dinv[li]);
            continue;
        }
#pragma unroll
        for (int lj = 0; lj < LC; ++lj) r[lj] = (VARIANT == 2) ? shfl_t(a[lk][lj], grp_base + gro * GC + gc) : a[lk][lj];
#pragma unroll
        for (int li = 0; li < LR; ++li) c[li] = (VARIANT == 2) ? shfl_t(a[li][ck], grp_base + gr * GC + gco) : a[li][ck];
        const float pv = (VARIANT == 2) ? shfl_t(a[lk][ck], grp_base + gro * GC + gco) : a[lk][ck];
        const float rinv = rcp_t(pv);
        float nf[LR];
#pragma unroll
        for (int li = 0; li < LR; ++li) nf[li] = -(c[li] * rinv);
#pragma unroll
        for (int li = 0; li < LR; ++li) row_update<LC>(a[li], r, nf[li]);
        dinv[lk] = rinv;
    }
}

template <int VARIANT, int MAXT, int GR = 4, int GC = 4>
__global__ void __launch_bounds__(MAXT, 1) elim_only(float* out, long long* cyc, int iters) {
    constexpr int G = GR * GC, LR = N / GR, LC = N / GC, CPL = LC / CH;
    const int lane = threadIdx.x & 31;
    const int g = lane % G, ml = lane / G, gr = g / GC, gc = g % GC, grp_base = ml * G;
    float a[LR][LC];
#pragma unroll
    for (int li = 0; li < LR; ++li)
#pragma unroll
        for (int lj = 0; lj < LC; ++lj) {
            const int i = li * GR + gr, j = gc * LC + lj;
            a[li][lj] = (i == j) ? 8.0f : 0.05f * (float)(((i * 37 + j * 11 + ml * 5 + (threadIdx.x >> 5)) % 17) - 8);
        }
    float dinv[LR];
#pragma unroll
    for (int li = 0; li < LR; ++li) dinv[li] = 1e-3f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (VARIANT == 0) gj_eliminate<float, N, GR, GC, CH, CPL, LR, LC>(a, dinv, gr, gc, grp_base);
        else if (VARIANT == 1) gj_eliminate_lean<float, N, GR, GC, CH, CPL, LR, LC>(a, dinv, gr, gc, grp_base);
        else step_minimal<VARIANT, GR, GC, LR, LC, CPL>(a, dinv, gr, gc, grp_base);
        if (VARIANT <= 1) {
#pragma unroll
            for (int li = 0; li < LR; ++li)
#pragma unroll
                for (int lj = 0; lj < LC; ++lj) a[li][lj] *= dinv[li];
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int li = 0; li < LR; ++li)
#pragma unroll
        for (int lj = 0; lj < LC; ++lj) s += a[li][lj];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int VARIANT, int MAXT, int GR = 4, int GC = 4>
void run(const char* name, int warps, int iters) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    elim_only<VARIANT, MAXT, GR, GC><<<148, warps * 32>>>(out, cyc, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    elim_only<VARIANT, MAXT, GR, GC><<<148, warps * 32>>>(out, cyc, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    float chk[4]; cudaMemcpy(chk, out, sizeof(chk), cudaMemcpyDeviceToHost);
    const double mats_per_sm = (double)iters * warps * (32 / (GR * GC));
    printf("{\"variant\": \"%s\", \"lane_grid\": \"%dx%d\", \"warps_per_sm\": %d, \"sm_cycles_per_matrix\": %.1f, \"ms_per_1e6_matrices_at_148sm\": %.3f, "
           "\"check\": %.4g, \"err\": \"%s\"}\n",
           name, GR, GC, warps, avg / mats_per_sm, ms * 1e6 / (mats_per_sm * 148), chk[0], cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char** argv) {
    const int iters = 400;
    if (argc > 1) {  // profiling aid: one variant at 12 warps per SM (ncu -k regex:elim_only ./step_floor <variant>)
        switch (argv[1][0]) {
            case '0': run<0, 384>("r1 step (FSEL fix-ups)", 12, iters); break;
            case '1': run<1, 384>("lean step (FMA-pipe fix-ups)", 12, iters); break;
            case '2': run<2, 384>("minimal step (no fix-ups)", 12, iters); break;
            case '3': run<3, 384>("minimal step, no shuffles", 12, iters); break;
            default: run<4, 384>("FFMA2 only", 12, iters); break;
        }
        return 0;
    }
    // other lane grids (fewer lanes per matrix = fewer exchanged words per flop, more registers per lane)
    run<1, 256, 4, 2>("lean step (FMA-pipe fix-ups)", 8, iters);
    run<0, 256, 4, 2>("r1 step (FSEL fix-ups)", 8, iters);
    run<2, 256, 4, 2>("minimal step (no fix-ups)", 8, iters);
    run<1, 256, 2, 4>("lean step (FMA-pipe fix-ups)", 8, iters);
    run<0, 256, 2, 4>("r1 step (FSEL fix-ups)", 8, iters);
    run<2, 256, 2, 4>("minimal step (no fix-ups)", 8, iters);
    run<1, 256, 8, 2>("lean step (FMA-pipe fix-ups)", 8, iters);
    run<1, 256, 2, 8>("lean step (FMA-pipe fix-ups)", 8, iters);
    run<1, 256, 8, 4>("lean step (FMA-pipe fix-ups)", 8, iters);
    for (int warps : {8, 12, 16}) {
        if (warps <= 8) {
            run<0, 256>("r1 step (FSEL fix-ups)", warps, iters);
            run<1, 256>("lean step (FMA-pipe fix-ups)", warps, iters);
            run<2, 256>("minimal step (no fix-ups)", warps, iters);
            run<3, 256>("minimal step, no shuffles", warps, iters);
            run<4, 256>("FFMA2 only", warps, iters);
        } else if (warps <= 12) {
            run<0, 384>("r1 step (FSEL fix-ups)", warps, iters);
            run<1, 384>("lean step (FMA-pipe fix-ups)", warps, iters);
            run<2, 384>("minimal step (no fix-ups)", warps, iters);
            run<3, 384>("minimal step, no shuffles", warps, iters);
            run<4, 384>("FFMA2 only", warps, iters);
        } else {
            run<0, 512>("r1 step (FSEL fix-ups)", warps, iters);
            run<1, 512>("lean step (FMA-pipe fix-ups)", warps, iters);
            run<2, 512>("minimal step (no fix-ups)", warps, iters);
            run<3, 512>("minimal step, no shuffles", warps, iters);
            run<4, 512>("FFMA2 only", warps, iters);
        }
    }
    return 0;
}
