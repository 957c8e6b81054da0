// Micro-benchmark (dev only): issue rate of the fp64 paths on B200 -- DFMA (the rank-1 update of the fp64 kernels),
// DMMA mma.sync.m8n8k4.f64 (north star: "optional DMMA fp64 trailing-update variant kept only if ncu shows a gain"),
// and the 64-bit shuffle (two SHFL.32) that today's fp64 kernels are bound by.  Output: SMSP-cycles per warp
// instruction and the FMA rate per SM and clock each implies (DFMA: 32 FMA per warp instruction, DMMA: 256).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 200
#define REP 8

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, long long* cyc, int p, int q) {
    const int lane = threadIdx.x & 31;
    double c[16][2], a[4], b[4], f[16], s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = i + lane; c[i][1] = i * q; f[i] = i + lane * p; s[i] = 1.0 + 1e-9 * (i * q + p); }
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = 1.0 + 1e-9 * (lane + i * q); b[i] = 1.0 - 1e-9 * (lane * p + i); }
    const int src = (lane + 5) & 31;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int rep = 0; rep < REP; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(s[i]), "d"(s[(i + 3) & 15]));
                if (MODE == 1) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                            : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i & 3]), "d"(b[(i >> 2) & 3]));
                if (MODE == 2) f[i] = __shfl_sync(0xffffffffu, f[i], src);
                if (MODE == 3) {  // the fp64 kernel's step mix per 8 DFMA: 8 DFMA + 3 x 64-bit shuffles (N = 32: 32 DFMA, 13 shuffles)
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(s[i]), "d"(s[(i + 3) & 15]));
                    if ((i & 7) < 3) s[i] = __shfl_sync(0xffffffffu, s[i], src);
                }
                if (MODE == 4) {  // DMMA variant of the same work: 1 DMMA replaces 8 DFMA per lane, ~1 64-bit shuffle per DMMA for the panels
                    if ((i & 7) == 0) {
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i & 3]), "d"(b[(i >> 2) & 3]));
                        a[i & 3] = __shfl_sync(0xffffffffu, a[i & 3], src);
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += c[i][0] + c[i][1] + f[i] + s[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) r += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, double inst_per_group, double fma_per_group) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, warps * 32>>>(out, cyc, 1, 3);
    k<MODE><<<148, warps * 32>>>(out, cyc, 1, 3);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double groups_per_smsp = (double)ITER * REP * 16 * warps / 4.0;
    const double cyc_per_group = avg / groups_per_smsp;
    printf("{\"mix\": \"%s\", \"warps_per_sm\": %d, \"smsp_cycles_per_group\": %.3f, \"fma_per_sm_per_clock\": %.1f, \"err\": \"%s\"}\n", name, warps,
           cyc_per_group * inst_per_group, fma_per_group > 0 ? 4.0 * fma_per_group / cyc_per_group : 0.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int warps : {4, 8, 16}) {
        run<0>("DFMA", warps, 1, 32);
        run<1>("DMMA m8n8k4", warps, 1, 256);
        run<2>("SHFL 64-bit", warps, 1, 0);
        run<3>("8 DFMA + 3 64-bit SHFL (fp64 step mix)", warps, 8, 32);      // per group of 1/8: report x8
        run<4>("1 DMMA + 1 64-bit SHFL (same 256 FMA)", warps, 8, 32);       // one DMMA per 8 groups = 256 FMA per 8 groups
    }
    return 0;
}
