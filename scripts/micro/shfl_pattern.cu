// SHFL.IDX throughput for the source-lane patterns the elimination uses, all SMs busy,
// 16 warps/SM (what the 4x4 layout runs at), several independent shuffles in flight per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int PAT>
__global__ void k(float* out, int iters) {
    const int lane = threadIdx.x & 31;
    int src;
    if (PAT == 0) src = 5;                                          // uniform broadcast
    else if (PAT == 1) src = (lane & ~15) | (2 * 4 + (lane & 3));   // 4x4 grid: row piece from lane-row 2
    else if (PAT == 2) src = (lane & ~3) | 1;                       // 4x4 grid: column piece from lane-col 1
    else if (PAT == 3) src = (lane * 7 + 3) & 31;                   // permutation
    else src = lane ^ 1;                                            // butterfly
    float v[8];
    for (int u = 0; u < 8; ++u) v[u] = threadIdx.x + u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] += __shfl_sync(0xffffffffu, v[u], src);
        src = (src + (PAT == 1 ? 4 : PAT == 2 ? 1 : 0)) & 31 | (PAT == 1 || PAT == 2 ? 0 : 0);
        if (PAT == 1) src = (lane & ~15) | (src & 15);
        if (PAT == 2) src = (lane & ~3) | (src & 3);
    }
    float s = 0; for (int u = 0; u < 8; ++u) s += v[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int PAT> void run(const char* name, int warps_per_sm) {
    float* d; cudaMalloc(&d, 148 * 64 * 32 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
    k<PAT><<<blocks, threads>>>(d, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<PAT><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // each shuffle is followed by one FADD; FADD cost is negligible next to the crossbar
    printf("%-28s %2d warps/SM: %.2f SM-cycles per warp-SHFL (at 1.92 GHz)\n", name, warps_per_sm,
           ms * 1e-3 * 1.92e9 / (double(iters) * 8 * warps_per_sm));
    cudaFree(d);
}
int main() {
    for (int w : {16, 32}) {
        run<0>("uniform source", w); run<1>("4x4 row-piece pattern", w); run<2>("4x4 column-piece pattern", w);
        run<3>("lane permutation", w); run<4>("butterfly xor 1", w);
    }
    return 0;
}
