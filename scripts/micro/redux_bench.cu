// Micro-benchmark: throughput (all SMs busy, many warps) and latency (1 warp) of the warp
// collectives the pivot pre-pass is built from.  nvcc -arch=sm_100a -O3 redux_bench.cu -o redux_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(unsigned* out, int iters) {
    unsigned v = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            unsigned r;
            if (OP == 0) r = __reduce_max_sync(0xffffffffu, v);
            else if (OP == 1) r = __shfl_xor_sync(0xffffffffu, v, 1);
            else if (OP == 2) r = __ballot_sync(0xffffffffu, v & 1);
            else if (OP == 3) r = __shfl_sync(0xffffffffu, v, (v >> 3) & 31);
            else if (OP == 4) r = __match_any_sync(0xffffffffu, v & 3);
            else r = v * 3 + 1;
            // dependent (latency) flavour: feed the result back; independent: just accumulate
            v = v + r + u;
            acc ^= r;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + v;
}
template <int OP>
__global__ void kind(unsigned* out, int iters) {  // 8 independent chains per warp
    unsigned v[8], acc = 0;
    for (int u = 0; u < 8; ++u) v[u] = threadIdx.x * 2654435761u + blockIdx.x + u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            unsigned r;
            if (OP == 0) r = __reduce_max_sync(0xffffffffu, v[u]);
            else if (OP == 1) r = __shfl_xor_sync(0xffffffffu, v[u], 1);
            else if (OP == 2) r = __ballot_sync(0xffffffffu, v[u] & 1);
            else if (OP == 3) r = __shfl_sync(0xffffffffu, v[u], (v[u] >> 3) & 31);
            else if (OP == 4) r = __match_any_sync(0xffffffffu, v[u] & 3);
            else r = v[u] * 3 + 1;
            v[u] += r;
            acc ^= r;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + v[0];
}
template <int OP>
void run(const char* name) {
    unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    float ms;
    // latency: one warp, dependent chain
    k<OP><<<1, 32>>>(d, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<OP><<<1, 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double lat = ms * 1e-3 * 1.9e9 / (iters * 8.0);
    // throughput: 148*4 blocks of 256 threads (32 warps/SM), independent chains
    kind<OP><<<148 * 4, 256>>>(d, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); kind<OP><<<148 * 4, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double per_sm_cycles = ms * 1e-3 * 1.9e9 / (iters * 8.0 * 32.0);  // cycles per warp-op per SM
    printf("%-12s dependent-chain latency ~%.1f cyc (incl. add)   throughput: %.2f cyc per warp-instr per SM\n", name, lat, per_sm_cycles);
    cudaFree(d);
}
int main() {
    run<0>("redux.max"); run<1>("shfl.xor"); run<2>("ballot"); run<3>("shfl.idx"); run<4>("match.any"); run<5>("imad");
    return 0;
}
