// Micro-benchmark (dev only): issue rate of the instruction mix of the Gauss-Jordan step on B200.
// Each kernel runs ITER iterations of an unrolled body of independent instructions on WARPS warps of
// one block per SM and reports SMSP-cycles per warp-instruction (4 schedulers share the warps evenly).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2000

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
#define FFMA2(acc, a, b) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b))
#define FFMA(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))
#define FSEL(x, y, p) asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; selp.f32 %0, %1, %0, q; }" : "+f"(x) : "f"(y), "r"(p))
#define FMUL(x, y) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(y))
#define SHFL(x, src) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+f"(x) : "r"(src))

template <int MODE>
__global__ void __launch_bounds__(1024, 1) mix(float* out, long long* cyc, int p) {
    const int lane = threadIdx.x & 31;
    unsigned long long acc[8], a2 = pack(1.0001f, 0.9999f), b2 = pack(0.5f, 0.25f);
    float f[8], s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = pack((float)i, (float)(i + threadIdx.x)); f[i] = i + lane; s[i] = lane * i; }
    const float m = 1.0001f;
    const int src = (lane + 5) & 31;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { FFMA(f[i], m, m); FFMA(s[i], m, m); }                       // 2 FFMA
            if (MODE == 1) { FFMA2(acc[i], a2, b2); }                                      // 1 FFMA2
            if (MODE == 2) { FFMA2(acc[i], a2, b2); FSEL(f[i], m, p); }                    // FFMA2 + FSEL
            if (MODE == 3) { FFMA2(acc[i], a2, b2); FFMA2(acc[(i + 4) & 7], b2, a2); SHFL(s[i], src); }  // 2 FFMA2 + SHFL
            if (MODE == 4) { FFMA2(acc[i], a2, b2); FFMA(f[i], m, m); }                    // FFMA2 + FFMA
            if (MODE == 5) { FSEL(f[i], m, p); }                                           // FSEL
            if (MODE == 6) { SHFL(s[i], src); }                                            // SHFL
            if (MODE == 7) { FFMA2(acc[i], a2, b2); FFMA2(acc[(i + 4) & 7], b2, a2); FFMA2(acc[(i + 2) & 7], b2, b2); FFMA2(acc[(i + 6) & 7], a2, a2);
                             SHFL(s[i], src); SHFL(f[i], src); FSEL(f[(i + 1) & 7], m, p); FMUL(s[(i + 1) & 7], m); }  // GJ-like: 4 FFMA2 + 2 SHFL + FSEL + FMUL
            if (MODE == 8) { FFMA2(acc[i], a2, b2); FMUL(f[i], m); }                       // FFMA2 + FMUL
            if (MODE == 9) { FFMA(f[i], m, m); FSEL(s[i], m, p); }                         // FFMA + FSEL
        }
    }
    const long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += f[i] + s[i] + __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int warps) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    mix<MODE><<<148, warps * 32>>>(out, cyc, 0);
    mix<MODE><<<148, warps * 32>>>(out, cyc, 0);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double inst_per_smsp = (double)ITER * 8 * per_iter * warps / 4.0;
    printf("%-44s warps=%2d  SMSP-cycles per warp-inst = %.3f   (%s)\n", name, warps, avg / inst_per_smsp, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int warps : {4, 8, 16}) {
        run<0>("FFMA", 2, warps);
        run<1>("FFMA2", 1, warps);
        run<2>("FFMA2 + FSEL", 2, warps);
        run<3>("2 FFMA2 + SHFL", 3, warps);
        run<4>("FFMA2 + FFMA", 2, warps);
        run<5>("FSEL", 1, warps);
        run<6>("SHFL", 1, warps);
        run<7>("4 FFMA2 + 2 SHFL + FSEL + FMUL (GJ-like)", 8, warps);
        run<8>("FFMA2 + FMUL", 2, warps);
        run<9>("FFMA + FSEL", 2, warps);
    }
    return 0;
}
