// Micro-benchmark (dev only): dispatch cost of the instruction classes the Gauss-Jordan step is made of, on one
// B200 SM sub-partition.  Round 1's issue_mix.cu had 8-instruction bodies, so its numbers carried ~0.3 cycles of
// loop overhead per instruction; here a loop iteration holds 256+ instructions of the class under test (checked
// in the SASS: cuobjdump -sass issue_costs | grep -c ...), on 16 independent registers per class, 3 or 4 warps per
// scheduler.  Output: SMSP-cycles per warp instruction (per group for the mixes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o issue_costs issue_costs.cu && ./issue_costs
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 200
#define REP 16  // unrolled repetitions of the 16-register body per loop iteration

static __constant__ float cOne = 1.0f;

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
// the forms the kernel uses: FFMA2 with a scalar (.F32) multiplier, accumulating in place
#define FFMA2S(acc, s, r2) asm volatile("{ .reg .b64 t; mov.b64 t, {%1, %1}; fma.rn.f32x2 %0, t, %2, %0; }" : "+l"(acc) : "f"(s), "l"(r2))
#define FFMA2P(acc, a2, r2) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a2), "l"(r2))
#define FFMA(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))
#define FMUL(x, y) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(y))
#define PFMUL(x, y, one, p) asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q mul.rn.f32 %0, %1, %2; }" : "+f"(x) : "f"(y), "f"(one), "r"(p))
#define FSEL(x, y, p) asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; selp.f32 %0, %1, %0, q; }" : "+f"(x) : "f"(y), "r"(p))
#define PMOV(x, y, p) asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q mov.f32 %0, %1; }" : "+f"(x) : "f"(y), "r"(p))
#define SHFL(x, src) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+f"(x) : "r"(src))
#define LOP(x, y) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define IADD(x, y) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define IMAD(x, y, z) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define LDS32(x, addr) asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr))
#define LDS128(x, y, z, w, addr) asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr))
#define STS32(addr, x) asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory")
#define REDUX(d, x) asm volatile("redux.sync.max.abs.f32 %0, %1, 0xffffffff;" : "=f"(d) : "f"(x))

template <int MODE>
__global__ void __launch_bounds__(512, 1) mix(float* out, long long* cyc, int p, int q) {
    __shared__ float sm[2048];
    const int lane = threadIdx.x & 31;
    unsigned long long acc[16], r2 = pack(1.0001f, 0.9999f);
    float f[16], s[16];
    unsigned long long rr[16];
    int w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = pack((float)i, (float)(i + threadIdx.x)); f[i] = i + lane; s[i] = 1.0f + 1e-6f * (float)(i * q + p); w[i] = lane * i + q; }
#pragma unroll
    for (int i = 0; i < 16; ++i) rr[i] = pack(1.0f + 1e-7f * (i + lane), 1.0f - 1e-7f * i);
    sm[threadIdx.x] = lane; sm[threadIdx.x + 512] = lane; sm[threadIdx.x + 1024] = 1; sm[threadIdx.x + 1536] = 2;
    const float m = 1.0001f + 1e-6f * (float)q, one = cOne;
    const int src = (lane + 5) & 31;
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + 16 * lane + (threadIdx.x >> 5) * 512 % 4096;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int rep = 0; rep < REP; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) FFMA(f[i], m, s[i]);
                if (MODE == 1) FFMA2S(acc[i], s[i], r2);
                if (MODE == 2) FMUL(f[i], m);
                if (MODE == 3) PFMUL(f[i], f[(i + 1) & 15], one, p);
                if (MODE == 4) FSEL(f[i], f[(i + 1) & 15], p);
                if (MODE == 5) PMOV(f[i], f[(i + 1) & 15], p);
                if (MODE == 6) SHFL(f[i], src);
                if (MODE == 7) LOP(w[i], w[(i + 1) & 15]);
                if (MODE == 8) IADD(w[i], w[(i + 1) & 15]);
                if (MODE == 9) IMAD(w[i], w[(i + 1) & 15], w[(i + 2) & 15]);
                if (MODE == 10) LDS32(f[i], saddr + 4 * i);
                if (MODE == 11) LDS128(f[i], f[(i + 1) & 15], f[(i + 2) & 15], f[(i + 3) & 15], saddr);
                if (MODE == 12) STS32(saddr + 4 * (i & 3), f[i]);
                if (MODE == 13) REDUX(f[i], f[i]);
                if (MODE == 14) FFMA2S(acc[i], s[i], rr[i]);                     // everything fresh: acc pair, scalar, r pair
                if (MODE == 15) FFMA2S(acc[i], s[i], rr[rep & 3]);                // r pair shared by 16 consecutive FFMA2 (the kernel's order)
                if (MODE == 16) FFMA2S(acc[i], s[rep & 7], rr[i]);                // scalar shared by 16 consecutive FFMA2
                if (MODE == 17) FFMA2P(acc[i], rr[(i + 5) & 15], rr[i]);          // three fresh pairs
                if (MODE == 18) { FFMA(f[i], s[i], s[(i + 3) & 15]); }            // FFMA, three fresh registers
                // mixes (cycles reported per group)
                if (MODE == 20) { FFMA2S(acc[i], s[i], r2); FFMA2S(acc[(i + 8) & 15], s[(i + 3) & 15], r2); SHFL(f[i], src); }  // 2 FFMA2 + SHFL
                if (MODE == 21) { FFMA2S(acc[i], s[i], r2); FSEL(f[i], f[(i + 1) & 15], p); }                                         // FFMA2 + FSEL
                if (MODE == 22) { FFMA2S(acc[i], s[i], r2); PFMUL(f[i], f[(i + 1) & 15], one, p); }                                   // FFMA2 + @P FMUL
                if (MODE == 23) { FFMA2S(acc[i], s[i], r2); PMOV(f[i], f[(i + 1) & 15], p); }                                         // FFMA2 + @P MOV
                if (MODE == 24) { FFMA(f[i], m, s[i]); FSEL(s[i], s[(i + 1) & 15], p); }                                                 // FFMA + FSEL
                if (MODE == 25) { FFMA(f[i], m, s[i]); SHFL(s[i], src); }                                                  // FFMA + SHFL
                if (MODE == 26) { FSEL(f[i], f[(i + 1) & 15], p); SHFL(s[i], src); }                                                  // FSEL + SHFL
                // lean step, per 1/8 of a step: 4 FFMA2 + 2 SHFL + 2 FMUL (nf + fix-up) ; r1 step: 4 FFMA2 + 2 SHFL + FMUL + FSEL + MOV-like
                if (MODE == 30) { FFMA2S(acc[i], s[i], r2); FFMA2S(acc[(i + 4) & 15], s[i], r2); FFMA2S(acc[(i + 8) & 15], s[i], r2); FFMA2S(acc[(i + 12) & 15], s[i], r2);
                                  SHFL(f[i], src); SHFL(f[(i + 5) & 15], src); FMUL(s[i], m); PFMUL(f[(i + 9) & 15], f[(i + 2) & 15], one, p); }
                if (MODE == 31) { FFMA2S(acc[i], s[i], r2); FFMA2S(acc[(i + 4) & 15], s[i], r2); FFMA2S(acc[(i + 8) & 15], s[i], r2); FFMA2S(acc[(i + 12) & 15], s[i], r2);
                                  SHFL(f[i], src); SHFL(f[(i + 5) & 15], src); FMUL(s[i], m); FSEL(f[(i + 9) & 15], f[(i + 2) & 15], p); PMOV(f[(i + 3) & 15], f[(i + 7) & 15], p); }
                if (MODE == 32) { FFMA2S(acc[i], s[i], r2); FFMA2S(acc[(i + 4) & 15], s[i], r2); FFMA2S(acc[(i + 8) & 15], s[i], r2); FFMA2S(acc[(i + 12) & 15], s[i], r2);
                                  SHFL(f[i], src); SHFL(f[(i + 5) & 15], src); FMUL(s[i], m); }
            }
        }
    }
    const long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += f[i] + s[i] + __uint_as_float((unsigned)rr[i]) + __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32)) + (float)w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    mix<MODE><<<148, warps * 32>>>(out, cyc, 1, 3);
    mix<MODE><<<148, warps * 32>>>(out, cyc, 1, 3);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double groups_per_smsp = (double)ITER * REP * 16 * warps / 4.0;
    printf("{\"mix\": \"%s\", \"warps_per_sm\": %d, \"smsp_cycles_per_group\": %.3f, \"err\": \"%s\"}\n", name, warps, avg / groups_per_smsp,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int warps : {4, 12, 16}) {
        run<0>("FFMA", warps);
        run<1>("FFMA2 (scalar multiplier)", warps);
        run<2>("FMUL", warps);
        run<3>("@P FMUL", warps);
        run<4>("FSEL", warps);
        run<5>("@P MOV", warps);
        run<6>("SHFL.IDX", warps);
        run<7>("LOP3", warps);
        run<8>("IADD3", warps);
        run<9>("IMAD", warps);
        run<10>("LDS.32", warps);
        run<11>("LDS.128", warps);
        run<12>("STS.32", warps);
        run<13>("CREDUX.MAXABS", warps);
        run<14>("FFMA2, all operands fresh (pair, scalar, pair)", warps);
        run<15>("FFMA2, r pair shared by 16 consecutive", warps);
        run<16>("FFMA2, scalar shared by 16 consecutive", warps);
        run<17>("FFMA2, three fresh pairs", warps);
        run<18>("FFMA, three fresh registers", warps);
        run<20>("2 FFMA2 + SHFL", warps);
        run<21>("FFMA2 + FSEL", warps);
        run<22>("FFMA2 + @P FMUL", warps);
        run<23>("FFMA2 + @P MOV", warps);
        run<24>("FFMA + FSEL", warps);
        run<25>("FFMA + SHFL", warps);
        run<26>("FSEL + SHFL", warps);
        run<30>("lean 1/8 step: 4 FFMA2 + 2 SHFL + FMUL + @P FMUL", warps);
        run<31>("r1 1/8 step: 4 FFMA2 + 2 SHFL + FMUL + FSEL + @P MOV", warps);
        run<32>("minimal 1/8 step: 4 FFMA2 + 2 SHFL + FMUL", warps);
    }
    return 0;
}
