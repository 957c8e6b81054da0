// Micro-benchmark (dev only): the rank-1 update of one Gauss-Jordan step, a[li][lj] += nf[li] * r[lj] on an 8 x 8
// register block, issued in different ORDERS and FORMS.  Round 2 found the FMA pipe's real limit for this kernel is
// not the pipe but the register file: an FFMA / FFMA2 whose three operands all come fresh from the register file
// costs ~1.45 / ~2.9 cycles, one with an operand served by the operand-reuse cache 1.0 / 2.0 (issue_costs.cu).
// The order decides which operand the reuse cache can serve.  asm volatile pins the order.
//   P1  FFMA2, row-major   (scalar nf[li] shared by 4 consecutive FFMA2, the 4 r pairs rotate)
//   P2  FFMA2, pair-major  (r pair shared by 8 consecutive FFMA2, the 8 scalars rotate)   <- what ptxas picks by itself
//   P3  FFMA,  row-major   (nf[li] shared by 8 consecutive FFMA)
//   P4  FFMA,  column-major (r[lj] shared by 8 consecutive FFMA)
//   P5  FFMA2, row-major, column pairs taken from the *rows* (pair = {a[li][lj], a[li+1][lj]}, scalar r[lj], nf pairs)
// Output: SMSP cycles per 8 x 8 update (64 FMA per lane; 64 = the FP32 pipe's peak).
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 100
#define REP 8

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
#define FFMA2S(acc, s, r2) asm volatile("{ .reg .b64 t; mov.b64 t, {%1, %1}; fma.rn.f32x2 %0, t, %2, %0; }" : "+l"(acc) : "f"(s), "l"(r2))
#define SHFL(x, src) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+f"(x) : "r"(src))
#define FMUL(x, y) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(y))
#define FSEL(x, y, p) asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; selp.f32 %0, %1, %0, q; }" : "+f"(x) : "f"(y), "r"(p))
#define FFMA(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))

template <int MODE>
__global__ void __launch_bounds__(512, 1) upd(float* out, long long* cyc, int p, int q) {
    const int lane = threadIdx.x & 31;
    unsigned long long a2[8][4], r2[4];
    float a[8][8], nf[8], r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        nf[i] = 1e-4f * (float)(i * q + p + lane);
        r[i] = 1.0f + 1e-5f * (float)(i * p + q + lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[i][j] = (float)(i * 8 + j + lane * q);
#pragma unroll
        for (int j = 0; j < 4; ++j) a2[i][j] = pack((float)(i * 8 + 2 * j + lane * q), (float)(i * 8 + 2 * j + 1 + p));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) r2[j] = pack(r[2 * j], r[2 * j + 1]);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int rep = 0; rep < REP; ++rep) {
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) FFMA2S(a2[i][j], nf[i], r2[j]);
            }
            if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) FFMA2S(a2[i][j], nf[i], r2[j]);
            }
            if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) FFMA(a[i][j], nf[i], r[j]);
            }
            if (MODE == 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) FFMA(a[i][j], nf[i], r[j]);
            }
            if (MODE >= 6 && MODE <= 11) {  // pair-major FFMA2 + 16 other instructions, inside the runs (even MODE) or at the 4 run boundaries (odd MODE)
                const int src = (lane + 5) & 31;
                const float m = nf[7];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        FFMA2S(a2[i][j], nf[i], r2[j]);
                        if ((MODE & 1) == 0 && (i & 1)) {
                            if (MODE == 6) SHFL(a[j][i >> 1], src);
                            if (MODE == 8) FMUL(a[j][i >> 1], m);
                            if (MODE == 10) FSEL(a[j][i >> 1], a[j][4 + (i >> 1)], p);
                        }
                    }
                    if (MODE & 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (MODE == 7) SHFL(a[j][i], src);
                            if (MODE == 9) FMUL(a[j][i], m);
                            if (MODE == 11) FSEL(a[j][i], a[j][4 + i], p);
                        }
                    }
                }
            }
            if (MODE == 5) {  // pairs along rows: a2[j][i/2] = {a[i][j], a[i+1][j]}; multiplier pair nf2 (in r2), scalar r[j]
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) FFMA2S(a2[j][i], r[j], r2[i]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s += a[i][j];
#pragma unroll
        for (int j = 0; j < 4; ++j) s += __uint_as_float((unsigned)a2[i][j]) + __uint_as_float((unsigned)(a2[i][j] >> 32));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    upd<MODE><<<148, warps * 32>>>(out, cyc, 1, 3);
    upd<MODE><<<148, warps * 32>>>(out, cyc, 1, 3);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double upd_per_smsp = (double)ITER * REP * warps / 4.0;
    printf("{\"order\": \"%s\", \"warps_per_sm\": %d, \"smsp_cycles_per_8x8_update\": %.2f, \"err\": \"%s\"}\n", name, warps, avg / upd_per_smsp,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int warps : {8, 12, 16}) {
        run<1>("FFMA2 row-major (scalar reused x4)", warps);
        run<2>("FFMA2 pair-major (r pair reused x8)", warps);
        run<3>("FFMA row-major (nf reused x8)", warps);
        run<4>("FFMA column-major (r reused x8)", warps);
        run<5>("FFMA2 row pairs (scalar r reused x4)", warps);
        run<6>("FFMA2 pair-major + 16 SHFL inside the runs", warps);
        run<7>("FFMA2 pair-major + 16 SHFL at the run boundaries", warps);
        run<8>("FFMA2 pair-major + 16 FMUL inside the runs", warps);
        run<9>("FFMA2 pair-major + 16 FMUL at the run boundaries", warps);
        run<10>("FFMA2 pair-major + 16 FSEL inside the runs", warps);
        run<11>("FFMA2 pair-major + 16 FSEL at the run boundaries", warps);
    }
    return 0;
}
