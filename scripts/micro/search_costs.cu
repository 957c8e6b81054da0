// Micro-benchmark (dev only): dispatch cost of the warp collectives and ALU-pipe instructions the pivot searches are
// built from (same harness as issue_costs.cu: 256-instruction loop bodies, SMSP-cycles per warp instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o search_costs search_costs.cu && ./search_costs
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 200
#define REP 16

#define CREDUX(d, x) asm volatile("redux.sync.max.abs.f32 %0, %1, 0xffffffff;" : "=f"(d) : "f"(x))
#define REDUXOR(d, x) asm volatile("redux.sync.or.b32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(x))
#define REDUXADD(d, x) asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(x))
#define REDUXXOR(d, x) asm volatile("redux.sync.xor.b32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(x))
#define REDUXMIN(d, x) asm volatile("redux.sync.min.s32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(x))
#define REDUXMAX(d, x) asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(x))
#define BALLOT(d, x) asm volatile("{ .reg .pred q; setp.ne.s32 q, %1, 0; vote.sync.ballot.b32 %0, q, 0xffffffff; }" : "=r"(d) : "r"(x))
#define VOTEANY(d, x) asm volatile("{ .reg .pred q, r; setp.ne.s32 q, %1, 0; vote.sync.any.pred r, q, 0xffffffff; selp.b32 %0, 1, 0, r; }" : "=r"(d) : "r"(x))
#define SELP(x, y, p) asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; selp.b32 %0, %1, %0, q; }" : "+r"(x) : "r"(y), "r"(p))
#define LOP(x, y) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define SHFLI(x, src) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(x) : "r"(src))
#define FMUL(x, y) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(y))

template <int MODE>
__global__ void __launch_bounds__(512, 1) mix(float* out, long long* cyc, int p, int q) {
    const int lane = threadIdx.x & 31;
    float f[16];
    int w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { f[i] = i + lane; w[i] = lane * i + q; }
    const float m = 1.0001f + 1e-6f * (float)q;
    const int src = (lane + 5) & 31;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int rep = 0; rep < REP; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) CREDUX(f[i], f[i]);
                if (MODE == 1) REDUXOR(w[i], w[i]);
                if (MODE == 2) REDUXMAX(w[i], w[i]);
                if (MODE == 3) BALLOT(w[i], w[i]);
                if (MODE == 4) VOTEANY(w[i], w[i]);
                if (MODE == 5) SELP(w[i], w[(i + 1) & 15], p);
                if (MODE == 6) SHFLI(w[i], src);
                if (MODE == 7) REDUXADD(w[i], w[i]);
                if (MODE == 8) REDUXXOR(w[i], w[i]);
                if (MODE == 9) REDUXMIN(w[i], w[i]);
                // the same step with REDUX.ADD as the exchange (equal to the OR whenever one lane contributes)
                if (MODE == 13) {
                    float v = ((w[i] & (0x4d8d9b80 >> (i & 7))) != 0) ? f[i] : 0.0f, mx;
                    CREDUX(mx, v);
                    const bool hit = fabsf(v) == mx;
                    int hp; REDUXADD(hp, hit ? w[i] : 0);
                    if (hit || w[i] == (1 << i)) w[i] ^= hp ^ (1 << i);
                }
                // one step of the position-aware search, as compiled (per matrix): select the key, CREDUX, compare, select the
                // position word, REDUX.OR, compare + predicated XOR
                if (MODE == 10) {
                    float v = ((w[i] & (0x4d8d9b80 >> (i & 7))) != 0) ? f[i] : 0.0f, mx;
                    CREDUX(mx, v);
                    const bool hit = fabsf(v) == mx;
                    int hp; REDUXOR(hp, hit ? w[i] : 0);
                    if (hit || w[i] == (1 << i)) w[i] ^= hp ^ (1 << i);
                }
                // one step of the row-wise search in floating point: FMUL, CREDUX, FSET, FFMA, FFMA
                if (MODE == 11) {
                    float v = f[i] * f[(i + 1) & 15], mx;
                    CREDUX(mx, v);
                    const float hit = (fabsf(v) == mx) ? 1.0f : 0.0f;
                    f[(i + 2) & 15] = fmaf(hit, (float)i, f[(i + 2) & 15]);
                    f[(i + 1) & 15] = fmaf(-hit, f[(i + 1) & 15], f[(i + 1) & 15]);
                }
                // the exchange done with a ballot, find-first-set and a shuffle instead of the REDUX.OR
                if (MODE == 12) {
                    float v = ((w[i] & (0x4d8d9b80 >> (i & 7))) != 0) ? f[i] : 0.0f, mx;
                    CREDUX(mx, v);
                    const bool hit = fabsf(v) == mx;
                    const unsigned b = __ballot_sync(0xffffffffu, hit);
                    const int hp = __shfl_sync(0xffffffffu, w[i], __ffs(b) - 1);
                    if (hit || w[i] == (1 << i)) w[i] ^= hp ^ (1 << i);
                }
            }
        }
    }
    const long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += f[i] + (float)w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + m;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// dependent chains, one warp per SM: cycles per link = latency
template <int MODE>
__global__ void chain(float* out, long long* cyc, int q) {
    const int lane = threadIdx.x & 31;
    float f = 1.0f + lane, g = 0.5f;
    unsigned w = 1u << lane, acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            if (MODE == 0) { float mx; CREDUX(mx, f); f = mx * 1.0001f + (float)lane; }                 // CREDUX.MAXABS + FFMA
            if (MODE == 1) { unsigned mx; REDUXMAX(mx, w); w = (mx ^ (unsigned)lane) | 1u; }                 // CREDUX.MAX + LOP3
            if (MODE == 2) { unsigned mx; REDUXOR(mx, w); w = (mx ^ (unsigned)lane) | 1u; }                  // REDUX.OR + LOP3
            if (MODE == 3) { f = __shfl_xor_sync(0xffffffffu, f, 1) * 1.0001f; }                             // SHFL + FMUL
            if (MODE == 4) { const unsigned b = __ballot_sync(0xffffffffu, f > g); f = (b & (1u << lane)) ? f * 1.0001f : f * 0.9999f; }  // VOTE + select
            if (MODE == 5) {  // one step of the position-aware search
                float v = ((w & (0x4d8d9b80u >> (i & 7))) != 0u) ? f : 0.0f, mx;
                CREDUX(mx, v);
                const bool hit = fabsf(v) == mx;
                const unsigned mine = hit ? w : 0u;
                unsigned hp; REDUXMAX(hp, mine);
                acc |= mine;
                if (hit || w == (1u << (i & 31))) w ^= hp ^ (1u << (i & 31));
                f = f * 1.0001f + (float)(w & 1u);
            }
            if (MODE == 6) {  // one step of the row-wise search
                float v = f * g, mx;
                CREDUX(mx, v);
                const float hit = (fabsf(v) == mx) ? 1.0f : 0.0f;
                g = fmaf(-hit, g, g) + 1.0f;
                f = fmaf(hit, (float)i, f);
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 32 + lane] = f + g + (float)w + (float)acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run_chain(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    chain<MODE><<<148, 32>>>(out, cyc, 3);
    chain<MODE><<<148, 32>>>(out, cyc, 3);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("{\"chain\": \"%s\", \"cycles_per_link\": %.1f, \"err\": \"%s\"}\n", name, avg / (ITER * 64.0), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
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
    run_chain<0>("CREDUX.MAXABS.F32 + FFMA");
    run_chain<1>("CREDUX.MAX.U32 + LOP3");
    run_chain<2>("REDUX.OR + LOP3");
    run_chain<3>("SHFL.BFLY + FMUL");
    run_chain<4>("VOTE.BALLOT + select");
    run_chain<5>("position-aware search step");
    run_chain<6>("row-wise search step");
    for (int warps : {4, 12}) {
        run<0>("CREDUX.MAXABS.F32", warps);
        run<1>("REDUX.OR", warps);
        run<2>("REDUX.MAX.U32", warps);
        run<3>("VOTE.BALLOT", warps);
        run<4>("VOTE.ANY (+SEL)", warps);
        run<5>("SEL", warps);
        run<6>("SHFL.IDX", warps);
        run<7>("REDUX.ADD", warps);
        run<8>("REDUX.XOR", warps);
        run<9>("REDUX.MIN.S32", warps);
        run<10>("position-aware search step (REDUX.OR exchange)", warps);
        run<11>("row-wise search step (FMA pipe)", warps);
        run<12>("position-aware search step (ballot + shuffle exchange)", warps);
        run<13>("position-aware search step (REDUX.ADD exchange)", warps);
    }
    return 0;
}
