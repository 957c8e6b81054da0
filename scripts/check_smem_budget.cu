// Host-only check that every bulk-copy staged configuration (inverse, factors only, pivot_mode 3; N = 2..32, both dtypes) asks for
// at most the 227 KB of dynamic shared memory an SM has:  nvcc -std=c++17 -Imatrixinversion_b200/csrc -Iinclude -gencode arch=compute_100a,code=sm_100a scripts/check_smem_budget.cu -o /tmp/q && /tmp/q

#include "lub_launch.cuh"
#include <cstdio>
using namespace lub;
template <typename T, int N> void chkL() {
    using BC = BulkCfg<T, N, kModeLapack>; using BL = BulkLayout<T, N, BC::GR, BC::GC, kModeLapack>;
    int s = BL::smem_bytes(BC::MAXT / 32, (BC::OPT & kBulkSingle) ? 1 : 2);
    if (s > 232448) printf("TOO BIG lapack es=%d N=%d smem=%d maxt=%d\n", (int)sizeof(T), N, s, BC::MAXT);
}
template <typename T, int N, int MODE> void chk() {
    using BC = BulkCfg<T, N, MODE>; using BL = BulkLayout<T, N, BC::GR, BC::GC, MODE>;
    if (!BC::ON) return;
    int s = BL::smem_bytes(BC::MAXT / 32, 2);
    if (s > 232448) printf("TOO BIG es=%d N=%d mode=%d smem=%d\n", (int)sizeof(T), N, MODE, s);
}
template <typename T, int N, int MODE> void chkLu() {
    using BC = BulkLuCfg<T, N, MODE>; using BL = BulkLayout<T, N, BC::GR, BC::GC, MODE>;
    int s = BL::smem_bytes(BC::CAP / 32, BC::NIMG);
    if (s > 232448) printf("TOO BIG LU es=%d N=%d mode=%d cap=%d smem=%d\n", (int)sizeof(T), N, MODE, BC::CAP, s);
}
template <int N> void all() { chkL<float,N>(); chkL<double,N>(); chk<float,N,0>(); chk<float,N,1>(); chk<float,N,2>(); chk<double,N,0>(); chk<double,N,1>(); chk<double,N,2>();
  chkLu<float,N,0>(); chkLu<float,N,2>(); chkLu<double,N,0>(); chkLu<double,N,2>(); if constexpr (N < 32) all<N+1>(); }
int main() { all<2>(); printf("done\n"); }
