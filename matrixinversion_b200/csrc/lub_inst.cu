// lub_inst.cu -- one translation unit per (dtype, pivot mode, block of eight N): the 192
// kernel instantiations are split 24 ways so they build in parallel.  Compile with
//   -DLUB_T=float|double -DLUB_TN=f32|f64 -DLUB_MODE=0|1|2 -DLUB_Q=0..3
#include "lub_launch.cuh"

#define LUB_CAT4_(a, b, c, d) lub_get_##a##_m##b##_q##c
#define LUB_CAT4(a, b, c) LUB_CAT4_(a, b, c, )
#define LUB_N(i) (LUB_Q * 8 + (i) + 1)

LUB_DEFINE_GETTER(LUB_CAT4(LUB_TN, LUB_MODE, LUB_Q), LUB_T, LUB_MODE, LUB_N(0), LUB_N(1), LUB_N(2), LUB_N(3),
                  LUB_N(4), LUB_N(5), LUB_N(6), LUB_N(7))
