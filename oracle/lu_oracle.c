/*
 * TEST INFRASTRUCTURE ONLY -- see lu_oracle_impl.h.  Builds liblu_oracle.so:
 *   gcc -O3 -fopenmp -ffp-contract=off -shared -fPIC lu_oracle.c -o liblu_oracle.so -lm
 * -ffp-contract=off matters: the FMA / non-FMA choice is made explicitly per call.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OT float
#define OSUF _f32
#define OFMA fmaf
#define OFABS fabsf
#include "lu_oracle_impl.h"
#undef OT
#undef OSUF
#undef OFMA
#undef OFABS

#define OT double
#define OSUF _f64
#define OFMA fma
#define OFABS fabs
#include "lu_oracle_impl.h"
#undef OT
#undef OSUF
#undef OFMA
#undef OFABS

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
