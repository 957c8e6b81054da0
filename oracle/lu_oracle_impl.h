/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's
 * luBatchedInplace hot path.  Nothing under matrixinversion_b200/ may include,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * reported CPU baseline.
 *
 * This header is included twice by lu_oracle.c, once per arithmetic type:
 *   #define OT float  / OSUF _f32 / OFMA fmaf / OFABS fabsf
 *   #define OT double / OSUF _f64 / OFMA fma  / OFABS fabs
 *
 * What it restates (reference file:line):
 *   step order of the kernel ............ parallel_pivot/luBatchedInplace.cuh:156-192
 *   comp_U (row k of U) ................. templated/luBatchedInplace.cuh:11-21
 *   comp_L (column k of L) .............. templated/luBatchedInplace.cuh:24-34
 *   inversion (mode none) ............... templated/luBatchedInplace.cuh:37-75
 *   find_pivot / swap_rows (serial) ..... serial_pivot/luBatchedInplace.cuh:12-36
 *   find_pivot_parallel (tree arg-max) .. parallel_pivot/luBatchedInplace.cuh:12-44
 *   perm-vector update .................. parallel_pivot/luBatchedInplace.cuh:161-168
 *   comp_inv (pivoted solve) ............ parallel_pivot/luBatchedInplace.cuh:84-125
 *   verifyInv predicate ................. templated/verify.hpp:50-103
 *   pivotedA (host pivot order) ......... parallel_pivot/verify.hpp:106-155
 *
 * Arithmetic notes.  nvcc contracts `sum += a*b` into an FMA (default -fmad=true),
 * so `use_fma != 0` accumulates with fma(); `use_fma == 0` rounds the product
 * first (what an x86 build of the same C++ would do); `use_fma == 2` is the plain expression
 * `sum + a*b` for the CPU-baseline timing leg (no libm fma() call, no volatile).  Division is IEEE here; the
 * reference sweep builds with --use_fast_math (templated/run.py:47), i.e. an
 * approximate division, which is why value parity is a tolerance and only pivot
 * parity is bit-exact.
 */

#define OCAT2(a, b) a##b
#define OCAT(a, b) OCAT2(a, b)
#define ONAME(base) OCAT(base, OSUF)

#ifndef ORACLE_MAXN
#define ORACLE_MAXN 64
#endif

/* serial_pivot/luBatchedInplace.cuh:22-36: seed (|A[k][k]|, k), strict '>' so the
 * lowest row index wins ties. */
static int ONAME(o_find_pivot_serial)(const OT *A, int n, int k)
{
    int p = k;
    OT m = OFABS(A[k * n + k]);
    for (int i = k + 1; i < n; ++i) {
        OT v = OFABS(A[i * n + k]);
        if (v > m) { m = v; p = i; }
    }
    return p;
}

/* parallel_pivot/luBatchedInplace.cuh:12-44 with threadsPerMatrix == tpm slots.
 * Every slot t seeds (|A[k][k]|, k), scans rows k+1+t, k+1+t+tpm, ... with strict
 * '>', then the shared-memory tree `for (stride = tpm/2; stride > 0; stride >>= 1)`
 * merges slot t+stride into slot t (t < stride) when vals[t] < vals[t+stride].
 * For a non-power-of-two tpm some slots never reach slot 0; this
 * literal emulation reproduces that. */
static int ONAME(o_find_pivot_parallel)(const OT *A, int n, int k, int tpm)
{
    OT vals[ORACLE_MAXN];
    int idx[ORACLE_MAXN];
    for (int t = 0; t < tpm; ++t) {
        OT m = OFABS(A[k * n + k]);
        int p = k;
        for (int i = k + 1 + t; i < n; i += tpm) {
            OT v = OFABS(A[i * n + k]);
            if (v > m) { m = v; p = i; }
        }
        vals[t] = m;
        idx[t] = p;
    }
    for (int stride = tpm / 2; stride > 0; stride >>= 1) {
        /* reads touch [stride, 2*stride), writes touch [0, stride): no overlap, so a
         * sequential sweep equals the GPU's simultaneous update. */
        for (int t = 0; t < stride; ++t) {
            if (vals[t] < vals[t + stride]) {
                vals[t] = vals[t + stride];
                idx[t] = idx[t + stride];
            }
        }
    }
    return idx[0];
}

static void ONAME(o_swap_rows)(OT *A, int n, int r1, int r2)
{
    for (int j = 0; j < n; ++j) {
        OT t = A[r1 * n + j];
        A[r1 * n + j] = A[r2 * n + j];
        A[r2 * n + j] = t;
    }
}

static inline OT ONAME(o_mac)(OT a, OT b, OT sum, int use_fma)
{
    if (use_fma == 2) return sum + a * b; /* timing leg: whatever the host compiler makes of it */
    if (use_fma) return OFMA(a, b, sum);
    volatile OT prod = a * b; /* volatile: forbid the host compiler from contracting */
    return sum + prod;
}

/* One matrix, in place.  mode 0 none / 1 serial / 2 parallel.  perm (may be NULL)
 * receives the permutation vector with reference semantics:
 * perm[i] = original row index sitting in row i after all swaps.  steps (may be
 * NULL) receives the pivot row chosen at every step k (the "pivot index sequence").
 * lu_only != 0 stops after the factorisation and leaves the packed LU in A. */
static void ONAME(o_invert_one)(OT *A, int n, int mode, int tpm, int use_fma,
                                int lu_only, int32_t *perm_out, int32_t *steps_out)
{
    int perm[ORACLE_MAXN];
    for (int i = 0; i < n; ++i) perm[i] = i;

    for (int k = 0; k < n; ++k) {
        if (mode != 0) {
            int p = (mode == 1) ? ONAME(o_find_pivot_serial)(A, n, k)
                                : ONAME(o_find_pivot_parallel)(A, n, k, tpm);
            if (steps_out) steps_out[k] = p;
            if (p != k) {
                int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
                ONAME(o_swap_rows)(A, n, k, p);
            }
        } else if (steps_out) {
            steps_out[k] = k;
        }
        /* comp_U: row k of U, lanes j >= k */
        for (int j = k; j < n; ++j) {
            OT sum = (OT)0;
            for (int l = 0; l < k; ++l) sum = ONAME(o_mac)(A[k * n + l], A[l * n + j], sum, use_fma);
            A[k * n + j] = A[k * n + j] - sum;
        }
        /* comp_L: column k of L, lanes i > k; uses the A[k][k] just written */
        for (int i = k + 1; i < n; ++i) {
            OT sum = (OT)0;
            for (int l = 0; l < k; ++l) sum = ONAME(o_mac)(A[i * n + l], A[l * n + k], sum, use_fma);
            A[i * n + k] = (A[i * n + k] - sum) / A[k * n + k];
        }
    }
    if (perm_out) for (int i = 0; i < n; ++i) perm_out[i] = perm[i];
    if (lu_only) return;

    /* inversion / comp_inv: every column is solved from the complete LU, then the
     * whole inverse replaces A (the GPU lanes all finish their solves before any
     * column is written back). */
    OT Xh[ORACLE_MAXN * ORACLE_MAXN];
    OT y[ORACLE_MAXN], x[ORACLE_MAXN];
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) {
            OT b = (perm[i] == c) ? (OT)1 : (OT)0;
            OT sum = (OT)0;
            for (int j = 0; j < i; ++j) sum = ONAME(o_mac)(A[i * n + j], y[j], sum, use_fma);
            y[i] = b - sum;
        }
        for (int i = n - 1; i >= 0; --i) {
            OT sum = (OT)0;
            for (int j = i + 1; j < n; ++j) sum = ONAME(o_mac)(A[i * n + j], x[j], sum, use_fma);
            x[i] = (y[i] - sum) / A[i * n + i];
        }
        for (int i = 0; i < n; ++i) Xh[i * n + c] = x[i];
    }
    memcpy(A, Xh, sizeof(OT) * (size_t)n * (size_t)n);
}

/* Batched entry: A is T[batch][n][n] row-major, overwritten by the inverses (or by
 * the packed LU when lu_only).  piv (may be NULL) is int32[batch][n] permutation
 * vectors; steps (may be NULL) int32[batch][n] per-step pivot rows.  threads <= 0
 * uses every OpenMP thread.  Returns the number of threads used. */
int ONAME(oracle_lu_batched)(OT *A, int32_t *piv, int32_t *steps, int n, int64_t batch,
                             int mode, int tpm, int use_fma, int lu_only, int threads)
{
    if (n < 1 || n > ORACLE_MAXN || mode < 0 || mode > 2) return -1;
    if (tpm <= 0) tpm = n;
    int used = 1;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel for num_threads(threads) schedule(static)
#endif
    for (int64_t b = 0; b < batch; ++b) {
        ONAME(o_invert_one)(A + b * (int64_t)n * n, n, mode, tpm, use_fma, lu_only,
                            piv ? piv + b * n : NULL, steps ? steps + b * n : NULL);
    }
    return used;
}

/* verifyInv (templated/verify.hpp:50-103): r(i,j) = sum_l A[j][l] * Ainv[l][i]
 * accumulated in T in increasing l; a matrix is correct iff every diagonal r has
 * |r - 1| < thr and every off-diagonal |r| < thr.  The reference hard-codes
 * thr = 1e-3 cast to T.  max_abs_dev (may be NULL) gets max |r - delta|. */
void ONAME(oracle_verify_inv)(const OT *A, const OT *Ainv, int n, int64_t batch, double thr,
                              int64_t *n_ok, int64_t *n_bad, double *max_abs_dev)
{
    OT threshold = (OT)thr;
    int64_t ok = 0, bad = 0;
    double worst = 0.0;
    for (int64_t k = 0; k < batch; ++k) {
        const OT *a = A + k * (int64_t)n * n, *x = Ainv + k * (int64_t)n * n;
        int idc = 0, offc = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                OT r = (OT)0;
                for (int l = 0; l < n; ++l) {
                    volatile OT prod = a[j * n + l] * x[l * n + i];
                    r += prod;
                }
                OT dev = (i == j) ? OFABS(r - (OT)1) : OFABS(r);
                if (i == j && dev < threshold) idc++;
                if (i != j && dev < threshold) offc++;
                if (!(dev <= worst)) worst = (double)dev; /* NaN propagates */
            }
        if (idc == n && offc == n * (n - 1)) ok++; else bad++;
    }
    if (n_ok) *n_ok = ok;
    if (n_bad) *n_bad = bad;
    if (max_abs_dev) *max_abs_dev = worst;
}

/* pivotedA (parallel_pivot/verify.hpp:106-155): host pivot order on the
 * UN-eliminated matrix.  Note the seed is A[i][i] WITHOUT fabs and the scan starts
 * at j = i; that is equivalent to find_pivot for every finite input (a negative or
 * zero seed is immediately replaced / kept by the j = i comparison). */
void ONAME(oracle_pivotedA)(const OT *A, OT *PA, int32_t *pivots, int n)
{
    for (int i = 0; i < n; ++i) pivots[i] = i;
    memcpy(PA, A, sizeof(OT) * (size_t)n * n);
    for (int i = 0; i < n; ++i) {
        OT max_val = PA[i * n + i];
        int max_idx = i;
        for (int j = i; j < n; ++j) {
            if (OFABS(PA[j * n + i]) > max_val) {
                max_val = OFABS(PA[j * n + i]);
                max_idx = j;
            }
        }
        if (max_idx != i) {
            int t = pivots[i]; pivots[i] = pivots[max_idx]; pivots[max_idx] = t;
            ONAME(o_swap_rows)(PA, n, i, max_idx);
        }
    }
}

#undef OCAT2
#undef OCAT
#undef ONAME
