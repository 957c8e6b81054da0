// lubatched_api.cu -- the C ABI of include/lubatched.h: argument checking, runtime dispatch
// over what the reference fixes at compile time (MATRIXSIZE / NUMTHREADS / FpType,
// templated/luBatchedInplace.cu:4-6, templated/verify.hpp:9-10), stream / timing state, the
// chunked host-pointer pipeline, the verify.hpp-compatible residual check and the text
// loader.  There is no CPU implementation of the inversion in this library.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <cctype>
#include <mutex>
#include <thread>
#include <pthread.h>
#include <sched.h>

#include "../../include/lubatched.h"
#include "lub_launch.cuh"

namespace lub {
#define LUB_DECL(TN, M) \
    LaunchFn lub_get_##TN##_m##M##_q0(int); LaunchFn lub_get_##TN##_m##M##_q1(int); \
    LaunchFn lub_get_##TN##_m##M##_q2(int); LaunchFn lub_get_##TN##_m##M##_q3(int);
LUB_DECL(f32, 0) LUB_DECL(f32, 1) LUB_DECL(f32, 2)
LUB_DECL(f64, 0) LUB_DECL(f64, 1) LUB_DECL(f64, 2)
#undef LUB_DECL

// pivot_mode 3 (lub_lapack_inst.cu)
cudaError_t launch_lapack_f32(void*, int32_t*, int32_t*, int, long long, int, cudaStream_t, LaunchInfo*, int, cudaEvent_t);
cudaError_t launch_lapack_f64(void*, int32_t*, int32_t*, int, long long, int, cudaStream_t, LaunchInfo*, int, cudaEvent_t);

// batch-interleaved layout (lub_interleaved_inst.cu)
cudaError_t launch_interleaved_f32(void*, int32_t*, int32_t*, int, long long, int, cudaStream_t, cudaEvent_t);
cudaError_t launch_interleaved_f64(void*, int32_t*, int32_t*, int, long long, int, cudaStream_t, cudaEvent_t);

static LaunchFn find_launcher(int n, int mode, int dtype) {
    using Getter = LaunchFn (*)(int);
#define LUB_ROW(TN, M) { lub_get_##TN##_m##M##_q0, lub_get_##TN##_m##M##_q1, lub_get_##TN##_m##M##_q2, lub_get_##TN##_m##M##_q3 }
    static const Getter table[2][3][4] = {
        { LUB_ROW(f32, 0), LUB_ROW(f32, 1), LUB_ROW(f32, 2) },
        { LUB_ROW(f64, 0), LUB_ROW(f64, 1), LUB_ROW(f64, 2) },
    };
#undef LUB_ROW
    return table[dtype][mode][(n - 1) / 8](n);
}
}  // namespace lub

namespace {

thread_local std::string g_err;
thread_local cudaStream_t g_stream = nullptr;
thread_local int g_threads = 0;
thread_local int g_opt_flags = 0;   // lub::kLaunchNoTma | lub::kLaunchNoDmma (lu_batched_set_option)
thread_local bool g_timing = false;
// start / stop events of the timed launch, one pair per device (events belong to the device they were created on)
thread_local cudaEvent_t g_ev[lub::kMaxDevices][2] = {};
thread_local int g_ev_dev = -1;  // device of the last timed launch, -1 = none

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    // a non-sticky error stays in the runtime's last-error slot until it is read: clear it, or the next launch's
    // cudaGetLastError() would report this failure again
    (void)cudaGetLastError();
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? LUB_ERR_NO_DEVICE : LUB_ERR_CUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

int check_args(int n, int64_t batch, int mode, int dtype) {
    if (n < 1 || n > 32) return fail(LUB_ERR_BAD_N, "n must be in [1, 32]");
    if (mode < LUB_PIVOT_NONE || mode > LUB_PIVOT_LAPACK) return fail(LUB_ERR_BAD_MODE, "pivot_mode must be 0 (none), 1 (serial), 2 (parallel) or 3 (LAPACK partial pivoting)");
    if (dtype != LUB_DTYPE_F32 && dtype != LUB_DTYPE_F64) return fail(LUB_ERR_BAD_DTYPE, "dtype must be 0 (fp32) or 1 (fp64)");
    if (batch < 0) return fail(LUB_ERR_BAD_ARG, "batch must be >= 0");
    return LUB_OK;
}

size_t esize(int dtype) { return dtype == LUB_DTYPE_F32 ? 4 : 8; }

// flags: lub::kLaunchDryRun | lub::kLaunchLuOnly
int launch_on(void* ptr, int32_t* piv, int n, int64_t batch, int mode, int dtype, cudaStream_t s,
              lub::LaunchInfo* info, int flags, int32_t* status = nullptr) {
    const bool dry = (flags & lub::kLaunchDryRun) != 0;
    int rc = check_args(n, batch, mode, dtype);
    if (rc != LUB_OK) return rc;
    if (!dry && batch > 0) {
        if (ptr == nullptr) return fail(LUB_ERR_BAD_ARG, "ptr is NULL");
        if (reinterpret_cast<uintptr_t>(ptr) % esize(dtype)) return fail(LUB_ERR_BAD_ARG, "ptr is not aligned to the element size");
    }
    lub::LaunchFn fn = (mode == LUB_PIVOT_LAPACK) ? nullptr : lub::find_launcher(n, mode, dtype);
    if (!fn && mode != LUB_PIVOT_LAPACK) return fail(LUB_ERR_BAD_N, "no kernel for this n");
    const bool timed = g_timing && !dry && batch > 0;
    int dev = 0;
    if (timed) {
        CU(cudaGetDevice(&dev));
        if (dev < 0 || dev >= lub::kMaxDevices) return fail(LUB_ERR_CUDA, "device index out of range");
        if (!g_ev[dev][0]) { CU(cudaEventCreate(&g_ev[dev][0])); CU(cudaEventCreate(&g_ev[dev][1])); }
        g_ev_dev = -1;
    }
    // the launcher records the start event itself, after its one-time preparation (function attributes,
    // occupancy query, tensor-map encode): the interval is the kernel's, also on the first call
    flags |= g_opt_flags;
    cudaError_t e;
    if (mode == LUB_PIVOT_LAPACK) {
        e = (dtype == LUB_DTYPE_F32 ? lub::launch_lapack_f32 : lub::launch_lapack_f64)(ptr, piv, status, n, (long long)batch, g_threads, s, info, flags,
                                                                                     timed ? g_ev[dev][0] : nullptr);
    } else {
        // the reference's variants have no numerical status: a status array given with modes 0-2 reads 0
        if (status && !dry && batch > 0) CU(cudaMemsetAsync(status, 0, (size_t)batch * sizeof(int32_t), s));
        e = fn(ptr, piv, (long long)batch, g_threads, s, info, flags, timed ? g_ev[dev][0] : nullptr);
    }
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    if (timed) { CU(cudaEventRecord(g_ev[dev][1], s)); g_ev_dev = dev; }
    return LUB_OK;
}

int launch_interleaved(void* ptr, int32_t* piv, int32_t* status, int n, int64_t batch, int mode, int dtype, cudaStream_t s) {
    int rc = check_args(n, batch, mode, dtype);
    if (rc != LUB_OK) return rc;
    if (n > 8) return fail(LUB_ERR_BAD_N, "the batch-interleaved layout holds one matrix per lane: n must be in [1, 8]");
    if (batch == 0) return LUB_OK;
    if (ptr == nullptr) return fail(LUB_ERR_BAD_ARG, "ptr is NULL");
    if (reinterpret_cast<uintptr_t>(ptr) % esize(dtype)) return fail(LUB_ERR_BAD_ARG, "ptr is not aligned to the element size");
    const bool timed = g_timing;
    int dev = 0;
    if (timed) {
        CU(cudaGetDevice(&dev));
        if (dev < 0 || dev >= lub::kMaxDevices) return fail(LUB_ERR_CUDA, "device index out of range");
        if (!g_ev[dev][0]) { CU(cudaEventCreate(&g_ev[dev][0])); CU(cudaEventCreate(&g_ev[dev][1])); }
        g_ev_dev = -1;
    }
    cudaError_t e = (dtype == LUB_DTYPE_F32 ? lub::launch_interleaved_f32 : lub::launch_interleaved_f64)(
        ptr, piv, status, n, (long long)batch, mode, s, timed ? g_ev[dev][0] : nullptr);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    if (timed) { CU(cudaEventRecord(g_ev[dev][1], s)); g_ev_dev = dev; }
    return LUB_OK;
}

// ---- verify.hpp-compatible residual check ---------------------------------------------------

template <typename T>
void verify_host(const T* A, const T* X, int n, int64_t batch, double thr, int64_t* ok, int64_t* bad, double* dev) {
    const T threshold = static_cast<T>(thr);
    int64_t good = 0;
    double worst = 0.0;
    bool saw_nan = false;
#pragma omp parallel for schedule(static) reduction(+ : good) reduction(max : worst) reduction(|| : saw_nan)
    for (int64_t k = 0; k < batch; ++k) {
        const T* a = A + k * (int64_t)n * n;
        const T* x = X + k * (int64_t)n * n;
        int id_cnt = 0, off_cnt = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                T r = T(0);
                for (int l = 0; l < n; ++l) r += a[j * n + l] * x[l * n + i];
                const T d = (i == j) ? std::fabs(r - T(1)) : std::fabs(r);
                if (d < threshold) (i == j ? id_cnt : off_cnt)++;
                if (d != d) saw_nan = true;
                else if ((double)d > worst) worst = (double)d;
            }
        if (id_cnt == n && off_cnt == n * (n - 1)) good++;
    }
    if (ok) *ok = good;
    if (bad) *bad = batch - good;
    if (dev) *dev = saw_nan ? NAN : worst;
}

// verifyLU (templated/verify.hpp:105-186) / verifyLUwithPivoting (parallel_pivot/verify.hpp:157-242):
// L = unit-lower part of LU, U = upper part, every |(P A)(i,j) - sum_l L(i,l) U(l,j)| < thr, the sum
// accumulated in T over ALL l (zeros included, as the reference does).  The reference builds P A once
// from pivotedA because all its matrices are equal; here each matrix brings its own permutation vector.
template <typename T>
void verify_lu_host(const T* A, const T* LU, const int32_t* piv, int n, int64_t batch, double thr, int64_t* ok, int64_t* bad, double* dev) {
    const T threshold = static_cast<T>(thr);
    int64_t good = 0;
    double worst = 0.0;
    bool saw_nan = false;
#pragma omp parallel for schedule(static) reduction(+ : good) reduction(max : worst) reduction(|| : saw_nan)
    for (int64_t k = 0; k < batch; ++k) {
        const T* a = A + k * (int64_t)n * n;
        const T* lu = LU + k * (int64_t)n * n;
        const int32_t* p = piv ? piv + k * (int64_t)n : nullptr;
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            const int src = p ? p[i] : i;
            for (int j = 0; j < n; ++j) {
                T r = T(0);
                for (int l = 0; l < n; ++l) {
                    const T lv = (l < i) ? lu[i * n + l] : (l == i ? T(1) : T(0));
                    const T uv = (l <= j) ? lu[l * n + j] : T(0);
                    r += lv * uv;
                }
                const T ref = (src >= 0 && src < n) ? a[src * n + j] : T(NAN);
                const T d = std::fabs(ref - r);
                if (d < threshold) cnt++;
                if (d != d) saw_nan = true;
                else if ((double)d > worst) worst = (double)d;
            }
        }
        if (cnt == n * n) good++;
    }
    if (ok) *ok = good;
    if (bad) *bad = batch - good;
    if (dev) *dev = saw_nan ? NAN : worst;
}

template <typename T>
__global__ void verify_kernel(const T* __restrict__ A, const T* __restrict__ X, int n, long long batch, T thr,
                              unsigned long long* good, unsigned int* worst_bits, unsigned int* nan_flag) {
    // one warp per matrix; lane handles entries e = lane, lane+32, ... of the n*n product
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long k = w; k < batch; k += nw) {
        const T* a = A + k * (long long)n * n;
        const T* x = X + k * (long long)n * n;
        bool all_ok = true;
        float worst = 0.f;
        bool nan = false;
        for (int e = lane; e < n * n; e += 32) {
            const int i = e / n, j = e % n;
            T r = T(0);
            for (int l = 0; l < n; ++l) r += a[j * n + l] * x[l * n + i];
            const T d = (i == j) ? fabs(r - T(1)) : fabs(r);
            if (!(d < thr)) all_ok = false;
            if (d != d) nan = true;
            else worst = fmaxf(worst, (float)d);
        }
        all_ok = __all_sync(0xffffffffu, all_ok);
        nan = __any_sync(0xffffffffu, nan);
        for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        if (lane == 0) {
            if (all_ok) atomicAdd(good, 1ull);
            atomicMax(worst_bits, __float_as_uint(worst));
            if (nan) atomicOr(nan_flag, 1u);
        }
    }
}

// ---- the host-pointer pipeline -------------------------------------------------------------

// Device-side staging of one device: three chunk buffers on three streams (H2D / kernel / D2H of consecutive chunks
// overlap).  One per device, created on first use, serialised by its mutex (two host threads may target one GPU).
struct HostPipe {
    static constexpr int kSlots = 3;
    std::mutex mu;
    void* buf[kSlots] = {nullptr, nullptr, nullptr};
    int32_t* pbuf[kSlots] = {nullptr, nullptr, nullptr};
    size_t cap = 0, pcap = 0;
    cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
};
HostPipe g_pipes[lub::kMaxDevices];

unsigned long long host_chunk_mib() {  // tuning knob, default 64 MiB (profiles/r01_tune_v6.md section 4)
    static const unsigned long long v = []() -> unsigned long long {
        const char* e = std::getenv("LUB_HOST_CHUNK_MIB");
        const long x = e ? std::atol(e) : 0;
        return (x >= 1 && x <= 4096) ? (unsigned long long)x : 64ull;
    }();
    return v;
}

// main()'s cudaMalloc + H2D + launch + D2H (parallel_pivot/luBatchedInplace.cu:112-135) for `batch` matrices that start
// at host_ptr, on the CURRENT device, chunked and pipelined.  Synchronous.
int host_pipeline(void* host_ptr, int32_t* host_piv, int n, int64_t batch, int mode, int dtype) {
    if (batch == 0) return LUB_OK;
    const size_t mat_bytes = (size_t)n * n * esize(dtype);
    // ~64 MiB chunks, a whole number of matrices, at least 3 chunks in flight when possible
    int64_t chunk = std::max<int64_t>(1, (int64_t)((host_chunk_mib() << 20) / mat_bytes));
    chunk = std::min<int64_t>(chunk, std::max<int64_t>(1, (batch + HostPipe::kSlots - 1) / HostPipe::kSlots));
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= lub::kMaxDevices) return fail(LUB_ERR_CUDA, "device index out of range");
    HostPipe& P = g_pipes[dev];
    std::lock_guard<std::mutex> lk(P.mu);
    const size_t need = (size_t)chunk * mat_bytes, pneed = host_piv ? (size_t)chunk * n * 4 : 0;
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        if (!P.st[i]) CU(cudaStreamCreateWithFlags(&P.st[i], cudaStreamNonBlocking));
        if (P.cap < need) { if (P.buf[i]) cudaFree(P.buf[i]); P.buf[i] = nullptr; CU(cudaMalloc(&P.buf[i], need)); }
        if (P.pcap < pneed) { if (P.pbuf[i]) cudaFree(P.pbuf[i]); P.pbuf[i] = nullptr; CU(cudaMalloc((void**)&P.pbuf[i], pneed)); }
    }
    P.cap = std::max(P.cap, need);
    P.pcap = std::max(P.pcap, pneed);
    char* h = static_cast<char*>(host_ptr);
    int slot = 0, rc = LUB_OK;
    for (int64_t b0 = 0; b0 < batch && rc == LUB_OK; b0 += chunk, slot = (slot + 1) % HostPipe::kSlots) {
        const int64_t nb = std::min<int64_t>(chunk, batch - b0);
        cudaStream_t s = P.st[slot];
        cudaError_t e = cudaMemcpyAsync(P.buf[slot], h + (size_t)b0 * mat_bytes, (size_t)nb * mat_bytes, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(H2D)"); break; }
        rc = launch_on(P.buf[slot], host_piv ? P.pbuf[slot] : nullptr, n, nb, mode, dtype, s, nullptr, 0);
        if (rc != LUB_OK) break;
        e = cudaMemcpyAsync(h + (size_t)b0 * mat_bytes, P.buf[slot], (size_t)nb * mat_bytes, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && host_piv) e = cudaMemcpyAsync(host_piv + b0 * n, P.pbuf[slot], (size_t)nb * n * 4, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync(D2H)");
    }
    // also on an error: copies already queued still read / write the caller's buffer, drain them first
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(P.st[i]);
        if (e != cudaSuccess && rc == LUB_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    return rc;
}

// CPUs next to a GPU: /sys/bus/pci/devices/<bus id>/local_cpulist ("0-31,64-95").  Empty set when unknown.
bool device_local_cpus(int dev, cpu_set_t* set) {
    CPU_ZERO(set);
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return false;
    for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist");
    std::string list;
    if (!f || !std::getline(f, list)) return false;
    int count = 0;
    size_t i = 0;
    while (i < list.size()) {
        char* end = nullptr;
        const long a = std::strtol(list.c_str() + i, &end, 10);
        if (end == list.c_str() + i) break;
        long b = a;
        i = (size_t)(end - list.c_str());
        if (i < list.size() && list[i] == '-') {
            b = std::strtol(list.c_str() + i + 1, &end, 10);
            i = (size_t)(end - list.c_str());
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++count; }
        if (i < list.size() && list[i] == ',') ++i;
    }
    return count > 0;
}

}  // namespace

extern "C" {

const char* lu_batched_version(void) { return "lubatched 0.1 (sm_100a)"; }
const char* lu_batched_last_error(void) { return g_err.c_str(); }

int lu_batched_set_stream(void* stream) {
    g_stream = static_cast<cudaStream_t>(stream);
    return LUB_OK;
}

int lu_batched_set_threads(int numthreads) {
    if (numthreads != 0 && (numthreads < 32 || numthreads > lub::kMaxThreads || numthreads % 32))
        return fail(LUB_ERR_BAD_ARG, "numthreads must be 0 or a multiple of 32 in [32, 256]");
    g_threads = numthreads;
    return LUB_OK;
}

int lu_batched_set_option(int option, int value) {
    if (option == LUB_OPT_STAGING) {
        if (value != 0 && value != 1) return fail(LUB_ERR_BAD_ARG, "LUB_OPT_STAGING: 0 (library choice) or 1 (LSU staging)");
        g_opt_flags = value ? (g_opt_flags | lub::kLaunchNoTma) : (g_opt_flags & ~lub::kLaunchNoTma);
    } else if (option == LUB_OPT_FP64_TENSOR) {
        if (value < 0 || value > 2) return fail(LUB_ERR_BAD_ARG, "LUB_OPT_FP64_TENSOR: 0 (library choice), 1 (never) or 2 (always)");
        g_opt_flags &= ~(lub::kLaunchNoDmma | lub::kLaunchForceDmma);
        if (value == 1) g_opt_flags |= lub::kLaunchNoDmma;
        if (value == 2) g_opt_flags |= lub::kLaunchForceDmma;
    } else {
        return fail(LUB_ERR_BAD_ARG, "unknown option");
    }
    return LUB_OK;
}

int lu_batched_get_option(int option) {
    if (option == LUB_OPT_STAGING) return (g_opt_flags & lub::kLaunchNoTma) ? 1 : 0;
    if (option == LUB_OPT_FP64_TENSOR) return (g_opt_flags & lub::kLaunchNoDmma) ? 1 : ((g_opt_flags & lub::kLaunchForceDmma) ? 2 : 0);
    return fail(LUB_ERR_BAD_ARG, "unknown option");
}

int lu_batched_get_threads(int n, int dtype) {
    if (g_threads) return g_threads;
    // the library's own default for this size and type (pivot_mode parallel, the reference's headline variant);
    // needs a device, like lu_batched_geometry
    lub::LaunchInfo info{};
    if (launch_on(nullptr, nullptr, n, 1, LUB_PIVOT_PARALLEL, dtype, nullptr, &info, lub::kLaunchDryRun) != LUB_OK) return -1;
    return info.threads_per_block;
}

int lu_batched_enable_timing(int on) {
    g_timing = on != 0;
    g_ev_dev = -1;
    return LUB_OK;
}

float lu_batched_last_kernel_ms(void) {
    if (g_ev_dev < 0) return -1.f;
    if (cudaEventSynchronize(g_ev[g_ev_dev][1]) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, g_ev[g_ev_dev][0], g_ev[g_ev_dev][1]) != cudaSuccess) return -1.f;
    return ms;
}

int lu_batched_inplace_stream(void* ptr, int32_t* piv, int n, int64_t batch, int pivot_mode, int dtype, void* stream) {
    return launch_on(ptr, piv, n, batch, pivot_mode, dtype, static_cast<cudaStream_t>(stream), nullptr, 0);
}

int lu_batched_inplace(void* ptr, int32_t* piv, int n, int64_t batch, int pivot_mode, int dtype) {
    return launch_on(ptr, piv, n, batch, pivot_mode, dtype, g_stream, nullptr, 0);
}

int lu_batched_inplace_ex(void* ptr, int32_t* piv, int32_t* info, int n, int64_t batch, int pivot_mode, int dtype, int layout, void* stream) {
    if (layout == LUB_LAYOUT_BATCH_INTERLEAVED) return launch_interleaved(ptr, piv, info, n, batch, pivot_mode, dtype, static_cast<cudaStream_t>(stream));
    if (layout != LUB_LAYOUT_MATRIX_MAJOR) return fail(LUB_ERR_BAD_ARG, "layout must be 0 (matrix-major) or 1 (batch-interleaved)");
    return launch_on(ptr, piv, n, batch, pivot_mode, dtype, static_cast<cudaStream_t>(stream), nullptr, 0, info);
}

int lu_batched_factor_inplace_ex(void* ptr, int32_t* piv, int32_t* info, int n, int64_t batch, int pivot_mode, int dtype, void* stream) {
    return launch_on(ptr, piv, n, batch, pivot_mode, dtype, static_cast<cudaStream_t>(stream), nullptr, lub::kLaunchLuOnly, info);
}

int lu_batched_ipiv_to_perm(const int32_t* ipiv, int32_t* perm, int n, int64_t batch) {
    if (n < 1 || batch < 0 || ((!ipiv || !perm) && batch)) return fail(LUB_ERR_BAD_ARG, "bad arguments");
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t b = 0; b < batch; ++b) {
        int32_t* p = perm + b * n;
        const int32_t* ip = ipiv + b * n;
        for (int i = 0; i < n; ++i) p[i] = i;
        for (int k = 0; k < n; ++k) {
            const int q = ip[k] - 1;
            if (q < k || q >= n) { bad = 1; continue; }
            const int32_t t = p[k]; p[k] = p[q]; p[q] = t;
        }
    }
    return bad ? fail(LUB_ERR_BAD_ARG, "ipiv holds an entry outside [k + 1, n]") : LUB_OK;
}

int lu_batched_factor_inplace_stream(void* ptr, int32_t* piv, int n, int64_t batch, int pivot_mode, int dtype, void* stream) {
    return launch_on(ptr, piv, n, batch, pivot_mode, dtype, static_cast<cudaStream_t>(stream), nullptr, lub::kLaunchLuOnly);
}

int lu_batched_factor_inplace(void* ptr, int32_t* piv, int n, int64_t batch, int pivot_mode, int dtype) {
    return launch_on(ptr, piv, n, batch, pivot_mode, dtype, g_stream, nullptr, lub::kLaunchLuOnly);
}

int lu_batched_verify_lu(const void* A, const void* LU, const int32_t* piv, int n, int64_t batch, int dtype, double thr,
                         int64_t* n_correct, int64_t* n_incorrect, double* max_abs_dev) {
    if (n < 1 || n > 1024) return fail(LUB_ERR_BAD_N, "n out of range");
    if (batch < 0 || (!A && batch) || (!LU && batch)) return fail(LUB_ERR_BAD_ARG, "bad buffers");
    if (dtype == LUB_DTYPE_F32) verify_lu_host(static_cast<const float*>(A), static_cast<const float*>(LU), piv, n, batch, thr, n_correct, n_incorrect, max_abs_dev);
    else if (dtype == LUB_DTYPE_F64) verify_lu_host(static_cast<const double*>(A), static_cast<const double*>(LU), piv, n, batch, thr, n_correct, n_incorrect, max_abs_dev);
    else return fail(LUB_ERR_BAD_DTYPE, "dtype must be 0 or 1");
    return LUB_OK;
}

int lu_batched_geometry(int n, int64_t batch, int pivot_mode, int dtype, int* threads_per_block,
                        int* threads_per_matrix, int* matrices_per_block, int64_t* num_blocks, int* dyn_smem_bytes) {
    lub::LaunchInfo info{};
    int rc = launch_on(nullptr, nullptr, n, batch, pivot_mode, dtype, nullptr, &info, 1);
    if (rc != LUB_OK) return rc;
    if (threads_per_block) *threads_per_block = info.threads_per_block;
    if (threads_per_matrix) *threads_per_matrix = info.threads_per_matrix;
    if (matrices_per_block) *matrices_per_block = info.matrices_per_block;
    if (num_blocks) *num_blocks = info.num_blocks;
    if (dyn_smem_bytes) *dyn_smem_bytes = info.dyn_smem_bytes;
    return LUB_OK;
}

const char* lu_batched_kernel_name(int n, int pivot_mode, int dtype) {
    lub::LaunchInfo info{};
    if (launch_on(nullptr, nullptr, n, 1, pivot_mode, dtype, nullptr, &info, 1) != LUB_OK) return nullptr;
    return info.kernel;
}

int lu_batched_inplace_host(void* host_ptr, int32_t* host_piv, int n, int64_t batch, int pivot_mode, int dtype) {
    int rc = check_args(n, batch, pivot_mode, dtype);
    if (rc != LUB_OK) return rc;
    if (batch == 0) return LUB_OK;
    if (!host_ptr) return fail(LUB_ERR_BAD_ARG, "host_ptr is NULL");
    return host_pipeline(host_ptr, host_piv, n, batch, pivot_mode, dtype);
}

int lu_batched_bind_thread_near_device(int device) {
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(LUB_ERR_BAD_ARG, "no such device");
    cpu_set_t set;
    if (!device_local_cpus(device, &set)) return fail(LUB_ERR_IO, "the PCI topology of the device is not exposed in /sys");
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return fail(LUB_ERR_IO, "sched_setaffinity failed");
    return LUB_OK;
}

int lu_batched_inplace_host_multi(void* host_ptr, int32_t* host_piv, int n, int64_t batch, int pivot_mode, int dtype,
                                  int n_devices, int flags) {
    int rc = check_args(n, batch, pivot_mode, dtype);
    if (rc != LUB_OK) return rc;
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (n_devices <= 0) n_devices = count;
    if (n_devices > count || n_devices > lub::kMaxDevices) return fail(LUB_ERR_BAD_ARG, "n_devices exceeds the devices visible to this process");
    if (batch == 0) return LUB_OK;
    if (!host_ptr) return fail(LUB_ERR_BAD_ARG, "host_ptr is NULL");
    const size_t mat_bytes = (size_t)n * n * esize(dtype);
    bool reg_a = false, reg_p = false;
    if (flags & LUB_HOST_REGISTER) {  // page-lock the caller's pageable buffers in place for the duration of the call
        reg_a = cudaHostRegister(host_ptr, (size_t)batch * mat_bytes, cudaHostRegisterPortable) == cudaSuccess;
        if (host_piv) reg_p = cudaHostRegister(host_piv, (size_t)batch * n * 4, cudaHostRegisterPortable) == cudaSuccess;
        cudaGetLastError();  // "already registered" / pinned by the caller is fine
    }
    int dev0 = 0;
    cudaGetDevice(&dev0);
    const int knob = g_threads;
    const int64_t per = (batch + n_devices - 1) / n_devices;   // contiguous shards, SURVEY.md 8(e)
    std::vector<int> rcs(n_devices, LUB_OK);
    std::vector<std::string> errs(n_devices);
    std::vector<std::thread> workers;
    for (int d = 0; d < n_devices; ++d) {
        workers.emplace_back([&, d]() {
            const int64_t lo = std::min<int64_t>(batch, d * per), hi = std::min<int64_t>(batch, lo + per);
            if (hi <= lo) return;
            if (cudaSetDevice(d) != cudaSuccess) { rcs[d] = LUB_ERR_CUDA; errs[d] = "cudaSetDevice failed"; return; }
            if (flags & LUB_HOST_BIND_THREADS) {
                cpu_set_t set;
                if (device_local_cpus(d, &set)) pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
            }
            g_threads = knob;
            rcs[d] = host_pipeline(static_cast<char*>(host_ptr) + (size_t)lo * mat_bytes, host_piv ? host_piv + lo * n : nullptr, n, hi - lo,
                                   pivot_mode, dtype);
            if (rcs[d] != LUB_OK) errs[d] = g_err;
        });
    }
    for (auto& w : workers) w.join();
    cudaSetDevice(dev0);
    if (reg_a) cudaHostUnregister(host_ptr);
    if (reg_p) cudaHostUnregister(host_piv);
    for (int d = 0; d < n_devices; ++d)
        if (rcs[d] != LUB_OK) return fail(rcs[d], "device " + std::to_string(d) + ": " + errs[d]);
    return LUB_OK;
}

int lu_batched_verify_inv(const void* A, const void* Ainv, int n, int64_t batch, int dtype, double thr,
                          int64_t* n_correct, int64_t* n_incorrect, double* max_abs_dev) {
    if (n < 1 || n > 1024) return fail(LUB_ERR_BAD_N, "n out of range");
    if (batch < 0 || (!A && batch) || (!Ainv && batch)) return fail(LUB_ERR_BAD_ARG, "bad buffers");
    if (dtype == LUB_DTYPE_F32) verify_host(static_cast<const float*>(A), static_cast<const float*>(Ainv), n, batch, thr, n_correct, n_incorrect, max_abs_dev);
    else if (dtype == LUB_DTYPE_F64) verify_host(static_cast<const double*>(A), static_cast<const double*>(Ainv), n, batch, thr, n_correct, n_incorrect, max_abs_dev);
    else return fail(LUB_ERR_BAD_DTYPE, "dtype must be 0 or 1");
    return LUB_OK;
}

namespace {
// device-side accumulator of lu_batched_verify_inv_device: one small allocation per (host thread, device), kept
struct VerifyAcc { unsigned long long good; unsigned int worst; unsigned int nan; };
thread_local VerifyAcc* g_vacc[lub::kMaxDevices] = {};

int verify_inv_device_on(const void* dA, const void* dAinv, int n, int64_t batch, int dtype, double thr, int64_t* n_correct,
                         int64_t* n_incorrect, double* max_abs_dev, cudaStream_t stream) {
    if (n < 1 || n > 1024) return fail(LUB_ERR_BAD_N, "n out of range");
    if (dtype != LUB_DTYPE_F32 && dtype != LUB_DTYPE_F64) return fail(LUB_ERR_BAD_DTYPE, "dtype must be 0 or 1");
    if (batch < 0 || (!dA && batch) || (!dAinv && batch)) return fail(LUB_ERR_BAD_ARG, "bad buffers");
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= lub::kMaxDevices) return fail(LUB_ERR_CUDA, "device index out of range");
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!g_vacc[dev]) CU(cudaMalloc((void**)&g_vacc[dev], sizeof(VerifyAcc)));
    VerifyAcc* d = g_vacc[dev];
    CU(cudaMemsetAsync(d, 0, sizeof(VerifyAcc), stream));
    if (batch > 0) {
        const int threads = 256;
        const long long want = (batch * 32 + threads - 1) / threads;
        const unsigned blocks = (unsigned)std::min<long long>(want, (long long)sms * 16);
        if (dtype == LUB_DTYPE_F32)
            verify_kernel<float><<<blocks, threads, 0, stream>>>(static_cast<const float*>(dA), static_cast<const float*>(dAinv), n, batch, (float)thr, &d->good, &d->worst, &d->nan);
        else
            verify_kernel<double><<<blocks, threads, 0, stream>>>(static_cast<const double*>(dA), static_cast<const double*>(dAinv), n, batch, thr, &d->good, &d->worst, &d->nan);
        CU(cudaGetLastError());
    }
    VerifyAcc h{};
    CU(cudaMemcpyAsync(&h, d, sizeof(VerifyAcc), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    if (n_correct) *n_correct = (int64_t)h.good;
    if (n_incorrect) *n_incorrect = batch - (int64_t)h.good;
    if (max_abs_dev) { float w; memcpy(&w, &h.worst, 4); *max_abs_dev = h.nan ? NAN : (double)w; }
    return LUB_OK;
}
}  // namespace

int lu_batched_verify_inv_device(const void* dA, const void* dAinv, int n, int64_t batch, int dtype, double thr,
                                 int64_t* n_correct, int64_t* n_incorrect, double* max_abs_dev) {
    return verify_inv_device_on(dA, dAinv, n, batch, dtype, thr, n_correct, n_incorrect, max_abs_dev, g_stream);
}

int lu_batched_verify_inv_device_stream(const void* dA, const void* dAinv, int n, int64_t batch, int dtype, double thr,
                                        int64_t* n_correct, int64_t* n_incorrect, double* max_abs_dev, void* stream) {
    return verify_inv_device_on(dA, dAinv, n, batch, dtype, thr, n_correct, n_incorrect, max_abs_dev, static_cast<cudaStream_t>(stream));
}

int lu_batched_read_tokens(const char* path, void* out, int64_t count, int dtype) {
    if (!path || (!out && count) || count < 0) return fail(LUB_ERR_BAD_ARG, "bad arguments");
    if (dtype != LUB_DTYPE_F32 && dtype != LUB_DTYPE_F64) return fail(LUB_ERR_BAD_DTYPE, "dtype must be 0 or 1");
    std::ifstream f(path);
    if (!f) return fail(LUB_ERR_IO, std::string("cannot open ") + path);
    // same extraction the reference uses: `file >> templateMatrix[i]` (templated/luBatchedInplace.cu:33)
    for (int64_t i = 0; i < count; ++i) {
        if (dtype == LUB_DTYPE_F32) f >> static_cast<float*>(out)[i];
        else f >> static_cast<double*>(out)[i];
        if (!f) return fail(LUB_ERR_IO, std::string(path) + ": fewer tokens than requested");
    }
    return LUB_OK;
}

// printMatrices / writeToFile (templated/verify.hpp:12-48): the FIRST matrix of the buffer only (both
// reference loops `break` after k = 0), default ostream formatting, "value<space>" per entry, one row
// per line; printMatrices adds an empty line after the matrix, writeToFile does not.
int lu_batched_write_matrix(const void* A, const char* path, int n, int dtype) {
    if (!A || n < 1) return fail(LUB_ERR_BAD_ARG, "bad arguments");
    if (dtype != LUB_DTYPE_F32 && dtype != LUB_DTYPE_F64) return fail(LUB_ERR_BAD_DTYPE, "dtype must be 0 or 1");
    std::ofstream file;
    if (path) {
        file.open(path);
        if (!file) return fail(LUB_ERR_IO, std::string("cannot open ") + path);
    }
    std::ostream& os = path ? static_cast<std::ostream&>(file) : std::cout;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            if (dtype == LUB_DTYPE_F32) os << static_cast<const float*>(A)[i * n + j] << ' ';
            else os << static_cast<const double*>(A)[i * n + j] << ' ';
        }
        os << '\n';
    }
    if (!path) { os << '\n'; os.flush(); }
    return LUB_OK;
}

int lu_batched_replicate(const void* tmpl, void* dst, int n, int64_t batch, int dtype) {
    if (n < 1 || batch < 0 || !tmpl || (!dst && batch)) return fail(LUB_ERR_BAD_ARG, "bad arguments");
    if (dtype != LUB_DTYPE_F32 && dtype != LUB_DTYPE_F64) return fail(LUB_ERR_BAD_DTYPE, "dtype must be 0 or 1");
    const size_t mb = (size_t)n * n * esize(dtype);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < batch; ++b) memcpy(static_cast<char*>(dst) + (size_t)b * mb, tmpl, mb);
    return LUB_OK;
}

int lu_batched_device_info(int* sm_count, int* max_smem_optin, int* clock_khz, int* cc_major, int* cc_minor) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) { CU(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (max_smem_optin) { CU(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)); *max_smem_optin = v; }
    if (clock_khz) { CU(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev)); *clock_khz = v; }
    if (cc_major) { CU(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { CU(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return LUB_OK;
}

}  // extern "C"
