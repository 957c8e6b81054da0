// Dev-only tuning harness (not part of the product library): explicit instantiations of the
// fast kernel for a handful of lane layouts, launched by index so that scripts/tune/run.py can
// time them side by side on the GPU box.  Build: make -C scripts/tune
#include <cstdio>
#include "lub_v5.cuh"
#include "lub_tma2.cuh"
#include "../../matrixinversion_b200/csrc/lub_dmma.cuh"
#include "lub_v6.cuh"
#include "../../matrixinversion_b200/csrc/lub_bulk.cuh"

using namespace lub;

struct Variant {
    const char* name;
    int mpw, warp_bytes, header;   // warp_bytes < 0: persistent kernel with -warp_bytes threads, header = total smem
    void (*set_attr)(int smem);
    void (*launch)(void*, int*, long long, unsigned, int, int, cudaStream_t);
    int (*occ)(int threads, int smem);
    int (*smem_of)(int warps) = nullptr;
};

template <typename T, int N, int GR, int GC, int MODE, int MINB, int DBG = 0, int MAXT = 256>
struct V3 {
    using L = V3Layout<T, N, GR, GC, MODE>;
    static void set_attr(int smem) {
        cudaFuncSetAttribute(lub_v3_kernel<T, N, GR, GC, MODE, MINB, (DBG & 8) != 0, (DBG & 16) != 0, (DBG & 32) != 0, (DBG & 64) != 0, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    static void launch(void* A, int* piv, long long batch, unsigned blocks, int threads, int smem, cudaStream_t s) {
        lub_v3_kernel<T, N, GR, GC, MODE, MINB, (DBG & 8) != 0, (DBG & 16) != 0, (DBG & 32) != 0, (DBG & 64) != 0, MAXT><<<blocks, threads, smem, s>>>((T*)A, piv, batch);
    }
    static int occ(int threads, int smem) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, lub_v3_kernel<T, N, GR, GC, MODE, MINB, (DBG & 8) != 0, (DBG & 16) != 0, (DBG & 32) != 0, (DBG & 64) != 0, MAXT>, threads, smem);
        return o;
    }
    static Variant make(const char* name) { return Variant{name, L::MPW, (DBG & 16) ? L::WARP_BYTES_PF : ((DBG & 32) ? L::WARP_BYTES_PFD : L::WARP_BYTES), L::HEADER_BYTES, set_attr, launch, occ}; }
};

template <typename T, int N, int GR, int GC, int MODE, int MINB, int DBG = 0>
struct V {
    using L = V4Layout<T, N, GR, GC, MODE>;
    static void set_attr(int smem) {
        cudaFuncSetAttribute(lub_v4_kernel<T, N, GR, GC, MODE, MINB, (DBG & 8) != 0, (DBG & 32) != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    static void launch(void* A, int* piv, long long batch, unsigned blocks, int threads, int smem, cudaStream_t s) {
        lub_v4_kernel<T, N, GR, GC, MODE, MINB, (DBG & 8) != 0, (DBG & 32) != 0><<<blocks, threads, smem, s>>>((T*)A, piv, batch);
    }
    static int occ(int threads, int smem) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, lub_v4_kernel<T, N, GR, GC, MODE, MINB, (DBG & 8) != 0, (DBG & 32) != 0>, threads, smem);
        return o;
    }
    static Variant make(const char* name) { return Variant{name, L::MPW, (DBG & 32) ? L::WARP_BYTES_PFD : L::WARP_BYTES, L::HEADER_BYTES, set_attr, launch, occ}; }
};

template <typename T, int N, int GR, int GC, int MODE, int NPW, int NCW, int NB, int PRE>
struct V5 {
    using L = V5Layout<T, N, GR, GC, MODE, NB>;
    static void set_attr(int smem) {
        cudaFuncSetAttribute(lub_v5_kernel<T, N, GR, GC, MODE, NPW, NCW, NB, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    static void launch(void* A, int* piv, long long batch, unsigned blocks, int threads, int smem, cudaStream_t s) {
        lub_v5_kernel<T, N, GR, GC, MODE, NPW, NCW, NB, PRE><<<blocks, threads, smem, s>>>((T*)A, piv, batch);
    }
    static int occ(int threads, int smem) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, lub_v5_kernel<T, N, GR, GC, MODE, NPW, NCW, NB, PRE>, threads, smem);
        return o;
    }
    static Variant make(const char* name) { return Variant{name, L::MPW, -(NPW + NCW) * 32, L::SMEM_BYTES, set_attr, launch, occ}; }
};

// TMA-staged kernel: warp_bytes == -1000000 marks it; launch builds the tensor map per call
// BS: bit 0 = BSYNC, 1 = PF, 2 = OUTIMG, 3 = ST256, 4 = lean step, 5 = double-buffered image, 6 = pivot search of the next tile fused into the elimination; MAXT = compiled block size
template <typename T, int N, int GR, int GC, int MODE, int MINB, int BS, int MAXT = 256>
struct VT {
    using L = TmaLayout<T, N, GR, GC, MODE>;
    static constexpr bool PF = (BS & 2) != 0;
    static constexpr int OPT = ((BS & 16) ? kTmaLean : 0) | ((BS & 32) ? kTmaDB : 0) | ((BS & 64) ? kTmaFused : 0);
    static constexpr auto kern() { return lub_tma_kernel<T, N, GR, GC, MODE, MINB, (BS & 1) != 0, PF, (BS & 4) != 0, (BS & 8) != 0, OPT, MAXT>; }
    static void set_attr(int smem) { cudaFuncSetAttribute(kern(), cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }
    static void launch(void* A, int* piv, long long batch, unsigned blocks, int threads, int smem, cudaStream_t s) {
        CUtensorMap map;
        if (make_batch_tmap<T>(&map, A, N, batch, L::MPW) != cudaSuccess) { printf("tensor map failed\n"); return; }
        kern()<<<blocks, threads, smem, s>>>(map, (T*)A, piv, batch, nullptr);
    }
    static int occ(int threads, int smem) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern(), threads, smem);
        return o;
    }
    static int smem_of(int warps) { return L::smem_bytes(warps, (BS & 32) ? 2 : 1); }
    static Variant make(const char* name) { return Variant{name, L::MPW, -1000000, MAXT, set_attr, launch, occ, &smem_of}; }
};
#define VART(T, N, GR, GC, MODE, MINB, BS) VT<T, N, GR, GC, MODE, MINB, BS>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " minb" #MINB " bs" #BS " tma")
#define VARTM(T, N, GR, GC, MODE, MINB, BS, MAXT) VT<T, N, GR, GC, MODE, MINB, BS, MAXT>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " minb" #MINB " bs" #BS " maxt" #MAXT " tma")

// round-2 fused kernel (lub_tma2.cuh): double-buffered image, next tile's pivot search inside the elimination
template <int N, int GR, int GC, int MODE, int MAXT, int K1>
struct VT2 {
    using L = Tma2Layout<N, GR, GC, MODE>;
    static constexpr auto kern() { return lub_tma2_kernel<N, GR, GC, MODE, MAXT, K1>; }
    static void set_attr(int smem) { cudaFuncSetAttribute(kern(), cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }
    static void launch(void* A, int* piv, long long batch, unsigned blocks, int threads, int smem, cudaStream_t s) {
        CUtensorMap map;
        if (make_batch_tmap<float>(&map, A, N, batch, L::MPW) != cudaSuccess) { printf("tensor map failed\n"); return; }
        kern()<<<blocks, threads, smem, s>>>(map, (float*)A, piv, batch);
    }
    static int occ(int threads, int smem) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern(), threads, smem);
        return o;
    }
    static int smem_of(int warps) { return L::smem_bytes(warps); }
    static Variant make(const char* name) { return Variant{name, L::MPW, -1000000, MAXT, set_attr, launch, occ, &smem_of}; }
};
#define VART2(N, GR, GC, MODE, MAXT, K1) VT2<N, GR, GC, MODE, MAXT, K1>::make("float N=" #N " " #GR "x" #GC " mode" #MODE " maxt" #MAXT " k1_" #K1 " tma2")

// fp64 N = 32 DMMA kernel (lub_dmma.cuh)
template <int MODE, int MINB, int BS>
struct VDM {
    using L = TmaLayout<double, 32, 8, 4, MODE>;
    static constexpr auto kern() { return lub_dmma_kernel<MODE, MINB, (BS & 1) != 0>; }
    static void set_attr(int smem) { cudaFuncSetAttribute(kern(), cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }
    static void launch(void* A, int* piv, long long batch, unsigned blocks, int threads, int smem, cudaStream_t s) {
        CUtensorMap map;
        if (make_batch_tmap<double>(&map, A, 32, batch, L::MPW) != cudaSuccess) { printf("tensor map failed\n"); return; }
        kern()<<<blocks, threads, smem, s>>>(map, (double*)A, piv, batch);
    }
    static int occ(int threads, int smem) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern(), threads, smem);
        return o;
    }
    static int smem_of(int warps) { return dmma_smem_bytes(warps, L::PERM_BYTES); }
    static Variant make(const char* name) { return Variant{name, L::MPW, -1000000, 256, set_attr, launch, occ, &smem_of}; }
};
#define VARDM(MODE, MINB, BS) VDM<MODE, MINB, BS>::make("double N=32 8x4 mode" #MODE " minb" #MINB " bs" #BS " dmma")

// warp-specialised TMA kernel: one persistent block per SM; warp_bytes == -2000000 marks it
template <typename T, int N, int GR, int GC, int MODE, int NCW, int NPW, int NB, int CREG, int PREG, int NCG, int LA, int DBG>
struct VT6 {
    using L = V6Layout<T, N, GR, GC, MODE, NB>;
    static constexpr auto kern() { return lub_v6_kernel<T, N, GR, GC, MODE, NCW, NPW, NB, CREG, PREG, NCG, LA, DBG>; }
    static void set_attr(int smem) { cudaFuncSetAttribute(kern(), cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }
    static void launch(void* A, int* piv, long long batch, unsigned blocks, int threads, int smem, cudaStream_t s) {
        CUtensorMap map;
        if (make_batch_tmap<T>(&map, A, N, batch, L::MPW) != cudaSuccess) { printf("tensor map failed\n"); return; }
        kern()<<<blocks, threads, smem, s>>>(map, (T*)A, piv, batch);
    }
    static int occ(int threads, int smem) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern(), threads, smem);
        return o;
    }
    static Variant make(const char* name) { return Variant{name, L::MPW, -(NCW + NPW) * 32, L::SMEM_BYTES, set_attr, launch, occ}; }
};
#define VART6(T, N, GR, GC, MODE, NCW, NPW, NB, CREG, PREG, NCG, LA, DBG) VT6<T, N, GR, GC, MODE, NCW, NPW, NB, CREG, PREG, NCG, LA, DBG>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " c" #NCW " p" #NPW " nb" #NB " creg" #CREG " preg" #PREG " ncg" #NCG " la" #LA " dbg" #DBG " v6")

// bulk-copy staged kernel (lub_bulk.cuh); OPT: 1 = lean step, 2 = old position-wise search, 4 = one image per warp, 8 = BSYNC, 16 = two-reduction search
template <typename T, int N, int GR, int GC, int MODE, int MINB, int OPT, int MAXT>
struct VB {
    using L = BulkLayout<T, N, GR, GC, MODE>;
    static constexpr auto kern() { return lub_bulk_kernel<T, N, GR, GC, MODE, MINB, (OPT & 8) != 0, (OPT & ~8), MAXT>; }
    static void set_attr(int smem) { cudaFuncSetAttribute(kern(), cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }
    static void launch(void* A, int* piv, long long batch, unsigned blocks, int threads, int smem, cudaStream_t s) {
        kern()<<<blocks, threads, smem, s>>>((T*)A, piv, batch, nullptr);
    }
    static int occ(int threads, int smem) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern(), threads, smem);
        return o;
    }
    static int smem_of(int warps) { return L::smem_bytes(warps, (OPT & 4) ? 1 : 2); }
    static Variant make(const char* name) { return Variant{name, L::MPW, -1000000, MAXT, set_attr, launch, occ, &smem_of}; }
};
#define VARB(T, N, GR, GC, MODE, MINB, OPT, MAXT) VB<T, N, GR, GC, MODE, MINB, OPT, MAXT>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " minb" #MINB " opt" #OPT " maxt" #MAXT " bulk")

#define VAR5(T, N, GR, GC, MODE, NPW, NCW, NB, PRE) V5<T, N, GR, GC, MODE, NPW, NCW, NB, PRE>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " p" #NPW " c" #NCW " nb" #NB " opt" #PRE " v5")

#define VAR(T, N, GR, GC, MODE, MINB) V<T, N, GR, GC, MODE, MINB>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " minb" #MINB)
#define VARD(T, N, GR, GC, MODE, MINB, DBG) V<T, N, GR, GC, MODE, MINB, DBG>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " minb" #MINB " dbg" #DBG " v4")
#define VAR3(T, N, GR, GC, MODE, MINB, DBG) V3<T, N, GR, GC, MODE, MINB, DBG>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " minb" #MINB " dbg" #DBG " v3")
#define VAR3M(T, N, GR, GC, MODE, MINB, DBG, MAXT) V3<T, N, GR, GC, MODE, MINB, DBG, MAXT>::make(#T " N=" #N " " #GR "x" #GC " mode" #MODE " minb" #MINB " dbg" #DBG " maxt" #MAXT " v3")

static Variant variants[] = {
#include "variants.inc"
};

extern "C" int tune_count() { return (int)(sizeof(variants) / sizeof(variants[0])); }
extern "C" const char* tune_name(int i) { return variants[i].name; }
extern "C" int tune_launch(int i, void* A, int* piv, long long batch, int threads, void* stream, int* occ_out, int* blocks_out) {
    Variant& v = variants[i];
    if (v.smem_of) {
        if (threads > v.header) return -3;  // compiled for at most v.header threads per block
        const int warps = threads / 32;
        const int smem = v.smem_of(warps);
        v.set_attr(smem);
        const int occ = v.occ(threads, smem);
        if (occ < 1) return -1;
        const long long ntiles = (batch + v.mpw - 1) / v.mpw;
        long long blocks = (ntiles + warps - 1) / warps;
        if (blocks > 148ll * occ) blocks = 148ll * occ;
        if (occ_out) *occ_out = occ;
        if (blocks_out) *blocks_out = (int)blocks;
        v.launch(A, piv, batch, (unsigned)blocks, threads, smem, (cudaStream_t)stream);
        return cudaGetLastError() == cudaSuccess ? 0 : -2;
    }
    if (v.warp_bytes < 0) {  // persistent producer/consumer kernel: one block per SM
        threads = -v.warp_bytes;
        const int smem = v.header;
        v.set_attr(smem);
        const int occ = v.occ(threads, smem);
        if (occ < 1) return -1;
        const long long ntiles = (batch + v.mpw - 1) / v.mpw;
        long long blocks = ntiles < 148 ? ntiles : 148;
        if (occ_out) *occ_out = occ;
        if (blocks_out) *blocks_out = (int)blocks;
        v.launch(A, piv, batch, (unsigned)blocks, threads, smem, (cudaStream_t)stream);
        return cudaGetLastError() == cudaSuccess ? 0 : -2;
    }
    const int warps = threads / 32;
    const int smem = v.header + warps * v.warp_bytes;
    v.set_attr(smem);
    const int occ = v.occ(threads, smem);
    if (occ < 1) return -1;
    const long long ntiles = (batch + v.mpw - 1) / v.mpw;
    long long blocks = (ntiles + warps - 1) / warps;
    if (blocks > 148ll * occ) blocks = 148ll * occ;
    if (occ_out) *occ_out = occ;
    if (blocks_out) *blocks_out = (int)blocks;
    v.launch(A, piv, batch, (unsigned)blocks, threads, smem, (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
