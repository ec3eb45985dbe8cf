// Shared device/host helpers for libmps_b200 (sm_100a only).
//
// Complex numbers are numpy-native interleaved complex128 (double2: x=re, y=im) everywhere, so
// host arrays cross the C ABI without any re-layout.  FP64 tensor-core work uses the warp-level
// mma.sync m8n8k4 f64 instruction (SASS DMMA.8x8x4); tcgen05 has no f64 kind.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmps_b200 is written for sm_100a only"
#endif

namespace mpsb {

typedef double2 cplx;

// ---- error plumbing (host) -------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define MPSB_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            mpsb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,             \
                            cudaGetErrorString(_e));                                       \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)
#define MPSB_REQUIRE(cond, ...)                                                            \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            mpsb::set_error(__VA_ARGS__);                                                  \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

// Launch counter: every kernel launched by the library bumps it (bench.py reports it as gpu_launches).
extern unsigned long long g_launch_count;
#define MPSB_COUNT_LAUNCH() (++mpsb::g_launch_count)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Raises cudaFuncAttributeMaxDynamicSharedMemorySize of `func` on the CURRENT device when a launch needs more than
// was granted so far (the attribute is per device, so the cache is keyed by (function, device)).  Returns 0 / 2.
int ensure_dynamic_smem(const void* func, size_t bytes);

// ---- device helpers --------------------------------------------------------------------------
#ifdef __CUDACC__

// D(8x8) += A(8x4, row) * B(4x8, col).  Fragment layout (lane = 4*g + t, g = lane/4, t = lane%4):
//   a = A[g][t],  b = B[t][g],  c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte cp.async (LDGSTS.128) with zero-fill when !pred.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool pred) {
    uint32_t d = smem_u32(smem_dst);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP) ----------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Bounded wait: returns false if the phase never completed (caller traps -> loud failure, no hang).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global bulk copy (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::); }
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}
// all committed bulk stores have finished READING their shared-memory source (the buffers may be refilled)
__device__ __forceinline__ void bulk_wait_read_all() {
    asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
// make generic-proxy smem writes visible to the async proxy (before bulk_s2g)
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {  // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// Scalar traits: the iDMRG kernels are written once for complex128 (cplx) and for real doubles - the reference's
// real path, MPO.tomatrix(dtype='double'), DMRG/MPO.py:68-76.
template <typename T> struct Sc;
template <> struct Sc<cplx> {
    static __host__ __device__ __forceinline__ cplx zero() { return make_double2(0.0, 0.0); }
    static __host__ __device__ __forceinline__ cplx from_re(double v) { return make_double2(v, 0.0); }
    static __host__ __device__ __forceinline__ double re(cplx a) { return a.x; }
    static __device__ __forceinline__ bool is_zero(cplx g) { return g.x == 0.0 && g.y == 0.0; }
    static __device__ __forceinline__ void mac(cplx& acc, cplx g, cplx x) {          // acc += g x
        acc.x = fma(g.x, x.x, fma(-g.y, x.y, acc.x));
        acc.y = fma(g.x, x.y, fma(g.y, x.x, acc.y));
    }
    static __device__ __forceinline__ void macc(cplx& acc, cplx a, cplx b) {         // acc += conj(a) b
        acc.x = fma(a.x, b.x, fma(a.y, b.y, acc.x));
        acc.y = fma(a.x, b.y, fma(-a.y, b.x, acc.y));
    }
    static __device__ __forceinline__ cplx add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
    static __device__ __forceinline__ cplx scale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
    static __device__ __forceinline__ cplx conj(cplx a) { return make_double2(a.x, -a.y); }
};
template <> struct Sc<double> {
    static __host__ __device__ __forceinline__ double zero() { return 0.0; }
    static __host__ __device__ __forceinline__ double from_re(double v) { return v; }
    static __host__ __device__ __forceinline__ double re(double a) { return a; }
    static __device__ __forceinline__ bool is_zero(double g) { return g == 0.0; }
    static __device__ __forceinline__ void mac(double& acc, double g, double x) { acc = fma(g, x, acc); }
    static __device__ __forceinline__ void macc(double& acc, double a, double b) { acc = fma(a, b, acc); }
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double scale(double a, double s) { return a * s; }
    static __device__ __forceinline__ double conj(double a) { return a; }
};
#endif  // __CUDACC__

}  // namespace mpsb
