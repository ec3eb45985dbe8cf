// Batched real (double) GEMM on FP64 tensor cores (DMMA m8n8k4): the real-arithmetic twin of zgemm.cu.
//
// The reference runs its whole iDMRG step in real arithmetic when the MPO is built with
// MPO.tomatrix(dtype='double') (DMRG/MPO.py:68-76; SURVEY section 6: 6-7 s instead of 9-12 s per step): L, W, R and
// the two-site tensor are then real and every contraction of DMRG/iDMRG.py:40,59,60 is a real product - a quarter of
// the flops and half of the bytes of the complex one.  This kernel serves those contractions.
//
// C[z] = op(A[z]) * op(B[z]) (+ C[z]), row-major doubles, op 0 / 3 = N, 1 / 2 = T.
// CTA tile BM x BN x 16 (128 x 128 with 16 warps, or 64 x 64 with 8 warps for launches that would not fill the GPU
// otherwise); operands arrive by TMA (cp.async.bulk.tensor, 4-D maps, 128-byte swizzle) through a ring of mbarriers,
// exactly as in zgemm_tma_kernel.  A row of a box is 16 doubles = 128 bytes = the swizzle span:
//   operand stored with k contiguous ("KC": A with op N, B with op T): ONE box of BM (BN) rows, 128 B per row;
//   operand stored with m/n contiguous ("MC": A with op T, B with op N): BM/16 boxes of 16 k-rows, 2 KiB each.
// The k index a lane feeds into its DMMA slot is permuted inside the 16-wide k tile (perm_k) so that the 8-byte
// fragment loads of every half-warp spread over all banks for both layouts.
// Operands a tensor map cannot describe (odd leading dimension or stride, misaligned base) take a plain SIMT kernel;
// those are the tiny odd-shaped products of the first growth steps.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "mpsb_internal.h"

namespace mpsb {

namespace {

typedef GemmArgsT<double> DArgs;

constexpr int DBK = 16;
constexpr int DSTAGES = 4;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
            "r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

// k handled by DMMA slot t in sub-step kk of a 16-wide k tile: a bijection of (kk, t) onto 0..15 chosen so that the
// 8-byte fragment loads of each HALF-warp (the unit in which 64-bit shared loads are served; lanes g = 0..3 x t = 0..3)
// touch all 32 banks exactly once in both tile layouts.  With k = b0 + 2 b1 + 4 b2 + 8 b3:
//   KC tiles (chunk = (k >> 1) ^ row, half = k & 1) need t -> (b0, b3) one-to-one,
//   MC tiles (chunk = column-chunk ^ (k & 7))       need t -> (b1, b2) one-to-one;
// b0 = t0 ^ kk0, b1 = t0, b2 = t1, b3 = t1 ^ kk1 satisfies both (ncu: shared-memory bank conflicts of the natural
// order k = 4 kk + t were two extra wavefronts per load).
__device__ __forceinline__ int perm_k(int kk, int t) {
    const int t0 = t & 1, t1 = t >> 1;
    return (t0 ^ (kk & 1)) | (t0 << 1) | (t1 << 2) | ((t1 ^ (kk >> 1)) << 3);
}
// byte offset of element (row r, k) in a KC tile (one box, rows of 128 B)
__device__ __forceinline__ int dkc_off(int r, int k) { return r * 128 + ((((k >> 1) ^ (r & 7))) << 4) + ((k & 1) << 3); }
// byte offset of element (k, c) in an MC tile (boxes of 16 doubles x 16 k-rows)
__device__ __forceinline__ int dmc_off(int k, int c) {
    return (c >> 4) * 2048 + k * 128 + (((((c & 15) >> 1)) ^ (k & 7)) << 4) + ((c & 1) << 3);
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB>
__global__ void __launch_bounds__(32 * WM * WN, (WM * WN <= 8) ? 2 : 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, DArgs p, int az1,
                 int az2, int bz1, int bz2) {
    constexpr int TI = BM / WM / 8, TJ = BN / WN / 8;
    constexpr int A_BYTES = BM * DBK * 8, B_BYTES = BN * DBK * 8, STAGE = A_BYTES + B_BYTES;
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[DSTAGES];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    const int z = blockIdx.z;
    const int z1 = z / p.batch2, z2 = z % p.batch2;
    double* C = p.C + (size_t)z1 * p.sC1 + (size_t)z2 * p.sC2;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkt = (p.K + DBK - 1) / DBK;

    if (tid == 0) {
        for (int s = 0; s < DSTAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int ca2 = az1 ? z1 : 0, ca3 = az2 ? z2 : 0, cb2 = bz1 ? z1 : 0, cb3 = bz2 ? z2 : 0;
    auto issue = [&](int kt) {
        const int s = kt % DSTAGES;
        unsigned char* as = tiles + (size_t)s * STAGE;
        unsigned char* bs = as + A_BYTES;
        const int k0 = kt * DBK;
        mbar_expect_tx(&full[s], STAGE);
        if (!TA) {
            tma_load_4d(as, &mapA, k0, m0, ca2, ca3, &full[s]);
        } else {
#pragma unroll
            for (int q = 0; q < BM / 16; ++q) tma_load_4d(as + q * 2048, &mapA, m0 + 16 * q, k0, ca2, ca3, &full[s]);
        }
        if (!TB) {
#pragma unroll
            for (int q = 0; q < BN / 16; ++q) tma_load_4d(bs + q * 2048, &mapB, n0 + 16 * q, k0, cb2, cb3, &full[s]);
        } else {
            tma_load_4d(bs, &mapB, k0, n0, cb2, cb3, &full[s]);
        }
    };
    if (tid == 0)
        for (int kt = 0; kt < DSTAGES - 1 && kt < nkt; ++kt) issue(kt);

    double acc[TI][TJ][2];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < nkt; ++kt) {
        const int s = kt % DSTAGES;
        if (tid == 0 && kt + DSTAGES - 1 < nkt) issue(kt + DSTAGES - 1);
        if (!mbar_wait(&full[s], (kt / DSTAGES) & 1)) __trap();
        const unsigned char* as = tiles + (size_t)s * STAGE;
        const unsigned char* bs = as + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < DBK / 4; ++kk) {
            const int k = perm_k(kk, t);
            double a[TI], b[TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i) {
                const int m = wm * (BM / WM) + i * 8 + g;
                a[i] = *reinterpret_cast<const double*>(as + (TA ? dmc_off(k, m) : dkc_off(m, k)));
            }
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                const int n = wn * (BN / WN) + j * 8 + g;
                b[j] = *reinterpret_cast<const double*>(bs + (TB ? dkc_off(n, k) : dmc_off(k, n)));
            }
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();            // stage s may be refilled from the next iteration on
    }

    const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < TI; ++i) {
        const int row = m0 + wm * (BM / WM) + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
            const int col = n0 + wn * (BN / WN) + j * 8 + 2 * t;
            double* dst = C + (size_t)row * p.ldc + col;
            if (vec_ok && col + 1 < p.N) {
                double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
                if (p.accumulate) { const double2 o = *reinterpret_cast<const double2*>(dst); v.x += o.x; v.y += o.y; }
                *reinterpret_cast<double2*>(dst) = v;
            } else {
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (col + e < p.N) dst[e] = acc[i][j][e] + (p.accumulate ? dst[e] : 0.0);
            }
        }
    }
}

// Plain SIMT product for operands no tensor map describes: one thread per element of C.
__global__ void __launch_bounds__(256) dgemm_simt_kernel(DArgs p, int ta, int tb) {
    const int z = blockIdx.z;
    const int z1 = z / p.batch2, z2 = z % p.batch2;
    const double* A = p.A + (size_t)z1 * p.sA1 + (size_t)z2 * p.sA2;
    const double* B = p.B + (size_t)z1 * p.sB1 + (size_t)z2 * p.sB2;
    double* C = p.C + (size_t)z1 * p.sC1 + (size_t)z2 * p.sC2;
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), row = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (row >= p.M || col >= p.N) return;
    double acc = 0.0;
    for (int k = 0; k < p.K; ++k) {
        const double a = ta ? A[(size_t)k * p.lda + row] : A[(size_t)row * p.lda + k];
        const double b = tb ? B[(size_t)col * p.ldb + k] : B[(size_t)k * p.ldb + col];
        acc = fma(a, b, acc);
    }
    double* dst = C + (size_t)row * p.ldc + col;
    *dst = acc + (p.accumulate ? *dst : 0.0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// 4-D map (inner doubles, rows, batch level 1, batch level 2) of a row-major real matrix of `rows` x `cols` doubles.
bool make_map(CUtensorMap* map, const double* base, long long rows, long long cols, long long ld, int n1, long long s1,
              int n2, long long s2, int box_rows, int* use1, int* use2) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || rows < 1 || cols < 1 || (ld & 1)) return false;
    *use1 = (n1 > 1 && s1 != 0) ? 1 : 0;
    *use2 = (n2 > 1 && s2 != 0) ? 1 : 0;
    if ((*use1 && (s1 & 1)) || (*use2 && (s2 & 1))) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(*use1 ? n1 : 1), (cuuint64_t)(*use2 ? n2 : 1)};
    const long long row_bytes = ld * 8;
    const long long b1 = *use1 ? s1 * 8 : row_bytes * rows, b2 = *use2 ? s2 * 8 : row_bytes * rows;
    if (row_bytes <= 0 || b1 <= 0 || b2 <= 0 || row_bytes >= (1ll << 40) || b1 >= (1ll << 40) || b2 >= (1ll << 40)) return false;
    if ((b1 & 15) || (b2 & 15)) return false;
    const cuuint64_t strides[3] = {(cuuint64_t)row_bytes, (cuuint64_t)b1, (cuuint64_t)b2};
    const cuuint32_t box[4] = {16, (cuuint32_t)box_rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB>
int launch_tma(const DArgs& p, cudaStream_t st) {
    CUtensorMap mapA, mapB;
    int a1, a2, b1, b2;
    if (!make_map(&mapA, p.A, TA ? p.K : p.M, TA ? p.M : p.K, p.lda, p.batch1, p.sA1, p.batch2, p.sA2, TA ? DBK : BM, &a1, &a2))
        return -1;
    if (!make_map(&mapB, p.B, TB ? p.N : p.K, TB ? p.K : p.N, p.ldb, p.batch1, p.sB1, p.batch2, p.sB2, TB ? BN : DBK, &b1, &b2))
        return -1;
    constexpr size_t SMEM = (size_t)DSTAGES * (BM + BN) * DBK * 8 + 1024;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(dgemm_tma_kernel<BM, BN, WM, WN, TA, TB>), SMEM)) return rc;
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.batch1 * p.batch2);
    dgemm_tma_kernel<BM, BN, WM, WN, TA, TB><<<grid, 32 * WM * WN, SMEM, st>>>(mapA, mapB, p, a1, a2, b1, b2);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <bool TA, bool TB>
int launch(const DArgs& p, cudaStream_t st) {
    const char* e = getenv("MPSB_DGEMM_TMA");
    if (!(e && *e == '0') && p.K > 0) {
        const long long big = (long long)((p.M + 127) / 128) * ((p.N + 127) / 128) * p.batch1 * p.batch2;
        const int rc = big >= 148 ? launch_tma<128, 128, 4, 4, TA, TB>(p, st) : launch_tma<64, 64, 4, 2, TA, TB>(p, st);
        if (rc >= 0) return rc;
    }
    dim3 grid((p.N + 31) / 32, (p.M + 7) / 8, p.batch1 * p.batch2);
    MPSB_REQUIRE(grid.y <= 65535, "dgemm: %d rows exceed the grid of the SIMT kernel", p.M);
    dgemm_simt_kernel<<<grid, 256, 0, st>>>(p, TA, TB);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int dgemm(const DArgs& p, cudaStream_t st) {
    MPSB_REQUIRE(p.M >= 0 && p.N >= 0 && p.K >= 0 && p.batch1 >= 0 && p.batch2 >= 1, "dgemm: negative dims");
    MPSB_REQUIRE(p.opA >= 0 && p.opA <= 3 && p.opB >= 0 && p.opB <= 3, "dgemm: bad op %d / %d", p.opA, p.opB);
    if (p.M == 0 || p.N == 0 || p.batch1 == 0) return 0;
    MPSB_REQUIRE((long long)p.batch1 * p.batch2 <= 65535, "dgemm: batch %d x %d exceeds grid.z", p.batch1, p.batch2);
    const bool ta = p.opA == 1 || p.opA == 2, tb = p.opB == 1 || p.opB == 2;
    if (ta) return tb ? launch<true, true>(p, st) : launch<true, false>(p, st);
    return tb ? launch<false, true>(p, st) : launch<false, false>(p, st);
}

}  // namespace mpsb
