// Batched complex128 GEMM on FP64 tensor cores (DMMA m8n8k4), the workhorse behind every dense
// contraction of the two-site update:
//   theta0 = B_i . B_{i+1}                      (reference DMRG/MPS.py:445, first two operands)
//   B_i'   = theta_B . Vh^H                     (reference DMRG/MPS.py:454)
//   carry  = Vh . M[i+1], M[i-1] . U            (reference DMRG/MPS.py:264,284)
//   transfer-matrix steps of dot/corr           (reference DMRG/MPS.py:351-355,384-388)
//   L.theta, WW.T1, T3.R of the matrix-free H_eff matvec and the environment updates
//                                               (reference DMRG/iDMRG.py:40,59,60)
// The reference evaluates all of these with np.einsum loops; here each is one launch of this kernel.
//
// C[z] = op(A[z]) * op(B[z]) (+ C[z] if accumulate), row-major, leading dimensions in elements.
// op: 0 = N, 1 = T (transpose), 2 = C (conjugate transpose), 3 = J (conjugate only).
// CTA tile 64x64x16, 8 warps (4x2), warp tile 16x32 = 2x4 DMMA tiles.  Two operand pipelines:
//   zgemm_tma_kernel   tiles arrive by TMA (cp.async.bulk.tensor, SASS UTMALDG) through 4-D tensor maps over the
//                      (re, im)-interleaved matrices, 128-byte swizzle, a 3-stage ring of mbarriers, one issuing
//                      thread; out-of-range parts of a tile are zero-filled by the TMA unit.  Default.
//   zgemm_kernel       cp.async double buffering (one 16-byte LDGSTS per complex element) - the fallback for operands
//                      a tensor map cannot describe (misaligned base, strides beyond its limits) and for MPSB_ZGEMM_TMA=0.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "mpsb_internal.h"

namespace mpsb {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int AS_LD = BK + 4;   // 20 cplx = 320 B  (== 64 mod 128: conflict-free A fragments)
constexpr int BS_LD = BN + 2;   // 66 cplx = 1056 B (== 32 mod 128: conflict-free B fragments)
constexpr int GEMM_THREADS = 256;
constexpr size_t GEMM_SMEM = sizeof(cplx) * 2 * (BM * AS_LD + BK * BS_LD);

template <bool TRANS_A, bool TRANS_B>
__device__ __forceinline__ void load_tiles(const GemmArgs& p, const cplx* __restrict__ A,
                                           const cplx* __restrict__ B, cplx* As, cplx* Bs, int m0,
                                           int n0, int k0, int tid) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / GEMM_THREADS; ++i) {
        int e = tid + i * GEMM_THREADS;
        int m, k;
        if (!TRANS_A) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
        int gm = m0 + m, gk = k0 + k;
        bool ok = (gm < p.M) && (gk < p.K);
        const cplx* src = ok ? (TRANS_A ? A + (size_t)gk * p.lda + gm : A + (size_t)gm * p.lda + gk) : A;
        cp_async16(As + m * AS_LD + k, src, ok);
    }
#pragma unroll
    for (int i = 0; i < (BK * BN) / GEMM_THREADS; ++i) {
        int e = tid + i * GEMM_THREADS;
        int k, n;
        if (!TRANS_B) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
        int gk = k0 + k, gn = n0 + n;
        bool ok = (gk < p.K) && (gn < p.N);
        const cplx* src = ok ? (TRANS_B ? B + (size_t)gn * p.ldb + gk : B + (size_t)gk * p.ldb + gn) : B;
        cp_async16(Bs + k * BS_LD + n, src, ok);
    }
}

template <bool TRANS_A, bool CONJ_A, bool TRANS_B, bool CONJ_B>
__global__ void __launch_bounds__(GEMM_THREADS, 2) zgemm_kernel(GemmArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* As = reinterpret_cast<cplx*>(smem_raw);            // [2][BM][AS_LD]
    cplx* Bs = As + 2 * BM * AS_LD;                          // [2][BK][BS_LD]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int z = blockIdx.z;
    const int z1 = z / p.batch2, z2 = z % p.batch2;
    const cplx* A = p.A + (size_t)z1 * p.sA1 + (size_t)z2 * p.sA2;
    const cplx* B = p.B + (size_t)z1 * p.sB1 + (size_t)z2 * p.sB2;
    cplx* C = p.C + (size_t)z1 * p.sC1 + (size_t)z2 * p.sC2;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    double cre[2][4][2], cim[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0; }

    const int nkt = (p.K + BK - 1) / BK;
    load_tiles<TRANS_A, TRANS_B>(p, A, B, As, Bs, m0, n0, 0, tid);
    cp_async_commit();
    for (int kt = 0; kt < nkt; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nkt) {
            load_tiles<TRANS_A, TRANS_B>(p, A, B, As + (cur ^ 1) * BM * AS_LD, Bs + (cur ^ 1) * BK * BS_LD,
                                         m0, n0, (kt + 1) * BK, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const cplx* as = As + cur * BM * AS_LD + (wm * 16 + g) * AS_LD + t;
        const cplx* bs = Bs + cur * BK * BS_LD + t * BS_LD + wn * 32 + g;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            cplx a[2], b[4];
            double nay[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                a[i] = as[i * 8 * AS_LD + kk * 4];
                if (CONJ_A) a[i].y = -a[i].y;
                nay[i] = -a[i].y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b[j] = bs[kk * 4 * BS_LD + j * 8];
                if (CONJ_B) b[j].y = -b[j].y;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(cre[i][j][0], cre[i][j][1], a[i].x, b[j].x);
                    dmma884(cre[i][j][0], cre[i][j][1], nay[i], b[j].y);
                    dmma884(cim[i][j][0], cim[i][j][1], a[i].x, b[j].y);
                    dmma884(cim[i][j][0], cim[i][j][1], a[i].y, b[j].x);
                }
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int row = m0 + wm * 16 + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int col = n0 + wn * 32 + j * 8 + 2 * t;
            cplx* dst = C + (size_t)row * p.ldc + col;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (col + e < p.N) {
                    cplx v = make_double2(cre[i][j][e], cim[i][j][e]);
                    if (p.accumulate) { cplx o = dst[e]; v.x += o.x; v.y += o.y; }
                    dst[e] = v;
                }
            }
        }
    }
}


// ---- TMA pipeline ------------------------------------------------------------------------------------------------
// A stage holds the A tile and the B tile, 16 KiB each, as 128-byte-swizzled boxes of 8 complex (16 doubles) by R rows:
//   operand stored with k contiguous ("KC": A with op N/J, B with op T/C): 2 boxes of 64 rows (rows = m or n), one per
//                                                                          half of the 16 k's, 8 KiB each;
//   operand stored with m/n contiguous ("MC": A with op T/C, B with op N/J): 8 boxes of 16 rows (rows = k), one per 8
//                                                                          consecutive m (n), 2 KiB each.
// Inside a box the 16-byte element (row r, column c) sits at r * 128 + ((c ^ (r & 7)) * 16)  (CU_TENSOR_MAP_SWIZZLE_128B),
// which spreads the 8 rows a quarter-warp touches over all banks.
constexpr int TMA_STAGES = 3;
constexpr int TILE_BYTES = BM * BK * (int)sizeof(cplx);                 // 16 KiB (BM == BN)
constexpr size_t TMA_SMEM = (size_t)TMA_STAGES * 2 * TILE_BYTES + 1024;  // + slack for the 1 KiB alignment swizzling needs

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
            "r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

// element (r, c) of a KC tile: r in [0, 64) (m or n), c in [0, 16) (k)
__device__ __forceinline__ int kc_off(int r, int c) { return (c >> 3) * 8192 + r * 128 + (((c & 7) ^ (r & 7)) << 4); }
// element (k, c) of an MC tile: k in [0, 16), c in [0, 64) (m or n)
__device__ __forceinline__ int mc_off(int k, int c) { return (c >> 3) * 2048 + k * 128 + (((c & 7) ^ (k & 7)) << 4); }

template <bool TRANS_A, bool CONJ_A, bool TRANS_B, bool CONJ_B>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
zgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmArgs p,
                 int az1, int az2, int bz1, int bz2) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[TMA_STAGES];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int z = blockIdx.z;
    const int z1 = z / p.batch2, z2 = z % p.batch2;
    cplx* C = p.C + (size_t)z1 * p.sC1 + (size_t)z2 * p.sC2;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkt = (p.K + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // batch coordinates of the two maps (a shared operand has extent 1 in that dimension)
    const int ca2 = az1 ? z1 : 0, ca3 = az2 ? z2 : 0, cb2 = bz1 ? z1 : 0, cb3 = bz2 ? z2 : 0;
    auto issue = [&](int kt) {          // one thread: all boxes of k-tile kt into stage kt % TMA_STAGES
        const int s = kt % TMA_STAGES;
        unsigned char* as = tiles + (size_t)s * 2 * TILE_BYTES;
        unsigned char* bs = as + TILE_BYTES;
        const int k0 = kt * BK;
        mbar_expect_tx(&full[s], 2 * TILE_BYTES);
        if (!TRANS_A) {                 // memory [M][K]: inner coordinate = 2 k (doubles), rows = m
            tma_load_4d(as, &mapA, 2 * k0, m0, ca2, ca3, &full[s]);
            tma_load_4d(as + 8192, &mapA, 2 * (k0 + 8), m0, ca2, ca3, &full[s]);
        } else {                        // memory [K][M]: inner = 2 m, rows = k
#pragma unroll
            for (int q = 0; q < 8; ++q) tma_load_4d(as + q * 2048, &mapA, 2 * (m0 + 8 * q), k0, ca2, ca3, &full[s]);
        }
        if (!TRANS_B) {                 // memory [K][N]: inner = 2 n, rows = k
#pragma unroll
            for (int q = 0; q < 8; ++q) tma_load_4d(bs + q * 2048, &mapB, 2 * (n0 + 8 * q), k0, cb2, cb3, &full[s]);
        } else {                        // memory [N][K]: inner = 2 k, rows = n
            tma_load_4d(bs, &mapB, 2 * k0, n0, cb2, cb3, &full[s]);
            tma_load_4d(bs + 8192, &mapB, 2 * (k0 + 8), n0, cb2, cb3, &full[s]);
        }
    };
    if (tid == 0)
        for (int kt = 0; kt < TMA_STAGES - 1 && kt < nkt; ++kt) issue(kt);

    double cre[2][4][2], cim[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0; }

    for (int kt = 0; kt < nkt; ++kt) {
        const int s = kt % TMA_STAGES;
        // the stage that k-tile kt + STAGES - 1 will use was read in iteration kt - 1; everybody is past that barrier
        if (tid == 0 && kt + TMA_STAGES - 1 < nkt) issue(kt + TMA_STAGES - 1);
        if (!mbar_wait(&full[s], (kt / TMA_STAGES) & 1)) __trap();
        const unsigned char* as = tiles + (size_t)s * 2 * TILE_BYTES;
        const unsigned char* bs = as + TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            cplx a[2], b[4];
            double nay[2];
            const int k = kk * 4 + t;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = wm * 16 + i * 8 + g;
                a[i] = *reinterpret_cast<const cplx*>(as + (TRANS_A ? mc_off(k, m) : kc_off(m, k)));
                if (CONJ_A) a[i].y = -a[i].y;
                nay[i] = -a[i].y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = wn * 32 + j * 8 + g;
                b[j] = *reinterpret_cast<const cplx*>(bs + (TRANS_B ? kc_off(n, k) : mc_off(k, n)));
                if (CONJ_B) b[j].y = -b[j].y;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(cre[i][j][0], cre[i][j][1], a[i].x, b[j].x);
                    dmma884(cre[i][j][0], cre[i][j][1], nay[i], b[j].y);
                    dmma884(cim[i][j][0], cim[i][j][1], a[i].x, b[j].y);
                    dmma884(cim[i][j][0], cim[i][j][1], a[i].y, b[j].x);
                }
        }
        __syncthreads();            // stage s may be refilled from the next iteration on
    }

#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int row = m0 + wm * 16 + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int col = n0 + wn * 32 + j * 8 + 2 * t;
            cplx* dst = C + (size_t)row * p.ldc + col;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (col + e < p.N) {
                    cplx v = make_double2(cre[i][j][e], cim[i][j][e]);
                    if (p.accumulate) { cplx o = dst[e]; v.x += o.x; v.y += o.y; }
                    dst[e] = v;
                }
            }
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
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

// 4-D map (inner doubles, rows, batch level 1, batch level 2) of a row-major complex matrix with `rows` x `cols`
// elements in memory; box = 16 doubles x box_rows.  Returns false when the operand cannot be described.
bool make_map(CUtensorMap* map, const cplx* base, long long rows, long long cols, long long ld, int n1, long long s1, int n2,
              long long s2, int box_rows, int* use1, int* use2) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || rows < 1 || cols < 1) return false;
    *use1 = (n1 > 1 && s1 != 0) ? 1 : 0;
    *use2 = (n2 > 1 && s2 != 0) ? 1 : 0;
    const cuuint64_t dims[4] = {(cuuint64_t)(2 * cols), (cuuint64_t)rows, (cuuint64_t)(*use1 ? n1 : 1), (cuuint64_t)(*use2 ? n2 : 1)};
    const long long row_bytes = ld * 16;
    const long long b1 = *use1 ? s1 * 16 : row_bytes * rows, b2 = *use2 ? s2 * 16 : row_bytes * rows;
    if (row_bytes <= 0 || b1 <= 0 || b2 <= 0 || row_bytes >= (1ll << 40) || b1 >= (1ll << 40) || b2 >= (1ll << 40)) return false;
    const cuuint64_t strides[3] = {(cuuint64_t)row_bytes, (cuuint64_t)b1, (cuuint64_t)b2};
    const cuuint32_t box[4] = {16, (cuuint32_t)box_rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<cplx*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns 0 when launched, -1 when the operands do not fit a tensor map (caller falls back), > 0 on errors.
template <bool TA, bool CA, bool TB, bool CB>
int launch_tma(const GemmArgs& p, cudaStream_t st) {
    CUtensorMap mapA, mapB;
    int a1, a2, b1, b2;
    // A in memory: op N/J -> [M][K], op T/C -> [K][M];  B in memory: op N/J -> [K][N], op T/C -> [N][K]
    if (!make_map(&mapA, p.A, TA ? p.K : p.M, TA ? p.M : p.K, p.lda, p.batch1, p.sA1, p.batch2, p.sA2, TA ? BK : BM, &a1, &a2))
        return -1;
    if (!make_map(&mapB, p.B, TB ? p.N : p.K, TB ? p.K : p.N, p.ldb, p.batch1, p.sB1, p.batch2, p.sB2, TB ? BN : BK, &b1, &b2))
        return -1;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(zgemm_tma_kernel<TA, CA, TB, CB>), TMA_SMEM)) return rc;
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.batch1 * p.batch2);
    zgemm_tma_kernel<TA, CA, TB, CB><<<grid, GEMM_THREADS, TMA_SMEM, st>>>(mapA, mapB, p, a1, a2, b1, b2);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <bool TA, bool CA, bool TB, bool CB>
int launch(const GemmArgs& p, cudaStream_t st) {
    const char* e = getenv("MPSB_ZGEMM_TMA");
    if (!(e && *e == '0') && p.K > 0) {
        const int rc = launch_tma<TA, CA, TB, CB>(p, st);
        if (rc >= 0) return rc;
    }
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(zgemm_kernel<TA, CA, TB, CB>), GEMM_SMEM)) return rc;
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.batch1 * p.batch2);
    zgemm_kernel<TA, CA, TB, CB><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(p);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <bool TA, bool CA>
int dispatch_b(const GemmArgs& p, cudaStream_t st) {
    switch (p.opB) {
        case 0: return launch<TA, CA, false, false>(p, st);
        case 1: return launch<TA, CA, true, false>(p, st);
        case 2: return launch<TA, CA, true, true>(p, st);
        case 3: return launch<TA, CA, false, true>(p, st);
    }
    set_error("zgemm: bad opB %d", p.opB);
    return 1;
}

}  // namespace

int zgemm(const GemmArgs& p, cudaStream_t st) {
    MPSB_REQUIRE(p.M >= 0 && p.N >= 0 && p.K >= 0 && p.batch1 >= 0 && p.batch2 >= 1, "zgemm: negative dims");
    if (p.M == 0 || p.N == 0 || p.batch1 == 0) return 0;
    MPSB_REQUIRE((long long)p.batch1 * p.batch2 <= 65535, "zgemm: batch %d x %d exceeds grid.z", p.batch1, p.batch2);
    switch (p.opA) {
        case 0: return dispatch_b<false, false>(p, st);
        case 1: return dispatch_b<true, false>(p, st);
        case 2: return dispatch_b<true, true>(p, st);
        case 3: return dispatch_b<false, true>(p, st);
    }
    set_error("zgemm: bad opA %d", p.opA);
    return 1;
}

}  // namespace mpsb
