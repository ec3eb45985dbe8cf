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
// CTA tile 64x64x16, 8 warps (4x2), warp tile 16x32 = 2x4 DMMA tiles, cp.async double buffering
// (one 16-byte LDGSTS per complex element, which also performs the transposes for free).
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

template <bool TA, bool CA, bool TB, bool CB>
int launch(const GemmArgs& p, cudaStream_t st) {
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
