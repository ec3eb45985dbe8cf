// TEBD two-site update (reference MPS.update_double, DMRG/MPS.py:428-457) and one-site gate
// (MPS.update_single, DMRG/MPS.py:413-425), batched over independent bonds / chains.
//
// Launch sequence of one batched update:
//   zgemm            theta0 = B_i . B_{i+1}                                   (first two operands of :445)
//   gate_scale       thetaB = U . theta0  and  X = Sl * thetaB               (third operand of :445, :446)
//   svd_orthogonalise / svd_finalise   (s, Vh) with the reference truncation  (:448-452)  -> B_{i+1}', Sr, rank
//   zgemm            B_i' = thetaB . Vh^H                                     (:454)
#include "common.cuh"
#include "mpsb_internal.h"

namespace mpsb {

namespace {

// One thread per output element (l, a, b, k); k is the fastest index so loads and stores coalesce.
// HBM-bound: reads d^2 * 16 B (L1/L2-cached across the d^2 threads sharing them) and writes 32 B.
__global__ void __launch_bounds__(256)
gate_scale_kernel(int chil, int chir, int d, const cplx* __restrict__ theta0, long long s_theta0,
                  const cplx* __restrict__ U, long long s_U, const double* __restrict__ Sl, long long s_Sl,
                  cplx* __restrict__ thetaB, long long s_thetaB, cplx* __restrict__ X, long long s_X, int ldX) {
    const int b = blockIdx.y;
    const long long total = (long long)chil * d * d * chir;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int k = (int)(e % chir);
    long long r = e / chir;
    const int pb = (int)(r % d); r /= d;
    const int pa = (int)(r % d);
    const int l = (int)(r / d);
    const cplx* t0 = theta0 + (size_t)b * s_theta0;
    const cplx* u = U + (size_t)b * s_U + (size_t)(pa * d + pb) * d * d;   // U[a][b][c][j]
    const int ncol = d * chir;
    double re = 0.0, im = 0.0;
    for (int c = 0; c < d; ++c)
        for (int j = 0; j < d; ++j) {
            const cplx g = u[c * d + j];
            const cplx x = t0[(size_t)(l * d + c) * ncol + (size_t)j * chir + k];
            re = fma(g.x, x.x, fma(-g.y, x.y, re));
            im = fma(g.x, x.y, fma(g.y, x.x, im));
        }
    const int row = l * d + pa, col = pb * chir + k;
    thetaB[(size_t)b * s_thetaB + (size_t)row * ncol + col] = make_double2(re, im);
    const double s = Sl[(size_t)b * s_Sl + l];
    X[(size_t)b * s_X + (size_t)row * ldX + col] = make_double2(s * re, s * im);
}

constexpr int ONE_SITE_DMAX = 16;
__global__ void __launch_bounds__(256)
one_site_kernel(int chil, int chir, int d, const cplx* M, long long sM, const cplx* __restrict__ U, long long sU,
                cplx* Mo, long long sMo) {      // M and Mo may be the same buffer (in-place gate): no __restrict__
    const int b = blockIdx.y;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)chil * chir) return;
    const int r = (int)(e % chir), l = (int)(e / chir);
    const cplx* m = M + (size_t)b * sM + (size_t)l * d * chir + r;
    const cplx* u = U + (size_t)b * sU;
    cplx in[ONE_SITE_DMAX];
    for (int c = 0; c < d; ++c) in[c] = m[(size_t)c * chir];
    cplx* o = Mo + (size_t)b * sMo + (size_t)l * d * chir + r;
    for (int q = 0; q < d; ++q) {
        double re = 0.0, im = 0.0;
        for (int c = 0; c < d; ++c) {
            const cplx g = u[q * d + c];
            re = fma(g.x, in[c].x, fma(-g.y, in[c].y, re));
            im = fma(g.x, in[c].y, fma(g.y, in[c].x, im));
        }
        o[(size_t)q * chir] = make_double2(re, im);
    }
}

// Site tensor times MPO tensor (BMPS.update_MPO, transmit/bhopping.py:94-104):
//   out[(i,l), o, (k,m)] = sum_j M[i,j,k] W[l,m,j,o].
// One thread per output element, (k,m) fastest so stores coalesce; M[i,:,k] is shared by the wl*wr*d threads
// around it and W is small, so both operands come from L1/L2.  HBM-bound on the store (16 B per element).
__global__ void __launch_bounds__(256)
mpo_site_kernel(int chil, int chir, int d, int wl, int wr, const cplx* __restrict__ M, long long sM,
                const cplx* __restrict__ W, long long sW, cplx* __restrict__ out, long long sOut) {
    const int b = blockIdx.y;
    const long long total = (long long)chil * wl * d * chir * wr;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int m = (int)(e % wr);
    long long r = e / wr;
    const int k = (int)(r % chir); r /= chir;
    const int o = (int)(r % d); r /= d;
    const int l = (int)(r % wl);
    const int i = (int)(r / wl);
    const cplx* src = M + (size_t)b * sM + (size_t)i * d * chir + k;
    const cplx* w = W + (size_t)b * sW + ((size_t)(l * wr + m) * d) * d + o;
    double re = 0.0, im = 0.0;
    for (int j = 0; j < d; ++j) {
        const cplx x = src[(size_t)j * chir];
        const cplx g = __ldg(w + (size_t)j * d);
        re = fma(g.x, x.x, fma(-g.y, x.y, re));
        im = fma(g.x, x.y, fma(g.y, x.x, im));
    }
    out[(size_t)b * sOut + e] = make_double2(re, im);
}

}  // namespace

int mpo_site(int batch, int chil, int chir, int d, int wl, int wr, const cplx* M, long long sM, const cplx* W,
             long long sW, cplx* out, long long sOut, cudaStream_t st) {
    MPSB_REQUIRE(chil >= 1 && chir >= 1 && d >= 1 && wl >= 1 && wr >= 1, "mpo_site: bad dims");
    MPSB_REQUIRE(batch <= 65535, "mpo_site: batch %d exceeds grid.y", batch);
    if (batch == 0) return 0;
    const long long total = (long long)chil * wl * d * chir * wr;
    MPSB_REQUIRE((total + 255) / 256 <= 2147483647LL, "mpo_site: output too large for one launch");
    dim3 grid((unsigned)((total + 255) / 256), batch);
    mpo_site_kernel<<<grid, 256, 0, st>>>(chil, chir, d, wl, wr, M, sM, W, sW, out, sOut);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int gate_scale(int batch, int chil, int chir, int d, const cplx* theta0, long long s_theta0, const cplx* U,
               long long s_U, const double* Sl, long long s_Sl, cplx* thetaB, long long s_thetaB, cplx* X,
               long long s_X, int ldX, cudaStream_t st) {
    const long long total = (long long)chil * d * d * chir;
    MPSB_REQUIRE(batch <= 65535, "gate_scale: batch %d exceeds grid.y", batch);
    dim3 grid((unsigned)((total + 255) / 256), batch);
    gate_scale_kernel<<<grid, 256, 0, st>>>(chil, chir, d, theta0, s_theta0, U, s_U, Sl, s_Sl, thetaB, s_thetaB, X,
                                            s_X, ldX);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int one_site(int batch, int chil, int chir, int d, const cplx* M, long long sM, const cplx* U, long long sU, cplx* Mo,
             long long sMo, cudaStream_t st) {
    MPSB_REQUIRE(d >= 1, "one_site: local dimension %d", d);
    MPSB_REQUIRE(batch <= 65535, "one_site: batch %d exceeds grid.y", batch);
    const long long total = (long long)chil * chir;
    if (total == 0 || batch == 0) return 0;
    if (d > ONE_SITE_DMAX) {        // large local spaces (boson chains, fused sites): one d x d GEMM per left index
        MPSB_REQUIRE(M != Mo, "one_site: the in-place form needs d <= %d (d = %d)", ONE_SITE_DMAX, d);
        GemmArgs g;
        g.A = U; g.B = M; g.C = Mo;
        g.M = d; g.N = chir; g.K = d; g.lda = d; g.ldb = chir; g.ldc = chir; g.opA = 0; g.opB = 0;
        g.batch1 = batch; g.batch2 = chil;
        g.sA1 = sU; g.sB1 = sM; g.sC1 = sMo; g.sA2 = 0; g.sB2 = (long long)d * chir; g.sC2 = (long long)d * chir;
        g.accumulate = 0;
        return zgemm(g, st);
    }
    dim3 grid((unsigned)((total + 255) / 256), batch);
    one_site_kernel<<<grid, 256, 0, st>>>(chil, chir, d, M, sM, U, sU, Mo, sMo);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mpsb
