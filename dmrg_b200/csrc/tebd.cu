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
one_site_kernel(int chil, int chir, int d, const cplx* __restrict__ M, long long sM, const cplx* __restrict__ U,
                long long sU, cplx* __restrict__ Mo, long long sMo) {
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

}  // namespace

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
    MPSB_REQUIRE(d >= 1 && d <= ONE_SITE_DMAX, "one_site: local dimension %d not in [1, %d]", d, ONE_SITE_DMAX);
    MPSB_REQUIRE(batch <= 65535, "one_site: batch %d exceeds grid.y", batch);
    const long long total = (long long)chil * chir;
    if (total == 0 || batch == 0) return 0;
    dim3 grid((unsigned)((total + 255) / 256), batch);
    one_site_kernel<<<grid, 256, 0, st>>>(chil, chir, d, M, sM, U, sU, Mo, sMo);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mpsb
