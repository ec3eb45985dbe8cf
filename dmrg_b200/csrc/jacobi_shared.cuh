// Pieces shared by the Jacobi kernels of jacobi_svd.cu and jacobi_colsplit.cu (internal, device + host).
#pragma once
#include "common.cuh"

namespace mpsb {

// The convergence test |x_p . x_q|^2 > tol^2 |x_p|^2 |x_q|^2 is relative, so that small singular values keep their
// relative accuracy.  Its right-hand side underflows for rows at the 1e-80 level (products of Schmidt values that are
// themselves rounding noise reach 1e-260 in an iDMRG run on a rank-deficient state): such pairs would rotate for ever.
// PAIR_FLOOR is added to the threshold (folded into the FMA that forms it, so it costs nothing): pairs with
// |x_p|^2 |x_q|^2 below ~1e-260 are left alone, and finalise_kernel emits rows with |x|^2 <= ROW_DEAD as zero vectors
// (singular values 1e-62 below the scale of a normalised state carry no information).
constexpr double PAIR_FLOOR = 1e-290;

// Arguments of one column-split Jacobi round (jacobi_colsplit.cu): grid = (nblk / 2, count).
struct ColsplitArgs {
    cplx* X;
    long long strideX;
    int mat0, count;
    int ld, ldot, ltot_r;       // row pitch, columns entering the inner products, columns rotated (multiple of 32)
    int nblk, round;
    int full16;                 // 1: complete tournament of the 16 rows of each block pair (round 0 of a sweep)
    int cpl;                    // complex columns per lane: 1 or 2
    int* rot;
    const int* done;
    double tol2;
    unsigned long long* phase_clk;
    int* stamps;
    int stamp_stride;
    int launch_id;
};
bool colsplit_supported(int bsz, int ltot_r, int nblk);
int launch_jacobi_colsplit(const ColsplitArgs& a, cudaStream_t st);

#ifdef __CUDACC__
// Round-robin (circle method) pairing of m (even) players: round r in [0, m-1), slot s in [0, m/2).
__device__ __forceinline__ void rr_pair(int m, int r, int s, int& p, int& q) {
    const int mm = m - 1;
    if (s == 0) { p = mm; q = r % mm; return; }
    int a = (r + s) % mm, b = (r - s) % mm;
    if (b < 0) b += mm;
    p = a; q = b;
}

// MUFU seeds (about 20 bits) and Newton refinements: each refinement is 2-3 dependent FP64 operations instead
// of the ~20-instruction IEEE division / square-root sequences.
__device__ __forceinline__ double rcp_approx(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double rsqrt_approx(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double newton_rcp(double x, double r) { return fma(r, fma(-x, r, 1.0), r); }
__device__ __forceinline__ double newton_rsqrt(double x, double y) {
    const double h = 0.5 * x * y;
    return fma(y, fma(-h, y, 0.5), y);
}
#endif

}  // namespace mpsb
