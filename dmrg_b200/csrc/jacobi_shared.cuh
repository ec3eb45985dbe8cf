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

#ifdef __CUDACC__
// Round-robin (circle method) pairing of m (even) players: round r in [0, m-1), slot s in [0, m/2).
__device__ __forceinline__ void rr_pair(int m, int r, int s, int& p, int& q) {
    const int mm = m - 1;
    if (s == 0) { p = mm; q = r % mm; return; }
    int a = (r + s) % mm, b = (r - s) % mm;
    if (b < 0) b += mm;
    p = a; q = b;
}

// Modulus ordering of the block pairs (Luk & Park): step k in [0, m) holds the pairs {I, J} with I + J = k (mod m).
// Unlike the round-robin ordering it behaves like the cyclic-by-rows ordering, and together with de Rijk's rule
// (after a rotation the row of the LOWER block keeps the larger norm) it needs 10 instead of 12 sweeps and 20 %
// fewer rotations on the chi = 128 benchmark matrices (numpy model: tools/jacobi_order_model.py).
// k even: slots t = 1 .. m/2-1 are cross pairs, blocks k/2 and k/2 + m/2 meet themselves (separate "self" tasks);
// k odd : slots t = 0 .. m/2-1 are all cross pairs.  `slot` counts the cross pairs of the step from 0; bI < bJ.
__host__ __device__ inline int mod_cross_pairs(int m, int k) { return (k & 1) ? m / 2 : m / 2 - 1; }
__host__ __device__ __forceinline__ void mod_pair(int m, int k, int slot, int& bI, int& bJ) {
    int a, b;
    if (k & 1) { a = (k - 1) / 2 - slot; b = (k + 1) / 2 + slot; }
    else { a = k / 2 - (slot + 1); b = k / 2 + (slot + 1); }
    a %= m; if (a < 0) a += m;
    b %= m;
    bI = a < b ? a : b;
    bJ = a < b ? b : a;
}
// Stamp slots of one problem (bsz = 8): [0, m) change stamps of the blocks | m + I m + J "pair (I, J) verified clean"
// | m + m^2 + u "self unit u (blocks u, u + m/2) verified clean".  The legacy round-robin kernels index
// m + round (m/2) + slot, which stays inside the second range.
__host__ __device__ inline int stamp_pair_index(int m, int bI, int bJ) { return m + bI * m + bJ; }
__host__ __device__ inline int stamp_self_index(int m, int u) { return m + m * m + u; }

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
