// Column-split Jacobi round: the hot kernel of the truncated SVD for rows of <= 256 columns.
//
// One CTA orthogonalises the 16 rows of a block pair (8 + 8).  Unlike jacobi_cross_kernel (one warp per row,
// partner rows streaming through shared memory) the COLUMNS are dealt out to the warps: warp w, lane l owns
// columns [CPL * (32 w + l), +CPL) of ALL 16 rows in registers for the whole task.  Consequences:
//   * no row data ever passes through shared memory (global -> registers -> global, 512-byte warp transactions);
//   * a tournament step handles its 8 disjoint pairs together: every lane forms the 8 partial inner products of its
//     columns, ONE transposing butterfly reduces the 16 real sums of a warp (32 shuffles for 8 pairs instead of
//     20 per pair), the warps' partials meet in shared memory;
//   * the rotation parameters of the 8 pairs are computed once, by 8 lanes of one warp side by side (SIMD over
//     pairs), instead of by all 32 lanes of a warp per pair;
//   * every lane then applies the 8 rotations to its own columns (8 FMA per complex column, scaled "fast Givens"
//     form exactly as in jacobi_cross_kernel: rows are kept as x = X / d).
// FP64 instructions per pair, summed over the CTA: 32 (inner product) + 16 (butterflies) + ~6 (parameters) + 64
// (rotation) against ~165 in the row-per-warp kernel.
//
// FULL16 = false: the 8 x 8 cross pairs, 8 steps (rounds >= 1 of a sweep).
// FULL16 = true : a complete 15-step tournament of the 16 rows (round 0 of a sweep covers the pairs inside a block).
//
// Replaces the arithmetic of scipy.linalg.svd inside svd_truncate (DMRG/MPS.py:41-45, called from update_double
// DMRG/MPS.py:448-452); the truncation itself is finalise_kernel (jacobi_svd.cu).
#include "common.cuh"
#include "mpsb_internal.h"
#include "jacobi_shared.cuh"

namespace mpsb {

namespace {

constexpr unsigned FULLMASK = 0xffffffffu;

// Sum 16 per-lane doubles over the warp with a transposing butterfly: 15 + 1 additions and 32 32-bit shuffles.
// On return lane l holds the warp total of v[(l >> 1) & 15].
__device__ __forceinline__ double reduce16(double (&v)[16], int lane) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int o = 16 >> k, m = 8 >> k;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < m; ++i) {
            const double keep = up ? v[i + m] : v[i];
            const double send = up ? v[i] : v[i + m];
            v[i] = keep + __shfl_xor_sync(FULLMASK, send, o);
        }
    }
    return v[0] + __shfl_xor_sync(FULLMASK, v[0], 1);
}

template <bool FULL16>
__device__ __forceinline__ void step_pair(int s, int j, int& p, int& q) {
    if (FULL16) {                       // circle method on 16 players, player 15 fixed
        if (j == 0) { p = s; q = 15; }
        else { p = (s + j) % 15; q = (s + 15 - j) % 15; }
    } else {                            // row j of block I meets row (j + s) mod 8 of block J
        p = j; q = 8 + ((j + s) & 7);
    }
}

struct CsSmem {
    double part[8][16];     // per-warp partial sums of the current step (8 complex inner products)
    double a[16];           // true squared norms of the rows (dot columns)
    double d[16];           // scales: X = d * x
    double e[16];           // d^2
    double prm[8][4];       // tau_p (re, im), sig_q (re, im) of the pairs of the current step
    int mask;               // bit j: pair j of the current step rotates
    int dirty;              // bit r: row r was rotated in this task
    int nrot;
};

// CPL: complex columns per lane (1: 8 warps cover 256 columns; 2: 4 warps).  EXACT: ldot == ltot_r == 32*CPL*warps,
// so no column predicates are needed anywhere.
template <int CPL, bool FULL16, bool EXACT>
__global__ void __launch_bounds__(256 / CPL, CPL == 1 ? 2 : 3)
jacobi_colsplit_kernel(cplx* __restrict__ Xall, long long strideX, int mat0, int ld, int ldot, int ltot_r, int nblk,
                       int round, int* __restrict__ rot, const int* __restrict__ done, double tol2,
                       unsigned long long* __restrict__ phase_clk, int* __restrict__ stamps, int stamp_stride,
                       int launch_id) {
    __shared__ CsSmem sm;
    constexpr int NSTEP = FULL16 ? 15 : 8;

    const int mat = mat0 + blockIdx.y;
    if (done[mat]) return;
    const long long t_start = phase_clk ? clock64() : 0;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    int bI, bJ;
    rr_pair(nblk, round, blockIdx.x, bI, bJ);
    int* stp = stamps ? stamps + (size_t)mat * stamp_stride : nullptr;
    const int chk_idx = nblk + round * (nblk >> 1) + blockIdx.x;
    if (stp) {                                               // verified clean since both blocks last changed
        const int chk = stp[chk_idx];
        if (chk > stp[bI] && chk > stp[bJ]) return;
    }
    cplx* X = Xall + (size_t)mat * strideX;

    // ---- load: columns c0 + 32*k (k < CPL) of the 16 rows, straight into registers ----------------------------
    const int c0 = w * (32 * CPL) + lane;
    bool live[CPL], indot[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        live[k] = EXACT || (c0 + 32 * k < ltot_r);
        indot[k] = EXACT || (c0 + 32 * k < ldot);
    }
    cplx x[16][CPL];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const cplx* g = X + (size_t)((r < 8 ? bI : bJ) * 8 + (r & 7)) * ld + c0;
#pragma unroll
        for (int k = 0; k < CPL; ++k) x[r][k] = live[k] ? __ldcg(g + 32 * k) : make_double2(0.0, 0.0);
    }
    if (tid == 0) { sm.dirty = 0; sm.nrot = 0; }
    {   // squared norms: per-lane partials -> warp totals -> part[w][r]; the parameter warp of step 0 adds them up
        double v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < CPL; ++k)
                if (indot[k]) acc = fma(x[r][k].x, x[r][k].x, fma(x[r][k].y, x[r][k].y, acc));
            v[r] = acc;
        }
        const double tot = reduce16(v, lane);
        if ((lane & 1) == 0) sm.part[w][lane >> 1] = tot;
    }
    __syncthreads();
    if (tid < 16) {
        double acc = 0.0;
        for (int ww = 0; ww < nw; ++ww) acc += sm.part[ww][tid];
        sm.a[tid] = acc;
        sm.d[tid] = 1.0;
        sm.e[tid] = 1.0;
    }
    __syncthreads();        // part[] is rewritten by step 0
    const long long t_loaded = phase_clk ? clock64() : 0;

#pragma unroll
    for (int s = 0; s < NSTEP; ++s) {
        // ---- A. partial inner products x_p . conj(x_q) of this lane's columns, all 8 pairs ------------------
        double v[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int p, q;
            step_pair<FULL16>(s, j, p, q);
            double cr = 0.0, ci = 0.0;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                if (indot[k]) {
                    cr = fma(x[p][k].x, x[q][k].x, fma(x[p][k].y, x[q][k].y, cr));
                    ci = fma(x[p][k].y, x[q][k].x, fma(-x[p][k].x, x[q][k].y, ci));
                }
            }
            v[2 * j] = cr;
            v[2 * j + 1] = ci;
        }
        const double tot = reduce16(v, lane);
        if ((lane & 1) == 0) sm.part[w][lane >> 1] = tot;
        __syncthreads();
        // ---- B. rotation parameters: lane j of one warp handles pair j ---------------------------------------
        if (w == s % nw) {
            int rotate = 0;
            if (lane < 8) {
                int p, q;
                step_pair<FULL16>(s, lane, p, q);
                double crs = 0.0, cis = 0.0;
                for (int ww = 0; ww < nw; ++ww) {
                    crs += sm.part[ww][2 * lane];
                    cis += sm.part[ww][2 * lane + 1];
                }
                const double a = sm.a[p], b = sm.a[q], ep = sm.e[p], eq = sm.e[q];
                const double ac2 = ep * eq * fma(crs, crs, cis * cis);          // |c|^2 of the true rows
                if (ac2 > fma(tol2 * a, b, PAIR_FLOOR)) {
                    rotate = 1;
                    // t/|c| = sgn(z) / (|z| + sqrt(z^2 + |c|^2)); see cross_pair() in jacobi_svd.cu for the scheme
                    const double z = 0.5 * (b - a);
                    const double v1 = fma(z, z, ac2);
                    const double denom = fabs(z) + v1 * newton_rsqrt(v1, rsqrt_approx(v1));
                    const double inv = newton_rcp(denom, rcp_approx(denom));
                    const double kk = z >= 0.0 ? inv : -inv;
                    const double kq = kk * eq, kp = kk * ep;
                    sm.prm[lane][0] = kq * crs;       // tau_p = tau d_q / d_p
                    sm.prm[lane][1] = kq * cis;
                    sm.prm[lane][2] = kp * crs;       // sig_q = conj(tau) d_p / d_q
                    sm.prm[lane][3] = -kp * cis;
                    const double wv = fma(inv * inv, ac2, 1.0);
                    const double cs = newton_rsqrt(wv, newton_rsqrt(wv, rsqrt_approx(wv)));
                    const double cs2 = cs * cs;
                    const double delta = kk * ac2;                               // t |c|
                    sm.a[p] = a - delta;
                    sm.a[q] = b + delta;
                    sm.d[p] *= cs;
                    sm.d[q] *= cs;
                    sm.e[p] = ep * cs2;
                    sm.e[q] = eq * cs2;
                }
            }
            const unsigned bal = __ballot_sync(FULLMASK, rotate);
            if (lane < 8 && rotate) {
                int p, q;
                step_pair<FULL16>(s, lane, p, q);
                atomicOr(&sm.dirty, (1 << p) | (1 << q));
            }
            if (lane == 0) {
                sm.mask = (int)bal;
                sm.nrot += __popc(bal);
            }
        }
        __syncthreads();
        // ---- C. apply the rotations of this step to this lane's columns --------------------------------------
        const int mask = sm.mask;
        if (mask) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (mask & (1 << j)) {
                    int p, q;
                    step_pair<FULL16>(s, j, p, q);
                    const double tpr = sm.prm[j][0], tpi = sm.prm[j][1], sqr = sm.prm[j][2], sqi = sm.prm[j][3];
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const cplx xo = x[p][k], yo = x[q][k];
                        x[p][k].x = fma(-tpr, yo.x, fma(tpi, yo.y, xo.x));
                        x[p][k].y = fma(-tpr, yo.y, fma(-tpi, yo.x, xo.y));
                        x[q][k].x = fma(sqr, xo.x, fma(-sqi, xo.y, yo.x));
                        x[q][k].y = fma(sqr, xo.y, fma(sqi, xo.x, yo.y));
                    }
                }
            }
        }
        // no barrier here: part[] is next written after this step's second barrier, prm[]/mask after the next
        // step's first barrier, which every warp reaches only after finishing C
    }
    const int dirty = sm.dirty;      // final: the last write happened before the last barrier
    const long long t_rotated = phase_clk ? clock64() : 0;
    if (!dirty) {
        if (tid == 0) {
            if (stp) stp[chk_idx] = launch_id;      // verified clean at this launch
            if (phase_clk) {
                atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
                atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
                atomicAdd(&phase_clk[3], 1ull);
            }
        }
        return;
    }
    // ---- store: fold the scales back into the rotated rows ---------------------------------------------------
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (dirty & (1 << r)) {
            const double dd = sm.d[r];
            cplx* g = X + (size_t)((r < 8 ? bI : bJ) * 8 + (r & 7)) * ld + c0;
#pragma unroll
            for (int k = 0; k < CPL; ++k)
                if (live[k]) __stcg(g + 32 * k, make_double2(x[r][k].x * dd, x[r][k].y * dd));
        }
    }
    if (tid == 0) {
        atomicAdd(&rot[mat], sm.nrot);
        if (stp) { stp[bI] = launch_id; stp[bJ] = launch_id; }
        if (phase_clk) {
            atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
            atomicAdd(&phase_clk[2], (unsigned long long)(clock64() - t_rotated));
            atomicAdd(&phase_clk[3], 1ull);
        }
    }
}

template <int CPL, bool FULL16, bool EXACT>
int launch_t(const ColsplitArgs& a, cudaStream_t st) {
    const int nw = (a.ltot_r + 32 * CPL - 1) / (32 * CPL);
    dim3 grid(a.nblk / 2, a.count);
    jacobi_colsplit_kernel<CPL, FULL16, EXACT><<<grid, 32 * nw, 0, st>>>(
        a.X, a.strideX, a.mat0, a.ld, a.ldot, a.ltot_r, a.nblk, a.round, a.rot, a.done, a.tol2, a.phase_clk, a.stamps,
        a.stamp_stride, a.launch_id);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int CPL>
int launch_c(const ColsplitArgs& a, cudaStream_t st) {
    const int nw = (a.ltot_r + 32 * CPL - 1) / (32 * CPL);
    const bool exact = (a.ldot == a.ltot_r) && (a.ltot_r == 32 * CPL * nw);
    if (a.full16) return exact ? launch_t<CPL, true, true>(a, st) : launch_t<CPL, true, false>(a, st);
    return exact ? launch_t<CPL, false, true>(a, st) : launch_t<CPL, false, false>(a, st);
}

}  // namespace

bool colsplit_supported(int bsz, int ltot_r, int nblk) {
    return bsz == 8 && ltot_r >= 32 && ltot_r <= 256 && nblk >= 4 && (nblk & 1) == 0;
}

int launch_jacobi_colsplit(const ColsplitArgs& a, cudaStream_t st) {
    MPSB_REQUIRE(colsplit_supported(8, a.ltot_r, a.nblk), "colsplit: unsupported shape ltot_r=%d nblk=%d", a.ltot_r,
                 a.nblk);
    return a.cpl == 2 ? launch_c<2>(a, st) : launch_c<1>(a, st);
}

}  // namespace mpsb
