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

// ---- sub-group version ------------------------------------------------------------------------------------
// The 16 rows are four quads Q0, Q1 (block I), Q2, Q3 (block J).  The CTA is two SUB-GROUPS of four warps; a
// sub-group keeps two quads (8 rows) in registers, its four warps covering the columns (CPL complex columns per lane),
// and plays 4-pair tournament steps on them with its own named barriers - two independent dependency chains per CTA,
// four per SM, which is what keeps the FP64 pipe fed while a chain sits in a butterfly or in the parameter
// computation.  Between phases the sub-groups swap a quad through shared memory (16 KiB each way):
//   cross round : (Q0 x Q2 | Q1 x Q3)  swap  (Q0 x Q3 | Q1 x Q2)                                   8 steps
//   full round  : (inside Q0, Q1, Q0 x Q1 | inside Q2, Q3, Q2 x Q3)  swap  then the cross round     15 steps
// The duty of computing a step's four rotation parameters (lanes 0-3, one pair each) rotates over the warps of
// the sub-group; the duty warp releases the others as soon as the parameters are in shared memory (bar.arrive) and
// finishes the bookkeeping (cosine, row scales, norms) afterwards.

// Sum 8 per-lane doubles over the warp: lane l ends up with the warp total of v[(l >> 2) & 7].
__device__ __forceinline__ double reduce8(double (&v)[8], int lane) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int o = 16 >> k, m = 4 >> k;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < m; ++i) {
            const double keep = up ? v[i + m] : v[i];
            const double send = up ? v[i] : v[i + m];
            v[i] = keep + __shfl_xor_sync(FULLMASK, send, o);
        }
    }
    double t = v[0] + __shfl_xor_sync(FULLMASK, v[0], 2);
    return t + __shfl_xor_sync(FULLMASK, t, 1);
}

// Named barriers of the sub-groups (ids 1-4; 0 is __syncthreads).  Immediate ids so that ptxas reserves only these.
// which: 0 = "partials posted", 1 = "parameters posted".  NT = threads of a sub-group.
template <int NT>
__device__ __forceinline__ void sg_bar_sync(int sg, int which) {
    if (sg == 0) {
        if (which == 0) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");
    } else {
        if (which == 0) asm volatile("bar.sync 3, %0;" ::"n"(NT) : "memory");
        else asm volatile("bar.sync 4, %0;" ::"n"(NT) : "memory");
    }
}
template <int NT>
__device__ __forceinline__ void sg_bar_arrive(int sg) {       // "parameters posted", without waiting
    if (sg == 0) asm volatile("bar.arrive 2, %0;" ::"n"(NT) : "memory");
    else asm volatile("bar.arrive 4, %0;" ::"n"(NT) : "memory");
}

// Local pair j (0..3) of a step among the 8 local rows (0-3 first quad, 4-7 second quad).
// KIND 0: cross step T (0..3): row j of the first quad meets row (j + T) mod 4 of the second.
// KIND 1: step T (0..2) of the tournaments inside both quads.
template <int KIND, int T>
__device__ __forceinline__ void sg_pair(int j, int& p, int& q) {
    if (KIND == 0) { p = j; q = 4 + ((j + T) & 3); }
    else {
        const int base = (j >> 1) * 4, jj = j & 1;
        if (T == 0) { p = base + 2 * jj; q = p + 1; }
        else if (T == 1) { p = base + jj; q = p + 2; }
        else { p = base + jj; q = base + 3 - jj; }
    }
}

struct SgSmem {
    cplx xch[2][4][256];    // quad exchange buffers, one per sub-group
    double part[2][4][8];   // [sub-group][warp][value] (rows of absent warps stay zero): partial sums of the current step (4 complex inner products)
    double prm[2][4][4];    // [sub-group][pair]: tau_p (re, im), sig_q (re, im)
    double a[16];           // true squared norms of the 16 rows (dot columns)
    double d[16];           // scales: X = d * x
    double e[16];           // d^2
    int mask[2];            // bit j: pair j of the sub-group's current step rotates
    int dirty[2];           // bit r: row r was rotated in this task
    int nrot[2];
};

// PROBE: accumulate SM clocks of the step's phases in pc[0..3] (dots + butterflies | wait for the partials of the
// other warps | wait for the parameters (other warps) | parameter computation (duty warp)).
template <int CPL, int WPS, bool EXACT, bool PROBE, int KIND, int T>
__device__ __forceinline__ void sg_step(cplx (&x)[8][CPL], const bool (&indot)[CPL], SgSmem& sm, int sg, int wsg, int lane,
                                        int duty_, int r0, int r1, double tol2, long long (&pc)[4]) {
    const int duty = duty_ % WPS;
    const long long c0 = PROBE ? clock64() : 0;
    // ---- A. partial inner products of this lane's columns, the 4 pairs of the step ------------------------------
    double v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int p, q;
        sg_pair<KIND, T>(j, p, q);
        double cr = 0.0, ci = 0.0;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            if (EXACT || indot[k]) {
                cr = fma(x[p][k].x, x[q][k].x, fma(x[p][k].y, x[q][k].y, cr));
                ci = fma(x[p][k].y, x[q][k].x, fma(-x[p][k].x, x[q][k].y, ci));
            }
        }
        v[2 * j] = cr;
        v[2 * j + 1] = ci;
    }
    const double tot = reduce8(v, lane);
    if ((lane & 3) == 0) sm.part[sg][wsg][(lane >> 2) & 7] = tot;
    const long long c1 = PROBE ? clock64() : 0;
    sg_bar_sync<32 * WPS>(sg, 0);
    const long long c2 = PROBE ? clock64() : 0;
    // ---- B. rotation parameters: lane j < 4 of the duty warp handles pair j ----------------------------------------
    if (wsg == duty) {
        int rotate = 0;
        double ac2 = 0.0, inv = 0.0, kk = 0.0, a = 0.0, b = 0.0, ep = 0.0, eq = 0.0;
        int gp = 0, gq = 0;
        if (lane < 4) {
            int p, q;
            sg_pair<KIND, T>(lane, p, q);
            gp = p < 4 ? r0 + p : r1 + p - 4;           // ids of the two rows among the 16 of the task
            gq = q < 4 ? r0 + q : r1 + q - 4;
            const double2 s0 = *reinterpret_cast<const double2*>(&sm.part[sg][0][2 * lane]);
            const double2 s1 = *reinterpret_cast<const double2*>(&sm.part[sg][1][2 * lane]);
            const double2 s2 = *reinterpret_cast<const double2*>(&sm.part[sg][2][2 * lane]);
            const double2 s3 = *reinterpret_cast<const double2*>(&sm.part[sg][3][2 * lane]);
            a = sm.a[gp]; b = sm.a[gq]; ep = sm.e[gp]; eq = sm.e[gq];
            const double crs = (s0.x + s1.x) + (s2.x + s3.x), cis = (s0.y + s1.y) + (s2.y + s3.y);
            ac2 = (ep * eq) * fma(crs, crs, cis * cis);              // |c|^2 of the true rows
            double tpr = 0.0, tpi = 0.0, sqr = 0.0, sqi = 0.0;
            if (ac2 > fma(tol2 * a, b, PAIR_FLOOR)) {
                rotate = 1;
                // t/|c| = sgn(z) / (|z| + sqrt(z^2 + |c|^2)); see cross_pair() in jacobi_svd.cu for the scheme
                const double z = 0.5 * (b - a);
                const double v1 = fma(z, z, ac2);
                const double denom = fma(v1, newton_rsqrt(v1, rsqrt_approx(v1)), fabs(z));
                inv = newton_rcp(denom, rcp_approx(denom));
                kk = z >= 0.0 ? inv : -inv;
                const double kq = kk * eq, kp = kk * ep;
                tpr = kq * crs; tpi = kq * cis;       // tau_p = tau d_q / d_p
                sqr = kp * crs; sqi = -kp * cis;      // sig_q = conj(tau) d_p / d_q
            }
            *reinterpret_cast<double2*>(&sm.prm[sg][lane][0]) = make_double2(tpr, tpi);
            *reinterpret_cast<double2*>(&sm.prm[sg][lane][2]) = make_double2(sqr, sqi);
        }
        const unsigned bal = __ballot_sync(FULLMASK, rotate) & 0xfu;
        if (lane == 0) sm.mask[sg] = (int)bal;
        __syncwarp();
        sg_bar_arrive<32 * WPS>(sg);                   // the other warps may rotate now
        if (PROBE) pc[3] += clock64() - c2;
        if (rotate) {                                // bookkeeping, off the critical path of the other warps
            const double wv = fma(inv * inv, ac2, 1.0);
            const double cs = newton_rsqrt(wv, newton_rsqrt(wv, rsqrt_approx(wv)));
            const double cs2 = cs * cs;
            const double delta = kk * ac2;                               // t |c|
            sm.a[gp] = a - delta;
            sm.a[gq] = b + delta;
            sm.d[gp] *= cs;
            sm.d[gq] *= cs;
            sm.e[gp] = ep * cs2;
            sm.e[gq] = eq * cs2;
        }
        unsigned bits = rotate ? ((1u << gp) | (1u << gq)) : 0u;
        bits |= __shfl_xor_sync(FULLMASK, bits, 1);
        bits |= __shfl_xor_sync(FULLMASK, bits, 2);
        if (lane == 0 && bal) {
            sm.dirty[sg] |= (int)bits;
            sm.nrot[sg] += __popc(bal);
        }
    } else {
        sg_bar_sync<32 * WPS>(sg, 1);
        if (PROBE) pc[2] += clock64() - c2;
    }
    if (PROBE) { pc[0] += c1 - c0; pc[1] += c2 - c1; }
    // ---- C. apply the rotations to this lane's columns ----------------------------------------------------------------
    const int mask = sm.mask[sg];
    if (mask) {          // pairs that do not rotate carry tau = sig = 0 (exact identity): one straight-line variant
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int p, q;
            sg_pair<KIND, T>(j, p, q);
            const double2 tp = *reinterpret_cast<const double2*>(&sm.prm[sg][j][0]);
            const double2 sq = *reinterpret_cast<const double2*>(&sm.prm[sg][j][2]);
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const cplx xo = x[p][k], yo = x[q][k];
                x[p][k].x = fma(-tp.x, yo.x, fma(tp.y, yo.y, xo.x));
                x[p][k].y = fma(-tp.x, yo.y, fma(-tp.y, yo.x, xo.y));
                x[q][k].x = fma(sq.x, xo.x, fma(-sq.y, xo.y, yo.x));
                x[q][k].y = fma(sq.x, xo.y, fma(sq.y, xo.x, yo.y));
            }
        }
    }
}

// Swap a quad between the two sub-groups through shared memory.  `mine`: which of my quads leaves (0: rows 0-3, 1:
// rows 4-7); the partner's leaving quad takes its place.  Both sub-groups call this together (CTA-wide barriers).
template <int CPL>
__device__ __forceinline__ void sg_swap(cplx (&x)[8][CPL], SgSmem& sm, int sg, int c0, int mine) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < CPL; ++k) sm.xch[sg][r][c0 + 32 * k] = mine ? x[4 + r][k] : x[r][k];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const cplx t = sm.xch[sg ^ 1][r][c0 + 32 * k];
            if (mine) x[4 + r][k] = t; else x[r][k] = t;
        }
    __syncthreads();
}

// WPS: warps per sub-group (its columns = 32 * CPL * WPS >= ltot_r); the CTA has 2 * WPS warps.
// MODE 0: the 8 x 8 cross pairs of a block pair | 1: complete tournament of its 16 rows | 2: only the pairs INSIDE
// the two blocks (the self tasks of the modulus ordering; blocks round/2 and round/2 + nblk/2, grid.x = 1).
template <int CPL, int WPS, int MODE, bool EXACT, bool PROBE>
__global__ void __launch_bounds__(64 * WPS, WPS == 2 ? 3 : 2)
jacobi_subgroup_kernel(cplx* __restrict__ Xall, long long strideX, int mat0, int ld, int ldot, int ltot_r, int nblk,
                       int round, int* __restrict__ rot, const int* __restrict__ done, double tol2,
                       unsigned long long* __restrict__ phase_clk, int* __restrict__ stamps, int stamp_stride,
                       int launch_id) {
    __shared__ SgSmem sm;

    const int mat = mat0 + blockIdx.y;
    if (done[mat]) return;
    const long long t_start = phase_clk ? clock64() : 0;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, sg = w / WPS, wsg = w % WPS;
    constexpr bool FULL16 = MODE != 0;          // sub-group g starts with the two quads of ONE block
    int bI, bJ;
    if (MODE == 2) { bI = round >> 1; bJ = bI + (nblk >> 1); }
    else rr_pair(nblk, round, blockIdx.x, bI, bJ);
    int* stp = stamps ? stamps + (size_t)mat * stamp_stride : nullptr;
    const int chk_idx = MODE == 2 ? stamp_self_index(nblk, bI) : nblk + round * (nblk >> 1) + blockIdx.x;
    if (stp) {                                               // verified clean since both blocks last changed
        const int chk = stp[chk_idx];
        if (chk > stp[bI] && chk > stp[bJ]) return;
    }
    cplx* X = Xall + (size_t)mat * strideX;

    // Rows of the task are numbered 0-7 (block I) and 8-15 (block J); quad Qk = rows 4k .. 4k+3.
    // r0 / r1: first row of the quad held in x[0..3] / x[4..7].
    int r0 = FULL16 ? 8 * sg : 4 * sg;
    int r1 = FULL16 ? 8 * sg + 4 : 8 + 4 * sg;
    const int c0 = wsg * (32 * CPL) + lane;
    bool live[CPL], indot[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        live[k] = EXACT || (c0 + 32 * k < ltot_r);
        indot[k] = EXACT || (c0 + 32 * k < ldot);
    }
    auto grow = [&](int r) { return X + (size_t)((r < 8 ? bI : bJ) * 8 + (r & 7)) * ld + c0; };
    cplx x[8][CPL];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const cplx* g = grow(r < 4 ? r0 + r : r1 + r - 4);
#pragma unroll
        for (int k = 0; k < CPL; ++k) x[r][k] = live[k] ? __ldcg(g + 32 * k) : make_double2(0.0, 0.0);
    }
    if (tid < 2) { sm.dirty[tid] = 0; sm.nrot[tid] = 0; }
    if (WPS == 2 && tid < 32) sm.part[tid >> 4][2 + ((tid >> 3) & 1)][tid & 7] = 0.0;   // rows of the absent warps 2, 3
    {   // squared norms of my 8 rows
        double v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < CPL; ++k)
                if (indot[k]) acc = fma(x[r][k].x, x[r][k].x, fma(x[r][k].y, x[r][k].y, acc));
            v[r] = acc;
        }
        const double tot = reduce8(v, lane);
        if ((lane & 3) == 0) sm.part[sg][wsg][(lane >> 2) & 7] = tot;
    }
    __syncthreads();
    if (wsg == 0 && lane < 8) {
        const double acc = (sm.part[sg][0][lane] + sm.part[sg][1][lane]) + (sm.part[sg][2][lane] + sm.part[sg][3][lane]);
        const int r = lane < 4 ? r0 + lane : r1 + lane - 4;
        sm.a[r] = acc;
        sm.d[r] = 1.0;
        sm.e[r] = 1.0;
    }
    __syncthreads();        // part[] is rewritten by the first step
    const long long t_loaded = phase_clk ? clock64() : 0;
    long long pc[4] = {0, 0, 0, 0};

    if (FULL16) {
        sg_step<CPL, WPS, EXACT, PROBE, 1, 0>(x, indot, sm, sg, wsg, lane, 0, r0, r1, tol2, pc);
        sg_step<CPL, WPS, EXACT, PROBE, 1, 1>(x, indot, sm, sg, wsg, lane, 1, r0, r1, tol2, pc);
        sg_step<CPL, WPS, EXACT, PROBE, 1, 2>(x, indot, sm, sg, wsg, lane, 2, r0, r1, tol2, pc);
        sg_step<CPL, WPS, EXACT, PROBE, 0, 0>(x, indot, sm, sg, wsg, lane, 3, r0, r1, tol2, pc);
        sg_step<CPL, WPS, EXACT, PROBE, 0, 1>(x, indot, sm, sg, wsg, lane, 0, r0, r1, tol2, pc);
        sg_step<CPL, WPS, EXACT, PROBE, 0, 2>(x, indot, sm, sg, wsg, lane, 1, r0, r1, tol2, pc);
        sg_step<CPL, WPS, EXACT, PROBE, 0, 3>(x, indot, sm, sg, wsg, lane, 2, r0, r1, tol2, pc);
    }
    if (MODE == 1) {
        // (Q0, Q1 | Q2, Q3) -> (Q0, Q2 | Q1, Q3): sub-group 0 gives its second quad, sub-group 1 its first
        sg_swap<CPL>(x, sm, sg, c0, sg == 0 ? 1 : 0);
        if (sg == 0) r1 = 8; else r0 = 4;
    }
    if (MODE != 2) {
    sg_step<CPL, WPS, EXACT, PROBE, 0, 0>(x, indot, sm, sg, wsg, lane, 3, r0, r1, tol2, pc);
    sg_step<CPL, WPS, EXACT, PROBE, 0, 1>(x, indot, sm, sg, wsg, lane, 0, r0, r1, tol2, pc);
    sg_step<CPL, WPS, EXACT, PROBE, 0, 2>(x, indot, sm, sg, wsg, lane, 1, r0, r1, tol2, pc);
    sg_step<CPL, WPS, EXACT, PROBE, 0, 3>(x, indot, sm, sg, wsg, lane, 2, r0, r1, tol2, pc);
    // (Q0, Q2 | Q1, Q3) -> (Q0, Q3 | Q1, Q2): the block-J quads change hands
    sg_swap<CPL>(x, sm, sg, c0, 1);
    r1 = (sg == 0) ? 12 : 8;
    sg_step<CPL, WPS, EXACT, PROBE, 0, 0>(x, indot, sm, sg, wsg, lane, 3, r0, r1, tol2, pc);
    sg_step<CPL, WPS, EXACT, PROBE, 0, 1>(x, indot, sm, sg, wsg, lane, 0, r0, r1, tol2, pc);
    sg_step<CPL, WPS, EXACT, PROBE, 0, 2>(x, indot, sm, sg, wsg, lane, 1, r0, r1, tol2, pc);
    sg_step<CPL, WPS, EXACT, PROBE, 0, 3>(x, indot, sm, sg, wsg, lane, 2, r0, r1, tol2, pc);
    }
    __syncthreads();         // all bookkeeping of both sub-groups is in shared memory

    const int dirty = sm.dirty[0] | sm.dirty[1];
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
    for (int r = 0; r < 8; ++r) {
        const int gr = r < 4 ? r0 + r : r1 + r - 4;
        if (dirty & (1 << gr)) {
            const double dd = sm.d[gr];
            cplx* g = grow(gr);
#pragma unroll
            for (int k = 0; k < CPL; ++k)
                if (live[k]) __stcg(g + 32 * k, make_double2(x[r][k].x * dd, x[r][k].y * dd));
        }
    }
    if (tid == 0) {
        atomicAdd(&rot[mat], sm.nrot[0] + sm.nrot[1]);
        if (stp) { stp[bI] = launch_id; stp[bJ] = launch_id; }
        if (phase_clk) {
            atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
            atomicAdd(&phase_clk[2], (unsigned long long)(clock64() - t_rotated));
            atomicAdd(&phase_clk[3], 1ull);
        }
    }
    if (PROBE && phase_clk && tid == 32 * (WPS - 1)) {     // the last warp of sub-group 0
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(&phase_clk[4 + i], (unsigned long long)pc[i]);
    }
}

template <int CPL, int WPS, int MODE, bool EXACT>
int launch_sg_t(const ColsplitArgs& a, cudaStream_t st) {
    dim3 grid(MODE == 2 ? 1 : a.nblk / 2, a.count);
    if (a.phase_clk && EXACT && MODE == 0)     // instrumented twin (profiling level 2 only)
        jacobi_subgroup_kernel<CPL, WPS, MODE, EXACT, true><<<grid, 64 * WPS, 0, st>>>(
            a.X, a.strideX, a.mat0, a.ld, a.ldot, a.ltot_r, a.nblk, a.round, a.rot, a.done, a.tol2, a.phase_clk, a.stamps,
            a.stamp_stride, a.launch_id);
    else
        jacobi_subgroup_kernel<CPL, WPS, MODE, EXACT, false><<<grid, 64 * WPS, 0, st>>>(
            a.X, a.strideX, a.mat0, a.ld, a.ldot, a.ltot_r, a.nblk, a.round, a.rot, a.done, a.tol2, a.phase_clk, a.stamps,
            a.stamp_stride, a.launch_id);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int CPL, int WPS>
int launch_sg_c(const ColsplitArgs& a, cudaStream_t st) {
    const bool exact = (a.ldot == a.ltot_r) && (a.ltot_r == 32 * CPL * WPS);
    if (a.full16 == 2) return exact ? launch_sg_t<CPL, WPS, 2, true>(a, st) : launch_sg_t<CPL, WPS, 2, false>(a, st);
    if (a.full16) return exact ? launch_sg_t<CPL, WPS, 1, true>(a, st) : launch_sg_t<CPL, WPS, 1, false>(a, st);
    return exact ? launch_sg_t<CPL, WPS, 0, true>(a, st) : launch_sg_t<CPL, WPS, 0, false>(a, st);
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
    if (a.cpl == 0) return a.ltot_r <= 128 ? launch_sg_c<1, 4>(a, st) : launch_sg_c<2, 4>(a, st);     // 4-warp sub-groups
    if (a.cpl == -1) {                                                                               // 2-warp sub-groups
        if (a.ltot_r <= 64) return launch_sg_c<1, 2>(a, st);
        return a.ltot_r <= 128 ? launch_sg_c<2, 2>(a, st) : launch_sg_c<4, 2>(a, st);
    }
    return a.cpl == 2 ? launch_c<2>(a, st) : launch_c<1>(a, st);
}

}  // namespace mpsb
