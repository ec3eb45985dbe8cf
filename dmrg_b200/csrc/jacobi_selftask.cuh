// Self tasks of the modulus ordering: the 28 + 28 pairs INSIDE two 8-row blocks, handled by one 128-thread CTA.
//
// The CTA is two SUB-GROUPS of two warps; sub-group g keeps the 8 rows of block g in registers, its two warps
// covering the columns (CPL complex columns per lane, 64 CPL >= ltot_r).  A tournament step handles 4 disjoint pairs
// together: every lane forms the 4 partial inner products of its columns, ONE transposing butterfly reduces the 8 real
// sums of a warp, the two warps' partials meet in shared memory, the 4 rotation parameters are computed side by side
// by lanes 0-3 of the step's duty warp (SIMD over pairs), and every lane applies the 4 rotations to its columns in
// scaled "fast Givens" form (rows kept as x = X / d, exactly as in jacobi_cross_kernel).  Seven steps: three inside
// the quads (rows 0-3, 4-7), four between them.  The duty warp releases its partner as soon as the parameters are in
// shared memory (bar.arrive) and finishes the bookkeeping (cosine, scales, norms) afterwards.
//
// A column-split kernel of this kind was also built for the cross pairs (16 rows per CTA, 4- and 2-warp sub-groups,
// quad swaps through shared memory): 112 FP64 instructions per pair against 165 in jacobi_cross_kernel, but its
// dependency chains (butterfly -> barrier -> parameters -> barrier -> rotation, ~2000 clocks per step) keep the
// FP64 pipe at 45 %; measured 107.7 ms against 96.3 ms for the Jacobi phase of the benchmark, so only the self
// tasks - which the register-resident cross kernel cannot express - use it.  DESIGN.md section 3 has the numbers.
#pragma once
#include "common.cuh"
#include "jacobi_shared.cuh"

namespace mpsb {

// Sum 8 per-lane doubles over the warp with a transposing butterfly (7 + 2 additions, 18 32-bit shuffles):
// lane l ends up with the warp total of v[(l >> 2) & 7].
__device__ __forceinline__ double reduce8(double (&v)[8], int lane) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int o = 16 >> k, m = 4 >> k;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < m; ++i) {
            const double keep = up ? v[i + m] : v[i];
            const double send = up ? v[i] : v[i + m];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    const double t = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 2);
    return t + __shfl_xor_sync(0xffffffffu, t, 1);
}

// Named barriers of the sub-groups (ids 1-4; 0 is __syncthreads), 64 threads each.
// which: 0 = "partials posted", 1 = "parameters posted".
__device__ __forceinline__ void sg_bar_sync(int sg, int which) {
    if (sg == 0) {
        if (which == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
        else asm volatile("bar.sync 2, 64;" ::: "memory");
    } else {
        if (which == 0) asm volatile("bar.sync 3, 64;" ::: "memory");
        else asm volatile("bar.sync 4, 64;" ::: "memory");
    }
}
__device__ __forceinline__ void sg_bar_arrive(int sg) {       // "parameters posted", without waiting
    if (sg == 0) asm volatile("bar.arrive 2, 64;" ::: "memory");
    else asm volatile("bar.arrive 4, 64;" ::: "memory");
}

// Pair j (0..3) of a step among the 8 rows.  KIND 1: step T (0..2) of the tournaments inside the quads;
// KIND 0: step T (0..3) between the quads - row j meets row 4 + (j + T) mod 4.
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

struct __align__(16) SelfSmem {
    double part[2][2][8];   // [sub-group][warp][value]: partial sums of the current step (4 complex inner products)
    double prm[2][4][4];    // [sub-group][pair]: tau_p (re, im), sig_q (re, im)
    double a[2][8];         // true squared norms of the rows (dot columns)
    double d[2][8];         // scales: X = d * x
    double e[2][8];         // d^2
    int mask[2];            // bit j: pair j of the sub-group's current step rotates
    int dirty[2];           // bit r: row r of the block was rotated
    int nrot[2];
};

template <int CPL, bool EXACT, int KIND, int T>
__device__ __forceinline__ void sg_step(cplx (&x)[8][CPL], const bool (&indot)[CPL], SelfSmem& sm, int sg, int wsg,
                                        int lane, int duty, double tol2, bool real) {
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
    sg_bar_sync(sg, 0);
    // ---- B. rotation parameters: lane j < 4 of the duty warp handles pair j ----------------------------------------
    if (wsg == duty) {
        int rotate = 0, p = 0, q = 0;
        double ac2 = 0.0, inv = 0.0, kk = 0.0, a = 0.0, b = 0.0, ep = 0.0, eq = 0.0;
        if (lane < 4) {
            sg_pair<KIND, T>(lane, p, q);
            const double2 s0 = *reinterpret_cast<const double2*>(&sm.part[sg][0][2 * lane]);
            const double2 s1 = *reinterpret_cast<const double2*>(&sm.part[sg][1][2 * lane]);
            a = sm.a[sg][p]; b = sm.a[sg][q]; ep = sm.e[sg][p]; eq = sm.e[sg][q];
            const double crs = s0.x + s1.x, cis = real ? 0.0 : s0.y + s1.y;   // real: packed real rows (rotate_pair)
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
            *reinterpret_cast<double2*>(&sm.prm[sg][lane][0]) = make_double2(tpr, tpi);     // zeros: exact identity
            *reinterpret_cast<double2*>(&sm.prm[sg][lane][2]) = make_double2(sqr, sqi);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, rotate) & 0xfu;
        if (lane == 0) sm.mask[sg] = (int)bal;
        __syncwarp();
        sg_bar_arrive(sg);                           // the partner warp may rotate now
        if (rotate) {                                // bookkeeping, off the partner's critical path
            const double wv = fma(inv * inv, ac2, 1.0);
            const double cs = newton_rsqrt(wv, newton_rsqrt(wv, rsqrt_approx(wv)));
            const double cs2 = cs * cs;
            const double delta = kk * ac2;                               // t |c|
            sm.a[sg][p] = a - delta;
            sm.a[sg][q] = b + delta;
            sm.d[sg][p] *= cs;
            sm.d[sg][q] *= cs;
            sm.e[sg][p] = ep * cs2;
            sm.e[sg][q] = eq * cs2;
        }
        unsigned bits = rotate ? ((1u << p) | (1u << q)) : 0u;
        bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
        bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
        if (lane == 0 && bal) {
            sm.dirty[sg] |= (int)bits;
            sm.nrot[sg] += __popc(bal);
        }
    } else {
        sg_bar_sync(sg, 1);
    }
    // ---- C. apply the rotations to this lane's columns ----------------------------------------------------------------
    if (sm.mask[sg]) {
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
    // part[] is next written after this step's second barrier, prm[] / mask after the next step's first barrier,
    // which both warps reach only after finishing C: no barrier needed here
}

// One CTA of 128 threads; blocks bA and bB of problem `mat`.  `unit` = stamp slot of the task.
template <int CPL, bool EXACT>
__device__ __forceinline__ void selftask_body(cplx* __restrict__ X, int ld, int ldot, int ltot_r, int bA, int bB,
                                              int* __restrict__ rot_mat, double tol2, int* __restrict__ stp, int unit,
                                              int launch_id, bool real) {
    __shared__ SelfSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, sg = w >> 1, wsg = w & 1;
    if (stp) {                                               // verified clean since both blocks last changed
        const int chk = stp[unit];
        if (chk > stp[bA] && chk > stp[bB]) return;
    }
    const int c0 = wsg * (32 * CPL) + lane;
    bool live[CPL], indot[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        live[k] = EXACT || (c0 + 32 * k < ltot_r);
        indot[k] = EXACT || (c0 + 32 * k < ldot);
    }
    cplx* rows = X + (size_t)((sg == 0 ? bA : bB) * 8) * ld + c0;
    cplx x[8][CPL];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < CPL; ++k) x[r][k] = live[k] ? __ldcg(rows + (size_t)r * ld + 32 * k) : make_double2(0.0, 0.0);
    if (tid < 2) { sm.dirty[tid] = 0; sm.nrot[tid] = 0; }
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
        sm.a[sg][lane] = sm.part[sg][0][lane] + sm.part[sg][1][lane];
        sm.d[sg][lane] = 1.0;
        sm.e[sg][lane] = 1.0;
    }
    __syncthreads();        // part[] is rewritten by the first step

    sg_step<CPL, EXACT, 1, 0>(x, indot, sm, sg, wsg, lane, 0, tol2, real);
    sg_step<CPL, EXACT, 1, 1>(x, indot, sm, sg, wsg, lane, 1, tol2, real);
    sg_step<CPL, EXACT, 1, 2>(x, indot, sm, sg, wsg, lane, 0, tol2, real);
    sg_step<CPL, EXACT, 0, 0>(x, indot, sm, sg, wsg, lane, 1, tol2, real);
    sg_step<CPL, EXACT, 0, 1>(x, indot, sm, sg, wsg, lane, 0, tol2, real);
    sg_step<CPL, EXACT, 0, 2>(x, indot, sm, sg, wsg, lane, 1, tol2, real);
    sg_step<CPL, EXACT, 0, 3>(x, indot, sm, sg, wsg, lane, 0, tol2, real);
    __syncthreads();         // all bookkeeping of both sub-groups is in shared memory

    const int dirty = sm.dirty[sg];
    if (!(sm.dirty[0] | sm.dirty[1])) {
        if (tid == 0 && stp) stp[unit] = launch_id;         // verified clean at this launch
        return;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {                            // fold the scales back into the rotated rows
        if (dirty & (1 << r)) {
            const double dd = sm.d[sg][r];
#pragma unroll
            for (int k = 0; k < CPL; ++k)
                if (live[k]) __stcg(rows + (size_t)r * ld + 32 * k, make_double2(x[r][k].x * dd, x[r][k].y * dd));
        }
    }
    if (tid == 0) {
        atomicAdd(rot_mat, sm.nrot[0] + sm.nrot[1]);
        if (stp) {
            if (sm.dirty[0]) stp[bA] = launch_id;
            if (sm.dirty[1]) stp[bB] = launch_id;
        }
    }
}

}  // namespace mpsb
