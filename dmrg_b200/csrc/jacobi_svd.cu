// Batched truncated SVD for the two-site update: Householder-QR preconditioning followed by a
// parallel-ordered one-sided (Hestenes) Jacobi iteration on the ROWS of the work matrix.
//
// Replaces  scipy.linalg.svd + truncate  in the reference:
//   svd_truncate            DMRG/MPS.py:41-45   (update_double :448, ortho_*_site :260,:280)
//   truncate                DMRG/MPS.py:14-38
//   halve                   DMRG/iDMRG.py:17-35
//
// Why rows: left rotations Q^H m = diag(s) Vh leave the singular values in the row norms and the
// right singular vectors in the normalised rows, which is all update_double needs (its `u` is never
// used, DMRG/MPS.py:451).  When U or a gauge carry is wanted the caller appends columns to the
// work matrix (identity, or the neighbouring site tensor); they receive the same rotations.
//
// Pipeline per call:  QR preconditioning (columns in descending-norm order; makes the row Gram matrix
// scaled-diagonally dominant so Jacobi needs ~11 sweeps even for exponentially decaying Schmidt spectra instead
// of 30+)  ->  Jacobi sweeps until a sweep applies no rotation  ->  finalise_kernel (row norms, bitonic sort,
// reference truncation rule, coalesced scatter into the MPS tensors).
//
// Which kernels run is decided by the plan (svd_make_plan) from the batch size and the shape:
//   QR      qr_panel_dmma_kernel   large batches: columns in pivot order, panels of 8, trailing update on DMMA;
//           qr_blocked_kernel      its predecessor (one thread per column), kept for shapes whose row buffer does not fit
//           qr_precond_kernel      batch <= 148: one SM per problem, 1024 threads (half the latency)
//           qr_grid_kernel         <= 4 large problems: columns of each dealt out to up to 148 CTAs (cooperative)
//   Jacobi  jacobi_cross_kernel    rounds >= 1, rows of <= 256 columns: register-resident rows, TMA-streamed partners
//           jacobi_gram_kernel     round 0 and rows of 257-730 columns: Gram block + rotations + update on DMMA
//           jacobi_gram_tiled_kernel   longer rows: the same through two 256-column tile buffers
//           jacobi_resident_kernel     tiny problems in small batches: whole matrix in shared memory, one launch
//           jacobi_round_kernel<B>     generic fallback (any block size, rows that fit shared memory)
//
// Data movement: rows travel between global and shared memory with the TMA engine's 1-D bulk copy
// (cp.async.bulk, SASS UBLKCP) completing on mbarriers.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include "common.cuh"
#include "mpsb_internal.h"
#include "jacobi_shared.cuh"
#include "jacobi_selftask.cuh"

namespace mpsb {

namespace {

constexpr int QR_THREADS = 1024;
constexpr int FIN_THREADS = 256;

// Householder columns with |x|^2 at or below this are treated as zero columns: tau = 2 / (v^H v) would overflow for
// a subnormal norm (LAPACK's zlarfg rescales instead; here such a column is 1e-145 below a normalised state).
constexpr double QR_TINY = 1e-290;
constexpr double ROW_DEAD = 1e-125;    // tol^2 * ROW_DEAD^2 stays well above PAIR_FLOOR: every pair of live rows is tested relatively

struct SvdState {           // device scratch, carved out of the caller's workspace
    int* rot;               // [batch] rotations applied in the current sweep
    int* done;              // [batch] 1 once a sweep applied no rotation
    double* frob2;          // [batch] squared Frobenius norm of the dot part
    int* n_active;          // [1]
    unsigned long long* rot_total;      // [1] rotations applied over the whole call
    unsigned long long* sweep_total;    // [1] sum over problems of sweeps executed
    unsigned long long* phase_clk;      // [4] debug: summed SM clocks of load / rotate / store phases + CTA count
    int* stamps;            // [batch][stamp_stride] change stamps of the row blocks, then "verified clean" stamps of
    int stamp_stride;       //  the (round, slot) block pairs; 0 ints per problem when the plan does not use them
};

// A block pair that was found clean (no rotation needed) after the last change of either block is still clean:
// its CTA can leave before loading a single row.  Stamps are launch numbers (monotone within one SVD call).
__host__ __device__ inline int stamp_ints(int n_pad, int bsz) {
    const int nblk = n_pad / bsz;
    return (bsz == 8 && nblk >= 4 && nblk <= 512) ? nblk + nblk * nblk + nblk / 2 : 0;
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

SvdState carve_state(void* state, int batch, int stamp_stride) {
    unsigned char* p = static_cast<unsigned char*>(state);
    SvdState s;
    s.stamp_stride = stamp_stride;
    s.rot = reinterpret_cast<int*>(p);            p += align_up(sizeof(int) * batch, 256);
    s.done = reinterpret_cast<int*>(p);           p += align_up(sizeof(int) * batch, 256);
    s.frob2 = reinterpret_cast<double*>(p);       p += align_up(sizeof(double) * batch, 256);
    s.n_active = reinterpret_cast<int*>(p);
    s.rot_total = reinterpret_cast<unsigned long long*>(p + 64);
    s.sweep_total = reinterpret_cast<unsigned long long*>(p + 128);
    s.phase_clk = reinterpret_cast<unsigned long long*>(p + 192);
    s.stamps = stamp_stride ? reinterpret_cast<int*>(p + 256) : nullptr;
    return s;
}

// ---- block reductions ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum over the CTA; every thread receives the same value.  `red` holds >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// ---- Householder QR, one CTA per problem ----------------------------------------------------------
// Reflectors are built from the dot columns in descending-norm order (a static stand-in for column
// pivoting) and applied to all ltot columns.  Thread (cg, cj) owns rows i == cg (mod G) of column
// cj (+ strips of CW columns), so global accesses are coalesced along a row.
__global__ void __launch_bounds__(QR_THREADS, 1)
qr_precond_kernel(cplx* __restrict__ Xall, long long strideX, int n, int ldot, int ltot, int ld, int G,
                  int CW, double* __restrict__ frob2) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const size_t nal = align_up(n, 8);
    cplx* vbuf = reinterpret_cast<cplx*>(smraw);                    // [2][nal]   reflector of this / the pending step
    cplx* wbuf = vbuf + 2 * nal;                                    // [2][G][ld] partial w = v^H X per row group
    cplx* wpart = wbuf;                                             // scratch for the column norms below
    double* cn = reinterpret_cast<double*>(wbuf + (size_t)2 * G * ld);   // [ldot]
    int* order = reinterpret_cast<int*>(cn + align_up(ldot, 8));    // [ldot]
    unsigned char* colDone = reinterpret_cast<unsigned char*>(order + align_up(ldot, 8));  // [ld]
    __shared__ double red[33];

    const int tid = threadIdx.x, NT = blockDim.x;
    cplx* X = Xall + (size_t)blockIdx.x * strideX;
    const int cg = tid / CW, cj = tid % CW;
    const bool active = cg < G;
    const int nstrips = (ltot + CW - 1) / CW;

    // column norms of the dot part
    double* cpart = reinterpret_cast<double*>(wpart);               // [G][ld] doubles
    for (int s = 0; s < nstrips; ++s) {
        const int j = s * CW + cj;
        if (active && j < ldot) {
            double acc = 0.0;
#pragma unroll 4
            for (int i = cg; i < n; i += G) {
                const cplx x = X[(size_t)i * ld + j];
                acc = fma(x.x, x.x, fma(x.y, x.y, acc));
            }
            cpart[cg * ld + j] = acc;
        }
    }
    for (int j = tid; j < ld; j += NT) colDone[j] = 0;
    __syncthreads();
    double fl = 0.0;
    for (int j = tid; j < ldot; j += NT) {
        double acc = 0.0;
        for (int g = 0; g < G; ++g) acc += cpart[g * ld + j];
        cn[j] = acc;
        fl += acc;
    }
    const double F2 = block_sum(fl, red);
    if (tid == 0) frob2[blockIdx.x] = F2;
    // rank sort, descending, ties by index
    for (int j = tid; j < ldot; j += NT) {
        const double me = cn[j];
        int r = 0;
        for (int q = 0; q < ldot; ++q) {
            const double o = cn[q];
            r += (o > me) || (o == me && q < j);
        }
        order[r] = j;
    }
    __syncthreads();

    // Householder steps with the rank-1 update of step k-1 fused into the pass that forms w_k = v_k^H X:
    // every trailing element is read and written once per step instead of read twice and written once.
    // Pending reflector: (vbuf[pv], tau_prev, wbuf[pv]) acting on rows >= kprev.
    const int nsteps = n < ldot ? n : ldot;
    bool have_prev = false;
    int kprev = 0, pv = 0;
    double tau_prev = 0.0;
    for (int k = 0; k < nsteps; ++k) {
        const int pc = order[k];
        const int cu = pv ^ 1;                       // buffers written in this step
        cplx* vcur = vbuf + (size_t)cu * nal;
        const cplx* vprev = vbuf + (size_t)pv * nal;
        const cplx* wprev = wbuf + (size_t)pv * G * ld;
        cplx* wcur = wbuf + (size_t)cu * G * ld;
        // A. pivot column with the pending update applied
        {
            cplx wp = make_double2(0.0, 0.0);
            if (have_prev) {
                for (int g = 0; g < G; ++g) { wp.x += wprev[(size_t)g * ld + pc].x; wp.y += wprev[(size_t)g * ld + pc].y; }
                wp.x *= tau_prev;
                wp.y *= tau_prev;
            }
            const int lo = have_prev ? kprev : k;
            for (int i = lo + tid; i < n; i += NT) {
                cplx x = X[(size_t)i * ld + pc];
                if (have_prev) {
                    const cplx vi = vprev[i];
                    x.x -= vi.x * wp.x - vi.y * wp.y;
                    x.y -= vi.x * wp.y + vi.y * wp.x;
                }
                if (i < k) X[(size_t)i * ld + pc] = x;      // finished entries of R above the diagonal
                else vcur[i] = x;
            }
        }
        __syncthreads();
        double loc = 0.0;
        for (int i = k + 1 + tid; i < n; i += NT) {
            const cplx x = vcur[i];
            loc = fma(x.x, x.x, fma(x.y, x.y, loc));
        }
        const cplx alpha = vcur[k];
        const double sigma = block_sum(loc, red);          // contains the barriers that order vcur[k] reads
        const double a2 = alpha.x * alpha.x + alpha.y * alpha.y;
        const double normx2 = a2 + sigma;
        if (!(normx2 > QR_TINY)) {                              // zero column: no reflector, the pending one stays pending
            for (int i = k + tid; i < n; i += NT) X[(size_t)i * ld + pc] = make_double2(0.0, 0.0);
            if (tid == 0) colDone[pc] = 1;
            __syncthreads();
            continue;
        }
        const double normx = sqrt(normx2), absa = sqrt(a2);
        cplx phase = make_double2(1.0, 0.0);
        if (absa > 0.0) phase = make_double2(alpha.x / absa, alpha.y / absa);
        const cplx beta = make_double2(-phase.x * normx, -phase.y * normx);
        const cplx vk = make_double2(phase.x * (absa + normx), phase.y * (absa + normx));
        const double vhv = (absa + normx) * (absa + normx) + sigma;
        const double tau = 2.0 / vhv;
        if (tid == 0) vcur[k] = vk;
        __syncthreads();
        // C. fused pass: apply the pending reflector, accumulate w_k = sum_{i >= k} conj(v_k[i]) X[i][j]
        const int lo = have_prev ? kprev : k;
        for (int s = 0; s < nstrips; ++s) {
            const int j = s * CW + cj;
            if (active && j < ltot && j != pc && !(j < ldot && colDone[j])) {
                double pr = 0.0, pi = 0.0;
                if (have_prev) {
                    for (int g = 0; g < G; ++g) { pr += wprev[(size_t)g * ld + j].x; pi += wprev[(size_t)g * ld + j].y; }
                    pr *= tau_prev;
                    pi *= tau_prev;
                }
                double wr = 0.0, wi = 0.0;
                int i = lo + cg;
                for (; i < k; i += G) {                       // rows finished by the pending reflector only
                    const cplx vi = vprev[i];
                    cplx x = X[(size_t)i * ld + j];
                    x.x -= vi.x * pr - vi.y * pi;
                    x.y -= vi.x * pi + vi.y * pr;
                    X[(size_t)i * ld + j] = x;
                }
                if (have_prev) {
#pragma unroll 8
                    for (; i < n; i += G) {
                        const cplx vi = vprev[i], vc = vcur[i];
                        cplx x = X[(size_t)i * ld + j];
                        x.x -= vi.x * pr - vi.y * pi;
                        x.y -= vi.x * pi + vi.y * pr;
                        X[(size_t)i * ld + j] = x;
                        wr = fma(vc.x, x.x, fma(vc.y, x.y, wr));
                        wi = fma(vc.x, x.y, fma(-vc.y, x.x, wi));
                    }
                } else {
#pragma unroll 8
                    for (; i < n; i += G) {
                        const cplx vc = vcur[i];
                        const cplx x = X[(size_t)i * ld + j];
                        wr = fma(vc.x, x.x, fma(vc.y, x.y, wr));
                        wi = fma(vc.x, x.y, fma(-vc.y, x.x, wi));
                    }
                }
                wcur[(size_t)cg * ld + j] = make_double2(wr, wi);
            }
        }
        // D. the pivot column itself: beta on the diagonal, zeros below
        for (int i = k + tid; i < n; i += NT)
            X[(size_t)i * ld + pc] = (i == k) ? beta : make_double2(0.0, 0.0);
        if (tid == 0) colDone[pc] = 1;
        have_prev = true;
        kprev = k;
        tau_prev = tau;
        pv = cu;
        __syncthreads();
    }
    if (have_prev) {   // last pending update
        const cplx* vprev = vbuf + (size_t)pv * nal;
        const cplx* wprev = wbuf + (size_t)pv * G * ld;
        for (int s = 0; s < nstrips; ++s) {
            const int j = s * CW + cj;
            if (active && j < ltot && !(j < ldot && colDone[j])) {
                double pr = 0.0, pi = 0.0;
                for (int g = 0; g < G; ++g) { pr += wprev[(size_t)g * ld + j].x; pi += wprev[(size_t)g * ld + j].y; }
                pr *= tau_prev;
                pi *= tau_prev;
#pragma unroll 8
                for (int i = kprev + cg; i < n; i += G) {
                    const cplx vi = vprev[i];
                    cplx x = X[(size_t)i * ld + j];
                    x.x -= vi.x * pr - vi.y * pi;
                    x.y -= vi.x * pi + vi.y * pr;
                    X[(size_t)i * ld + j] = x;
                }
            }
        }
    }
}

// ---- Householder QR of ONE large problem spread over the whole GPU (cooperative launch) -------------------------
// The single-CTA kernels above take 0.2 s (1024 x 1024) to 1.3 s (2048 x 2048) when only one problem is in flight
// (iDMRG at chi >= 256): every reflector makes one SM stream the whole trailing matrix.  Here the COLUMNS of the
// problem are dealt out to up to 148 CTAs.  Per step the CTA that owns the pivot column finishes that column, forms
// the reflector and publishes it in global memory; after a grid barrier every CTA applies the previous (pending)
// reflector to its own columns and accumulates w = v^H x for the new one in the same pass, exactly like
// qr_precond_kernel, so the only inter-CTA traffic is one vector and one barrier per step.  Same pivot order
// (descending column norm, ties by index) and the same reflectors as the single-CTA kernels.
constexpr int QG_THREADS = 512;
struct QrGridMeta { double tau; int valid; int pad; };

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned seen = 0, it = 0;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen < target && ++it > (1u << 27)) __trap();      // a lost CTA must fail loudly, not hang the GPU
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

// grid = (ncta, batch), launched cooperatively.  scratch per problem: counter (256 B) | meta[2] | colnorm[ld] | v[2][nal]
__global__ void __launch_bounds__(QG_THREADS, 1)
qr_grid_kernel(cplx* __restrict__ Xall, long long strideX, int n, int ldot, int ltot, int ld, int Lc, int TC,
               double* __restrict__ frob2, unsigned char* __restrict__ scratch, size_t scratch_stride) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int nal = (int)align_up(n, 8);
    cplx* vbuf = reinterpret_cast<cplx*>(smraw);                              // [2][nal]
    cplx* wbuf = vbuf + 2 * (size_t)nal;                                      // [2][Lc]   w of the pending / new reflector
    cplx* wpart = wbuf + 2 * (size_t)Lc;                                      // [RG][Lc]
    const int RG = QG_THREADS / TC;
    double* cn = reinterpret_cast<double*>(wpart + (size_t)RG * Lc);          // [ldot]
    int* order = reinterpret_cast<int*>(cn + align_up(ldot, 8));              // [ldot]
    unsigned char* colDone = reinterpret_cast<unsigned char*>(order + align_up(ldot, 8));   // [Lc] own columns only
    __shared__ double red[33];

    const int tid = threadIdx.x, prob = blockIdx.y, ncta = gridDim.x;
    cplx* X = Xall + (size_t)prob * strideX;
    unsigned char* sc = scratch + (size_t)prob * scratch_stride;
    unsigned* counter = reinterpret_cast<unsigned*>(sc);
    QrGridMeta* meta = reinterpret_cast<QrGridMeta*>(sc + 256);
    double* colnorm = reinterpret_cast<double*>(sc + 512);
    cplx* vglob = reinterpret_cast<cplx*>(sc + 512 + align_up((size_t)ld * sizeof(double), 256));   // [2][nal]
    unsigned epoch = 0;

    const int c_lo = blockIdx.x * Lc, c_hi = min(ltot, c_lo + Lc);
    const int rg = tid / TC, cj = tid % TC;

    // column norms of the own dot columns, Frobenius norm, pivot order (identical in every CTA)
    for (int jj = cj; c_lo + jj < c_hi; jj += TC) {
        const int j = c_lo + jj;
        double acc = 0.0;
        if (j < ldot)
            for (int i = rg; i < n; i += RG) {
                const cplx x = X[(size_t)i * ld + j];
                acc = fma(x.x, x.x, fma(x.y, x.y, acc));
            }
        reinterpret_cast<double*>(wpart)[rg * Lc + jj] = acc;
    }
    for (int jj = tid; jj < Lc; jj += QG_THREADS) colDone[jj] = 0;
    __syncthreads();
    for (int jj = tid; c_lo + jj < c_hi; jj += QG_THREADS) {
        double acc = 0.0;
        for (int g = 0; g < RG; ++g) acc += reinterpret_cast<double*>(wpart)[g * Lc + jj];
        if (c_lo + jj < ldot) colnorm[c_lo + jj] = acc;
    }
    grid_barrier(counter, ++epoch * ncta);
    double fl = 0.0;
    for (int j = tid; j < ldot; j += QG_THREADS) {
        double v;
        asm volatile("ld.relaxed.gpu.f64 %0, [%1];" : "=d"(v) : "l"(colnorm + j) : "memory");
        cn[j] = v;
        fl += v;
    }
    const double F2 = block_sum(fl, red);
    if (blockIdx.x == 0 && tid == 0) frob2[prob] = F2;
    for (int j = tid; j < ldot; j += QG_THREADS) {
        const double me = cn[j];
        int r = 0;
        for (int q = 0; q < ldot; ++q) {
            const double o = cn[q];
            r += (o > me) || (o == me && q < j);
        }
        order[r] = j;
    }
    __syncthreads();

    const int nsteps = n < ldot ? n : ldot;
    bool have_prev = false;
    int kprev = 0, pv = 0;
    double tau_prev = 0.0;
    for (int k = 0; k < nsteps; ++k) {
        const int pc = order[k];
        const int cu = pv ^ 1;
        cplx* vcur = vbuf + (size_t)cu * nal;
        const cplx* vprev = vbuf + (size_t)pv * nal;
        const cplx* wprev = wbuf + (size_t)pv * Lc;
        cplx* wcur = wbuf + (size_t)cu * Lc;
        cplx* vg = vglob + (size_t)cu * nal;
        const bool owner = pc >= c_lo && pc < c_hi;
        if (owner) {
            // A. pivot column with the pending update applied
            cplx wp = make_double2(0.0, 0.0);
            if (have_prev) { wp = wprev[pc - c_lo]; wp.x *= tau_prev; wp.y *= tau_prev; }
            const int lo = have_prev ? kprev : k;
            for (int i = lo + tid; i < n; i += QG_THREADS) {
                cplx x = X[(size_t)i * ld + pc];
                if (have_prev) {
                    const cplx vi = vprev[i];
                    x.x -= vi.x * wp.x - vi.y * wp.y;
                    x.y -= vi.x * wp.y + vi.y * wp.x;
                }
                if (i < k) X[(size_t)i * ld + pc] = x;
                else vcur[i] = x;
            }
            __syncthreads();
            double loc = 0.0;
            for (int i = k + 1 + tid; i < n; i += QG_THREADS) {
                const cplx x = vcur[i];
                loc = fma(x.x, x.x, fma(x.y, x.y, loc));
            }
            const cplx alpha = vcur[k];
            const double sigma = block_sum(loc, red);
            const double a2 = alpha.x * alpha.x + alpha.y * alpha.y;
            const double normx2 = a2 + sigma;
            if (!(normx2 > QR_TINY)) {
                for (int i = k + tid; i < n; i += QG_THREADS) X[(size_t)i * ld + pc] = make_double2(0.0, 0.0);
                if (tid == 0) { meta[cu].tau = 0.0; meta[cu].valid = 0; colDone[pc - c_lo] = 1; }
            } else {
                const double normx = sqrt(normx2), absa = sqrt(a2);
                cplx phase = make_double2(1.0, 0.0);
                if (absa > 0.0) phase = make_double2(alpha.x / absa, alpha.y / absa);
                const cplx beta = make_double2(-phase.x * normx, -phase.y * normx);
                const cplx vk = make_double2(phase.x * (absa + normx), phase.y * (absa + normx));
                const double vhv = (absa + normx) * (absa + normx) + sigma;
                if (tid == 0) { vcur[k] = vk; meta[cu].tau = 2.0 / vhv; meta[cu].valid = 1; colDone[pc - c_lo] = 1; }
                __syncthreads();
                for (int i = k + tid; i < n; i += QG_THREADS) {
                    vg[i] = vcur[i];
                    X[(size_t)i * ld + pc] = (i == k) ? beta : make_double2(0.0, 0.0);
                }
            }
        }
        grid_barrier(counter, ++epoch * ncta);
        double tau;
        int valid;
        asm volatile("ld.relaxed.gpu.f64 %0, [%1];" : "=d"(tau) : "l"(&meta[cu].tau) : "memory");
        asm volatile("ld.relaxed.gpu.s32 %0, [%1];" : "=r"(valid) : "l"(&meta[cu].valid) : "memory");
        if (!valid) continue;                               // zero column: the pending reflector stays pending
        if (!owner) {
            for (int i = k + tid; i < n; i += QG_THREADS) {
                cplx v;
                asm volatile("ld.relaxed.gpu.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(vg + i) : "memory");
                vcur[i] = v;
            }
        }
        __syncthreads();
        // C. fused pass over the own columns
        const int lo = have_prev ? kprev : k;
        for (int jj = cj; jj < Lc; jj += TC) {
            const int j = c_lo + jj;
            double wr = 0.0, wi = 0.0;
            if (j < c_hi && j != pc && !(j < ldot && colDone[jj])) {
                double pr = 0.0, pi = 0.0;
                if (have_prev) { pr = wprev[jj].x * tau_prev; pi = wprev[jj].y * tau_prev; }
                int i = lo + rg;
                for (; i < k; i += RG) {
                    const cplx vi = vprev[i];
                    cplx x = X[(size_t)i * ld + j];
                    x.x -= vi.x * pr - vi.y * pi;
                    x.y -= vi.x * pi + vi.y * pr;
                    X[(size_t)i * ld + j] = x;
                }
                if (have_prev) {
#pragma unroll 4
                    for (; i < n; i += RG) {
                        const cplx vi = vprev[i], vc = vcur[i];
                        cplx x = X[(size_t)i * ld + j];
                        x.x -= vi.x * pr - vi.y * pi;
                        x.y -= vi.x * pi + vi.y * pr;
                        X[(size_t)i * ld + j] = x;
                        wr = fma(vc.x, x.x, fma(vc.y, x.y, wr));
                        wi = fma(vc.x, x.y, fma(-vc.y, x.x, wi));
                    }
                } else {
#pragma unroll 4
                    for (; i < n; i += RG) {
                        const cplx vc = vcur[i];
                        const cplx x = X[(size_t)i * ld + j];
                        wr = fma(vc.x, x.x, fma(vc.y, x.y, wr));
                        wi = fma(vc.x, x.y, fma(-vc.y, x.x, wi));
                    }
                }
            }
            wpart[(size_t)rg * Lc + jj] = make_double2(wr, wi);
        }
        __syncthreads();
        for (int jj = tid; jj < Lc; jj += QG_THREADS) {
            double sr = 0.0, si = 0.0;
            for (int g = 0; g < RG; ++g) { sr += wpart[(size_t)g * Lc + jj].x; si += wpart[(size_t)g * Lc + jj].y; }
            wcur[jj] = make_double2(sr, si);
        }
        have_prev = true;
        kprev = k;
        tau_prev = tau;
        pv = cu;
        __syncthreads();
    }
    if (have_prev) {   // last pending update
        const cplx* vprev = vbuf + (size_t)pv * nal;
        const cplx* wprev = wbuf + (size_t)pv * Lc;
        for (int jj = cj; jj < Lc; jj += TC) {
            const int j = c_lo + jj;
            if (j < c_hi && !(j < ldot && colDone[jj])) {
                const double pr = wprev[jj].x * tau_prev, pi = wprev[jj].y * tau_prev;
#pragma unroll 4
                for (int i = kprev + rg; i < n; i += RG) {
                    const cplx vi = vprev[i];
                    cplx x = X[(size_t)i * ld + j];
                    x.x -= vi.x * pr - vi.y * pi;
                    x.y -= vi.x * pi + vi.y * pr;
                    X[(size_t)i * ld + j] = x;
                }
            }
        }
    }
}

// ---- blocked Householder QR (compact WY, panels of 8 columns) -------------------------------------------
// Same transformation as qr_precond_kernel, but the trailing matrix is touched once per PANEL:
//   1. gather the NB pivot columns of the panel into shared memory and factor them there (Householder steps with
//      one vector block-reduction each: the sums conj(p_b) . p_c do not depend on beta, so norm and all dot
//      products of a step come out of a single reduction);
//   2. build the upper-triangular T of  H_0 ... H_{nb-1} = I - V T V^H  from V^H V;
//   3. trailing update, one thread per column:  W = V^H x,  Z = T^H W,  x -= V Z   (V broadcast from smem).
// L2 traffic per problem drops from ~120 MB (fused rank-1 updates) to ~3/(2 NB) of that; 256 threads and ~40 KB
// of shared memory per problem let ~5 problems share an SM.
constexpr int QB_THREADS = 256;
constexpr int QB_NB = 8;

// Sum NV doubles over the CTA; every thread receives the sums.  scratch: [QB_THREADS/32][NV] + [NV] doubles.
template <int NV>
__device__ __forceinline__ void block_sum_vec(double (&v)[NV], int nv, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = QB_THREADS / 32;
#pragma unroll
    for (int q = 0; q < NV; ++q)
        if (q < nv) v[q] = warp_sum(v[q]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q)
            if (q < nv) scratch[warp * NV + q] = v[q];
    __syncthreads();
    if (threadIdx.x < nv) {
        double t = 0.0;
        for (int ww = 0; ww < NW; ++ww) t += scratch[ww * NV + threadIdx.x];
        scratch[NW * NV + threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q)
        if (q < nv) v[q] = scratch[NW * NV + q];
}

__global__ void __launch_bounds__(QB_THREADS, 4)
qr_blocked_kernel(cplx* __restrict__ Xall, long long strideX, int n, int ldot, int ltot, int ld,
                  double* __restrict__ frob2) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int NB = QB_NB, NT = QB_THREADS, NW = QB_THREADS / 32;
    const size_t nal = align_up(n, 8);
    cplx* Vp = reinterpret_cast<cplx*>(smraw);                          // [nal][NB] panel / reflectors
    double* cn = reinterpret_cast<double*>(Vp + nal * NB);              // [ldot]
    int* order = reinterpret_cast<int*>(cn + align_up(ldot, 8));        // [ldot]
    unsigned char* colDone = reinterpret_cast<unsigned char*>(order + align_up(ldot, 8));   // [ld]
    __shared__ double scratch[(NW + 1) * 2 * NB];
    __shared__ cplx Tm[NB][NB];
    __shared__ cplx vdiag[NB], betas[NB];
    __shared__ double taus[NB];

    const int tid = threadIdx.x;
    cplx* X = Xall + (size_t)blockIdx.x * strideX;

    // column norms of the dot part, descending-norm pivot order (ties by index)
    double fl = 0.0;
    for (int j = tid; j < ld; j += NT) colDone[j] = 0;
    for (int j = tid; j < ldot; j += NT) {
        double acc = 0.0;
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const cplx x = X[(size_t)i * ld + j];
            acc = fma(x.x, x.x, fma(x.y, x.y, acc));
        }
        cn[j] = acc;
        fl += acc;
    }
    {
        double v1[1] = {fl};
        block_sum_vec<1>(v1, 1, scratch);
        if (tid == 0) frob2[blockIdx.x] = v1[0];
    }
    for (int j = tid; j < ldot; j += NT) {
        const double me = cn[j];
        int r = 0;
        for (int q = 0; q < ldot; ++q) {
            const double o = cn[q];
            r += (o > me) || (o == me && q < j);
        }
        order[r] = j;
    }
    __syncthreads();

    const int nsteps = n < ldot ? n : ldot;
    for (int k = 0; k < nsteps; k += NB) {
        const int nbp = (nsteps - k) < NB ? (nsteps - k) : NB;
        // 1. gather the panel (rows k..n-1 of the nbp pivot columns)
        for (int i = k + tid; i < n; i += NT)
#pragma unroll
            for (int b = 0; b < NB; ++b)
                Vp[(size_t)i * NB + b] = b < nbp ? X[(size_t)i * ld + order[k + b]] : make_double2(0.0, 0.0);
        __syncthreads();
        // 2. factor the panel in shared memory
        for (int b = 0; b < nbp; ++b) {
            const int kb = k + b;
            // S_c = sum_{i > kb} conj(p_b[i]) p_c[i], c = b..nbp-1   (S_b = sigma, real)
            double acc[2 * NB];
#pragma unroll
            for (int q = 0; q < 2 * NB; ++q) acc[q] = 0.0;
            for (int i = kb + 1 + tid; i < n; i += NT) {
                const cplx pb = Vp[(size_t)i * NB + b];
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c >= b && c < nbp) {
                        const cplx pc = Vp[(size_t)i * NB + c];
                        acc[2 * c] = fma(pb.x, pc.x, fma(pb.y, pc.y, acc[2 * c]));
                        acc[2 * c + 1] = fma(pb.x, pc.y, fma(-pb.y, pc.x, acc[2 * c + 1]));
                    }
            }
            block_sum_vec<2 * NB>(acc, 2 * nbp, scratch);
            const double sigma = scratch[NW * 2 * NB + 2 * b];      // = acc[2 b]; read from smem to keep acc in registers
            const cplx alpha = Vp[(size_t)kb * NB + b];
            const double a2 = alpha.x * alpha.x + alpha.y * alpha.y;
            const double normx2 = a2 + sigma;
            __syncthreads();                                   // everybody has read alpha / row kb
            if (!(normx2 > QR_TINY)) {                             // zero column: identity reflector
                if (tid == 0) { taus[b] = 0.0; vdiag[b] = make_double2(0.0, 0.0); betas[b] = make_double2(0.0, 0.0); }
                __syncthreads();
                continue;
            }
            const double normx = sqrt(normx2), absa = sqrt(a2);
            cplx phase = make_double2(1.0, 0.0);
            if (absa > 0.0) phase = make_double2(alpha.x / absa, alpha.y / absa);
            const cplx beta = make_double2(-phase.x * normx, -phase.y * normx);
            const cplx vk = make_double2(phase.x * (absa + normx), phase.y * (absa + normx));
            const double tau = 2.0 / ((absa + normx) * (absa + normx) + sigma);
            // apply H_b to the later panel columns: p_c -= tau v_b (v_b^H p_c),  v_b^H p_c = conj(vk) p_c[kb] + S_c
            for (int i = kb + tid; i < n; i += NT) {
                const cplx vi = (i == kb) ? vk : Vp[(size_t)i * NB + b];
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c > b && c < nbp) {
                        const cplx top = Vp[(size_t)kb * NB + c];
                        const double wr = tau * (vk.x * top.x + vk.y * top.y + acc[2 * c]);
                        const double wi = tau * (vk.x * top.y - vk.y * top.x + acc[2 * c + 1]);
                        cplx pc = Vp[(size_t)i * NB + c];
                        // row kb of column c is read by every thread above: update it last (after the barrier)
                        if (i != kb) {
                            pc.x -= vi.x * wr - vi.y * wi;
                            pc.y -= vi.x * wi + vi.y * wr;
                            Vp[(size_t)i * NB + c] = pc;
                        }
                    }
            }
            __syncthreads();
            if (tid == 0) {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c > b && c < nbp) {
                        const cplx top = Vp[(size_t)kb * NB + c];
                        const double wr = tau * (vk.x * top.x + vk.y * top.y + acc[2 * c]);
                        const double wi = tau * (vk.x * top.y - vk.y * top.x + acc[2 * c + 1]);
                        Vp[(size_t)kb * NB + c] = make_double2(top.x - (vk.x * wr - vk.y * wi), top.y - (vk.x * wi + vk.y * wr));
                    }
                taus[b] = tau;
                vdiag[b] = vk;
                betas[b] = beta;
            }
            __syncthreads();
        }
        // write the finished panel columns back (R entries above / on the diagonal, zeros below) and turn Vp
        // into V proper: zeros above the diagonal of each reflector, vk on it
        for (int i = k + tid; i < n; i += NT)
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (b < nbp) {
                    const int kb = k + b;
                    cplx* dst = X + (size_t)i * ld + order[k + b];
                    if (i < kb) { *dst = Vp[(size_t)i * NB + b]; Vp[(size_t)i * NB + b] = make_double2(0.0, 0.0); }
                    else if (i == kb) {
                        const bool ident = taus[b] == 0.0;
                        *dst = ident ? Vp[(size_t)i * NB + b] : betas[b];
                        Vp[(size_t)i * NB + b] = vdiag[b];
                    } else {
                        if (taus[b] == 0.0) Vp[(size_t)i * NB + b] = make_double2(0.0, 0.0);   // identity reflector
                        *dst = make_double2(0.0, 0.0);
                    }
                }
        __syncthreads();
        // T: T[b][b] = tau_b ;  T[0:b, b] = -tau_b T[0:b, 0:b] (V[:, 0:b]^H v_b)
        if (tid == 0) Tm[0][0] = make_double2(taus[0], 0.0);
        __syncthreads();
        for (int b = 1; b < nbp; ++b) {
            double acc[2 * NB];
#pragma unroll
            for (int q = 0; q < 2 * NB; ++q) acc[q] = 0.0;
            for (int i = k + b + tid; i < n; i += NT) {
                const cplx vb = Vp[(size_t)i * NB + b];
#pragma unroll
                for (int a = 0; a < NB; ++a)
                    if (a < b) {
                        const cplx va = Vp[(size_t)i * NB + a];
                        acc[2 * a] = fma(va.x, vb.x, fma(va.y, vb.y, acc[2 * a]));            // conj(va) vb
                        acc[2 * a + 1] = fma(va.x, vb.y, fma(-va.y, vb.x, acc[2 * a + 1]));
                    }
            }
            block_sum_vec<2 * NB>(acc, 2 * b, scratch);
            if (tid == 0) {
                const double tb = taus[b];
                for (int a = 0; a < b; ++a) {
                    double re = 0.0, im = 0.0;
                    for (int c = a; c < b; ++c) {                // T[a][c] (upper triangular) times g_c
                        const cplx tt = Tm[a][c];
                        const double gr = scratch[NW * 2 * NB + 2 * c], gi = scratch[NW * 2 * NB + 2 * c + 1];
                        re += tt.x * gr - tt.y * gi;
                        im += tt.x * gi + tt.y * gr;
                    }
                    Tm[a][b] = make_double2(-tb * re, -tb * im);
                }
                Tm[b][b] = make_double2(tb, 0.0);
            }
            __syncthreads();
        }
        if (tid < nbp) colDone[order[k + tid]] = 1;
        __syncthreads();
        // 3. trailing update, one thread per column
        for (int j = tid; j < ltot; j += NT) {
            if (j < ldot && colDone[j]) continue;
            cplx W[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) W[b] = make_double2(0.0, 0.0);
#pragma unroll 4
            for (int i = k; i < n; ++i) {
                const cplx x = X[(size_t)i * ld + j];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const cplx v = Vp[(size_t)i * NB + b];
                    W[b].x = fma(v.x, x.x, fma(v.y, x.y, W[b].x));
                    W[b].y = fma(v.x, x.y, fma(-v.y, x.x, W[b].y));
                }
            }
            cplx Z[NB];                                            // Z = T^H W
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                double re = 0.0, im = 0.0;
#pragma unroll
                for (int a = 0; a < NB; ++a)
                    if (a <= b && b < nbp) {
                        const cplx tt = Tm[a][b];
                        re += tt.x * W[a].x + tt.y * W[a].y;       // conj(T[a][b]) W[a]
                        im += tt.x * W[a].y - tt.y * W[a].x;
                    }
                Z[b] = make_double2(re, im);
            }
#pragma unroll 4
            for (int i = k; i < n; ++i) {
                cplx x = X[(size_t)i * ld + j];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const cplx v = Vp[(size_t)i * NB + b];
                    x.x -= v.x * Z[b].x - v.y * Z[b].y;
                    x.y -= v.x * Z[b].y + v.y * Z[b].x;
                }
                X[(size_t)i * ld + j] = x;
            }
        }
        __syncthreads();
    }
}

// ---- blocked Householder QR with the trailing update on FP64 tensor cores -------------------------------------------
// qr_blocked_kernel above spends its time in step 3: one thread per column needs 8 complex accumulators (W) and 8
// coefficients (Z), which at the 64 registers that 4 CTAs per SM allow spill, and the scattered pivot columns make
// every panel gather a strided read.  Here
//   0. the dot columns are first moved into pivot order IN PLACE (one pass over the matrix), so panels are contiguous
//      column ranges and the trailing matrix is "all columns to the right of the panel"; the inverse permutation is
//      applied at the end, so callers see the same R (up to rounding) in the original column order;
//   1./2. panel factorisation and T exactly as in qr_blocked_kernel (shared-memory panel, row pitch QD_VS = 9 elements
//      so that both fragment patterns below are conflict-free);
//   3. trailing update by DMMA, one warp per group of 8 columns:
//         W = V^H X   (8 x 8 accumulator tile, K = rows, V^H fragments from the shared panel, X fragments from global)
//         Z = T^H W   (8 x 8 x 8, W passes from accumulator to operand layout through a per-warp staging tile)
//         X -= V Z    (per 8-row tile: X is loaded in accumulator layout, 8 DMMAs, stored back)
//      four registers of accumulators instead of 32, no spills, a quarter of the issue slots.
constexpr int QD_VS = 9;               // row pitch of the shared panel in elements (8 columns + 1 pad)
#ifndef MPSB_QD_TILES
#define MPSB_QD_TILES 2
#endif
#ifndef MPSB_QD_WU
#define MPSB_QD_WU 4
#endif
constexpr int QD_TILES = MPSB_QD_TILES;   // 8-row tiles of X in flight per warp in the update pass
constexpr int QD_WU = MPSB_QD_WU;         // 4-row slices of X in flight per warp in the W pass

__global__ void __launch_bounds__(QB_THREADS, 4)
qr_panel_dmma_kernel(cplx* __restrict__ Xall, long long strideX, int n, int ldot, int ltot, int ld,
                     double* __restrict__ frob2) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int NB = QB_NB, NT = QB_THREADS, NW = QB_THREADS / 32, VS = QD_VS;
    const size_t nal = align_up(n, 8);
    cplx* Vp = reinterpret_cast<cplx*>(smraw);                          // [nal][VS] panel / reflectors; row buffer in step 0
    double* cn = reinterpret_cast<double*>(Vp + nal * VS);              // [ldot]
    int* order = reinterpret_cast<int*>(cn + align_up(ldot, 8));        // [ldot]
    __shared__ double scratch[(NW + 1) * 2 * NB];
    __shared__ cplx Tm[NB][NB];
    __shared__ cplx vdiag[NB], betas[NB];
    __shared__ double taus[NB];
    __shared__ cplx stage[NW][8][VS];                                   // per-warp fragment transposes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    cplx* X = Xall + (size_t)blockIdx.x * strideX;

    // column norms of the dot part, descending-norm pivot order (ties by index)
    double fl = 0.0;
    for (int j = tid; j < ldot; j += NT) {
        double acc = 0.0;
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const cplx x = X[(size_t)i * ld + j];
            acc = fma(x.x, x.x, fma(x.y, x.y, acc));
        }
        cn[j] = acc;
        fl += acc;
    }
    {
        double v1[1] = {fl};
        block_sum_vec<1>(v1, 1, scratch);
        if (tid == 0) frob2[blockIdx.x] = v1[0];
    }
    for (int j = tid; j < ldot; j += NT) {
        const double me = cn[j];
        int r = 0;
        for (int q = 0; q < ldot; ++q) {
            const double o = cn[q];
            r += (o > me) || (o == me && q < j);
        }
        order[r] = j;
    }
    __syncthreads();
    // 0. dot columns into pivot order: X[i][j] <- X[i][order[j]], ROWS_PER rows at a time through the panel buffer
    const int rows_per = (int)((nal * VS) / (size_t)ldot);             // >= 1 (checked by the launcher)
    for (int i0 = 0; i0 < n; i0 += rows_per) {
        const int nr = min(rows_per, n - i0);
        for (int e = tid; e < nr * ldot; e += NT) {
            const int r = e / ldot, j = e - r * ldot;
            Vp[e] = X[(size_t)(i0 + r) * ld + order[j]];
        }
        __syncthreads();
        for (int e = tid; e < nr * ldot; e += NT) {
            const int r = e / ldot, j = e - r * ldot;
            X[(size_t)(i0 + r) * ld + j] = Vp[e];
        }
        __syncthreads();
    }

    const int nsteps = n < ldot ? n : ldot;
    for (int k = 0; k < nsteps; k += NB) {
        const int nbp = (nsteps - k) < NB ? (nsteps - k) : NB;
        // 1. the panel: rows k..nal-1 of columns k..k+nbp-1 (rows >= n and columns >= nbp are zero)
        for (int e = tid; e < (int)(nal - k) * NB; e += NT) {
            const int i = k + e / NB, b = e % NB;
            Vp[(size_t)i * VS + b] = (i < n && b < nbp) ? X[(size_t)i * ld + k + b] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        // 2. factor the panel in shared memory
        for (int b = 0; b < nbp; ++b) {
            const int kb = k + b;
            double acc[2 * NB];
#pragma unroll
            for (int q = 0; q < 2 * NB; ++q) acc[q] = 0.0;
            for (int i = kb + 1 + tid; i < n; i += NT) {
                const cplx pb = Vp[(size_t)i * VS + b];
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c >= b && c < nbp) {
                        const cplx pc = Vp[(size_t)i * VS + c];
                        acc[2 * c] = fma(pb.x, pc.x, fma(pb.y, pc.y, acc[2 * c]));
                        acc[2 * c + 1] = fma(pb.x, pc.y, fma(-pb.y, pc.x, acc[2 * c + 1]));
                    }
            }
            block_sum_vec<2 * NB>(acc, 2 * nbp, scratch);
            const double sigma = scratch[NW * 2 * NB + 2 * b];
            const cplx alpha = Vp[(size_t)kb * VS + b];
            const double a2 = alpha.x * alpha.x + alpha.y * alpha.y;
            const double normx2 = a2 + sigma;
            __syncthreads();                                   // everybody has read alpha / row kb
            if (!(normx2 > QR_TINY)) {                             // zero column: identity reflector
                if (tid == 0) { taus[b] = 0.0; vdiag[b] = make_double2(0.0, 0.0); betas[b] = make_double2(0.0, 0.0); }
                __syncthreads();
                continue;
            }
            const double normx = sqrt(normx2), absa = sqrt(a2);
            cplx phase = make_double2(1.0, 0.0);
            if (absa > 0.0) phase = make_double2(alpha.x / absa, alpha.y / absa);
            const cplx beta = make_double2(-phase.x * normx, -phase.y * normx);
            const cplx vk = make_double2(phase.x * (absa + normx), phase.y * (absa + normx));
            const double tau = 2.0 / ((absa + normx) * (absa + normx) + sigma);
            for (int i = kb + tid; i < n; i += NT) {
                const cplx vi = (i == kb) ? vk : Vp[(size_t)i * VS + b];
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c > b && c < nbp) {
                        const cplx top = Vp[(size_t)kb * VS + c];
                        const double wr = tau * (vk.x * top.x + vk.y * top.y + acc[2 * c]);
                        const double wi = tau * (vk.x * top.y - vk.y * top.x + acc[2 * c + 1]);
                        cplx pc = Vp[(size_t)i * VS + c];
                        if (i != kb) {
                            pc.x -= vi.x * wr - vi.y * wi;
                            pc.y -= vi.x * wi + vi.y * wr;
                            Vp[(size_t)i * VS + c] = pc;
                        }
                    }
            }
            __syncthreads();
            if (tid == 0) {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c > b && c < nbp) {
                        const cplx top = Vp[(size_t)kb * VS + c];
                        const double wr = tau * (vk.x * top.x + vk.y * top.y + acc[2 * c]);
                        const double wi = tau * (vk.x * top.y - vk.y * top.x + acc[2 * c + 1]);
                        Vp[(size_t)kb * VS + c] = make_double2(top.x - (vk.x * wr - vk.y * wi), top.y - (vk.x * wi + vk.y * wr));
                    }
                taus[b] = tau;
                vdiag[b] = vk;
                betas[b] = beta;
            }
            __syncthreads();
        }
        // finished panel columns back to X (R on and above the diagonal, zeros below); Vp becomes V proper
        for (int e = tid; e < (n - k) * NB; e += NT) {
            const int i = k + e / NB, b = e % NB;
            if (b < nbp) {
                const int kb = k + b;
                cplx* dst = X + (size_t)i * ld + kb;
                cplx& v = Vp[(size_t)i * VS + b];
                if (i < kb) { *dst = v; v = make_double2(0.0, 0.0); }
                else if (i == kb) {
                    const bool ident = taus[b] == 0.0;
                    *dst = ident ? v : betas[b];
                    v = vdiag[b];
                } else {
                    if (taus[b] == 0.0) v = make_double2(0.0, 0.0);   // identity reflector
                    *dst = make_double2(0.0, 0.0);
                }
            }
        }
        __syncthreads();
        // T: T[b][b] = tau_b ;  T[0:b, b] = -tau_b T[0:b, 0:b] (V[:, 0:b]^H v_b); everything else zero
        for (int e = tid; e < NB * NB; e += NT) Tm[e / NB][e % NB] = make_double2(0.0, 0.0);
        __syncthreads();
        if (tid == 0) Tm[0][0] = make_double2(taus[0], 0.0);
        __syncthreads();
        for (int b = 1; b < nbp; ++b) {
            double acc[2 * NB];
#pragma unroll
            for (int q = 0; q < 2 * NB; ++q) acc[q] = 0.0;
            for (int i = k + b + tid; i < n; i += NT) {
                const cplx vb = Vp[(size_t)i * VS + b];
#pragma unroll
                for (int a = 0; a < NB; ++a)
                    if (a < b) {
                        const cplx va = Vp[(size_t)i * VS + a];
                        acc[2 * a] = fma(va.x, vb.x, fma(va.y, vb.y, acc[2 * a]));            // conj(va) vb
                        acc[2 * a + 1] = fma(va.x, vb.y, fma(-va.y, vb.x, acc[2 * a + 1]));
                    }
            }
            block_sum_vec<2 * NB>(acc, 2 * b, scratch);
            if (tid == 0) {
                const double tb = taus[b];
                for (int a = 0; a < b; ++a) {
                    double re = 0.0, im = 0.0;
                    for (int c = a; c < b; ++c) {                // T[a][c] (upper triangular) times g_c
                        const cplx tt = Tm[a][c];
                        const double gr = scratch[NW * 2 * NB + 2 * c], gi = scratch[NW * 2 * NB + 2 * c + 1];
                        re += tt.x * gr - tt.y * gi;
                        im += tt.x * gi + tt.y * gr;
                    }
                    Tm[a][b] = make_double2(-tb * re, -tb * im);
                }
                Tm[b][b] = make_double2(tb, 0.0);
            }
            __syncthreads();
        }
        // 3. trailing update on DMMA: warp -> groups of 8 columns to the right of the panel
        const int c_first = k + NB;                                       // k is a multiple of 8
        const int ngroups = (ltot - c_first + 7) / 8;
        // A fragments of T^H (m = b', k = a): conj(T[a][b']), two k-steps
        cplx th0 = Tm[t][g], th1 = Tm[t + 4][g];
        th0.y = -th0.y; th1.y = -th1.y;
        for (int grp = warp; grp < ngroups; grp += NW) {
            const int j0 = c_first + grp * 8;
            const bool colB = (j0 + g) < ltot;                            // operand layout: column g of the group
            const bool colC0 = (j0 + 2 * t) < ltot, colC1 = (j0 + 2 * t + 1) < ltot;      // accumulator layout
            // --- W = V^H X over rows k .. n-1 (rows >= n: V is zero there, X is not read) ---
            double wre0 = 0.0, wre1 = 0.0, wim0 = 0.0, wim1 = 0.0;
            const cplx* xcol = X + j0 + g;
            for (int i0 = k; i0 < n; i0 += 4 * QD_WU) {
                cplx xb[QD_WU];
#pragma unroll
                for (int u = 0; u < QD_WU; ++u) {
                    const int i = i0 + 4 * u + t;
                    xb[u] = (colB && i < n) ? xcol[(size_t)i * ld] : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < QD_WU; ++u) {
                    const int i = i0 + 4 * u + t;
                    const cplx v = (i < (int)nal) ? Vp[(size_t)i * VS + g] : make_double2(0.0, 0.0);
                    dmma884(wre0, wre1, v.x, xb[u].x);
                    dmma884(wre0, wre1, v.y, xb[u].y);
                    dmma884(wim0, wim1, v.x, xb[u].y);
                    dmma884(wim0, wim1, -v.y, xb[u].x);
                }
            }
            // --- Z = T^H W: W from accumulator layout (row g, columns 2t, 2t+1) to operand layout (row t, column g) ---
            cplx (*sg)[VS] = stage[warp];
            sg[g][2 * t] = make_double2(wre0, wim0);
            sg[g][2 * t + 1] = make_double2(wre1, wim1);
            __syncwarp();
            const cplx w0 = sg[t][g], w1 = sg[t + 4][g];
            __syncwarp();
            double zre0 = 0.0, zre1 = 0.0, zim0 = 0.0, zim1 = 0.0;
            dmma884(zre0, zre1, th0.x, w0.x); dmma884(zre0, zre1, -th0.y, w0.y);
            dmma884(zim0, zim1, th0.x, w0.y); dmma884(zim0, zim1, th0.y, w0.x);
            dmma884(zre0, zre1, th1.x, w1.x); dmma884(zre0, zre1, -th1.y, w1.y);
            dmma884(zim0, zim1, th1.x, w1.y); dmma884(zim0, zim1, th1.y, w1.x);
            sg[g][2 * t] = make_double2(zre0, zim0);
            sg[g][2 * t + 1] = make_double2(zre1, zim1);
            __syncwarp();
            const cplx z0 = sg[t][g], z1 = sg[t + 4][g];                  // Z[b = t][col g], Z[b = t + 4][col g]
            __syncwarp();
            // --- X -= V Z, 8-row tiles, QD_TILES tiles in flight ---
            for (int i0 = k; i0 < n; i0 += 8 * QD_TILES) {
                cplx x0[QD_TILES], x1[QD_TILES];
#pragma unroll
                for (int u = 0; u < QD_TILES; ++u) {
                    const int i = i0 + 8 * u + g;
                    const cplx* src = X + (size_t)i * ld + j0 + 2 * t;
                    x0[u] = (i < n && colC0) ? src[0] : make_double2(0.0, 0.0);
                    x1[u] = (i < n && colC1) ? src[1] : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < QD_TILES; ++u) {
                    const int i = i0 + 8 * u + g;
                    if (i0 + 8 * u < n) {                                    // warp-uniform
                        const cplx va = (i < (int)nal) ? Vp[(size_t)i * VS + t] : make_double2(0.0, 0.0);
                        const cplx vb = (i < (int)nal) ? Vp[(size_t)i * VS + t + 4] : make_double2(0.0, 0.0);
                        // x -= v z:  re -= v.x z.x - v.y z.y,  im -= v.x z.y + v.y z.x
                        dmma884(x0[u].x, x1[u].x, -va.x, z0.x); dmma884(x0[u].x, x1[u].x, va.y, z0.y);
                        dmma884(x0[u].y, x1[u].y, -va.x, z0.y); dmma884(x0[u].y, x1[u].y, -va.y, z0.x);
                        dmma884(x0[u].x, x1[u].x, -vb.x, z1.x); dmma884(x0[u].x, x1[u].x, vb.y, z1.y);
                        dmma884(x0[u].y, x1[u].y, -vb.x, z1.y); dmma884(x0[u].y, x1[u].y, -vb.y, z1.x);
                    }
                }
#pragma unroll
                for (int u = 0; u < QD_TILES; ++u) {
                    const int i = i0 + 8 * u + g;
                    cplx* dst = X + (size_t)i * ld + j0 + 2 * t;
                    if (i < n && colC0) dst[0] = x0[u];
                    if (i < n && colC1) dst[1] = x1[u];
                }
            }
        }
        __syncthreads();
    }
    // back to the original column order: X[i][order[j]] <- X[i][j]
    for (int i0 = 0; i0 < n; i0 += rows_per) {
        const int nr = min(rows_per, n - i0);
        for (int e = tid; e < nr * ldot; e += NT) {
            const int r = e / ldot, j = e - r * ldot;
            Vp[e] = X[(size_t)(i0 + r) * ld + j];
        }
        __syncthreads();
        for (int e = tid; e < nr * ldot; e += NT) {
            const int r = e / ldot, j = e - r * ldot;
            X[(size_t)(i0 + r) * ld + order[j]] = Vp[e];
        }
        __syncthreads();
    }
}

// ---- Jacobi -----------------------------------------------------------------------------------------
// Rotation parameters from the 2x2 Gram block [[a, c], [conj(c), b]]: the small-angle Jacobi rotation
//   P' = cs P - g Q,   Q' = conj(g) P + cs Q,   g = sn c/|c|,
// with t = sgn(z)|c| / (|z| + sqrt(z^2 + |c|^2)), z = (b - a)/2.  One sqrt, one reciprocal, one rsqrt
// (FP64 div/sqrt are ~20-40 instruction sequences, executed by every lane).
__device__ __forceinline__ void jacobi_angle(double a, double b, double cr, double ci, double ac2, double& cs,
                                             double& gr, double& gi) {
    const double z = 0.5 * (b - a);
    const double denom = fabs(z) + sqrt(fma(z, z, ac2));       // > 0 because ac2 > 0
    cs = denom * rsqrt(fma(denom, denom, ac2));
    const double f = (z >= 0.0 ? cs : -cs) / denom;
    gr = f * cr;
    gi = f * ci;
}

// One warp makes rows P and Q orthogonal over their first `ldot` entries and applies the rotation to
// all `ltot_r` entries.  Returns 1 if a rotation was applied.  All lanes take the same branch because
// the butterfly reduction leaves bit-identical sums in every lane.
// Fast path (ltot_r <= 256): both rows are pulled into registers once (8 complex per lane per row, all
// loads in flight together), the inner products and the rotation run from registers.
// real: the rows hold a REAL matrix with adjacent column pairs packed into one complex element; a left rotation by a
// real matrix acts on both halves alike, the real inner product is the real part of the complex one, and a zero
// imaginary part makes every term with gi below vanish - the real Jacobi rotation, bit for bit.
__device__ __forceinline__ int rotate_pair(cplx* __restrict__ P, cplx* __restrict__ Q, int ldot, int ltot_r,
                                           double F2, double tol2, double eabs2, int lane, bool real = false) {
    (void)F2; (void)eabs2;
    if (ltot_r <= 256) {
        cplx x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (c < ltot_r) { x[j] = P[c]; y[j] = Q[c]; }
            else { x[j] = make_double2(0.0, 0.0); y[j] = make_double2(0.0, 0.0); }
        }
        double a0 = 0.0, b0 = 0.0, cr0 = 0.0, ci0 = 0.0, a1 = 0.0, b1 = 0.0, cr1 = 0.0, ci1 = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            if (lane + 32 * j < ldot) {
                a0 = fma(x[j].x, x[j].x, fma(x[j].y, x[j].y, a0));
                b0 = fma(y[j].x, y[j].x, fma(y[j].y, y[j].y, b0));
                cr0 = fma(x[j].x, y[j].x, fma(x[j].y, y[j].y, cr0));
                ci0 = fma(x[j].y, y[j].x, fma(-x[j].x, y[j].y, ci0));
            }
            if (lane + 32 * (j + 1) < ldot) {
                a1 = fma(x[j + 1].x, x[j + 1].x, fma(x[j + 1].y, x[j + 1].y, a1));
                b1 = fma(y[j + 1].x, y[j + 1].x, fma(y[j + 1].y, y[j + 1].y, b1));
                cr1 = fma(x[j + 1].x, y[j + 1].x, fma(x[j + 1].y, y[j + 1].y, cr1));
                ci1 = fma(x[j + 1].y, y[j + 1].x, fma(-x[j + 1].x, y[j + 1].y, ci1));
            }
        }
        const double a = warp_sum(a0 + a1), b = warp_sum(b0 + b1);
        const double cr = warp_sum(cr0 + cr1), ci = real ? 0.0 : warp_sum(ci0 + ci1);
        const double ac2 = cr * cr + ci * ci;
        if (!(ac2 > fma(tol2 * a, b, PAIR_FLOOR))) return 0;
        double cs, gr, gi;
        jacobi_angle(a, b, cr, ci, ac2, cs, gr, gi);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (c < ltot_r) {
                cplx xn, yn;
                xn.x = fma(cs, x[j].x, fma(gi, y[j].y, -gr * y[j].x));
                xn.y = fma(cs, x[j].y, -fma(gr, y[j].y, gi * y[j].x));
                yn.x = fma(cs, y[j].x, fma(gr, x[j].x, gi * x[j].y));
                yn.y = fma(cs, y[j].y, fma(gr, x[j].y, -gi * x[j].x));
                P[c] = xn;
                Q[c] = yn;
            }
        }
        return 1;
    }
    double a = 0.0, b = 0.0, cr = 0.0, ci = 0.0;
    for (int c = lane; c < ldot; c += 32) {
        const cplx x = P[c], y = Q[c];
        a = fma(x.x, x.x, fma(x.y, x.y, a));
        b = fma(y.x, y.x, fma(y.y, y.y, b));
        cr = fma(x.x, y.x, fma(x.y, y.y, cr));     // x * conj(y)
        ci = fma(x.y, y.x, fma(-x.x, y.y, ci));
    }
    a = warp_sum(a); b = warp_sum(b); cr = warp_sum(cr); ci = real ? 0.0 : warp_sum(ci);
    const double ac2 = cr * cr + ci * ci;
    if (!(ac2 > fma(tol2 * a, b, PAIR_FLOOR))) return 0;
    double cs, gr, gi;
    jacobi_angle(a, b, cr, ci, ac2, cs, gr, gi);
    for (int c = lane; c < ltot_r; c += 32) {
        const cplx x = P[c], y = Q[c];
        cplx xn, yn;
        xn.x = fma(cs, x.x, fma(gi, y.y, -gr * y.x));
        xn.y = fma(cs, x.y, -fma(gr, y.y, gi * y.x));
        yn.x = fma(cs, y.x, fma(gr, x.x, gi * x.y));
        yn.y = fma(cs, y.y, fma(gr, x.y, -gi * x.x));
        P[c] = xn;
        Q[c] = yn;
    }
    return 1;
}

// grid = (nblk/2, problems in the chunk); block = 32*B threads.
// full != 0: the CTA plays a complete tournament among its 2B rows (used for round 0 of a sweep, so
// intra-block pairs are covered); full == 0: only the B*B cross pairs between the two blocks.
template <int B>
__global__ void __launch_bounds__(32 * B)
jacobi_round_kernel(cplx* __restrict__ Xall, long long strideX, int mat0, int ld, int ldot, int ltot_r,
                    int nblk, int round, int full, int inner_sweeps, const double* __restrict__ frob2,
                    int* __restrict__ rot, const int* __restrict__ done, double tol2, double eabs2,
                    unsigned long long* __restrict__ phase_clk, int real) {
    extern __shared__ __align__(128) unsigned char smraw[];
    cplx* rows = reinterpret_cast<cplx*>(smraw);            // [2B][ltot_r]
    const long long t_start = phase_clk ? clock64() : 0;
    __shared__ __align__(8) uint64_t bar;

    const int mat = mat0 + blockIdx.y;
    if (done[mat]) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int bI, bJ;
    rr_pair(nblk, round, blockIdx.x, bI, bJ);
    cplx* X = Xall + (size_t)mat * strideX;
    const uint32_t rowbytes = (uint32_t)ltot_r * sizeof(cplx);

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, rowbytes * 2 * B);
        for (int r = 0; r < B; ++r) {
            bulk_g2s(rows + (size_t)r * ltot_r, X + (size_t)(bI * B + r) * ld, rowbytes, &bar);
            bulk_g2s(rows + (size_t)(B + r) * ltot_r, X + (size_t)(bJ * B + r) * ld, rowbytes, &bar);
        }
    }
    if (!mbar_wait(&bar, 0)) __trap();
    const long long t_loaded = phase_clk ? clock64() : 0;

    const double F2 = frob2[mat];
    int nrot_total = 0;
    for (int sw = 0; sw < inner_sweeps; ++sw) {
        int nrot = 0;
        if (full) {
            for (int s = 0; s < 2 * B - 1; ++s) {
                int p, q;
                rr_pair(2 * B, s, w, p, q);
                nrot += rotate_pair(rows + (size_t)p * ltot_r, rows + (size_t)q * ltot_r, ldot, ltot_r, F2, tol2,
                                    eabs2, lane, real != 0);
                __syncthreads();
            }
        } else {
            for (int s = 0; s < B; ++s) {
                const int p = w, q = B + ((w + s) % B);
                nrot += rotate_pair(rows + (size_t)p * ltot_r, rows + (size_t)q * ltot_r, ldot, ltot_r, F2, tol2,
                                    eabs2, lane, real != 0);
                __syncthreads();
            }
        }
        nrot_total += nrot;
        if (!__syncthreads_or(nrot > 0)) break;
    }
    const int any = __syncthreads_or(nrot_total > 0);
    const long long t_rotated = phase_clk ? clock64() : 0;
    if (!any) {
        if (phase_clk && tid == 0) {
            atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
            atomicAdd(&phase_clk[3], 1ull);
        }
        return;
    }
    if (lane == 0 && nrot_total > 0) atomicAdd(&rot[mat], nrot_total);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        for (int r = 0; r < B; ++r) {
            bulk_s2g(X + (size_t)(bI * B + r) * ld, rows + (size_t)r * ltot_r, rowbytes);
            bulk_s2g(X + (size_t)(bJ * B + r) * ld, rows + (size_t)(B + r) * ltot_r, rowbytes);
        }
        bulk_commit();
        bulk_wait_all();
        if (phase_clk) {
            atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
            atomicAdd(&phase_clk[2], (unsigned long long)(clock64() - t_rotated));
            atomicAdd(&phase_clk[3], 1ull);
        }
    }
}

// ---- cross round, register-resident (rounds >= 1 of a sweep, rows of <= 256 columns) ------------------
// The 8 x 8 cross pairs between two blocks of 8 rows.  Warp w owns row p = w of block I in REGISTERS for the whole
// task (8 complex per lane, loaded straight from global, stored straight back); the rows of block J come through
// shared memory by TMA bulk copy and circulate: in step s warp w meets row q = (w + s) mod 8.
// Work per pair and column is cut to the minimum:
//   * row norms are computed once per task and then updated analytically (a' = a - t|c|, b' = b + t|c|), so only
//     the cross inner product c (4 FMA) is formed per pair;
//   * rotations are applied in scaled ("fast Givens") form: with rows X = d * x the update is
//         x_p' = x_p - tau_p x_q,   x_q' = x_q + sig_q x_p,   d' = cs * d        (8 FMA per column, not 12),
//     tau_p = tau d_q/d_p, sig_q = conj(tau) d_p/d_q, tau = t c/|c|; the scales are folded back into the rows at
//     the end of the task (at most 8 rotations per row, so d >= cs_min^8 stays far from underflow).
constexpr int CROSS_THREADS = 128;     // 4 warps, each owning TWO rows of block I in registers
#ifndef MPSB_CROSS_MINBLOCKS
#define MPSB_CROSS_MINBLOCKS 3
#endif

// State of one register-resident row: data x = X / d, true squared norm a, scale d and d^2.
struct RegRow {
    cplx x[8];
    double a, d, e;
    bool dirty;
};

// One Jacobi pair between a register row and the streamed row y (scale dq, eq = dq^2, true norm^2 bq).
// Returns 1 if a rotation was applied.
// FULL: all 256 columns take part in the inner product (no per-column predicates in the hot loop).
// sort: de Rijk's rule - if the rotation leaves the register row (lower block) with the smaller norm, the two
// results trade places (free: only the names of the registers change).
// REAL: rows of a real matrix with column pairs packed into complex elements (see rotate_pair): the inner product
// keeps its real part only and the rotation parameters are real - 6 instead of 12 FMAs per packed column.
template <bool FULL, bool REAL>
__device__ __forceinline__ int cross_pair(RegRow& r, cplx (&y)[8], double& dq, double& eq, double& bq, int ldot,
                                          double tol2, int lane, bool sort) {
    double cr0 = 0.0, ci0 = 0.0, cr1 = 0.0, ci1 = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        if (FULL || lane + 32 * j < ldot) {
            cr0 = fma(r.x[j].x, y[j].x, fma(r.x[j].y, y[j].y, cr0));
            if (!REAL) ci0 = fma(r.x[j].y, y[j].x, fma(-r.x[j].x, y[j].y, ci0));
        }
        if (FULL || lane + 32 * (j + 1) < ldot) {
            cr1 = fma(r.x[j + 1].x, y[j + 1].x, fma(r.x[j + 1].y, y[j + 1].y, cr1));
            if (!REAL) ci1 = fma(r.x[j + 1].y, y[j + 1].x, fma(-r.x[j + 1].x, y[j + 1].y, ci1));
        }
    }
    const double crs = warp_sum(cr0 + cr1), cis = REAL ? 0.0 : warp_sum(ci0 + ci1);   // scaled inner product x_p . x_q^H
    const double ac2 = r.e * eq * fma(crs, crs, cis * cis);                  // |c|^2 of the true rows
    if (!(ac2 > fma(tol2 * r.a, bq, PAIR_FLOOR))) return 0;
    // t/|c| = sgn(z) / (|z| + sqrt(z^2 + |c|^2)) from MUFU seeds + one Newton step each; the tangent may be
    // 1e-12 off without harm because ...
    const double z = 0.5 * (bq - r.a);
    const double v1 = fma(z, z, ac2);
    const double denom = fabs(z) + v1 * newton_rsqrt(v1, rsqrt_approx(v1));
    const double inv = newton_rcp(denom, rcp_approx(denom));
    const double k = z >= 0.0 ? inv : -inv;
    // tau_p = tau d_q/d_p = k d_q^2 c~ ;  sig_q = conj(tau) d_p/d_q = k d_p^2 conj(c~)   (no divisions)
    const double kq = k * eq, kp = k * r.e;
    const double tpr = kq * crs, tpi = kq * cis;
    const double sqr = kp * crs, sqi = -kp * cis;
    // ... the scale cs is computed from the tangent actually applied, to full precision, so the combined
    // transformation cs * [[1, -tau], [conj(tau), 1]] stays unitary to rounding
    const double wv = fma(inv * inv, ac2, 1.0);
    const double cs = newton_rsqrt(wv, newton_rsqrt(wv, rsqrt_approx(wv)));
    const double cs2 = cs * cs;
    const double delta = k * ac2;                                            // t |c|
    r.d *= cs;
    r.e *= cs2;
    r.a -= delta;
    r.dirty = true;
    dq *= cs;
    eq *= cs2;
    bq += delta;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const cplx xo = r.x[j], yo = y[j];
        if (REAL) {
            r.x[j].x = fma(-tpr, yo.x, xo.x);
            r.x[j].y = fma(-tpr, yo.y, xo.y);
            y[j].x = fma(sqr, xo.x, yo.x);
            y[j].y = fma(sqr, xo.y, yo.y);
        } else {
            r.x[j].x = fma(-tpr, yo.x, fma(tpi, yo.y, xo.x));
            r.x[j].y = fma(-tpr, yo.y, fma(-tpi, yo.x, xo.y));
            y[j].x = fma(sqr, xo.x, fma(-sqi, xo.y, yo.x));
            y[j].y = fma(sqr, xo.y, fma(sqi, xo.x, yo.y));
        }
    }
    // de Rijk's rule: the register row (lower block) keeps the larger norm.  With rows that start sorted (pivoted QR)
    // this fires for well under 1 % of the rotations of the first sweeps and never later; the caller then exchanges
    // the two rows through the shared-memory slot of y (return value 2), here only the bookkeeping trades places.
    if (sort && r.a < bq) {
        double t;
        t = r.d; r.d = dq; dq = t;
        t = r.e; r.e = eq; eq = t;
        t = r.a; r.a = bq; bq = t;
        return 2;
    }
    return 1;
}

// Exchange the register row with the streamed row y (cold path of de Rijk's rule): through y's shared-memory slot, so
// that no temporary registers are needed in a kernel that sits at its register limit.
template <bool FULL>
__device__ __forceinline__ void swap_rows_via_smem(RegRow& r, cplx (&y)[8], cplx* rq, int ltot_r, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        if (FULL || c < ltot_r) { rq[c] = r.x[j]; r.x[j] = y[j]; }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        if (FULL || c < ltot_r) y[j] = rq[c];
    }
}

template <bool FULL, bool REAL>
__global__ void __launch_bounds__(CROSS_THREADS, MPSB_CROSS_MINBLOCKS)
jacobi_cross_kernel(cplx* __restrict__ Xall, long long strideX, int mat0, int ld, int ldot, int ltot_r, int nblk,
                    int round, int* __restrict__ rot, const int* __restrict__ done, double tol2,
                    unsigned long long* __restrict__ phase_clk, int* __restrict__ stamps, int stamp_stride,
                    int launch_id, int modulus) {
    extern __shared__ __align__(128) unsigned char smraw[];
    cplx* rowsJ = reinterpret_cast<cplx*>(smraw);            // [8][ltot_r]
    __shared__ double nq[8], dqs[8], eqs[8];                  // true norm^2, scale d and d^2 of the rows of block J
    __shared__ __align__(8) uint64_t bar;

    const int mat = mat0 + blockIdx.y;
    if (done[mat]) return;
    const long long t_start = phase_clk ? clock64() : 0;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int bI, bJ;
    if (modulus) mod_pair(nblk, round, blockIdx.x, bI, bJ);      // bI < bJ: the register rows belong to the lower block
    else rr_pair(nblk, round, blockIdx.x, bI, bJ);
    int* stp = stamps ? stamps + (size_t)mat * stamp_stride : nullptr;
    const int chk_idx = modulus ? stamp_pair_index(nblk, bI, bJ) : nblk + round * (nblk >> 1) + blockIdx.x;
    if (stp) {                                               // verified clean since both blocks last changed
        const int chk = stp[chk_idx];
        if (chk > stp[bI] && chk > stp[bJ]) return;
    }
    cplx* X = Xall + (size_t)mat * strideX;
    const uint32_t rowbytes = (uint32_t)ltot_r * sizeof(cplx);

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, rowbytes * 8);
        for (int r = 0; r < 8; ++r) bulk_g2s(rowsJ + (size_t)r * ltot_r, X + (size_t)(bJ * 8 + r) * ld, rowbytes, &bar);
    }
    // rows 2w, 2w+1 of block I: global -> registers (512 contiguous bytes per warp load)
    cplx* gp0 = X + (size_t)(bI * 8 + 2 * w) * ld;
    cplx* gp1 = gp0 + ld;
    RegRow r0, r1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        r0.x[j] = (FULL || c < ltot_r) ? gp0[c] : make_double2(0.0, 0.0);
        r1.x[j] = (FULL || c < ltot_r) ? gp1[c] : make_double2(0.0, 0.0);
    }
    {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (FULL || lane + 32 * j < ldot) {
                a0 = fma(r0.x[j].x, r0.x[j].x, fma(r0.x[j].y, r0.x[j].y, a0));
                a1 = fma(r1.x[j].x, r1.x[j].x, fma(r1.x[j].y, r1.x[j].y, a1));
            }
        r0.a = warp_sum(a0); r1.a = warp_sum(a1);
        r0.d = r1.d = r0.e = r1.e = 1.0;
        r0.dirty = r1.dirty = false;
    }
    if (!mbar_wait(&bar, 0)) __trap();
    for (int rr = w; rr < 8; rr += 4) {   // norms of the rows of block J
        const cplx* r = rowsJ + (size_t)rr * ltot_r;
        double b = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (FULL || c < ldot) { const cplx y = r[c]; b = fma(y.x, y.x, fma(y.y, y.y, b)); }
        }
        b = warp_sum(b);
        if (lane == 0) { nq[rr] = b; dqs[rr] = 1.0; eqs[rr] = 1.0; }
    }
    __syncthreads();
    const long long t_loaded = phase_clk ? clock64() : 0;

    int nrot = 0;
#pragma unroll 1
    for (int s = 0; s < 8; ++s) {
        const int q = (2 * w + s) & 7;                    // the 4 warps meet 4 distinct rows of block J in every step
        cplx* rq = rowsJ + (size_t)q * ltot_r;
        cplx y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            y[j] = (FULL || c < ltot_r) ? rq[c] : make_double2(0.0, 0.0);
        }
        double dq = dqs[q], eq = eqs[q], bq = nq[q];
        int n0 = cross_pair<FULL, REAL>(r0, y, dq, eq, bq, ldot, tol2, lane, modulus != 0);
        if (n0 == 2) { swap_rows_via_smem<FULL>(r0, y, rq, ltot_r, lane); n0 = 1; }
        int n1 = cross_pair<FULL, REAL>(r1, y, dq, eq, bq, ldot, tol2, lane, modulus != 0);
        if (n1 == 2) { swap_rows_via_smem<FULL>(r1, y, rq, ltot_r, lane); n1 = 1; }
        nrot += n0 + n1;
        // Even rows of block J are visited in even steps, odd rows in odd steps, so steps 6 and 7 are the last
        // visits: fold the accumulated scale back into the row there.
        bool store = (n0 + n1) > 0;
        if (s >= 6 && dq != 1.0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { y[j].x *= dq; y[j].y *= dq; }
            dq = 1.0;
            eq = 1.0;
            store = true;
        }
        if (store) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                if (FULL || c < ltot_r) rq[c] = y[j];
            }
            if (lane == 0) { dqs[q] = dq; eqs[q] = eq; nq[q] = bq; }
        }
        __syncthreads();
    }
    const int any = __syncthreads_or(nrot > 0);
    const long long t_rotated = phase_clk ? clock64() : 0;
    if (!any) {
        if (phase_clk && tid == 0) {
            atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
            atomicAdd(&phase_clk[3], 1ull);
        }
        if (stp && tid == 0) stp[chk_idx] = launch_id;      // verified clean at this launch
        return;
    }
    if (lane == 0 && nrot > 0) atomicAdd(&rot[mat], nrot);
    if (stp && tid == 0) { stp[bI] = launch_id; stp[bJ] = launch_id; }
    // fold the scales back: register rows on their way to global memory, J rows in shared memory
    if (r0.dirty) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (FULL || c < ltot_r) gp0[c] = make_double2(r0.x[j].x * r0.d, r0.x[j].y * r0.d);
        }
    }
    if (r1.dirty) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (FULL || c < ltot_r) gp1[c] = make_double2(r1.x[j].x * r1.d, r1.x[j].y * r1.d);
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        for (int r = 0; r < 8; ++r) bulk_s2g(X + (size_t)(bJ * 8 + r) * ld, rowsJ + (size_t)r * ltot_r, rowbytes);
        bulk_commit();
        bulk_wait_all();
        if (phase_clk) {
            atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
            atomicAdd(&phase_clk[2], (unsigned long long)(clock64() - t_rotated));
            atomicAdd(&phase_clk[3], 1ull);
        }
    }
}

// ---- modulus ordering on SUPER blocks of 16 rows: four cross tasks per CTA ------------------------------------------
// jacobi_cross_kernel streams 16 rows in and out of the SM for 64 pairs; with 512 problems in a call that is 1 GiB per
// step and ~330 GB per SVD batch - at ~4 TB/s the sweeps sit closer to the HBM roofline than to the FP64 one.  Here the
// blocks are paired up: super block S = blocks {2S, 2S+1}, step K in [0, nblk/2) holds the super pairs with
// SI + SJ = K (mod nblk/2), and ONE CTA plays the four cross tasks (2SI+a, 2SJ+b): the 16 rows of the J side stay in
// shared memory for all four, the I side passes through the registers one 8-row block after the other - 32 rows moved
// for 256 pairs, half the traffic per pair, half the launches.  inner = 1: the cross task INSIDE the two super blocks
// that meet themselves on even steps, (2u, 2u+1) for u = K/2 and K/2 + nblk/4 (their self tasks run in jacobi_self_kernel
// before it on the same side stream).  Stamps: task t of the CTA writes launch_id + t, so that a pair verified clean
// after an earlier task of the same CTA changed one of its blocks still counts as clean next time.
template <bool FULL, bool REAL>
__global__ void __launch_bounds__(CROSS_THREADS, MPSB_CROSS_MINBLOCKS)
jacobi_quad_kernel(cplx* __restrict__ Xall, long long strideX, int mat0, int ld, int ldot, int ltot_r, int nblk, int kstep,
                   int inner, int* __restrict__ rot, const int* __restrict__ done, double tol2, int* __restrict__ stamps,
                   int stamp_stride, int launch_id) {
    extern __shared__ __align__(128) unsigned char smraw[];
    cplx* rowsJ = reinterpret_cast<cplx*>(smraw);            // [16][ltot_r] (8 rows when inner)
    __shared__ double nq[16], dqs[16], eqs[16];              // true norm^2, scale d and d^2 of the rows of the J side
    __shared__ __align__(8) uint64_t bar;

    const int mat = mat0 + blockIdx.y;
    if (done[mat]) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nsup = nblk >> 1;
    int bI0, bJ0, nsub;
    if (inner) {
        bI0 = 2 * ((kstep >> 1) + (int)blockIdx.x * (nsup >> 1));
        bJ0 = bI0 + 1;
        nsub = 1;
    } else {
        int SI, SJ;
        mod_pair(nsup, kstep, blockIdx.x, SI, SJ);
        bI0 = 2 * SI; bJ0 = 2 * SJ; nsub = 2;
    }
    int* stp = stamps ? stamps + (size_t)mat * stamp_stride : nullptr;
    // pairs verified clean since both blocks last changed need no visit - unless this CTA changes a block first
    // (bit 2a + b of `need`: task (bI0 + a, bJ0 + b); chI / chJ: blocks this CTA has rotated)
    unsigned need = nsub == 2 ? 0xfu : 0x1u;
    if (stp) {
        unsigned m = 0;
        for (int t = 0; t < 4; ++t) {
            if (!(need >> t & 1u)) continue;
            const int bi = bI0 + (t >> 1), bj = bJ0 + (t & 1);
            const int chk = stp[stamp_pair_index(nblk, bi, bj)];
            if (!(chk > stp[bi] && chk > stp[bj])) m |= 1u << t;
        }
        need = m;
        if (!need) return;
    }
    unsigned chI = 0, chJ = 0;
    cplx* X = Xall + (size_t)mat * strideX;
    const uint32_t rowbytes = (uint32_t)ltot_r * sizeof(cplx);
    const int nrowsJ = 8 * nsub;

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, rowbytes * nrowsJ);
        for (int r = 0; r < nrowsJ; ++r) bulk_g2s(rowsJ + (size_t)r * ltot_r, X + (size_t)(bJ0 * 8 + r) * ld, rowbytes, &bar);
    }
    bool waitedJ = false;
    int nrot = 0;
#pragma unroll 1
    for (int a = 0; a < nsub; ++a) {
        if (!((need >> (2 * a) & 3u) | chJ)) continue;
        // rows 2w, 2w+1 of block bI0 + a: global -> registers (512 contiguous bytes per warp load)
        cplx* gp0 = X + (size_t)((bI0 + a) * 8 + 2 * w) * ld;
        cplx* gp1 = gp0 + ld;
        RegRow r0, r1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            r0.x[j] = (FULL || c < ltot_r) ? gp0[c] : make_double2(0.0, 0.0);
            r1.x[j] = (FULL || c < ltot_r) ? gp1[c] : make_double2(0.0, 0.0);
        }
        {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (FULL || lane + 32 * j < ldot) {
                    a0 = fma(r0.x[j].x, r0.x[j].x, fma(r0.x[j].y, r0.x[j].y, a0));
                    a1 = fma(r1.x[j].x, r1.x[j].x, fma(r1.x[j].y, r1.x[j].y, a1));
                }
            r0.a = warp_sum(a0); r1.a = warp_sum(a1);
            r0.d = r1.d = r0.e = r1.e = 1.0;
            r0.dirty = r1.dirty = false;
        }
        if (!waitedJ) {
            if (!mbar_wait(&bar, 0)) __trap();
            for (int rr = w; rr < nrowsJ; rr += 4) {          // norms of the rows of the J side
                const cplx* r = rowsJ + (size_t)rr * ltot_r;
                double b = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = lane + 32 * j;
                    if (FULL || c < ldot) { const cplx y = r[c]; b = fma(y.x, y.x, fma(y.y, y.y, b)); }
                }
                b = warp_sum(b);
                if (lane == 0) { nq[rr] = b; dqs[rr] = 1.0; eqs[rr] = 1.0; }
            }
            __syncthreads();
            waitedJ = true;
        }
#pragma unroll 1
        for (int b = 0; b < nsub; ++b) {
            if (!((need >> (2 * a + b) & 1u) | (chI >> a & 1u) | (chJ >> b & 1u))) continue;
            cplx* blockJ = rowsJ + (size_t)(8 * b) * ltot_r;
            double* nqb = nq + 8 * b;
            double* dqb = dqs + 8 * b;
            double* eqb = eqs + 8 * b;
            int trot = 0;
#pragma unroll 1
            for (int s = 0; s < 8; ++s) {
                const int q = (2 * w + s) & 7;                // the 4 warps meet 4 distinct rows of the block in every step
                cplx* rq = blockJ + (size_t)q * ltot_r;
                cplx y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = lane + 32 * j;
                    y[j] = (FULL || c < ltot_r) ? rq[c] : make_double2(0.0, 0.0);
                }
                double dq = dqb[q], eq = eqb[q], bq = nqb[q];
                int n0 = cross_pair<FULL, REAL>(r0, y, dq, eq, bq, ldot, tol2, lane, true);
                if (n0 == 2) { swap_rows_via_smem<FULL>(r0, y, rq, ltot_r, lane); n0 = 1; }
                int n1 = cross_pair<FULL, REAL>(r1, y, dq, eq, bq, ldot, tol2, lane, true);
                if (n1 == 2) { swap_rows_via_smem<FULL>(r1, y, rq, ltot_r, lane); n1 = 1; }
                trot += n0 + n1;
                bool store = (n0 + n1) > 0;
                if (s >= 6 && dq != 1.0) {                    // last visit of this row in the task: fold its scale back
#pragma unroll
                    for (int j = 0; j < 8; ++j) { y[j].x *= dq; y[j].y *= dq; }
                    dq = 1.0;
                    eq = 1.0;
                    store = true;
                }
                if (store) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = lane + 32 * j;
                        if (FULL || c < ltot_r) rq[c] = y[j];
                    }
                    if (lane == 0) { dqb[q] = dq; eqb[q] = eq; nqb[q] = bq; }
                }
                __syncthreads();
            }
            const int any = __syncthreads_or(trot > 0);
            nrot += trot;
            const int id = launch_id + 2 * a + b;
            if (any) {
                chI |= 1u << a;
                chJ |= 1u << b;
                if (stp && tid == 0) { stp[bI0 + a] = id; stp[bJ0 + b] = id; }
            } else if (stp && tid == 0) {
                stp[stamp_pair_index(nblk, bI0 + a, bJ0 + b)] = id;      // verified clean at this task
            }
        }
        // fold the scales back: register rows on their way to global memory
        if (r0.dirty) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                if (FULL || c < ltot_r) gp0[c] = make_double2(r0.x[j].x * r0.d, r0.x[j].y * r0.d);
            }
        }
        if (r1.dirty) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                if (FULL || c < ltot_r) gp1[c] = make_double2(r1.x[j].x * r1.d, r1.x[j].y * r1.d);
            }
        }
    }
    if (lane == 0 && nrot > 0) atomicAdd(&rot[mat], nrot);
    if (!waitedJ) {                                           // nothing visited: the loads must still land before exit
        if (!mbar_wait(&bar, 0)) __trap();
        return;
    }
    if (!chJ) return;
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        for (int b = 0; b < nsub; ++b)
            if (chJ >> b & 1u)
                for (int r = 0; r < 8; ++r)
                    bulk_s2g(X + (size_t)((bJ0 + b) * 8 + r) * ld, rowsJ + (size_t)(8 * b + r) * ltot_r, rowbytes);
        bulk_commit();
        bulk_wait_all();
    }
}

// ---- self tasks of the modulus ordering (jacobi_selftask.cuh) --------------------------------------------------------
// Even steps: blocks round/2 and round/2 + nblk/2 meet nobody, one 128-thread CTA per problem orthogonalises the pairs
// inside them.  A kernel of its own: compiled into jacobi_cross_kernel its 4.5 k straight-line instructions cost the
// cross CTAs registers (spills in the hot loop) and instruction-cache hits (no-instruction stalls 0.08 -> 0.47 per issue).
template <int CPL, bool EXACT>
__global__ void __launch_bounds__(128, 3)
jacobi_self_kernel(cplx* __restrict__ Xall, long long strideX, int mat0, int ld, int ldot, int ltot_r, int nblk, int round,
                   int* __restrict__ rot, const int* __restrict__ done, double tol2, int* __restrict__ stamps,
                   int stamp_stride, int launch_id, int real) {
    const int mat = mat0 + blockIdx.x;
    if (done[mat]) return;
    cplx* X = Xall + (size_t)mat * strideX;
    int* stp = stamps ? stamps + (size_t)mat * stamp_stride : nullptr;
    const int bA = round >> 1, bB = bA + (nblk >> 1);
    selftask_body<CPL, EXACT>(X, ld, ldot, ltot_r, bA, bB, rot + mat, tol2, stp, stamp_self_index(nblk, bA), launch_id,
                              real != 0);
}

// ---- persistent sweeps for a few problems (latency mode of the modulus ordering) --------------------------------------
// One SVD of a 256 x 256 problem is ~330 steps of 16 block-pair tasks.  Launched step by step that is ~500 launches of
// 16-CTA kernels per problem: with only a few problems in the call (one iDMRG chain, a short TEBD chain) every step
// costs a launch latency.  Here ONE cooperative launch runs all sweeps: CTA (slot, problem) plays the slot's task of
// every step - the cross pair mod_pair(nblk, k, slot), or on even steps in the last slot the two self tasks - and the
// nblk/2 CTAs of a problem meet at a barrier of their own (a counter in global memory) between steps.  A sweep that
// rotated nothing ends the problem; no host round trip until the call returns.
// Rows move between CTAs through global memory (L2): every read below bypasses L1 (ld.global.cg / TMA), writes are
// fenced before the barrier.  The change stamps of the multi-launch path are not used (their reads would have to
// bypass L1 too, and skipping clean pairs saves little when the steps are latency-bound anyway).
__device__ __forceinline__ void problem_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned seen = 0, it = 0;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen < target && ++it > (1u << 26)) __trap();      // a lost CTA must fail loudly, not hang the GPU
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

// The body of jacobi_cross_kernel for one block pair, callable step after step: the mbarrier is re-armed per use
// (`parity` flips), block-I rows are read around L1.  Returns (to every thread) whether anything rotated.
template <bool FULL, bool REAL>
__device__ __forceinline__ void cross_task_persistent(cplx* __restrict__ X, int ld, int ldot, int ltot_r, int bI, int bJ,
                                                      int* __restrict__ rot_mat, double tol2, cplx* rowsJ, double* nq,
                                                      double* dqs, double* eqs, uint64_t* bar, uint32_t parity) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t rowbytes = (uint32_t)ltot_r * sizeof(cplx);
    if (tid == 0) {
        mbar_expect_tx(bar, rowbytes * 8);
        for (int r = 0; r < 8; ++r) bulk_g2s(rowsJ + (size_t)r * ltot_r, X + (size_t)(bJ * 8 + r) * ld, rowbytes, bar);
    }
    cplx* gp0 = X + (size_t)(bI * 8 + 2 * w) * ld;
    cplx* gp1 = gp0 + ld;
    RegRow r0, r1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        r0.x[j] = (FULL || c < ltot_r) ? __ldcg(gp0 + c) : make_double2(0.0, 0.0);
        r1.x[j] = (FULL || c < ltot_r) ? __ldcg(gp1 + c) : make_double2(0.0, 0.0);
    }
    {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (FULL || lane + 32 * j < ldot) {
                a0 = fma(r0.x[j].x, r0.x[j].x, fma(r0.x[j].y, r0.x[j].y, a0));
                a1 = fma(r1.x[j].x, r1.x[j].x, fma(r1.x[j].y, r1.x[j].y, a1));
            }
        r0.a = warp_sum(a0); r1.a = warp_sum(a1);
        r0.d = r1.d = r0.e = r1.e = 1.0;
        r0.dirty = r1.dirty = false;
    }
    if (!mbar_wait(bar, parity)) __trap();
    for (int rr = w; rr < 8; rr += 4) {   // norms of the rows of block J
        const cplx* r = rowsJ + (size_t)rr * ltot_r;
        double b = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (FULL || c < ldot) { const cplx y = r[c]; b = fma(y.x, y.x, fma(y.y, y.y, b)); }
        }
        b = warp_sum(b);
        if (lane == 0) { nq[rr] = b; dqs[rr] = 1.0; eqs[rr] = 1.0; }
    }
    __syncthreads();
    int nrot = 0;
#pragma unroll 1
    for (int s = 0; s < 8; ++s) {
        const int q = (2 * w + s) & 7;
        cplx* rq = rowsJ + (size_t)q * ltot_r;
        cplx y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            y[j] = (FULL || c < ltot_r) ? rq[c] : make_double2(0.0, 0.0);
        }
        double dq = dqs[q], eq = eqs[q], bq = nq[q];
        int n0 = cross_pair<FULL, REAL>(r0, y, dq, eq, bq, ldot, tol2, lane, true);
        if (n0 == 2) { swap_rows_via_smem<FULL>(r0, y, rq, ltot_r, lane); n0 = 1; }
        int n1 = cross_pair<FULL, REAL>(r1, y, dq, eq, bq, ldot, tol2, lane, true);
        if (n1 == 2) { swap_rows_via_smem<FULL>(r1, y, rq, ltot_r, lane); n1 = 1; }
        nrot += n0 + n1;
        bool store = (n0 + n1) > 0;
        if (s >= 6 && dq != 1.0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { y[j].x *= dq; y[j].y *= dq; }
            dq = 1.0;
            eq = 1.0;
            store = true;
        }
        if (store) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                if (FULL || c < ltot_r) rq[c] = y[j];
            }
            if (lane == 0) { dqs[q] = dq; eqs[q] = eq; nq[q] = bq; }
        }
        __syncthreads();
    }
    const int any = __syncthreads_or(nrot > 0);
    if (!any) return;
    if (lane == 0 && nrot > 0) atomicAdd(rot_mat, nrot);
    if (r0.dirty) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (FULL || c < ltot_r) __stcg(gp0 + c, make_double2(r0.x[j].x * r0.d, r0.x[j].y * r0.d));
        }
    }
    if (r1.dirty) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (FULL || c < ltot_r) __stcg(gp1 + c, make_double2(r1.x[j].x * r1.d, r1.x[j].y * r1.d));
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        for (int r = 0; r < 8; ++r) bulk_s2g(X + (size_t)(bJ * 8 + r) * ld, rowsJ + (size_t)r * ltot_r, rowbytes);
        bulk_commit();
        bulk_wait_all();
    }
}

// grid = (nblk / 2, batch), cooperative.  sync: [batch] barrier counters | [2][batch] rotation counters of the even / odd
// sweeps, all zero on entry.
template <int CPL, bool FULL, bool REAL>
__global__ void __launch_bounds__(CROSS_THREADS, 2)
jacobi_persistent_kernel(cplx* __restrict__ Xall, long long strideX, int ld, int ldot, int ltot_r, int nblk, int max_sweeps,
                         double tol2, unsigned* __restrict__ sync, unsigned long long* __restrict__ rot_total,
                         unsigned long long* __restrict__ sweep_total, int* __restrict__ max_sweeps_seen) {
    extern __shared__ __align__(128) unsigned char smraw[];
    cplx* rowsJ = reinterpret_cast<cplx*>(smraw);            // [8][ltot_r]
    __shared__ double nq[8], dqs[8], eqs[8];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int mat = blockIdx.y, slot = blockIdx.x, nslots = gridDim.x, batch = gridDim.y;
    cplx* X = Xall + (size_t)mat * strideX;
    unsigned* counter = sync + mat;
    int* rotbuf = reinterpret_cast<int*>(sync) + batch;      // [2][batch]
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t uses = 0;
    unsigned epoch = 0;
    unsigned long long rot_sum = 0ull;
    int sweeps = 0;
    while (sweeps < max_sweeps) {
        int* rot_mat = rotbuf + (size_t)(sweeps & 1) * batch + mat;
        for (int k = 0; k < nblk; ++k) {
            const int ncross = mod_cross_pairs(nblk, k);
            if (slot < ncross) {
                int bI, bJ;
                mod_pair(nblk, k, slot, bI, bJ);
                cross_task_persistent<FULL, REAL>(X, ld, ldot, ltot_r, bI, bJ, rot_mat, tol2, rowsJ, nq, dqs, eqs, &bar, uses & 1u);
                ++uses;
            } else {
                const int bA = k >> 1, bB = bA + (nblk >> 1);
                selftask_body<CPL, FULL>(X, ld, ldot, ltot_r, bA, bB, rot_mat, tol2, nullptr, 0, 0, REAL);
            }
            if (k == 1 && slot == 0 && tid == 0) rotbuf[(size_t)((sweeps + 1) & 1) * batch + mat] = 0;   // next sweep's counter
            problem_barrier(counter, ++epoch * (unsigned)nslots);
        }
        int nrot;
        asm volatile("ld.relaxed.gpu.s32 %0, [%1];" : "=r"(nrot) : "l"(rot_mat) : "memory");
        ++sweeps;
        rot_sum += (unsigned long long)nrot;
        if (nrot == 0) break;
    }
    if (slot == 0 && tid == 0) {
        atomicAdd(rot_total, rot_sum);
        atomicAdd(sweep_total, (unsigned long long)sweeps);
        atomicMax(max_sweeps_seen, sweeps);
    }
}

// ---- Gram-assisted round (block of 16 rows per CTA, FP64 tensor cores) -------------------------------
// Same outer schedule as jacobi_round_kernel<8>, but the long-row arithmetic runs on DMMA:
//   A. G = P P^H (16x16 Hermitian) over the first ldot columns            -- DMMA, K split over the 8 warps
//   B. one full round-robin tournament (15 steps x 8 disjoint pairs) of two-sided Jacobi rotations on G,
//      accumulated into the unitary Q (16x16); angles from the maintained G, relative test as above
//   C. P <- Q^H P over all ltot_r columns                                   -- DMMA, columns split over the warps
// No inner products by warp shuffles, no per-pair pass over the long rows.  Rotations computed from the
// maintained Gram block keep relative accuracy (errors scale with sqrt(g_pp g_qq), as in two-sided Jacobi on a
// scaled-diagonally-dominant matrix), and the convergence test of every later visit uses a freshly computed G.
constexpr int GR = 16;                 // rows per CTA
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_PAD = 2;            // row pitch = ltot_r + 2 elements: conflict-free B fragments in phase C

struct GramSmem {
    cplx G[GR][GR];                    // Gram block; also the K-half-0 partial during phase A
    cplx Q[GR][GR];                    // accumulated rotations; also the K-half-1 partial during phase A
    double cs[8], gr[8], gi[8];
    int partner[GR], role[GR], rix[GR];
    int nrot;
    uint64_t bar;
    uint64_t tile_bar[2];              // tiled kernel: one barrier per tile buffer
};

// Division-free rotation parameters (see jacobi_angle): f = +-1/sqrt(denom^2 + |c|^2), cs = denom * |f|.
__device__ __forceinline__ void jacobi_angle_fast(double a, double b, double cr, double ci, double ac2, double& cs,
                                                  double& gr, double& gi) {
    const double z = 0.5 * (b - a);
    const double denom = fabs(z) + sqrt(fma(z, z, ac2));
    const double h = rsqrt(fma(denom, denom, ac2));
    cs = denom * h;
    const double f = z >= 0.0 ? h : -h;
    gr = f * cr;
    gi = f * ci;
}

// Phase B of the Gram kernels: tournament of two-sided rotations on the 16 x 16 Gram block sm.G (nsteps = 15: all
// pairs; nsteps = 8: the cross pairs between rows 0-7 and 8-15), accumulated into sm.Q.  Thread (ei, ej) owns one
// element of G and Q.  Returns the number of rotations applied (same value in every thread).
__device__ __forceinline__ int gram_tournament(GramSmem& sm, int nsteps, double tol2, int tid, int ei, int ej) {
    int my_rot = 0;
    for (int s = 0; s < nsteps; ++s) {
        if (tid < 8) {
            int p, q;
            if (nsteps == GR - 1) rr_pair(GR, s, tid, p, q);
            else { p = tid; q = 8 + ((tid + s) & 7); }
            const double a = sm.G[p][p].x, b = sm.G[q][q].x;
            const cplx c = sm.G[p][q];
            const double ac2 = c.x * c.x + c.y * c.y;
            double cs = 1.0, gr = 0.0, gi = 0.0;
            if (ac2 > fma(tol2 * a, b, PAIR_FLOOR)) {
                jacobi_angle_fast(a, b, c.x, c.y, ac2, cs, gr, gi);
                ++my_rot;
            }
            sm.cs[tid] = cs; sm.gr[tid] = gr; sm.gi[tid] = gi;
            sm.partner[p] = q; sm.role[p] = 0; sm.rix[p] = tid;
            sm.partner[q] = p; sm.role[q] = 1; sm.rix[q] = tid;
        }
        __syncthreads();
        cplx gn, qn;
        {
            const int ip = sm.partner[ei], jp = sm.partner[ej];
            const int ri = sm.rix[ei], rj = sm.rix[ej];
            const double csi = sm.cs[ri], csj = sm.cs[rj];
            // rows:    P' = cs P - g Q,  Q' = conj(g) P + cs Q   -> partner coefficient w_i = -g | conj(g)
            // columns: Hermitian conjugate of the row transform     -> partner coefficient conj(w_j)
            const double wri = sm.role[ei] == 0 ? -sm.gr[ri] : sm.gr[ri];
            const double wii = -sm.gi[ri];
            const double wrj = sm.role[ej] == 0 ? -sm.gr[rj] : sm.gr[rj];
            const double wij = sm.gi[rj];
            const cplx g00 = sm.G[ei][ej], g10 = sm.G[ip][ej], g01 = sm.G[ei][jp], g11 = sm.G[ip][jp];
            cplx Tj, Tjp;
            Tj.x = fma(csi, g00.x, fma(wri, g10.x, -wii * g10.y));
            Tj.y = fma(csi, g00.y, fma(wri, g10.y, wii * g10.x));
            Tjp.x = fma(csi, g01.x, fma(wri, g11.x, -wii * g11.y));
            Tjp.y = fma(csi, g01.y, fma(wri, g11.y, wii * g11.x));
            gn.x = fma(csj, Tj.x, fma(wrj, Tjp.x, -wij * Tjp.y));
            gn.y = fma(csj, Tj.y, fma(wrj, Tjp.y, wij * Tjp.x));
            if (ei == ej) gn.y = 0.0;
            const cplx q0 = sm.Q[ei][ej], q1 = sm.Q[ei][jp];
            qn.x = fma(csj, q0.x, fma(wrj, q1.x, -wij * q1.y));
            qn.y = fma(csj, q0.y, fma(wrj, q1.y, wij * q1.x));
        }
        __syncthreads();
        sm.G[ei][ej] = gn;
        sm.Q[ei][ej] = qn;
        __syncthreads();
    }
    if (tid < 8 && my_rot > 0) atomicAdd(&sm.nrot, my_rot);
    __syncthreads();
    const int nrot = sm.nrot;
    __syncthreads();
    if (tid == 0) sm.nrot = 0;
    return nrot;
}

// REAL (both Gram kernels): packed real rows (see rotate_pair) - G keeps its real part (2 of the 4 DMMAs of phase A),
// the tournament then produces a real Q, and phase C drops the 2 DMMAs that carry Im Q.
template <bool REAL>
__global__ void __launch_bounds__(GRAM_THREADS, 3)
jacobi_gram_kernel(cplx* __restrict__ Xall, long long strideX, int mat0, int ld, int ldot, int ltot_r, int nblk,
                   int round, int inner_sweeps, int* __restrict__ rot, const int* __restrict__ done, double tol2,
                   unsigned long long* __restrict__ phase_clk, int* __restrict__ stamps, int stamp_stride,
                   int launch_id) {
    extern __shared__ __align__(128) unsigned char smraw[];
    GramSmem& sm = *reinterpret_cast<GramSmem*>(smraw);
    cplx* rows = reinterpret_cast<cplx*>(smraw + ((sizeof(GramSmem) + 127) / 128) * 128);   // [GR][pitch]
    const int pitch = ltot_r + GRAM_PAD;

    const int mat = mat0 + blockIdx.y;
    if (done[mat]) return;
    const long long t_start = phase_clk ? clock64() : 0;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    int bI, bJ;
    rr_pair(nblk, round, blockIdx.x, bI, bJ);
    int* stp = stamps ? stamps + (size_t)mat * stamp_stride : nullptr;
    const int chk_idx = nblk + round * (nblk >> 1) + blockIdx.x;
    if (stp) {                                               // verified clean since both blocks last changed
        const int chk = stp[chk_idx];
        if (chk > stp[bI] && chk > stp[bJ]) return;
    }
    cplx* X = Xall + (size_t)mat * strideX;
    const uint32_t rowbytes = (uint32_t)ltot_r * sizeof(cplx);

    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
        sm.nrot = 0;
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&sm.bar, rowbytes * GR);
        for (int r = 0; r < 8; ++r) {
            bulk_g2s(rows + (size_t)r * pitch, X + (size_t)(bI * 8 + r) * ld, rowbytes, &sm.bar);
            bulk_g2s(rows + (size_t)(8 + r) * pitch, X + (size_t)(bJ * 8 + r) * ld, rowbytes, &sm.bar);
        }
    }
    if (!mbar_wait(&sm.bar, 0)) __trap();
    const long long t_loaded = phase_clk ? clock64() : 0;

    // Round 0 of a sweep (and the single-CTA case) plays the complete 16-row tournament so that pairs inside a
    // block are covered; later rounds only need the 8 x 8 cross pairs between the two blocks.
    const int nsteps = (round == 0 || nblk == 2) ? GR - 1 : 8;
    const int b_reps = (inner_sweeps >> 16) > 0 ? (inner_sweeps >> 16) : 1;      // high half: tournament repetitions
    inner_sweeps &= 0xffff;
    const int ei = tid >> 4, ej = tid & 15;          // element of G / Q owned by this thread
    int total_rot = 0;
    for (int sw = 0; sw < inner_sweeps; ++sw) {
        // ---- A. Gram matrix: warp -> (tile, K half); partials land in sm.G (half 0) and sm.Q (half 1) ----
        const long long tA = phase_clk ? clock64() : 0;
        {
            const int tile = w & 3, khalf = w >> 2;
            const int ti = tile >> 1, tj = tile & 1;
            double re0 = 0.0, re1 = 0.0, im0 = 0.0, im1 = 0.0;
            const int ksteps = (ldot + 3) >> 2;
            const cplx* ra = rows + (size_t)(8 * ti + g) * pitch + t;
            const cplx* rb = rows + (size_t)(8 * tj + g) * pitch + t;
#pragma unroll 4
            for (int ks = khalf; ks < ksteps; ks += 2) {
                cplx a = ra[4 * ks], b = rb[4 * ks];
                if (4 * ks + t >= ldot) { a = make_double2(0.0, 0.0); b = make_double2(0.0, 0.0); }
                dmma884(re0, re1, a.x, b.x);
                dmma884(re0, re1, a.y, b.y);
                if (!REAL) {
                    dmma884(im0, im1, a.y, b.x);
                    dmma884(im0, im1, -a.x, b.y);
                }
            }
            cplx* dst = khalf == 0 ? &sm.G[8 * ti + g][8 * tj + 2 * t] : &sm.Q[8 * ti + g][8 * tj + 2 * t];
            dst[0] = make_double2(re0, im0);
            dst[1] = make_double2(re1, im1);
        }
        __syncthreads();
        {
            cplx v = sm.G[ei][ej];
            const cplx u = sm.Q[ei][ej];
            v.x += u.x;
            v.y = (ei == ej) ? 0.0 : v.y + u.y;
            sm.G[ei][ej] = v;
            sm.Q[ei][ej] = make_double2(ei == ej ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
        // ---- B. tournament on G ------------------------------------------------------------------------
        const long long tB = phase_clk ? clock64() : 0;
        // the tournament may be repeated on the maintained Gram block (G <- Q^H G Q stays current), so that a visit
        // orthogonalises its 16 rows further without another pass over the long rows
        int nrot = 0;
        for (int rep = 0; rep < b_reps; ++rep) {
            const int r = gram_tournament(sm, nsteps, tol2, tid, ei, ej);
            nrot += r;
            if (r == 0) break;
        }
        const long long tC = phase_clk ? clock64() : 0;
        if (phase_clk && tid == 0) {
            atomicAdd(&phase_clk[4], (unsigned long long)(tB - tA));
            atomicAdd(&phase_clk[5], (unsigned long long)(tC - tB));
        }
        if (nrot == 0) break;
        total_rot += nrot;
        // ---- C. P <- Q^H P ------------------------------------------------------------------------------
        {
            cplx af[2][4];            // A[m][k] = conj(Q[k][m]); row tiles m0 = 0, 8; k steps k0 = 0, 4, 8, 12
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const cplx v = sm.Q[4 * ks + t][8 * mt + g];
                    af[mt][ks] = make_double2(v.x, -v.y);
                }
            const int ntiles = ltot_r >> 3;
            for (int nt = w; nt < ntiles; nt += 8) {
                const int n0 = nt * 8;
                cplx bf[4];
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) bf[ks] = rows[(size_t)(4 * ks + t) * pitch + n0 + g];
                double cre[2][2], cim[2][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) { cre[mt][0] = cre[mt][1] = cim[mt][0] = cim[mt][1] = 0.0; }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        dmma884(cre[mt][0], cre[mt][1], af[mt][ks].x, bf[ks].x);
                        dmma884(cim[mt][0], cim[mt][1], af[mt][ks].x, bf[ks].y);
                        if (!REAL) {
                            dmma884(cre[mt][0], cre[mt][1], -af[mt][ks].y, bf[ks].y);
                            dmma884(cim[mt][0], cim[mt][1], af[mt][ks].y, bf[ks].x);
                        }
                    }
                __syncwarp();
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    cplx* dst = rows + (size_t)(8 * mt + g) * pitch + n0 + 2 * t;
                    dst[0] = make_double2(cre[mt][0], cim[mt][0]);
                    dst[1] = make_double2(cre[mt][1], cim[mt][1]);
                }
            }
        }
        __syncthreads();
        if (phase_clk && tid == 0) atomicAdd(&phase_clk[6], (unsigned long long)(clock64() - tC));
    }
    const long long t_rotated = phase_clk ? clock64() : 0;
    if (total_rot == 0) {
        if (phase_clk && tid == 0) {
            atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
            atomicAdd(&phase_clk[3], 1ull);
        }
        if (stp && tid == 0) stp[chk_idx] = launch_id;
        return;
    }
    if (tid == 0) atomicAdd(&rot[mat], total_rot);
    if (stp && tid == 0) { stp[bI] = launch_id; stp[bJ] = launch_id; }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        for (int r = 0; r < 8; ++r) {
            bulk_s2g(X + (size_t)(bI * 8 + r) * ld, rows + (size_t)r * pitch, rowbytes);
            bulk_s2g(X + (size_t)(bJ * 8 + r) * ld, rows + (size_t)(8 + r) * pitch, rowbytes);
        }
        bulk_commit();
        bulk_wait_all();
        if (phase_clk) {
            atomicAdd(&phase_clk[0], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&phase_clk[1], (unsigned long long)(t_rotated - t_loaded));
            atomicAdd(&phase_clk[2], (unsigned long long)(clock64() - t_rotated));
            atomicAdd(&phase_clk[3], 1ull);
        }
    }
}

// ---- Gram rounds for rows that do not fit shared memory (chi = 1024: 2048 x 4096 work matrices) -------------------
// Same three phases as jacobi_gram_kernel, but the 16 rows stream through two 256-column tile buffers: phase A
// accumulates G over the tiles of the dot columns (next tile in flight while the current one feeds the DMMAs), phase C
// re-reads every tile, applies Q^H and writes it back with an asynchronous bulk store that overlaps the next tile.
// Traffic per CTA and round: 16 rows x (ldot + 2 ltot) x 16 B; one CTA per SM (140 KB of shared memory).
constexpr int TILE_W = 256;
template <bool REAL>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
jacobi_gram_tiled_kernel(cplx* __restrict__ Xall, long long strideX, int mat0, int ld, int ldot, int ltot_r, int nblk,
                         int round, int inner_sweeps, int* __restrict__ rot, const int* __restrict__ done, double tol2,
                         int* __restrict__ stamps, int stamp_stride, int launch_id) {
    extern __shared__ __align__(128) unsigned char smraw[];
    GramSmem& sm = *reinterpret_cast<GramSmem*>(smraw);
    cplx* tiles = reinterpret_cast<cplx*>(smraw + ((sizeof(GramSmem) + 127) / 128) * 128);   // [2][GR][pitch]
    constexpr int pitch = TILE_W + GRAM_PAD;

    const int mat = mat0 + blockIdx.y;
    if (done[mat]) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    int bI, bJ;
    rr_pair(nblk, round, blockIdx.x, bI, bJ);
    int* stp = stamps ? stamps + (size_t)mat * stamp_stride : nullptr;
    const int chk_idx = nblk + round * (nblk >> 1) + blockIdx.x;
    if (stp) {
        const int chk = stp[chk_idx];
        if (chk > stp[bI] && chk > stp[bJ]) return;
    }
    cplx* X = Xall + (size_t)mat * strideX;
    if (tid == 0) {
        mbar_init(&sm.tile_bar[0], 1);
        mbar_init(&sm.tile_bar[1], 1);
        mbar_fence_init();
        sm.nrot = 0;
    }
    __syncthreads();
    uint32_t uses[2] = {0u, 0u};                          // completed phases of each tile barrier (uniform)
    auto row_ptr = [&](int r) { return X + (size_t)(r < 8 ? bI * 8 + r : bJ * 8 + r - 8) * ld; };
    auto tile_ptr = [&](int buf) { return tiles + (size_t)buf * GR * pitch; };
    auto issue_load = [&](int buf, int c0) {              // thread 0: columns [c0, c0 + width) of the 16 rows
        const int width = min(TILE_W, ltot_r - c0);
        const uint32_t bytes = (uint32_t)width * sizeof(cplx);
        mbar_expect_tx(&sm.tile_bar[buf], bytes * GR);
        for (int r = 0; r < GR; ++r) bulk_g2s(tile_ptr(buf) + (size_t)r * pitch, row_ptr(r) + c0, bytes, &sm.tile_bar[buf]);
    };
    auto wait_tile = [&](int buf) {
        if (!mbar_wait(&sm.tile_bar[buf], uses[buf] & 1u)) __trap();
        ++uses[buf];
    };

    const int nsteps = (round == 0 || nblk == 2) ? GR - 1 : 8;
    const int b_reps = (inner_sweeps >> 16) > 0 ? (inner_sweeps >> 16) : 1;
    inner_sweeps &= 0xffff;
    const int ei = tid >> 4, ej = tid & 15;
    const int ntA = (ldot + TILE_W - 1) / TILE_W, ntC = (ltot_r + TILE_W - 1) / TILE_W;
    int total_rot = 0;
    for (int sw = 0; sw < inner_sweeps; ++sw) {
        // ---- A. Gram matrix over the tiles of the dot columns ------------------------------------------------
        {
            const int qd = w & 3, khalf = w >> 2;
            const int ti = qd >> 1, tj = qd & 1;
            double re0 = 0.0, re1 = 0.0, im0 = 0.0, im1 = 0.0;
            if (tid == 0) issue_load(0, 0);
            for (int k = 0; k < ntA; ++k) {
                const int buf = k & 1, c0 = k * TILE_W;
                if (tid == 0 && k + 1 < ntA) issue_load(buf ^ 1, c0 + TILE_W);
                wait_tile(buf);
                const int lim = min(TILE_W, ldot - c0);
                const int ksteps = (lim + 3) >> 2;
                const cplx* ra = tile_ptr(buf) + (size_t)(8 * ti + g) * pitch + t;
                const cplx* rb = tile_ptr(buf) + (size_t)(8 * tj + g) * pitch + t;
#pragma unroll 4
                for (int ks = khalf; ks < ksteps; ks += 2) {
                    cplx a = ra[4 * ks], b = rb[4 * ks];
                    if (4 * ks + t >= lim) { a = make_double2(0.0, 0.0); b = make_double2(0.0, 0.0); }
                    dmma884(re0, re1, a.x, b.x);
                    dmma884(re0, re1, a.y, b.y);
                    if (!REAL) {
                        dmma884(im0, im1, a.y, b.x);
                        dmma884(im0, im1, -a.x, b.y);
                    }
                }
                __syncthreads();                          // buffer free for the load issued two tiles ahead
            }
            cplx* dst = khalf == 0 ? &sm.G[8 * ti + g][8 * tj + 2 * t] : &sm.Q[8 * ti + g][8 * tj + 2 * t];
            dst[0] = make_double2(re0, im0);
            dst[1] = make_double2(re1, im1);
        }
        __syncthreads();
        {
            cplx v = sm.G[ei][ej];
            const cplx u = sm.Q[ei][ej];
            v.x += u.x;
            v.y = (ei == ej) ? 0.0 : v.y + u.y;
            sm.G[ei][ej] = v;
            sm.Q[ei][ej] = make_double2(ei == ej ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
        // ---- B. tournament on G ------------------------------------------------------------------------------
        // the tournament may be repeated on the maintained Gram block (G <- Q^H G Q stays current), so that a visit
        // orthogonalises its 16 rows further without another pass over the long rows
        int nrot = 0;
        for (int rep = 0; rep < b_reps; ++rep) {
            const int r = gram_tournament(sm, nsteps, tol2, tid, ei, ej);
            nrot += r;
            if (r == 0) break;
        }
        if (nrot == 0) break;
        total_rot += nrot;
        // ---- C. P <- Q^H P, tile by tile ------------------------------------------------------------------------
        {
            cplx af[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const cplx v = sm.Q[4 * ks + t][8 * mt + g];
                    af[mt][ks] = make_double2(v.x, -v.y);
                }
            if (tid == 0) issue_load(0, 0);
            for (int k = 0; k < ntC; ++k) {
                const int buf = k & 1, c0 = k * TILE_W;
                const int width = min(TILE_W, ltot_r - c0);
                if (tid == 0 && k + 1 < ntC) {
                    bulk_wait_read_all();                 // the store of tile k-1 has finished reading that buffer
                    issue_load(buf ^ 1, c0 + TILE_W);
                }
                wait_tile(buf);
                cplx* rows = tile_ptr(buf);
                const int ntiles = width >> 3;
                for (int nt = w; nt < ntiles; nt += 8) {
                    const int n0 = nt * 8;
                    cplx bf[4];
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) bf[ks] = rows[(size_t)(4 * ks + t) * pitch + n0 + g];
                    double cre[2][2], cim[2][2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) { cre[mt][0] = cre[mt][1] = cim[mt][0] = cim[mt][1] = 0.0; }
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            dmma884(cre[mt][0], cre[mt][1], af[mt][ks].x, bf[ks].x);
                            dmma884(cim[mt][0], cim[mt][1], af[mt][ks].x, bf[ks].y);
                            if (!REAL) {
                                dmma884(cre[mt][0], cre[mt][1], -af[mt][ks].y, bf[ks].y);
                                dmma884(cim[mt][0], cim[mt][1], af[mt][ks].y, bf[ks].x);
                            }
                        }
                    __syncwarp();
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        cplx* dst = rows + (size_t)(8 * mt + g) * pitch + n0 + 2 * t;
                        dst[0] = make_double2(cre[mt][0], cim[mt][0]);
                        dst[1] = make_double2(cre[mt][1], cim[mt][1]);
                    }
                }
                fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    const uint32_t bytes = (uint32_t)width * sizeof(cplx);
                    for (int r = 0; r < GR; ++r) bulk_s2g(row_ptr(r) + c0, rows + (size_t)r * pitch, bytes);
                    bulk_commit();
                }
            }
            if (tid == 0) bulk_wait_all();                // rows are back in global memory before they are read again
            __syncthreads();
        }
    }
    if (total_rot == 0) {
        if (stp && tid == 0) stp[chk_idx] = launch_id;
        return;
    }
    if (tid == 0) {
        atomicAdd(&rot[mat], total_rot);
        if (stp) { stp[bI] = launch_id; stp[bJ] = launch_id; }
    }
}

// ---- resident iteration: one CTA keeps the whole work matrix in shared memory ----------------------------
// For matrices up to ~220 KB (100 x 100 with pitch 128, 64 x 64 with carried columns, ...) the multi-launch
// iteration above is bound by launch latency and the per-sweep convergence poll: ~15 launches per sweep for a few
// microseconds of work.  Here a single launch runs every sweep of a problem: the matrix comes in by TMA bulk copies,
// one warp rotates one row pair per step of the round-robin tournament over all n_pad rows (inner products by warp
// shuffles, both rows in registers, NJ = row length / 32 columns per lane), the CTA synchronises between steps and
// stops after the first sweep that applied no rotation.
template <int NJ>
__device__ __forceinline__ int rotate_rows(cplx* __restrict__ P, cplx* __restrict__ Q, int ldot, double tol2, int lane,
                                           bool real) {
    cplx x[NJ], y[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { x[j] = P[lane + 32 * j]; y[j] = Q[lane + 32 * j]; }     // pitch padding is zero
    double a = 0.0, b = 0.0, cr = 0.0, ci = 0.0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (lane + 32 * j < ldot) {
            a = fma(x[j].x, x[j].x, fma(x[j].y, x[j].y, a));
            b = fma(y[j].x, y[j].x, fma(y[j].y, y[j].y, b));
            cr = fma(x[j].x, y[j].x, fma(x[j].y, y[j].y, cr));
            ci = fma(x[j].y, y[j].x, fma(-x[j].x, y[j].y, ci));
        }
    }
    a = warp_sum(a); b = warp_sum(b); cr = warp_sum(cr); ci = real ? 0.0 : warp_sum(ci);
    const double ac2 = cr * cr + ci * ci;
    if (!(ac2 > fma(tol2 * a, b, PAIR_FLOOR))) return 0;
    double cs, gr, gi;
    jacobi_angle(a, b, cr, ci, ac2, cs, gr, gi);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        cplx xn, yn;
        xn.x = fma(cs, x[j].x, fma(gi, y[j].y, -gr * y[j].x));
        xn.y = fma(cs, x[j].y, -fma(gr, y[j].y, gi * y[j].x));
        yn.x = fma(cs, y[j].x, fma(gr, x[j].x, gi * x[j].y));
        yn.y = fma(cs, y[j].y, fma(gr, x[j].y, -gi * x[j].x));
        P[lane + 32 * j] = xn;
        Q[lane + 32 * j] = yn;
    }
    return 1;
}

template <int NJ>      // 1..8: columns per lane; 0: rows of any length, streamed from shared memory
__global__ void __launch_bounds__((NJ >= 1 && NJ <= 4) ? 1024 : 512, 1)
jacobi_resident_kernel(cplx* __restrict__ Xall, long long strideX, int n_pad, int ld, int ldot, int max_sweeps,
                       double tol2, unsigned long long* __restrict__ rot_total,
                       unsigned long long* __restrict__ sweep_total, int* __restrict__ max_sweeps_seen, int real) {
    extern __shared__ __align__(128) unsigned char smraw[];
    cplx* rows = reinterpret_cast<cplx*>(smraw);            // [n_pad][ld]
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned long long s_rot;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    cplx* X = Xall + (size_t)blockIdx.x * strideX;
    const uint32_t rowbytes = (uint32_t)ld * sizeof(cplx);

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        s_rot = 0ull;
    }
    __syncthreads();
    if (w == 0) {
        if (lane == 0) mbar_expect_tx(&bar, rowbytes * (uint32_t)n_pad);
        __syncwarp();
        for (int r = lane; r < n_pad; r += 32) bulk_g2s(rows + (size_t)r * ld, X + (size_t)r * ld, rowbytes, &bar);
    }
    if (!mbar_wait(&bar, 0)) __trap();

    const int half = n_pad >> 1, steps = n_pad - 1;
    int sweeps = 0;
    unsigned long long my_rot = 0ull;
    while (sweeps < max_sweeps) {
        int nrot = 0;
        for (int s = 0; s < steps; ++s) {
            for (int slot = w; slot < half; slot += nw) {
                int p, q;
                rr_pair(n_pad, s, slot, p, q);
                cplx* P = rows + (size_t)p * ld;
                cplx* Q = rows + (size_t)q * ld;
                if (NJ > 0) nrot += rotate_rows<(NJ > 0 ? NJ : 1)>(P, Q, ldot, tol2, lane, real != 0);
                else nrot += rotate_pair(P, Q, ldot, ld, 0.0, tol2, 0.0, lane, real != 0);
            }
            __syncthreads();
        }
        ++sweeps;
        my_rot += (unsigned long long)nrot;
        if (!__syncthreads_or(nrot > 0)) break;
    }
    if (lane == 0 && my_rot) atomicAdd(&s_rot, my_rot);
    fence_async_smem();
    __syncthreads();
    if (w == 0) {
        for (int r = lane; r < n_pad; r += 32) bulk_s2g(X + (size_t)r * ld, rows + (size_t)r * ld, rowbytes);
        bulk_commit();
        bulk_wait_all();
    }
    if (tid == 0) {
        atomicAdd(rot_total, s_rot);
        atomicAdd(sweep_total, (unsigned long long)sweeps);
        atomicMax(max_sweeps_seen, sweeps);
    }
}

__global__ void sweep_begin_kernel(int* rot, int* done, int* n_active, unsigned long long* rot_total,
                                   unsigned long long* sweep_total, int batch, int reset_done) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { *n_active = 0; *rot_total = 0ull; *sweep_total = 0ull; }
    if (i < 8) rot_total[16 + i] = 0ull;   // phase_clk lives 128 bytes after rot_total
    if (i < batch) {
        rot[i] = 0;
        if (reset_done) done[i] = 0;
    }
}
__global__ void sweep_end_kernel(int* rot, int* done, int* n_active, unsigned long long* rot_total,
                                 unsigned long long* sweep_total, int mat0, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const int m = mat0 + i;
        if (!done[m]) {
            atomicAdd(sweep_total, 1ull);
            if (rot[m] == 0) done[m] = 1;
            else { atomicAdd(n_active, 1); atomicAdd(rot_total, (unsigned long long)rot[m]); }
        }
        rot[m] = 0;
    }
}

// ---- finalise: row norms -> sort -> truncate -> scatter ------------------------------------------------
__global__ void __launch_bounds__(FIN_THREADS)
finalise_kernel(const cplx* __restrict__ Xall, long long strideX, int n, int n_pad, int npow2, int ldot, int ltot,
                int ld, SvdOutputs o) {
    extern __shared__ __align__(128) unsigned char smraw[];
    double* key = reinterpret_cast<double*>(smraw);              // [npow2]
    int* idx = reinterpret_cast<int*>(key + npow2);              // [npow2]
    __shared__ int s_k;
    __shared__ double s_scale;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = FIN_THREADS / 32;
    const int mat = blockIdx.x;
    const cplx* X = Xall + (size_t)mat * strideX;

    for (int i = warp; i < npow2; i += nw) {
        double acc = 0.0;
        if (i < n_pad) {
            const cplx* row = X + (size_t)i * ld;
            for (int c = lane; c < ldot; c += 32) {
                const cplx x = row[c];
                acc = fma(x.x, x.x, fma(x.y, x.y, acc));
            }
            acc = warp_sum(acc);
        }
        if (lane == 0) {
            key[i] = (i < n_pad) ? acc : -1.0;   // padding sorts last
            idx[i] = i;
        }
    }
    __syncthreads();
    // bitonic sort, descending by key
    for (int size = 2; size <= npow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < npow2; i += FIN_THREADS) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = ((i & size) == 0);
                    const double ki = key[i], kj = key[j];
                    const bool swap = desc ? (ki < kj) : (ki > kj);
                    if (swap) {
                        key[i] = kj; key[j] = ki;
                        const int t = idx[i]; idx[i] = idx[j]; idx[j] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    // truncation rule.  Reference `truncate` (DMRG/MPS.py:33-38): s /= s[0]; k = min(#(s > err), trun);
    // s = s[:k] / ||s[:k]||.   Reference `halve` (DMRG/iDMRG.py:26-31): k = min(trunc, S.size), raw S.
    const int lcols = o.real ? o.real_cols : ldot;      // packed real rows: two real columns per element
    const int nsv = n < lcols ? n : lcols;
    if (tid == 0) {
        const double s0 = key[0] > 0.0 ? sqrt(key[0]) : 0.0;
        int k = 0;
        if (o.normalise) {
            if (s0 > 0.0)
                for (int i = 0; i < nsv; ++i) k += (sqrt(fmax(key[i], 0.0)) / s0 > o.err);
        } else {
            k = nsv;
        }
        if (k > o.trun) k = o.trun;
        if (k > o.k_cap) k = o.k_cap;
        double nrm2 = 0.0;
        for (int i = 0; i < k; ++i) {
            const double r = sqrt(fmax(key[i], 0.0)) / s0;
            nrm2 += r * r;
        }
        s_k = k;
        s_scale = (k > 0) ? 1.0 / (s0 * sqrt(nrm2)) : 0.0;
        o.rank[mat] = k;
    }
    __syncthreads();
    const int k = s_k;
    for (int i = tid; i < o.k_cap; i += FIN_THREADS) {
        const double si = (i < k) ? sqrt(fmax(key[i], 0.0)) : 0.0;
        o.s[(size_t)mat * o.s_stride + i] = o.normalise ? si * s_scale : si;
    }
    if (o.s_raw)
        for (int i = tid; i < o.n_raw; i += FIN_THREADS)
            o.s_raw[(size_t)mat * o.s_raw_stride + i] = (i < nsv) ? sqrt(fmax(key[i], 0.0)) : 0.0;
    const int naug = ltot - ldot;
    for (int i = warp; i < o.k_cap; i += nw) {
        const bool keep = i < k && key[i] > ROW_DEAD;
        const cplx* row = X + (size_t)(keep ? idx[i] : 0) * ld;
        const double inv = keep ? 1.0 / sqrt(key[i]) : 0.0;
        if (o.real) {
            // packed real rows: double c of a row is real column c; outputs are arrays of doubles (strides in doubles)
            const double* rr = reinterpret_cast<const double*>(row);
            if (o.vh) {
                double* dst = reinterpret_cast<double*>(o.vh) + (size_t)mat * o.vh_stride + (size_t)i * o.vh_ld;
                for (int c = lane; c < o.real_cols; c += 32) dst[c] = keep ? rr[c] * inv : 0.0;
            }
            if (o.aug && o.real_naug > 0) {
                double* dst = reinterpret_cast<double*>(o.aug) + (size_t)mat * o.aug_stride + (size_t)i * o.aug_ld;
                for (int c = lane; c < o.real_naug; c += 32) dst[c] = keep ? rr[2 * ldot + c] : 0.0;
            }
            continue;
        }
        if (o.vh) {
            cplx* dst = o.vh + (size_t)mat * o.vh_stride + (size_t)i * o.vh_ld;
            for (int c = lane; c < ldot; c += 32) {
                cplx x = make_double2(0.0, 0.0);
                if (keep) { x = row[c]; x.x *= inv; x.y *= inv; }
                dst[c] = x;
            }
        }
        if (o.aug && naug > 0) {
            cplx* dst = o.aug + (size_t)mat * o.aug_stride + (size_t)i * o.aug_ld;
            for (int c = lane; c < naug; c += 32) {
                cplx x = make_double2(0.0, 0.0);
                if (keep) { x = row[ldot + c]; if (o.aug_conj) x.y = -x.y; }
                dst[c] = x;
            }
        }
    }
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    return atoi(s);
}

template <int B>
int launch_round(const SvdPlan& pl, cplx* X, const SvdState& s, int mat0, int count, int round, int full,
                 int inner, double tol2, double eabs2, cudaStream_t st) {
    const int ltot_r = round_up(pl.ltot, 32);
    const size_t smem = (size_t)2 * B * ltot_r * sizeof(cplx);
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(jacobi_round_kernel<B>), smem)) return rc;
    const int nblk = pl.n_pad / pl.bsz;
    dim3 grid(nblk / 2, count);
    jacobi_round_kernel<B><<<grid, 32 * B, smem, st>>>(X, (long long)pl.n_pad * pl.ld, mat0, pl.ld, pl.ldot, ltot_r,
                                                      nblk, round, full, inner, s.frob2, s.rot, s.done, tol2, eabs2,
                                                      g_prof.enabled > 1 ? s.phase_clk : nullptr, pl.real);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

size_t qr_grid_scratch_bytes(const SvdPlan& pl) {      // per problem
    return align_up(512 + align_up((size_t)pl.ld * sizeof(double), 256) + 2 * align_up(pl.n, 8) * sizeof(cplx), 256);
}
// Returns 0 when the cooperative kernel ran, -1 when it cannot be used here (the caller falls back), > 0 on errors.
int launch_qr_grid(const SvdPlan& pl, cplx* X, double* frob2, unsigned char* scratch, cudaStream_t st) {
    int dev = 0, sms = 0, coop = 0;
    MPSB_CHECK_CUDA(cudaGetDevice(&dev));
    MPSB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MPSB_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop || sms < 1) return -1;
    const int per_problem = std::max(1, sms / pl.batch);
    int Lc = round_up((pl.ltot + per_problem - 1) / per_problem, 2);
    int TC = 1;
    while (TC < Lc && TC < 32) TC <<= 1;
    const int ncta = (pl.ltot + Lc - 1) / Lc;
    const int RG = QG_THREADS / TC;
    const size_t smem = 2 * align_up(pl.n, 8) * sizeof(cplx) + (size_t)(2 + RG) * Lc * sizeof(cplx) +
                        align_up(pl.ldot, 8) * (sizeof(double) + sizeof(int)) + Lc + 64;
    if (smem > 200 * 1024) return -1;
    const void* fn = reinterpret_cast<const void*>(qr_grid_kernel);
    if (int rc = ensure_dynamic_smem(fn, smem)) return rc;
    int per_sm = 0;
    MPSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qr_grid_kernel, QG_THREADS, smem));
    if ((long long)per_sm * sms < (long long)ncta * pl.batch) return -1;
    const size_t stride = qr_grid_scratch_bytes(pl);
    MPSB_CHECK_CUDA(cudaMemsetAsync(scratch, 0, stride * pl.batch, st));
    long long strideX = (long long)pl.n_pad * pl.ld;
    int n = pl.n, ldot = pl.ldot, ltot = pl.ltot, ld = pl.ld;
    size_t stride_arg = stride;
    void* args[] = {&X, &strideX, &n, &ldot, &ltot, &ld, &Lc, &TC, &frob2, &scratch, &stride_arg};
    MPSB_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(ncta, pl.batch), dim3(QG_THREADS), args, smem, st));
    MPSB_COUNT_LAUNCH();
    return 0;
}

int gram_reps() {
    static const int r = std::max(1, std::min(16, env_int("MPSB_GRAM_REPS", 1)));
    return r;
}
size_t gram_smem_bytes(int ltot_r) {
    return ((sizeof(GramSmem) + 127) / 128) * 128 + (size_t)GR * (ltot_r + GRAM_PAD) * sizeof(cplx);
}
bool use_gram_kernel(const SvdPlan& pl) {
    static const int enabled = env_int("MPSB_JACOBI_GRAM", 1);
    return enabled && pl.bsz == 8 && gram_smem_bytes(round_up(pl.ltot, 32)) <= 220 * 1024;
}
int launch_gram(const SvdPlan& pl, cplx* X, const SvdState& s, int mat0, int count, int round, int inner,
                double tol2, int launch_id, cudaStream_t st) {
    const int ltot_r = round_up(pl.ltot, 32);
    const size_t smem = gram_smem_bytes(ltot_r);
    auto kern = pl.real ? jacobi_gram_kernel<true> : jacobi_gram_kernel<false>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem)) return rc;
    const int nblk = pl.n_pad / 8;
    dim3 grid(nblk / 2, count);
    kern<<<grid, GRAM_THREADS, smem, st>>>(X, (long long)pl.n_pad * pl.ld, mat0, pl.ld, pl.ldot, ltot_r,
                                          nblk, round, inner | (gram_reps() << 16), s.rot, s.done, tol2,
                                          g_prof.enabled > 1 ? s.phase_clk : nullptr,
                                          nblk == pl.n_pad / pl.bsz ? s.stamps : nullptr, s.stamp_stride,
                                          launch_id);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

size_t gram_tiled_smem_bytes() {
    return ((sizeof(GramSmem) + 127) / 128) * 128 + (size_t)2 * GR * (TILE_W + GRAM_PAD) * sizeof(cplx);
}
int launch_gram_tiled(const SvdPlan& pl, cplx* X, const SvdState& s, int mat0, int count, int round, int inner,
                      double tol2, int launch_id, cudaStream_t st) {
    const int ltot_r = round_up(pl.ltot, 32);
    const size_t smem = gram_tiled_smem_bytes();
    auto kern = pl.real ? jacobi_gram_tiled_kernel<true> : jacobi_gram_tiled_kernel<false>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem)) return rc;
    const int nblk = pl.n_pad / 8;
    dim3 grid(nblk / 2, count);
    kern<<<grid, GRAM_THREADS, smem, st>>>(X, (long long)pl.n_pad * pl.ld, mat0, pl.ld, pl.ldot,
                                          ltot_r, nblk, round, inner | (gram_reps() << 16), s.rot, s.done, tol2, s.stamps,
                                          s.stamp_stride, launch_id);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Side stream of the calling host thread on the current device (created on first use, lives as long as the process).
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };
SideStream* side_stream() {
    static thread_local std::map<int, SideStream> table;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    auto it = table.find(dev);
    if (it != table.end()) return &it->second;
    SideStream ss;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&ss.stream, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess)
        return nullptr;
    return &table.emplace(dev, ss).first->second;
}

typedef void (*CrossKernel)(cplx*, long long, int, int, int, int, int, int, int*, const int*, double, unsigned long long*,
                            int*, int, int, int);
CrossKernel cross_kernel_for(bool full, bool real) {
    if (real) return full ? jacobi_cross_kernel<true, true> : jacobi_cross_kernel<false, true>;
    return full ? jacobi_cross_kernel<true, false> : jacobi_cross_kernel<false, false>;
}

// Modulus ordering (rows of <= 256 columns, 8-row blocks): step `round` in [0, nblk) = one launch of the register-
// resident cross kernel over the step's cross pairs; on even steps a second, small launch plays the two self tasks
// (pairs inside blocks round/2 and round/2 + nblk/2, which no cross pair of the step touches).
bool use_modulus_order(const SvdPlan& pl) {
    const int nblk = pl.n_pad / 8, ltot_r = round_up(pl.ltot, 32);
    return env_int("MPSB_JACOBI_ORDER", 1) == 1 && pl.bsz == 8 && ltot_r <= 256 && nblk >= 4 && (nblk & 1) == 0 &&
           !pl.resident;
}
int launch_modulus_step(const SvdPlan& pl, cplx* X, const SvdState& s, int mat0, int count, int round, double tol2,
                        int launch_id, cudaStream_t st) {
    const int ltot_r = round_up(pl.ltot, 32), nblk = pl.n_pad / 8;
    const long long strideX = (long long)pl.n_pad * pl.ld;
    unsigned long long* pc = g_prof.enabled > 1 ? s.phase_clk : nullptr;
    const size_t smem = (size_t)8 * ltot_r * sizeof(cplx);
    const bool full = (ltot_r == 256 && pl.ldot == 256);
    auto kern = cross_kernel_for(full, pl.real != 0);
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem)) return rc;
    dim3 grid(mod_cross_pairs(nblk, round), count);
    if ((round & 1) == 0) {
        // The two self tasks of an even step touch blocks no cross pair of the step touches, and their kernel is bound
        // by latency (one 128-thread CTA per problem, ~2000 clocks per tournament step): it runs on a side stream,
        // forked before and joined after the cross launch, so that its CTAs share the SMs with the cross CTAs
        // instead of following them.
        static const int overlap = env_int("MPSB_SELF_OVERLAP", 1);
        SideStream* side = overlap ? side_stream() : nullptr;
        cudaStream_t ss = st;
        if (side) {
            MPSB_CHECK_CUDA(cudaEventRecord(side->fork, st));
            MPSB_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
            ss = side->stream;
        }
#define MPSB_SELF(CPL, EXACT)                                                                                              \
    jacobi_self_kernel<CPL, EXACT><<<count, 128, 0, ss>>>(X, strideX, mat0, pl.ld, pl.ldot, ltot_r, nblk, round, s.rot, s.done, \
                                                          tol2, s.stamps, s.stamp_stride, launch_id, pl.real)
        if (full) MPSB_SELF(4, true);
        else if (ltot_r > 128) MPSB_SELF(4, false);
        else if (ltot_r > 64) MPSB_SELF(2, false);
        else MPSB_SELF(1, false);
#undef MPSB_SELF
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
        if (side) MPSB_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
        kern<<<grid, CROSS_THREADS, smem, st>>>(X, strideX, mat0, pl.ld, pl.ldot, ltot_r, nblk, round, s.rot, s.done, tol2, pc,
                                               s.stamps, s.stamp_stride, launch_id, 1);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
        if (side) MPSB_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
        return 0;
    }
    kern<<<grid, CROSS_THREADS, smem, st>>>(X, strideX, mat0, pl.ld, pl.ldot, ltot_r, nblk, round, s.rot, s.done, tol2, pc,
                                           s.stamps, s.stamp_stride, launch_id, 1);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Super-block form of the modulus ordering (jacobi_quad_kernel) for calls with enough problems to fill the GPU with
// CTAs that live four times as long: step K in [0, nblk/2) = one launch over the step's super pairs; on even steps the
// two super blocks K/2 and K/2 + nblk/4 meet themselves on the side stream - the self tasks of their four 8-row blocks
// (the existing kernel, old rounds 2K and 2K+2) followed by the two cross tasks inside them.
// launch_id is a multiple of 8: tasks of a CTA stamp launch_id + 0..3, the side-stream launches + 0, 1, 2.
typedef void (*QuadKernel)(cplx*, long long, int, int, int, int, int, int, int, int*, const int*, double, int*, int, int);
QuadKernel quad_kernel_for(bool full, bool real) {
    if (real) return full ? jacobi_quad_kernel<true, true> : jacobi_quad_kernel<false, true>;
    return full ? jacobi_quad_kernel<true, false> : jacobi_quad_kernel<false, false>;
}
bool use_quad_order(const SvdPlan& pl, int count) {
    // 1: rows of 256 complex columns only - the shape it was measured on (512 two-site updates at chi = 128: Jacobi
    // 83.2 -> 81.2 ms); the packed real rows of the 256-chain iDMRG split run 10 % SLOWER through it (split 62 -> 70 ms),
    // so they stay on jacobi_cross_kernel.  2: every shape the modulus ordering serves (tests).  0: off.
    static const int enabled = env_int("MPSB_JACOBI_QUAD", 1);
    static const int min_ctas = env_int("MPSB_JACOBI_QUAD_MIN_CTAS", 1776);      // four waves of 3 CTAs on 148 SMs
    const int nblk = pl.n_pad / 8;
    const bool measured_shape = round_up(pl.ltot, 32) == 256 && pl.ldot == 256 && !pl.real;
    return enabled && (enabled > 1 || measured_shape) && use_modulus_order(pl) && (nblk & 3) == 0 && nblk >= 8 &&
           stamp_ints(pl.n_pad, pl.bsz) > 0 && (long long)count * (nblk / 4) >= min_ctas;
}
int launch_quad_step(const SvdPlan& pl, cplx* X, const SvdState& s, int mat0, int count, int kstep, double tol2,
                     int launch_id, cudaStream_t st) {
    const int ltot_r = round_up(pl.ltot, 32), nblk = pl.n_pad / 8, nsup = nblk / 2;
    const long long strideX = (long long)pl.n_pad * pl.ld;
    const size_t smem = (size_t)16 * ltot_r * sizeof(cplx);
    const bool full = (ltot_r == 256 && pl.ldot == 256);
    auto kern = quad_kernel_for(full, pl.real != 0);
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem)) return rc;
    dim3 grid(mod_cross_pairs(nsup, kstep), count);
    SideStream* side = nullptr;
    if ((kstep & 1) == 0) {
        side = side_stream();
        cudaStream_t ss = st;
        if (side) {
            MPSB_CHECK_CUDA(cudaEventRecord(side->fork, st));
            MPSB_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
            ss = side->stream;
        }
        for (int half = 0; half < 2; ++half) {
            const int round = 2 * kstep + 2 * half;          // blocks kstep + half and kstep + half + nblk/2
#define MPSB_SELF(CPL, EXACT)                                                                                              \
    jacobi_self_kernel<CPL, EXACT><<<count, 128, 0, ss>>>(X, strideX, mat0, pl.ld, pl.ldot, ltot_r, nblk, round, s.rot, s.done, \
                                                          tol2, s.stamps, s.stamp_stride, launch_id + half, pl.real)
            if (full) MPSB_SELF(4, true);
            else if (ltot_r > 128) MPSB_SELF(4, false);
            else if (ltot_r > 64) MPSB_SELF(2, false);
            else MPSB_SELF(1, false);
#undef MPSB_SELF
            MPSB_COUNT_LAUNCH();
            MPSB_CHECK_CUDA(cudaGetLastError());
        }
        kern<<<dim3(2, count), CROSS_THREADS, smem, ss>>>(X, strideX, mat0, pl.ld, pl.ldot, ltot_r, nblk, kstep, 1, s.rot, s.done,
                                                          tol2, s.stamps, s.stamp_stride, launch_id + 2);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
        if (side) MPSB_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
    }
    if (grid.x > 0) {
        kern<<<grid, CROSS_THREADS, smem, st>>>(X, strideX, mat0, pl.ld, pl.ldot, ltot_r, nblk, kstep, 0, s.rot, s.done, tol2,
                                               s.stamps, s.stamp_stride, launch_id);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
    }
    if (side) MPSB_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    return 0;
}

int launch_round_dyn(const SvdPlan& pl, cplx* X, const SvdState& s, int mat0, int count, int round, int full,
                     int inner, double tol2, double eabs2, int launch_id, cudaStream_t st) {
    static const int cross_enabled = env_int("MPSB_JACOBI_CROSS", 1);
    const int ltot_r = round_up(pl.ltot, 32);
    if (cross_enabled && pl.bsz == 8 && ltot_r <= 256 && round > 0 && pl.n_pad / 8 > 2) {
        const size_t smem = (size_t)8 * ltot_r * sizeof(cplx);
        const bool full = (ltot_r == 256 && pl.ldot == 256);
        auto kern = cross_kernel_for(full, pl.real != 0);
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem)) return rc;
        const int nblk = pl.n_pad / 8;
        dim3 grid(nblk / 2, count);
        unsigned long long* pc = g_prof.enabled > 1 ? s.phase_clk : nullptr;
        kern<<<grid, CROSS_THREADS, smem, st>>>(X, (long long)pl.n_pad * pl.ld, mat0, pl.ld, pl.ldot, ltot_r, nblk, round,
                                               s.rot, s.done, tol2, pc, s.stamps, s.stamp_stride, launch_id, 0);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (use_gram_kernel(pl)) return launch_gram(pl, X, s, mat0, count, round, inner, tol2, launch_id, st);
    if (pl.tiled) return launch_gram_tiled(pl, X, s, mat0, count, round, inner, tol2, launch_id, st);
    switch (pl.bsz) {
        case 1: return launch_round<1>(pl, X, s, mat0, count, round, full, inner, tol2, eabs2, st);
        case 2: return launch_round<2>(pl, X, s, mat0, count, round, full, inner, tol2, eabs2, st);
        case 4: return launch_round<4>(pl, X, s, mat0, count, round, full, inner, tol2, eabs2, st);
        case 8: return launch_round<8>(pl, X, s, mat0, count, round, full, inner, tol2, eabs2, st);
        case 16: return launch_round<16>(pl, X, s, mat0, count, round, full, inner, tol2, eabs2, st);
    }
    set_error("svd: unsupported block size %d", pl.bsz);
    return 1;
}

// Cooperative single-launch sweeps (jacobi_persistent_kernel): returns 0 when launched, -1 when the call does not
// qualify (the caller then takes the multi-launch path), > 0 on errors.
template <int CPL, bool FULL, bool REAL>
int launch_persistent_t(const SvdPlan& pl, cplx* X, const SvdState& s, double tol2, cudaStream_t st, bool dry_run) {
    const int ltot_r = round_up(pl.ltot, 32), nblk = pl.n_pad / 8;
    const size_t smem = (size_t)8 * ltot_r * sizeof(cplx);
    auto kern = jacobi_persistent_kernel<CPL, FULL, REAL>;
    const void* fn = reinterpret_cast<const void*>(kern);
    if (int rc = ensure_dynamic_smem(fn, smem)) return rc;
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    MPSB_CHECK_CUDA(cudaGetDevice(&dev));
    MPSB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MPSB_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return -1;
    MPSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CROSS_THREADS, smem));
    if ((long long)per_sm * sms < (long long)(nblk / 2) * pl.batch) return -1;
    if (dry_run) return 0;
    unsigned* sync = reinterpret_cast<unsigned*>(s.stamps);              // zeroed by the caller; 3 ints per problem
    cplx* Xp = X;
    long long strideX = (long long)pl.n_pad * pl.ld;
    int ld = pl.ld, ldot = pl.ldot, ltr = ltot_r, nb = nblk, ms = pl.max_sweeps;
    double t2 = tol2;
    unsigned long long* rt = s.rot_total;
    unsigned long long* swt = s.sweep_total;
    int* seen = s.n_active;
    void* args[] = {&Xp, &strideX, &ld, &ldot, &ltr, &nb, &ms, &t2, &sync, &rt, &swt, &seen};
    MPSB_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(nblk / 2, pl.batch), dim3(CROSS_THREADS), args, smem, st));
    MPSB_COUNT_LAUNCH();
    return 0;
}
int launch_persistent(const SvdPlan& pl, cplx* X, const SvdState& s, double tol2, cudaStream_t st, bool dry_run) {
    static const int enabled = env_int("MPSB_SVD_PERSISTENT", 1);
    static const int max_ctas = env_int("MPSB_SVD_PERSISTENT_CTAS", 148);   // beyond one CTA per SM the launches win
    const int nblk = pl.n_pad / 8, ltot_r = round_up(pl.ltot, 32);
    if (!enabled || !s.stamps || s.stamp_stride < 3 || (nblk / 2) * pl.batch > max_ctas) return -1;
    const bool full = (ltot_r == 256 && pl.ldot == 256);
    if (pl.real) {
        if (full) return launch_persistent_t<4, true, true>(pl, X, s, tol2, st, dry_run);
        if (ltot_r > 128) return launch_persistent_t<4, false, true>(pl, X, s, tol2, st, dry_run);
        if (ltot_r > 64) return launch_persistent_t<2, false, true>(pl, X, s, tol2, st, dry_run);
        return launch_persistent_t<1, false, true>(pl, X, s, tol2, st, dry_run);
    }
    if (full) return launch_persistent_t<4, true, false>(pl, X, s, tol2, st, dry_run);
    if (ltot_r > 128) return launch_persistent_t<4, false, false>(pl, X, s, tol2, st, dry_run);
    if (ltot_r > 64) return launch_persistent_t<2, false, false>(pl, X, s, tol2, st, dry_run);
    return launch_persistent_t<1, false, false>(pl, X, s, tol2, st, dry_run);
}

size_t resident_smem_bytes(int n_pad, int ld) { return (size_t)n_pad * ld * sizeof(cplx); }

template <int NJ>
int launch_resident_t(const SvdPlan& pl, cplx* X, const SvdState& s, double tol2, cudaStream_t st) {
    const size_t smem = resident_smem_bytes(pl.n_pad, pl.ld);
    const void* fn = reinterpret_cast<const void*>(jacobi_resident_kernel<NJ>);
    if (int rc = ensure_dynamic_smem(fn, smem)) return rc;
    const int cap = (NJ >= 1 && NJ <= 4) ? 32 : 16;
    const int nw = std::max(1, std::min(cap, pl.n_pad / 2));
    jacobi_resident_kernel<NJ><<<pl.batch, 32 * nw, smem, st>>>(X, (long long)pl.n_pad * pl.ld, pl.n_pad, pl.ld, pl.ldot,
                                                               pl.max_sweeps, tol2, s.rot_total, s.sweep_total,
                                                               s.n_active, pl.real);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int launch_resident(const SvdPlan& pl, cplx* X, const SvdState& s, double tol2, cudaStream_t st) {
    switch (pl.ld <= 256 ? pl.ld / 32 : 0) {
        case 1: return launch_resident_t<1>(pl, X, s, tol2, st);
        case 2: return launch_resident_t<2>(pl, X, s, tol2, st);
        case 3: return launch_resident_t<3>(pl, X, s, tol2, st);
        case 4: return launch_resident_t<4>(pl, X, s, tol2, st);
        case 5: return launch_resident_t<5>(pl, X, s, tol2, st);
        case 6: return launch_resident_t<6>(pl, X, s, tol2, st);
        case 7: return launch_resident_t<7>(pl, X, s, tol2, st);
        case 8: return launch_resident_t<8>(pl, X, s, tol2, st);
    }
    return launch_resident_t<0>(pl, X, s, tol2, st);
}

}  // namespace

int svd_make_plan(int batch, int n, int ldot, int ltot, SvdPlan* plan) {
    MPSB_REQUIRE(batch >= 1 && n >= 1 && ldot >= 1 && ltot >= ldot, "svd: bad dims batch=%d n=%d ldot=%d ltot=%d",
                 batch, n, ldot, ltot);
    SvdPlan p;
    p.batch = batch;
    p.n = n;
    p.ldot = ldot;
    p.ltot = ltot;
    p.ld = round_up(ltot, 32);
    const size_t rowbytes = (size_t)p.ld * sizeof(cplx);
    const size_t smem_budget = 200 * 1024;
    const int rows_fit = (int)(smem_budget / rowbytes);
    int bcap = env_int("MPSB_JACOBI_BSZ", 8);
    if (bcap < 1) bcap = 1;
    if (bcap > 16) bcap = 16;
    int b = 1;
    while (b * 2 <= bcap && b < (n + 1) / 2) b *= 2;
    // Rows too long for 16 of them to sit in shared memory stream through the tiled Gram kernel (8-row blocks);
    // other block sizes fall back to however many whole rows fit.
    static const int tiled_enabled = env_int("MPSB_JACOBI_TILED", 1);
    p.tiled = (tiled_enabled && b == 8 && gram_smem_bytes(p.ld) > 220 * 1024) ? 1 : 0;
    if (!p.tiled) {
        MPSB_REQUIRE(rows_fit >= 2, "svd: row of %d elements does not fit shared memory", p.ld);
        while (b > 1 && 2 * b > rows_fit) b /= 2;
    }
    p.bsz = b;
    p.n_pad = round_up(n, 2 * b);
    // Tiny problems in small batches: one launch per SVD with the matrix resident in one CTA's shared memory.  One SM
    // has 1/148 of the FP64 throughput, so this only wins while the multi-launch path is bound by launch latency
    // (measured cross-over near 64 x 64: tools/svd_small_probe.py).
    static const int resident_max_batch = env_int("MPSB_SVD_RESIDENT_BATCH", 296);
    static const int resident_max_work = env_int("MPSB_SVD_RESIDENT_WORK", 160000);
    const int n_even = round_up(n, 2);
    p.resident = 0;
    if (batch <= resident_max_batch && (long long)n_even * n_even * p.ld <= resident_max_work &&
        resident_smem_bytes(n_even, p.ld) <= 216 * 1024) {
        p.resident = 1;
        p.bsz = 1;
        p.n_pad = n_even;
    }
    const size_t matbytes = (size_t)p.n_pad * rowbytes;
    const size_t l2_budget = (size_t)env_int("MPSB_SVD_L2_MB", 1 << 20) * 1024 * 1024;     // default: no chunking
    int chunk = (int)std::max<size_t>(1, l2_budget / matbytes);
    chunk = env_int("MPSB_SVD_CHUNK", chunk);
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    if (chunk > 65535) chunk = 65535;
    p.chunk = chunk;
    p.max_sweeps = env_int("MPSB_SVD_MAX_SWEEPS", 40);
    static const int qr_grid_enabled = env_int("MPSB_SVD_QR_GRID", 1);
    static const int qr_grid_min = env_int("MPSB_SVD_QR_GRID_MIN", 160);
    p.qr_grid = (qr_grid_enabled && batch <= 4 && std::min(n, ldot) >= qr_grid_min) ? 1 : 0;
    if (p.qr_grid) p.chunk = batch;          // the cooperative QR factors the whole (tiny) batch in one launch
    p.real = 0;
    p.phase = 0;
    *plan = p;
    return 0;
}

size_t svd_state_bytes(const SvdPlan& plan) {
    return align_up(sizeof(int) * plan.batch, 256) * 2 + align_up(sizeof(double) * plan.batch, 256) + 256 +
           align_up(sizeof(int) * (size_t)plan.batch * stamp_ints(plan.n_pad, plan.bsz), 256) +
           (plan.qr_grid ? qr_grid_scratch_bytes(plan) * plan.batch : 0);
}

namespace {

// QR preconditioning of problems [mat0, mat0 + count)
int launch_qr(const SvdPlan& pl, cplx* X, const SvdState& s, void* state, int mat0, int count, cudaStream_t st) {
    const long long strideX = (long long)pl.n_pad * pl.ld;
    cplx* Xc = X + (size_t)mat0 * strideX;
    double* frob = s.frob2 + mat0;
    int CW = std::min(pl.ld, QR_THREADS);
    int G = QR_THREADS / CW;
    const size_t wbudget = 48 * 1024;       // per w buffer (two are kept)
    G = std::max(1, std::min<int>(G, (int)(wbudget / ((size_t)pl.ld * sizeof(cplx)))));
    G = std::min(G, std::max(1, pl.n));
    const size_t smem = 2 * align_up(pl.n, 8) * sizeof(cplx) + (size_t)2 * G * pl.ld * sizeof(cplx) +
                        align_up(pl.ldot, 8) * sizeof(double) + align_up(pl.ldot, 8) * sizeof(int) + pl.ld + 64;
    static const int qr_enabled = env_int("MPSB_SVD_QR", 1);
    static const int qr_blocked = env_int("MPSB_SVD_QR_BLOCKED", 1);
    const size_t smem_b = align_up(pl.n, 8) * QB_NB * sizeof(cplx) + align_up(pl.ldot, 8) * (sizeof(double) + sizeof(int)) +
                          pl.ld + 64;
    static const int qr_dmma = env_int("MPSB_SVD_QR_DMMA", 1);
    const size_t smem_d = align_up(pl.n, 8) * QD_VS * sizeof(cplx) + align_up(pl.ldot, 8) * (sizeof(double) + sizeof(int)) + 64;
    // Very long rows (chi >= 1024 with carried columns) do not leave room for the reflector scratch; the
    // Jacobi iteration is then started from the raw matrix (more sweeps, same result).
    // The blocked kernel is built for throughput (256 threads, 4 problems per SM); when every problem can have an SM
    // to itself the 1024-thread fused kernel has half the latency (tools/tebd_small_probe.py).
    static const int latency_batch = env_int("MPSB_SVD_QR_LATENCY_BATCH", 148);
    // ... up to ~128 rows; from 256 rows on the DMMA trailing update wins at every batch size (tools/qr_latency_probe.py:
    // pack + QR + finalise of 256 x 256 problems, 16 / 64 / 148 of them: 3.9 / 3.9 / 6.6 ms against 2.8 / 2.8 / 3.4 ms).
    static const int latency_max_n = env_int("MPSB_SVD_QR_LATENCY_MAXN", 191);
    const bool latency_mode = pl.batch <= latency_batch && pl.n <= latency_max_n && smem <= 220 * 1024;
    int grid_rc = -1;
    if (qr_enabled && pl.qr_grid) {          // only for batch <= 4: never chunked
        SvdPlan base = pl;
        base.qr_grid = 0;
        unsigned char* scratch = static_cast<unsigned char*>(state) + svd_state_bytes(base);
        grid_rc = launch_qr_grid(pl, X, s.frob2, scratch, st);
        if (grid_rc > 0) return grid_rc;
    }
    if (grid_rc == 0) {
    } else if (qr_enabled && qr_blocked && qr_dmma && !latency_mode && smem_d <= 200 * 1024 &&
               align_up(pl.n, 8) * QD_VS >= (size_t)pl.ldot) {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(qr_panel_dmma_kernel), smem_d)) return rc;
        qr_panel_dmma_kernel<<<count, QB_THREADS, smem_d, st>>>(Xc, strideX, pl.n, pl.ldot, pl.ltot, pl.ld, frob);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
    } else if (qr_enabled && qr_blocked && !latency_mode && smem_b <= 200 * 1024) {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(qr_blocked_kernel), smem_b)) return rc;
        qr_blocked_kernel<<<count, QB_THREADS, smem_b, st>>>(Xc, strideX, pl.n, pl.ldot, pl.ltot, pl.ld, frob);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
    } else if (qr_enabled && smem <= 220 * 1024) {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(qr_precond_kernel), smem)) return rc;
        qr_precond_kernel<<<count, QR_THREADS, smem, st>>>(Xc, strideX, pl.n, pl.ldot, pl.ltot, pl.ld, G, CW, frob);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace

// QR -> sweeps until a sweep applies no rotation, chunk by chunk (one chunk unless MPSB_SVD_L2_MB / MPSB_SVD_CHUNK ask
// for more: cutting the batch into L2-sized chunks, even with several chunks in flight on separate streams, was
// measured to be slower - a round of a small chunk cannot fill the GPU and ~3700 dependent launches per chunk add up).
// Convergence is polled (4-byte D2H + stream sync) only from two sweeps before the count the previous call needed.
int svd_orthogonalise(const SvdPlan& pl, cplx* X, void* state, cudaStream_t st, int* sweeps_out) {
    const SvdState s = carve_state(state, pl.batch, stamp_ints(pl.n_pad, pl.bsz));
    if (s.stamps) MPSB_CHECK_CUDA(cudaMemsetAsync(s.stamps, 0, sizeof(int) * (size_t)pl.batch * s.stamp_stride, st));
    sweep_begin_kernel<<<(pl.batch + 255) / 256, 256, 0, st>>>(s.rot, s.done, s.n_active, s.rot_total, s.sweep_total,
                                                               pl.batch, 1);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());

    const double eps = 2.220446049250313e-16;
    static const double tol_factor = (double)env_int("MPSB_SVD_TOL_X10", 10) / 10.0;
    const double tol = tol_factor * sqrt((double)pl.ldot) * eps;
    const double eabs = 0.0;   // purely relative criterion: kept right singular vectors must be orthonormal to ~tol
    const double tol2 = tol * tol, eabs2 = eabs * eabs;
    const int nblk = pl.n_pad / pl.bsz;
    const bool modulus = use_modulus_order(pl);
    const bool quad = modulus && use_quad_order(pl, std::min(pl.chunk, pl.batch));
    const int rounds = quad ? nblk / 2 : modulus ? nblk : nblk - 1;
    const int inner = (nblk == 2) ? pl.max_sweeps : 1;
    int max_sweeps_seen = 0;
    int launch_id = 1;
    const int trace = env_int("MPSB_SVD_TRACE", 0);

    if (pl.phase == 1) {                                     // QR only (real path: the sweeps run on the packed copy)
        if (int rc = launch_qr(pl, X, s, state, 0, pl.batch, st)) return rc;
        prof_mark(3, st);
        if (sweeps_out) *sweeps_out = 0;
        return 0;
    }
    const bool with_qr = pl.phase == 0;
    if (pl.resident) {
        if (with_qr) {
            if (int rc = launch_qr(pl, X, s, state, 0, pl.batch, st)) return rc;
            prof_mark(3, st);
        }
        if (int rc = launch_resident(pl, X, s, tol2, st)) return rc;
        MPSB_CHECK_CUDA(cudaMemcpyAsync(&max_sweeps_seen, s.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
        MPSB_CHECK_CUDA(cudaStreamSynchronize(st));
    } else if (modulus && pl.chunk >= pl.batch && !trace && launch_persistent(pl, X, s, tol2, st, true) == 0) {
        // few problems: all sweeps in one cooperative launch (no per-step launch latency, no host poll per sweep)
        if (with_qr) {
            if (int rc = launch_qr(pl, X, s, state, 0, pl.batch, st)) return rc;
            prof_mark(3, st);
        }
        MPSB_CHECK_CUDA(cudaMemsetAsync(s.n_active, 0, sizeof(int), st));
        if (int rc = launch_persistent(pl, X, s, tol2, st, false)) {
            if (rc < 0) set_error("svd: the persistent sweep kernel became unavailable between check and launch");
            return rc < 0 ? 2 : rc;
        }
        MPSB_CHECK_CUDA(cudaMemcpyAsync(&max_sweeps_seen, s.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
        MPSB_CHECK_CUDA(cudaStreamSynchronize(st));
    } else {
        const int nchunks = (pl.batch + pl.chunk - 1) / pl.chunk;
        // With one chunk the QR of the batch is timed as its own phase; otherwise QR and sweeps of different chunks
        // alternate and the "qr" phase of the profile reads zero (its time is inside "jacobi").
        if (nchunks > 1 && with_qr) prof_mark(3, st);
        static int last_sweeps = 0;                       // sweeps the previous call needed (any shape: a hint only)
        const int poll_from = std::max(0, last_sweeps - 2);
        for (int mat0 = 0; mat0 < pl.batch; mat0 += pl.chunk) {
            const int count = std::min(pl.chunk, pl.batch - mat0);
            if (with_qr) {
                if (int rc = launch_qr(pl, X, s, state, mat0, count, st)) return rc;
                if (nchunks == 1) prof_mark(3, st);
            }
            int sweep = 0;
            for (; sweep < pl.max_sweeps; ++sweep) {
                for (int r = 0; r < rounds; ++r) {
                    int rc = quad ? launch_quad_step(pl, X, s, mat0, count, r, tol2, launch_id += 8, st)
                             : modulus ? launch_modulus_step(pl, X, s, mat0, count, r, tol2, ++launch_id, st)
                                       : launch_round_dyn(pl, X, s, mat0, count, r, r == 0, inner, tol2, eabs2, ++launch_id, st);
                    if (rc) return rc;
                }
                MPSB_CHECK_CUDA(cudaMemsetAsync(s.n_active, 0, sizeof(int), st));
                sweep_end_kernel<<<(count + 255) / 256, 256, 0, st>>>(s.rot, s.done, s.n_active, s.rot_total, s.sweep_total,
                                                                      mat0, count);
                MPSB_COUNT_LAUNCH();
                MPSB_CHECK_CUDA(cudaGetLastError());
                if (!(sweep >= poll_from || sweep + 1 == pl.max_sweeps || trace)) continue;
                int n_active = 0;
                MPSB_CHECK_CUDA(cudaMemcpyAsync(&n_active, s.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
                MPSB_CHECK_CUDA(cudaStreamSynchronize(st));
                if (trace) {
                    unsigned long long rt = 0;
                    MPSB_CHECK_CUDA(cudaMemcpy(&rt, s.rot_total, sizeof(rt), cudaMemcpyDeviceToHost));
                    fprintf(stderr, "[mpsb] svd trace: chunk %d sweep %2d: %d of %d problems still rotating, %llu rotations "
                                    "so far\n", mat0 / pl.chunk, sweep, n_active, count, rt);
                }
                if (n_active == 0) { ++sweep; break; }
            }
            max_sweeps_seen = std::max(max_sweeps_seen, sweep);
        }
        last_sweeps = max_sweeps_seen;
    }
    if (sweeps_out) *sweeps_out = max_sweeps_seen;
    prof_mark(4, st);
    if (g_prof.enabled) {
        unsigned long long h[2] = {0ull, 0ull};
        MPSB_CHECK_CUDA(cudaMemcpyAsync(&h[0], s.rot_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        MPSB_CHECK_CUDA(cudaMemcpyAsync(&h[1], s.sweep_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        MPSB_CHECK_CUDA(cudaStreamSynchronize(st));
        g_prof.rotations = h[0];
        g_prof.problem_sweeps = h[1];
        if (g_prof.enabled > 1) {
            unsigned long long c[8];
            MPSB_CHECK_CUDA(cudaMemcpy(c, s.phase_clk, sizeof(c), cudaMemcpyDeviceToHost));
            const double nc = (double)(c[3] ? c[3] : 1);
            fprintf(stderr, "[mpsb] jacobi CTA phases (mean SM clocks): load %.0f rotate %.0f store %.0f over %llu CTAs"
                            " | gram %.0f tournament %.0f update %.0f | %.0f\n",
                    (double)c[0] / nc, (double)c[1] / nc, (double)c[2] / nc, c[3], (double)c[4] / nc, (double)c[5] / nc,
                    (double)c[6] / nc, (double)c[7] / nc);
        }
    }
    return 0;
}

// Host-side listing of one sweep of the block orderings, built from the same pairing helpers the kernels and launchers
// use (mod_pair, mod_cross_pairs, the self-task rounds of launch_modulus_step / launch_quad_step): entries
// (step, stream, block I, block J), stream 0 = the step's main launch, 1 = its side stream, I == J = self task of a
// block.  For the CPU tests of the schedule (every block pair once per sweep, no block touched twice by concurrent
// launches of a step); returns the number of entries, or -1 if `cap` is too small.
int sweep_schedule(int nblk, int quad, int* out, int cap) {
    int n = 0;
    auto emit = [&](int step, int stream, int bI, int bJ) {
        if (n < cap) { out[4 * n] = step; out[4 * n + 1] = stream; out[4 * n + 2] = bI; out[4 * n + 3] = bJ; }
        ++n;
    };
    if (!quad) {
        for (int k = 0; k < nblk; ++k) {
            if ((k & 1) == 0) { emit(k, 1, k >> 1, k >> 1); emit(k, 1, (k >> 1) + (nblk >> 1), (k >> 1) + (nblk >> 1)); }
            for (int slot = 0; slot < mod_cross_pairs(nblk, k); ++slot) {
                int bI, bJ;
                mod_pair(nblk, k, slot, bI, bJ);
                emit(k, 0, bI, bJ);
            }
        }
    } else {
        const int nsup = nblk / 2;
        for (int k = 0; k < nsup; ++k) {
            if ((k & 1) == 0) {
                for (int half = 0; half < 2; ++half) {               // jacobi_self_kernel with round = 2k + 2 half
                    const int bA = (2 * k + 2 * half) >> 1;
                    emit(k, 1, bA, bA);
                    emit(k, 1, bA + (nblk >> 1), bA + (nblk >> 1));
                }
                for (int x = 0; x < 2; ++x) {                        // jacobi_quad_kernel, inner = 1
                    const int b0 = 2 * ((k >> 1) + x * (nsup >> 1));
                    emit(k, 1, b0, b0 + 1);
                }
            }
            for (int slot = 0; slot < mod_cross_pairs(nsup, k); ++slot) {
                int SI, SJ;
                mod_pair(nsup, k, slot, SI, SJ);
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) emit(k, 0, 2 * SI + a, 2 * SJ + b);
            }
        }
    }
    return n <= cap ? n : -1;
}

int svd_finalise(const SvdPlan& pl, const cplx* X, void* state, const SvdOutputs& out, cudaStream_t st) {
    (void)state;
    MPSB_REQUIRE(out.s && out.rank && out.k_cap >= 1, "svd_finalise: s, rank and k_cap are required");
    int npow2 = 1;
    while (npow2 < pl.n_pad) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * (sizeof(double) + sizeof(int));
    MPSB_REQUIRE(smem <= 200 * 1024, "svd_finalise: %d rows exceed the sort buffer", pl.n_pad);
    if (smem > 48 * 1024)
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(finalise_kernel), smem)) return rc;
    finalise_kernel<<<pl.batch, FIN_THREADS, smem, st>>>(X, (long long)pl.n_pad * pl.ld, pl.n, pl.n_pad, npow2,
                                                        pl.ldot, pl.ltot, pl.ld, out);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mpsb
