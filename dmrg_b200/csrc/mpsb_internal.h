// Internal (C++) interfaces between the translation units of libmps_b200.  The public surface is
// include/mpsb.h; nothing here crosses the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "common.cuh"

namespace mpsb {

// ---- optional phase timing (api.cu); bench.py reads it through mpsb_profile_* -------------------
// Phase boundaries of mpsb_tebd_two_site: 0 start | 1 theta GEMM | 2 gate | 3 QR | 4 Jacobi sweeps |
// 5 finalise | 6 back-projection GEMM.  CUDA events on the caller's stream, only when enabled.
struct Profile {
    int enabled;
    cudaEvent_t ev[8];
    int have_events;
    unsigned recorded;                  // bit i: event i was recorded since profiling was switched on
    unsigned long long rotations;       // Jacobi rotations applied in the last SVD
    unsigned long long problem_sweeps;  // sum over problems of sweeps executed in the last SVD
};
extern Profile g_prof;
void prof_mark(int idx, cudaStream_t st);

// ---- batched complex128 GEMM (zgemm.cu) ------------------------------------------------------
// C[z1,z2] = op(A[z1,z2]) * op(B[z1,z2]) (+ C[z1,z2]); op: 0 = N, 1 = T, 2 = C (conj-transpose),
// 3 = J (conjugate, no transpose).  M, N, K are the dimensions of the product (after op).
template <typename T>
struct GemmArgsT {
    const T* A;
    const T* B;
    T* C;
    int M, N, K;
    int lda, ldb, ldc;
    int opA, opB;
    int batch1, batch2;            // grid.z = batch1 * batch2, z1 = z / batch2, z2 = z % batch2
    long long sA1, sA2, sB1, sB2, sC1, sC2;  // element strides of the two batch levels
    int accumulate;
};
typedef GemmArgsT<cplx> GemmArgs;
int zgemm(const GemmArgs& p, cudaStream_t st);
// ---- batched real (double) GEMM on the same DMMA tiles (dgemm.cu) --------------------------------
// The real-arithmetic twin used when the MPO and the state are real (the reference's
// MPO.tomatrix(dtype='double') path, DMRG/MPO.py:68-76): a quarter of the flops of the complex product.
// op: 0 / 3 = N, 1 / 2 = T (conjugation is a no-op).
int dgemm(const GemmArgsT<double>& p, cudaStream_t st);
inline int gemm(const GemmArgsT<cplx>& p, cudaStream_t st) { return zgemm(p, st); }
inline int gemm(const GemmArgsT<double>& p, cudaStream_t st) { return dgemm(p, st); }

// ---- truncated SVD by QR-preconditioned one-sided Jacobi (jacobi_svd.cu) -----------------------
// The work matrix X of one problem is n_pad x ld complex128, row-major.  Its first `ldot` columns
// hold the matrix m whose rows get orthogonalised (Q^H m = S Vh); columns [ldot, ltot) ride along
// (they receive the same left rotations), which is how U (identity carried) or the gauge carry
// Vh.M[i+1] are produced without a separate accumulation.
struct SvdPlan {
    int batch;       // number of problems
    int n;           // real rows of m
    int n_pad;       // rows of X, multiple of 2*bsz
    int ldot;        // columns entering the inner products
    int ltot;        // columns rotated (>= ldot)
    int ld;          // row pitch of X in elements, multiple of 32, >= ltot
    int bsz;         // rows per Jacobi block (power of two, 1..16)
    int chunk;       // problems swept together (L2 residency)
    int max_sweeps;
    int qr_grid;     // 1: few large problems - the QR of each is spread over many CTAs (cooperative launch)
    int tiled;       // 1: rows longer than shared memory allows stream through the tiled Gram kernel
    int resident;    // 1: the whole work matrix of a problem lives in one CTA's shared memory (single launch)
    int real;        // 1: rows hold a REAL matrix, adjacent column pairs packed into complex elements; rotations are real
    int phase;       // 0: QR + sweeps; 1: QR only; 2: sweeps only (the real path runs the QR on the unpacked matrix)
};
int svd_make_plan(int batch, int n, int ldot, int ltot, SvdPlan* plan);
size_t svd_state_bytes(const SvdPlan& plan);   // per-call scratch besides X (flags, counters, order)
// Runs QR preconditioning + Jacobi sweeps on X in place.  `state` is svd_state_bytes() of scratch.
int svd_orthogonalise(const SvdPlan& plan, cplx* X, void* state, cudaStream_t st, int* sweeps_out);
// Sorts rows by norm, applies the reference truncation rule (DMRG/MPS.py:33-38) and scatters:
//   vh_out[b][i][0:ldot]   = X[row_i][0:ldot] / s_i          (i < k; rows k..k_cap-1 zeroed)
//   aug_out[b][i][0:ltot-ldot] = X[row_i][ldot:ltot]         (i < k; rows k..k_cap-1 zeroed)
//   s_out[b][i]            = normalised kept values (zeros beyond k),  s_raw[b][i] = raw row norms
struct SvdOutputs {
    cplx* vh;   long long vh_stride;  int vh_ld;    // may be null
    cplx* aug;  long long aug_stride; int aug_ld;   // may be null
    int aug_conj;                                    // conjugate the carried part on output
    double* s;  long long s_stride;                  // k_cap values per problem
    double* s_raw; long long s_raw_stride; int n_raw;  // may be null; n_raw values per problem
    int* rank;                                       // one int per problem
    int k_cap;                                       // rows written per problem (>= trun not required)
    int trun; double err;
    int normalise;                                   // 1: reference `truncate`; 0: iDMRG `halve` (no eps, raw S)
    int real;                                        // packed real rows: vh / aug are arrays of doubles, strides in doubles
    int real_cols, real_naug;                        // real columns of the dot part / of the carried part
};
int svd_finalise(const SvdPlan& plan, const cplx* X, void* state, const SvdOutputs& out, cudaStream_t st);
int sweep_schedule(int nblk, int quad, int* out, int cap);      // host listing of one sweep's block pairs (tests)

// ---- small fused element-wise kernels (tebd.cu, gauge.cu, heff.cu) ------------------------------
int gate_scale(int batch, int chil, int chir, int d, const cplx* theta0, long long s_theta0,
               const cplx* U, long long s_U, const double* Sl, long long s_Sl, cplx* thetaB,
               long long s_thetaB, cplx* X, long long s_X, int ldX, cudaStream_t st);

int one_site(int batch, int chil, int chir, int d, const cplx* M, long long sM, const cplx* U, long long sU,
             cplx* Mo, long long sMo, cudaStream_t st);

int mpo_site(int batch, int chil, int chir, int d, int wl, int wr, const cplx* M, long long sM, const cplx* W,
             long long sW, cplx* out, long long sOut, cudaStream_t st);

// ---- layout helpers (aux.cu) ---------------------------------------------------------------------
// Describes how an (i, j) element of a logical matrix is fetched from memory:
//   value = conj?( src[i * rs + j * cs] ) * (scale ? scale[(j / sdiv) % smod] : 1)
struct MatSrc {
    const cplx* src;
    long long rs, cs;
    int conj;
    const double* scale;   // may be null; indexed by column j as (j / sdiv) % smod
    int sdiv, smod;
    long long bs = 0;       // element stride of `src` between the problems of a batch
    long long scale_bs = 0; // element stride of `scale` between the problems of a batch
};
// X[i][0:ldot] <- dot (rows x ldot), X[i][ldot:ltot] <- aug (rows x (ltot-ldot)) or identity when
// aug.src == nullptr and identity != 0; everything else in the n_pad x ld buffer is zeroed.
// Batched: problem b reads dot.src + b * dot.bs (scale + b * scale_bs), writes X + b * n_pad * ld - one launch.
int pack_work_matrix(cplx* X, int n_pad, int ld, int rows, int ldot, int ltot, const MatSrc& dot, const MatSrc& aug,
                     int identity, cudaStream_t st, int batch = 1);
// Real path: X[b][i][j] = (A[b][i][j], 0) for j < cols, (1, 0) at j = ldot + i when identity, zero elsewhere.
int pack_real_matrix(cplx* X, int n_pad, int ld, int rows, int cols, int ldot, const double* A, int lda, long long sA,
                     int identity, cudaStream_t st, int batch);
// Xp[b][i][j] = (Re X[b][i][2j], Re X[b][i][2j+1]) for i < rows, j < ltot_p; zero in the padding of the packed matrix.
int compact_real_pairs(const cplx* X, int n_pad, int ld, int rows, cplx* Xp, int n_pad_p, int ld_p, int ltot_p,
                       cudaStream_t st, int batch);
// out[j * ldo + i] = conj?(in[i * ldi + j]),  i < rows, j < cols; batched with element strides s_in / s_out
int transpose_conj(const cplx* in, int ldi, int rows, int cols, cplx* out, int ldo, int conj, cudaStream_t st,
                   int batch = 1, long long s_in = 0, long long s_out = 0);
int transpose_conj(const double* in, int ldi, int rows, int cols, double* out, int ldo, int conj, cudaStream_t st,
                   int batch = 1, long long s_in = 0, long long s_out = 0);
// out[i][j] = s[i / group] * in[i][j]
int scale_rows(int rows, int cols, const cplx* in, int ldi, const double* s, int group, cplx* out, int ldo,
               cudaStream_t st);
// out[c][a][b] = in[a][b][c]  (n_ab = a*b merged)   and the inverse
int perm_last_to_first(const cplx* in, int n_ab, int n_c, cplx* out, cudaStream_t st);
int perm_first_to_last(const cplx* in, int n_c, int n_ab, cplx* out, cudaStream_t st);
// The BLAS-1 helpers below take `batch` independent problems in one launch (the chains of a batched Lanczos run):
// problem b uses V + b sV, w + b sw, the coefficient / result vector + b sc and the scratch + b spart.
struct BlasBatch {
    int batch = 1;
    long long sV = 0, sw = 0, sc = 0, spart = 0;
};
// y[j] = sum_i conj(V[j][i]) w[i]  for j < m (deterministic, single launch).  part: m_cap * nchunk partial sums
// followed by m_cap ticket counters (unsigned int) that must be zero on entry; they are zero again on exit.
int zdotc_multi(const cplx* V, long long ldv, int m, long long n, const cplx* w, cplx* y, cplx* part, int nchunk,
                int m_cap, cudaStream_t st, BlasBatch bb = BlasBatch());
int zdotc_multi(const double* V, long long ldv, int m, long long n, const double* w, double* y, double* part, int nchunk,
                int m_cap, cudaStream_t st, BlasBatch bb = BlasBatch());      // real twins of the three helpers
// w[i] += sign * sum_j c[j] V[j][i]   (c on device, m entries)
int zaxpy_multi(const cplx* V, long long ldv, int m, long long n, const cplx* c, double sign, cplx* w,
                cudaStream_t st, BlasBatch bb = BlasBatch());
int zaxpy_multi(const double* V, long long ldv, int m, long long n, const double* c, double sign, double* w,
                cudaStream_t st, BlasBatch bb = BlasBatch());
// w *= 1/sqrt(nrm2[0].x) with nrm2 on the device (zero vector if the norm vanishes)
int znormalise(long long n, const cplx* nrm2, cplx* w, cudaStream_t st, BlasBatch bb = BlasBatch());
int znormalise(long long n, const double* nrm2, double* w, cudaStream_t st, BlasBatch bb = BlasBatch());
// w[i] = alpha * w[i]  (alpha real, on host)
int zdscal(long long n, double alpha, cplx* w, cudaStream_t st);

}  // namespace mpsb
