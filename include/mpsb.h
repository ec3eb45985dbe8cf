/* libmps_b200 — C ABI of the B200-native matrix-product-state two-site update.
 *
 * The reference (peijunz-archive/DMRG) is pure Python and has no FFI of its own; the boundary it
 * offers is its Python API.  Each entry point below replaces the numerical body of
 * one reference function; the Python shim in dmrg_b200/ keeps the reference's names and signatures
 * and reaches these symbols through ctypes (see INTEGRATION.md for the binding a maintainer of the
 * reference would add).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; mpsb_last_error() gives the message;
 *  - all pointers are DEVICE pointers unless the name ends in `_host`; complex numbers are numpy's
 *    complex128 layout (interleaved re, im doubles); tensors are C-contiguous unless a leading
 *    dimension / stride is passed; strides are in ELEMENTS of the pointed-to type;
 *  - the library never allocates device memory: scratch comes from the caller (`ws`, sized by the
 *    matching *_workspace() query) — PyTorch tensors are the allocator;
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous
 *    except where noted (the SVD polls a convergence counter once per Jacobi sweep);
 *  - thread-safety: one host thread per (device, stream).
 */
#ifndef MPSB_H
#define MPSB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

int mpsb_version(void);
const char* mpsb_last_error(void);
/* number of kernels launched by the library in this process (bench.py: gpu_launches) */
unsigned long long mpsb_launch_count(void);

/* Batched complex128 GEMM on FP64 tensor cores:  C[b] = op(A[b]) op(B[b]) (+ C[b]).
 * op codes: 0 = N, 1 = T, 2 = C (conjugate transpose), 3 = J (conjugate).  Two batch levels
 * (batch1 x batch2) with independent element strides.  Replaces the reference's multi-operand
 * np.einsum loops: DMRG/MPS.py:264,284,325,351-355,384-388,423,445,454; DMRG/iDMRG.py:40,59,60. */
int mpsb_zgemm(int opA, int opB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C,
               int ldc, int batch1, int batch2, long long sA1, long long sA2, long long sB1, long long sB2,
               long long sC1, long long sC2, int accumulate, void* stream);

/* The same for real matrices of doubles (op 0 / 3 = N, 1 / 2 = T): the real-arithmetic path the reference takes when
 * the MPO is built with MPO.tomatrix(dtype='double') (DMRG/MPO.py:68-76) - L, W, R and the two-site tensor are then
 * real and the contractions of DMRG/iDMRG.py:40,59,60 cost a quarter of the complex flops. */
int mpsb_dgemm(int opA, int opB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C,
               int ldc, int batch1, int batch2, long long sA1, long long sA2, long long sB1, long long sB2,
               long long sC1, long long sC2, int accumulate, void* stream);

/* Truncated SVD  m = U diag(s) Vh  of `batch` matrices (rows x cols, row pitch lda, batch stride sA).
 * Replaces svd_truncate / truncate (DMRG/MPS.py:14-45) when normalise = 1:
 *     s /= s[0];  k = min(#{s > err}, trun);  s = s[:k] / ||s[:k]||
 * and halve's SVD (DMRG/iDMRG.py:23-31) when normalise = 0:  k = min(trun, min(rows, cols)), raw s.
 * Outputs (k_cap = rows of capacity per problem; entries beyond the rank k are zero):
 *     Vh  [batch][k_cap][cols]   (pitch cols)           — always
 *     Uh  [batch][k_cap][rows]   = U^H (pitch rows)      — if Uh != NULL
 *     S   [batch][k_cap]          kept values (normalised when normalise = 1)
 *     S_raw [batch][min(rows,cols)] all singular values, unnormalised — if S_raw != NULL
 *     rank[batch]
 * Singular vectors carry an arbitrary phase per singular value (as LAPACK's do). */
size_t mpsb_svd_trunc_workspace(int batch, int rows, int cols, int want_u);
int mpsb_svd_trunc(int batch, int rows, int cols, const void* A, int lda, long long sA, int trun, double err,
                   int normalise, int k_cap, void* Vh, void* Uh, double* S, double* S_raw, int* rank, void* ws,
                   size_t ws_bytes, void* stream);

/* The same for REAL matrices (A, Vh, Uh are arrays of doubles; lda, sA in doubles): the split of the real iDMRG path
 * (halve on a real two-site tensor, DMRG/iDMRG.py:17-35 with MPO.tomatrix(dtype='double'), DMRG/MPO.py:68-76).
 * Householder QR on the matrix itself, then the Jacobi sweeps on rows whose adjacent real columns are packed
 * pairwise into complex elements (half the row length, real rotations). */
size_t mpsb_svd_trunc_real_workspace(int batch, int rows, int cols, int want_u);
int mpsb_svd_trunc_real(int batch, int rows, int cols, const void* A, int lda, long long sA, int trun, double err,
                        int normalise, int k_cap, void* Vh, void* Uh, double* S, double* S_raw, int* rank, void* ws,
                        size_t ws_bytes, void* stream);

/* TEBD two-site update, batched over independent bonds (MPS.update_double, DMRG/MPS.py:428-457):
 *     thetaB[l,a,b,k] = sum_{c,j,r} Bi[l,c,r] Bj[r,j,k] U[a,b,c,j]          (:445)
 *     m = Sl[l] * thetaB;  (s, Vh) = svd_truncate(m as (chil*d) x (d*chir), trun, err)   (:446-449)
 *     Bi'[l,a,m] = sum_{b,k} thetaB[l,a,b,k] conj(Vh[m,b,k]);   Bj' = Vh;   Sr = s       (:453-455)
 * Shapes (k_cap = capacity of the new middle bond, k_cap >= min(trun, d*chil, d*chir) recommended):
 *     Bi [chil][d][chi]  Bj [chi][d][chir]  Sl [chil]  U [d][d][d][d]      (batch strides sBi, sBj, sSl, sU;
 *                                                                            sU = 0 shares one gate)
 *     Bi_out [chil][d][k_cap]   Bj_out [k_cap][d][chir]   Sr_out [k_cap]   rank [1]   (strides sBio, sBjo, sSr)
 * Outputs may alias the inputs (the inputs are consumed before anything is written).
 * Entries beyond the new rank are written as zeros, so zero-padded tensors stay zero-padded and the
 * result does not depend on the padding. */
size_t mpsb_tebd_two_site_workspace(int batch, int chil, int chi, int chir, int d);
int mpsb_tebd_two_site(int batch, int chil, int chi, int chir, int d, const void* Bi, long long sBi, const void* Bj,
                       long long sBj, const double* Sl, long long sSl, const void* U, long long sU, int trun,
                       double err, int k_cap, void* Bi_out, long long sBio, void* Bj_out, long long sBjo,
                       double* Sr_out, long long sSr, int* rank, void* ws, size_t ws_bytes, void* stream);

/* One-site gate (MPS.update_single, DMRG/MPS.py:423):  M'[l,e,r] = sum_c U[e,c] M[l,c,r].  In place allowed. */
int mpsb_one_site(int batch, int chil, int chir, int d, const void* M, long long sM, const void* U, long long sU,
                  void* M_out, long long sMo, void* stream);

/* Matrix product operator applied to one site tensor (BMPS.update_MPO, transmit/bhopping.py:94-104, the einsum
 * "ijk,lmjo->ilokm" followed by the reshape to (chil*wl, d, chir*wr)):
 *     out[(i,l), o, (k,m)] = sum_j M[i,j,k] W[l,m,j,o]
 * M [chil][d][chir], W [wl][wr][d][d] (left MPO bond, right MPO bond, input leg, output leg), out
 * [chil*wl][d][chir*wr]; strides in elements between batch items (sW = 0 shares one W).  out must not alias M. */
int mpsb_mpo_site(int batch, int chil, int chir, int d, int wl, int wr, const void* M, long long sM, const void* W,
                  long long sW, void* out, long long sOut, void* stream);

/* Gauge steps of MPS.canon (DMRG/MPS.py:246-284).  Left step on site i:
 *     m = (Sl * M_i) as (chil*d) x chir;  u, s, vh = svd_truncate(m);  M_i <- u;  Sr <- s;  M_next <- vh . M_next
 * Right step on site i:
 *     m = (M_i * Sr) as chil x (d*chir);  M_i <- vh;  Sl <- s;  M_prev <- M_prev . u
 * M_next is [chir][nnext] (any trailing shape flattened), M_prev is [nprev][chil].
 * Outputs are written with the new bond dimension padded to k_cap (zeros beyond the rank). */
size_t mpsb_gauge_workspace(int chil, int chir, int d, int nother, int k_cap);
int mpsb_gauge_left(int chil, int chir, int d, const void* M, const double* Sl, const void* M_next, int nnext,
                    int trun, double err, int k_cap, void* M_out, double* Sr_out, void* M_next_out, int* rank,
                    void* ws, size_t ws_bytes, void* stream);
int mpsb_gauge_right(int chil, int chir, int d, const void* M, const double* Sr, const void* M_prev, int nprev,
                     int trun, double err, int k_cap, void* M_out, double* Sl_out, void* M_prev_out, int* rank,
                     void* ws, size_t ws_bytes, void* stream);

/* Batched forms for `batch` independent chains whose tensors share one (zero-padded) shape - ChainBatch.canon
 * (the reference canonicalises chain by chain: DMRG/MPS.py:295-302 inside a Python loop over MPS objects).  s* are
 * element strides between consecutive chains; Sr_out / Sl_out / rank hold one entry (row) per chain. */
size_t mpsb_gauge_batched_workspace(int batch, int chil, int chir, int d, int nother, int k_cap);
int mpsb_gauge_left_batched(int batch, int chil, int chir, int d, const void* M, long long sM, const double* Sl,
                            long long sSl, const void* M_next, long long sMn, int nnext, int trun, double err, int k_cap,
                            void* M_out, long long sMo, double* Sr_out, long long sSr, void* M_next_out, long long sMno,
                            int* rank, void* ws, size_t ws_bytes, void* stream);
int mpsb_gauge_right_batched(int batch, int chil, int chir, int d, const void* M, long long sM, const double* Sr,
                             long long sSr, const void* M_prev, long long sMp, int nprev, int trun, double err, int k_cap,
                             void* M_out, long long sMo, double* Sl_out, long long sSl, void* M_prev_out, long long sMpo,
                             int* rank, void* ws, size_t ws_bytes, void* stream);

/* Transfer-matrix step shared by MPS.dot and MPS.corr (DMRG/MPS.py:351-355, 384-388):
 *     E'[o,r] = sum_{k,l,i,j} E[k,l] conj(A[k,i,o]) Op[i,j] B[l,j,r]      (Op = NULL means identity)
 * A is [ka][d][oa], B is [kb][d][rb], E is [ka][kb], E' is [oa][rb]. */
size_t mpsb_transfer_workspace(int ka, int kb, int d, int oa, int rb);
int mpsb_transfer_step(int ka, int kb, int d, int oa, int rb, const void* E, const void* A, const void* B,
                       const void* Op, void* E_out, void* ws, size_t ws_bytes, void* stream);

/* The same for `batch` chains (ChainBatch.corr / measure, DMRG/MPS.py:358-410 per chain in the reference);
 * sOp = 0 shares one operator between the chains. */
size_t mpsb_transfer_batched_workspace(int batch, int ka, int kb, int d, int oa, int rb);
int mpsb_transfer_step_batched(int batch, int ka, int kb, int d, int oa, int rb, const void* E, long long sE, const void* A,
                               long long sA, const void* B, long long sB, const void* Op, long long sOp, void* E_out,
                               long long sEo, void* ws, size_t ws_bytes, void* stream);

/* Matrix-free effective-Hamiltonian product (replaces building DMRG/iDMRG.py:38-40 and the dense
 * matvec inside eigsh, :57):
 *     out[A,B,C,D] = sum L[A,a,i] W[i,j,B,b] W[j,k,C,c] R[D,d,k] theta[a,b,c,d]
 * L [chil][chil][w], R [chir][chir][w], W [w][w][d][d], theta/out [chil][d][d][chir]. */
size_t mpsb_heff_workspace(int chil, int chir, int d, int w);
int mpsb_heff_matvec(int chil, int chir, int d, int w, const void* L, const void* W, const void* R, const void* theta,
                     void* out, void* ws, size_t ws_bytes, void* stream);

/* `batch` independent chains of equal shape in one call (coupling sweeps over h/J, BASELINE config 4; the reference
 * runs one next_LR chain per Python process, DMRG/iDMRG.py:64-76).  s* = element strides between chains; sW = 0
 * shares one MPO between all chains. */
size_t mpsb_heff_batched_workspace(int batch, int chil, int chir, int d, int w);
int mpsb_heff_matvec_batched(int batch, int chil, int chir, int d, int w, const void* L, long long sL, const void* W,
                             long long sW, const void* R, long long sR, const void* theta, long long s_theta, void* out,
                             long long s_out, void* ws, size_t ws_bytes, void* stream);

/* Real-arithmetic twin (L, W, R, theta, out are tensors of doubles; strides in doubles; the workspace query of the
 * complex call is an upper bound). */
int mpsb_heff_matvec_batched_real(int batch, int chil, int chir, int d, int w, const void* L, long long sL, const void* W,
                                  long long sW, const void* R, long long sR, const void* theta, long long s_theta,
                                  void* out, long long s_out, void* ws, size_t ws_bytes, void* stream);

/* Lowest eigenpair of the effective Hamiltonian by Lanczos with full reorthogonalisation and
 * restarts (replaces eigsh(k=1, which='SA'), DMRG/iDMRG.py:57).  theta holds the start vector on
 * entry (must be non-zero) and the normalised ground vector on exit.  Results (host pointers):
 * energy_host, resid_host (||H v - E v||), n_matvec_host.  Synchronous. */
size_t mpsb_lanczos_workspace(int chil, int chir, int d, int w, int krylov);
int mpsb_lanczos_ground(int chil, int chir, int d, int w, const void* L, const void* W, const void* R, void* theta,
                        int krylov, int max_restarts, double tol, double* energy_host, double* resid_host,
                        int* n_matvec_host, void* ws, size_t ws_bytes, void* stream);

/* The same for `batch` chains in lock step: every product, orthogonalisation pass and update is one batched launch;
 * a chain that has converged is frozen at the next restart (its vector is final) and leaves the batch: the chains still
 * iterating move up into consecutive slots, so later cycles only pay for them.  energy_host and resid_host receive
 * `batch` values, n_matvec_host the number of products of the slowest chain. */
size_t mpsb_lanczos_batched_workspace(int batch, int chil, int chir, int d, int w, int krylov);
int mpsb_lanczos_ground_batched(int batch, int chil, int chir, int d, int w, const void* L, long long sL, const void* W,
                                long long sW, const void* R, long long sR, void* theta, long long s_theta, int krylov,
                                int max_restarts, double tol, double* energy_host, double* resid_host, int* n_matvec_host,
                                void* ws, size_t ws_bytes, void* stream);

int mpsb_lanczos_ground_batched_real(int batch, int chil, int chir, int d, int w, const void* L, long long sL,
                                     const void* W, long long sW, const void* R, long long sR, void* theta,
                                     long long s_theta, int krylov, int max_restarts, double tol, double* energy_host,
                                     double* resid_host, int* n_matvec_host, void* ws, size_t ws_bytes, void* stream);

/* H_eff products formed by the last (batched) Lanczos call of this process, summed over the chains that took part in
 * each: chains that have converged are frozen at a restart and leave the batched launches, so this is less than
 * batch * n_matvec for a coupling sweep whose chains need different numbers of products. */
unsigned long long mpsb_last_lanczos_chain_matvecs(void);

/* Environment growth (DMRG/iDMRG.py:59-60, with the conjugation on the bra leg — SURVEY App. B1):
 *     L'[A,B,C] = sum L[i,j,k] conj(U[i,l,A]) U[j,m,B] W[k,C,l,m]        U [chi][d][chin]
 *     R'[A,B,C] = sum R[i,j,k] conj(V[A,l,i]) V[B,m,j] W[C,k,l,m]        V [chin][d][chi]
 * swap_conj = 1 reproduces the reference's as-written conjugation (identical for real U, V). */
size_t mpsb_env_workspace(int chi, int chin, int d, int w);
int mpsb_env_left(int chi, int chin, int d, int w, const void* L, const void* U, const void* W, int swap_conj,
                  void* L_out, void* ws, size_t ws_bytes, void* stream);
int mpsb_env_right(int chi, int chin, int d, int w, const void* R, const void* V, const void* W, int swap_conj,
                   void* R_out, void* ws, size_t ws_bytes, void* stream);

size_t mpsb_env_batched_workspace(int batch, int chi, int chin, int d, int w);
int mpsb_env_left_batched(int batch, int chi, int chin, int d, int w, const void* L, long long sL, const void* U, long long sU,
                          const void* W, long long sW, int swap_conj, void* L_out, long long sLo, void* ws, size_t ws_bytes,
                          void* stream);
int mpsb_env_right_batched(int batch, int chi, int chin, int d, int w, const void* R, long long sR, const void* V, long long sV,
                           const void* W, long long sW, int swap_conj, void* R_out, long long sRo, void* ws, size_t ws_bytes,
                           void* stream);

/* real twins (no conjugation to choose) */
int mpsb_env_left_batched_real(int batch, int chi, int chin, int d, int w, const void* L, long long sL, const void* U,
                               long long sU, const void* W, long long sW, void* L_out, long long sLo, void* ws,
                               size_t ws_bytes, void* stream);
int mpsb_env_right_batched_real(int batch, int chi, int chin, int d, int w, const void* R, long long sR, const void* V,
                                long long sV, const void* W, long long sW, void* R_out, long long sRo, void* ws,
                                size_t ws_bytes, void* stream);

/* Left singular vectors from the right ones, for the split of two-site tensors whose rows are too long to carry an
 * identity through the SVD (halve, DMRG/iDMRG.py:17-35, at chi d > 256):  with Vh [batch][k][cols] and the singular
 * values S (first k of each chain, stride sS) of A [batch][rows][cols],
 *     Y = A Vh^H diag(1/S),     U = Y after two Newton-Schulz steps  Y <- Y (3 I - Y^H Y) / 2
 * U_out [batch][rows][k] holds the Schmidt partners of the rows of Vh, orthonormal to rounding; the symmetric
 * re-orthonormalisation moves every column (the dominant ones too) by ~eps S_0 / S_k, so use it while S_k / S_0 >~ 1e-5
 * and take U from an SVD of A^H otherwise.  real = 1: arrays of doubles. */
size_t mpsb_left_vectors_workspace(int batch, int rows, int cols, int k);
int mpsb_left_vectors(int batch, int rows, int cols, int k, const void* A, int lda, long long sA, const void* Vh,
                      const double* S, long long sS, void* U_out, int real, void* ws, size_t ws_bytes, void* stream);

/* Start vector for the next growth step from the tensors of the last one (new; the reference restarts ARPACK from a
 * random vector at DMRG/iDMRG.py:57).  With theta_n = A S B from the last split (A [chip][d][chi] left-orthonormal,
 * B [chi][d][chip] right-orthonormal, S [chi] normalised, all from ONE SVD so that their gauges match) and the
 * Schmidt values S_prev [chip] of the step before,
 *     theta[a,s,t,b] = sum_c S[a] B[a,s,c] S_prev_inv[c] A[c,t,b] S[b]          ([chi][d][d][chi], not normalised)
 * is the two-site tensor of the enlarged chain shifted by one site (McCulloch, arXiv:0804.2509).  S_prev_inv holds
 * 1/S_prev (0 where the caller cuts the pseudo-inverse). */
size_t mpsb_predict_workspace(int chi, int chip, int d);
int mpsb_predict_theta(int chi, int chip, int d, const void* A, const void* B, const double* S, const double* S_prev_inv,
                       void* theta, void* ws, size_t ws_bytes, void* stream);

size_t mpsb_predict_batched_workspace(int batch, int chi, int chip, int d);
int mpsb_predict_theta_batched(int batch, int chi, int chip, int d, const void* A, const void* B, const double* S,
                               const double* S_prev_inv, void* theta, void* ws, size_t ws_bytes, void* stream);

int mpsb_predict_theta_batched_real(int batch, int chi, int chip, int d, const void* A, const void* B, const double* S,
                                    const double* S_prev_inv, void* theta, void* ws, size_t ws_bytes, void* stream);

/* out[i * ldo + j] = s[i / group] * in[i * ldi + j]: the Lambda_l scaling `self.Sl[i][:, newaxis, ...] * theta`
 * (DMRG/MPS.py:446) of a tensor viewed as a (chi_l * group) x cols matrix.  Used by the multi-site update
 * (MPS.update_k, DMRG/MPS.py:459-461), whose splits have unequal local dimensions on the two sides. */
int mpsb_scale_rows(int rows, int cols, const void* in, int ldi, const double* s, int group, void* out, int ldo,
                    void* stream);

/* out[j * ldo + i] = conj?(in[i * ldi + j]) for i < rows, j < cols (layout helper of the shim). */
int mpsb_transpose(int rows, int cols, const void* in, int ldi, void* out, int ldo, int conj, void* stream);
int mpsb_transpose_batched(int batch, int rows, int cols, const void* in, int ldi, long long s_in, void* out, int ldo,
                           long long s_out, int conj, void* stream);

int mpsb_transpose_batched_real(int batch, int rows, int cols, const void* in, int ldi, long long s_in, void* out, int ldo,
                                long long s_out, void* stream);

/* Diagnostics of the last SVD call on this thread: number of Jacobi sweeps of the slowest problem. */
int mpsb_last_svd_sweeps(void);

/* Diagnostics (host only, no GPU needed): one sweep of the Jacobi block ordering as the launchers and kernels compute it,
 * for nblk 8-row blocks (nblk even; quad = 1: the 16-row super-block form, nblk a multiple of 4).  Writes up to `cap`
 * entries (step, stream, block I, block J) into out_host - stream 0 = the step's main launch, 1 = its side stream,
 * I == J = the pairs inside a block - and returns the number of entries (-1 if cap is too small).  Every pair of
 * blocks appears exactly once per sweep and the launches of one step touch disjoint blocks (tests/test_host.py). */
int mpsb_debug_sweep_schedule(int nblk, int quad, int* out_host, int cap);

/* Phase timing of mpsb_tebd_two_site for the roofline report (bench.py).  When enabled, CUDA events are
 * recorded on the caller's stream at the phase boundaries; mpsb_profile_read returns the 6 phase
 * durations in ms (theta GEMM, gate, QR, Jacobi sweeps, finalise, back-projection GEMM) of the last
 * call plus the Jacobi work counters of its SVD; a phase the last call did not run reads -1 (an SVD-only call
 * reports just the Jacobi sweeps). */
void mpsb_profile_enable(int on);
int mpsb_profile_read(double* phase_ms, unsigned long long* rotations, unsigned long long* problem_sweeps);

#ifdef __cplusplus
}
#endif
#endif /* MPSB_H */
