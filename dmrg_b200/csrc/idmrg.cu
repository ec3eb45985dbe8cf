// iDMRG growth step (reference DMRG/iDMRG.py:38-61) without ever forming the dense effective
// Hamiltonian:
//   mpsb_heff_matvec      out = H_eff theta as  GEMM(L) -> fused W.W kernel -> GEMM(R)      (replaces :38-40 + the
//                         dense matvec ARPACK performs inside eigsh, :57)
//   mpsb_lanczos_ground   restarted Lanczos with full (twice-applied) reorthogonalisation   (replaces eigsh, :57)
//   mpsb_env_left/right   environment growth as GEMM -> W kernel -> GEMM                    (replaces :59-60)
// Index conventions: L[A,a,i] (bra, ket, mpo), R[D,d,k], W[i,j,B,b] (mpo-row, mpo-col, bra, ket),
// theta[a,b,c,d] (left bond, phys1, phys2, right bond).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/mpsb.h"
#include "common.cuh"
#include "mpsb_internal.h"

namespace mpsb {

namespace {

inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

// WW[i][k][B][C][b][c] = sum_j W[i][j][B][b] W[j][k][C][c]
template <typename T>
__global__ void ww_build_kernel(int w, int d, const T* __restrict__ W, long long sW, T* __restrict__ WW) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = w * w * d * d * d * d;
    if (e >= total) return;
    W += (size_t)blockIdx.y * sW;
    WW += (size_t)blockIdx.y * total;
    int r = e;
    const int c = r % d; r /= d;
    const int b = r % d; r /= d;
    const int C = r % d; r /= d;
    const int B = r % d; r /= d;
    const int k = r % w;
    const int i = r / w;
    T acc = Sc<T>::zero();
    for (int j = 0; j < w; ++j)
        Sc<T>::mac(acc, W[((size_t)(i * w + j) * d + B) * d + b], W[((size_t)(j * w + k) * d + C) * d + c]);
    WW[e] = acc;
}

// T3[A][B][C][dr][k] = sum_{i,b,c} WW[i][k][B][C][b][c] T1[i][A][b][c][dr]
// One thread per (A, dr): it reads its w d^2 inputs once and produces all d^2 w outputs (instantiated for d = 2 with
// w = 3 and w = 5; other MPOs take one (B, C) per thread, below), so T1 and T3 cross the memory system exactly once.  Zero blocks of
// the (lower-triangular, mostly empty) MPO are skipped; the test is uniform across the warp.
template <typename T, int W, int D>
__global__ void __launch_bounds__(128)
ww_apply_fused_kernel(int chil, int chir, const T* __restrict__ WW, long long sWW, const T* __restrict__ T1,
                      T* __restrict__ T3) {
    constexpr int DD = D * D;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)chil * chir;
    if (e >= total) return;
    const long long per_chain = (long long)chil * DD * chir * W;
    WW += (size_t)blockIdx.y * sWW;
    T1 += (size_t)blockIdx.y * per_chain;
    T3 += (size_t)blockIdx.y * per_chain;
    const int dr = (int)(e % chir), A = (int)(e / chir);
    T x[W][DD];
    const size_t t1_row = (size_t)DD * chir;             // elements per (i, A)
#pragma unroll
    for (int i = 0; i < W; ++i)
#pragma unroll
        for (int bc = 0; bc < DD; ++bc) x[i][bc] = T1[((size_t)i * chil + A) * t1_row + (size_t)bc * chir + dr];
#pragma unroll
    for (int BC = 0; BC < DD; ++BC) {
        T* out = T3 + (((size_t)A * DD + BC) * chir + dr) * W;
#pragma unroll
        for (int k = 0; k < W; ++k) {
            T acc = Sc<T>::zero();
#pragma unroll
            for (int i = 0; i < W; ++i) {
                const T* g0 = WW + (((size_t)i * W + k) * DD + BC) * DD;
#pragma unroll
                for (int bc = 0; bc < DD; ++bc) {
                    const T g = g0[bc];
                    if (Sc<T>::is_zero(g)) continue;
                    Sc<T>::mac(acc, g, x[i][bc]);
                }
            }
            out[k] = acc;
        }
    }
}

// General form: one thread per (A, B, C, dr) producing the w values of k (contiguous store).
template <typename T>
__global__ void __launch_bounds__(256)
ww_apply_kernel(int chil, int chir, int d, int w, const T* __restrict__ WW, long long sWW,
                const T* __restrict__ T1, T* __restrict__ T3) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)chil * d * d * chir;
    if (e >= total) return;
    WW += (size_t)blockIdx.y * sWW;                      // chain of the batch
    T1 += (size_t)blockIdx.y * total * w;
    T3 += (size_t)blockIdx.y * total * w;
    const int dr = (int)(e % chir);
    long long r = e / chir;
    const int C = (int)(r % d); r /= d;
    const int B = (int)(r % d);
    const int A = (int)(r / d);
    const size_t t1_row = (size_t)d * d * chir;          // elements per (i, A)
    for (int k = 0; k < w; ++k) {
        T acc = Sc<T>::zero();
        for (int i = 0; i < w; ++i) {
            const T* t1 = T1 + ((size_t)i * chil + A) * t1_row + dr;
            const T* g0 = WW + ((((size_t)i * w + k) * d + B) * d + C) * d * d;
            for (int b = 0; b < d; ++b)
                for (int c = 0; c < d; ++c) {
                    const T g = g0[b * d + c];
                    if (Sc<T>::is_zero(g)) continue;
                    Sc<T>::mac(acc, g, t1[(size_t)(b * d + c) * chir]);
                }
        }
        T3[(size_t)e * w + k] = acc;
    }
}

// Left environment:  T2[C][i][l][B] = sum_{k,m} W[k][C][l][m] T1[k][i][m][B]
// Right environment: T2[C][l][i][B] = sum_{k,m} W[C][k][l][m] T1[k][i][B][m]
template <typename T>
__global__ void __launch_bounds__(256)
env_w_kernel(int right, int chi, int chin, int d, int w, const T* __restrict__ W, long long sW,
             const T* __restrict__ T1, T* __restrict__ T2) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)w * chi * d * chin;
    if (e >= total) return;
    W += (size_t)blockIdx.y * sW;
    T1 += (size_t)blockIdx.y * total;
    T2 += (size_t)blockIdx.y * total;
    const int B = (int)(e % chin);
    long long r = e / chin;
    int i, l;
    if (!right) { l = (int)(r % d); r /= d; i = (int)(r % chi); r /= chi; }
    else        { i = (int)(r % chi); r /= chi; l = (int)(r % d); r /= d; }
    const int C = (int)r;
    T acc = Sc<T>::zero();
    for (int k = 0; k < w; ++k)
        for (int m = 0; m < d; ++m) {
            const T g = right ? W[((size_t)(C * w + k) * d + l) * d + m] : W[((size_t)(k * w + C) * d + l) * d + m];
            if (Sc<T>::is_zero(g)) continue;
            Sc<T>::mac(acc, g, right ? T1[(((size_t)k * chi + i) * chin + B) * d + m]
                                     : T1[(((size_t)k * chi + i) * d + m) * chin + B]);
        }
    T2[e] = acc;
}

template <typename T>
struct Heff {
    int chil, chir, d, w;
    int batch;          // independent chains, every tensor below has a leading chain index
    const T* R;         // as given: [D][dr][k]
    long long sR;       // chain stride of R
    long long sWW;      // chain stride of WW (0: one MPO shared by all chains)
    T *Lt, *WW, *T1, *T3;
    size_t n;           // chil*d*d*chir
};
// Wavefunction prediction (new; no counterpart in the reference, which restarts ARPACK from a random vector):
//   XL[a,s,c] = S[a] B[a,s,c] Sinv[c]        XR[c,t,b] = A[c,t,b] S[b]
// so that theta_trial[a,s,t,b] = sum_c XL[a,s,c] XR[c,t,b] = (S B) Sprev^-1 (A S).  Element-wise, HBM-bound.
template <typename T>
__global__ void __launch_bounds__(256)
predict_scale_kernel(int chi, int chip, int d, const T* __restrict__ B, const T* __restrict__ A,
                     const double* __restrict__ S, const double* __restrict__ Sinv, T* __restrict__ XL,
                     T* __restrict__ XR) {
    const long long total = (long long)chi * d * chip;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    B += (size_t)blockIdx.y * total; A += (size_t)blockIdx.y * total;         // chain of the batch
    XL += (size_t)blockIdx.y * total; XR += (size_t)blockIdx.y * total;
    S += (size_t)blockIdx.y * chi; Sinv += (size_t)blockIdx.y * chip;
    {   // B[a][s][c]
        const int c = (int)(e % chip);
        const int a = (int)(e / ((long long)d * chip));
        XL[e] = Sc<T>::scale(B[e], S[a] * Sinv[c]);
    }
    {   // A[c][t][b]
        const int b = (int)(e % chi);
        XR[e] = Sc<T>::scale(A[e], S[b]);
    }
}

size_t heff_bytes(int chil, int chir, int d, int w, int batch = 1, size_t esz = sizeof(cplx)) {
    const size_t n = (size_t)chil * d * d * chir;
    return al((size_t)batch * w * chil * chil * esz) + al((size_t)batch * w * w * d * d * d * d * esz) +
           al((size_t)batch * n * w * esz) * 2;
}
// L, R: one environment per chain (strides sL, sR in elements); W: one MPO per chain (sW) or a shared one (sW = 0).
template <typename T>
int heff_prepare(Heff<T>& h, int chil, int chir, int d, int w, const T* L, const T* W, const T* R, void* ws,
                 cudaStream_t st, int batch = 1, long long sL = 0, long long sW = 0, long long sR = 0) {
    h.chil = chil; h.chir = chir; h.d = d; h.w = w; h.R = R; h.batch = batch; h.sR = sR;
    h.n = (size_t)chil * d * d * chir;
    const int ww = w * w * d * d * d * d;
    unsigned char* p = static_cast<unsigned char*>(ws);
    h.Lt = reinterpret_cast<T*>(p); p += al((size_t)batch * w * chil * chil * sizeof(T));
    h.WW = reinterpret_cast<T*>(p); p += al((size_t)batch * ww * sizeof(T));
    h.T1 = reinterpret_cast<T*>(p); p += al((size_t)batch * h.n * w * sizeof(T));
    h.T3 = reinterpret_cast<T*>(p);
    // Lt[i][A][a] = L[A][a][i]
    int rc = transpose_conj(L, w, chil * chil, w, h.Lt, chil * chil, 0, st, batch, sL, (long long)w * chil * chil);
    if (rc) return rc;
    const int nww = (sW == 0) ? 1 : batch;
    h.sWW = (sW == 0) ? 0 : ww;
    dim3 grid((ww + 127) / 128, nww);
    ww_build_kernel<T><<<grid, 128, 0, st>>>(w, d, W, sW, h.WW);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
// theta / out: [batch][n] with chain strides s_theta / s_out
template <typename T>
int heff_apply(const Heff<T>& h, const T* theta, T* out, cudaStream_t st, long long s_theta = 0, long long s_out = 0) {
    const int ncol = h.d * h.d * h.chir;
    GemmArgsT<T> g;
    // T1[(i,A)][(b,c,dr)] = sum_a Lt[(i,A)][a] theta[a][(b,c,dr)]
    g.A = h.Lt; g.B = theta; g.C = h.T1;
    g.M = h.w * h.chil; g.N = ncol; g.K = h.chil; g.lda = h.chil; g.ldb = ncol; g.ldc = ncol; g.opA = 0; g.opB = 0;
    g.batch1 = h.batch; g.batch2 = 1;
    g.sA1 = (long long)h.w * h.chil * h.chil; g.sB1 = s_theta; g.sC1 = (long long)h.n * h.w;
    g.sA2 = g.sB2 = g.sC2 = 0; g.accumulate = 0;
    int rc = gemm(g, st);
    if (rc) return rc;
    dim3 fgrid((unsigned)(((long long)h.chil * h.chir + 127) / 128), h.batch);
    if (h.d == 2 && h.w == 3) {            // TFIM-like MPOs: the inputs of a thread fit its registers
        ww_apply_fused_kernel<T, 3, 2><<<fgrid, 128, 0, st>>>(h.chil, h.chir, h.WW, h.sWW, h.T1, h.T3);
    } else if (h.d == 2 && h.w == 5) {     // Heisenberg / XXZ
        ww_apply_fused_kernel<T, 5, 2><<<fgrid, 128, 0, st>>>(h.chil, h.chir, h.WW, h.sWW, h.T1, h.T3);
    } else {
        dim3 grid((unsigned)((h.n + 255) / 256), h.batch);
        ww_apply_kernel<T><<<grid, 256, 0, st>>>(h.chil, h.chir, h.d, h.w, h.WW, h.sWW, h.T1, h.T3);
    }
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    // out[(A,B,C)][D] = sum_{(dr,k)} T3[(A,B,C)][(dr,k)] R[D][(dr,k)]
    g.A = h.T3; g.B = h.R; g.C = out;
    g.M = h.chil * h.d * h.d; g.N = h.chir; g.K = h.chir * h.w; g.lda = g.K; g.ldb = g.K; g.ldc = h.chir;
    g.opA = 0; g.opB = 1;
    g.sA1 = (long long)h.n * h.w; g.sB1 = h.sR; g.sC1 = s_out;
    return gemm(g, st);
}

// Symmetric eigenproblem of a small dense real matrix by cyclic Jacobi (host; m <= a few dozen).
void sym_eig_host(int m, std::vector<double>& a, std::vector<double>& v) {
    v.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i) v[(size_t)i * m + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) (i == j ? diag : off) += a[(size_t)i * m + j] * a[(size_t)i * m + j];
        if (off <= 1e-29 * (diag + off)) break;      // off-norm / norm <= 3e-15: quadratically convergent from here
        for (int p = 0; p < m - 1; ++p)
            for (int q = p + 1; q < m; ++q) {
                const double apq = a[(size_t)p * m + q];
                if (std::fabs(apq) <= 1e-17 * std::sqrt(std::fabs(a[(size_t)p * m + p] * a[(size_t)q * m + q])) ) continue;
                const double tau = (a[(size_t)q * m + q] - a[(size_t)p * m + p]) / (2.0 * apq);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < m; ++k) {
                    const double akp = a[(size_t)k * m + p], akq = a[(size_t)k * m + q];
                    a[(size_t)k * m + p] = c * akp - s * akq;
                    a[(size_t)k * m + q] = s * akp + c * akq;
                }
                for (int k = 0; k < m; ++k) {
                    const double apk = a[(size_t)p * m + k], aqk = a[(size_t)q * m + k];
                    a[(size_t)p * m + k] = c * apk - s * aqk;
                    a[(size_t)q * m + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < m; ++k) {
                    const double vkp = v[(size_t)k * m + p], vkq = v[(size_t)k * m + q];
                    v[(size_t)k * m + p] = c * vkp - s * vkq;
                    v[(size_t)k * m + q] = s * vkp + c * vkq;
                }
            }
    }
}


template <typename T>
int heff_matvec_impl(int batch, int chil, int chir, int d, int w, const void* L, long long sL, const void* W,
                     long long sW, const void* R, long long sR, const void* theta, long long s_theta, void* out,
                     long long s_out, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && chil >= 1 && chir >= 1 && d >= 1 && w >= 1, "heff_matvec: bad dims");
    const size_t need = heff_bytes(chil, chir, d, w, batch, sizeof(T));
    MPSB_REQUIRE(ws && ws_bytes >= need, "heff_matvec: workspace %zu B < required %zu B", ws_bytes, need);
    Heff<T> h;
    int rc = heff_prepare<T>(h, chil, chir, d, w, static_cast<const T*>(L), static_cast<const T*>(W),
                             static_cast<const T*>(R), ws, st, batch, sL, sW, sR);
    if (rc) return rc;
    return heff_apply<T>(h, static_cast<const T*>(theta), static_cast<T*>(out), st, s_theta, s_out);
}


unsigned long long g_last_chain_matvecs = 0ull;

constexpr int LANCZOS_KEEP = 8;        // Ritz vectors carried across a (thick) restart
constexpr int LANCZOS_NCHUNK = 64;

size_t lanczos_bytes(int batch, int chil, int chir, int d, int w, int krylov, size_t esz) {
    const size_t n = (size_t)chil * d * d * chir;
    const size_t kk = (size_t)krylov + 2;
    const size_t b = (size_t)batch;
    return heff_bytes(chil, chir, d, w, batch, esz) + b * al(n * esz) * (kk + LANCZOS_KEEP) +
           al(b * kk * (LANCZOS_NCHUNK + 1) * esz) + al(b * kk * kk * esz) * 2 + al(b * kk * esz) +
           al(b * kk * LANCZOS_KEEP * esz) +
           // compacted copies of the per-chain operators for the chains still iterating, and two index lists
           al(b * w * chil * chil * esz) + al(b * w * chir * chir * esz) + al(b * w * w * d * d * d * d * esz) +
           al(2 * b * sizeof(int));
}

// dst[j][0:len) = src[idx[j]][0:len): moves the data of the chains that are still iterating into consecutive slots
template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_kernel(T* __restrict__ dst, long long sd, const T* __restrict__ src, long long ss, const int* __restrict__ idx,
                   long long len) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    dst[(size_t)blockIdx.y * sd + e] = src[(size_t)idx[blockIdx.y] * ss + e];
}
template <typename T>
int gather_rows(T* dst, long long sd, const T* src, long long ss, const int* idx, long long len, int count, cudaStream_t st) {
    if (count == 0 || len == 0) return 0;
    dim3 grid((unsigned)((len + 255) / 256), count);
    gather_rows_kernel<T><<<grid, 256, 0, st>>>(dst, sd, src, ss, idx, len);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Thick-restart Lanczos with full (twice-applied) reorthogonalisation for `batch` independent chains in lock step
// (same shapes; per-chain environments, start vectors and - with sW != 0 - MPOs).
//  * A cycle extends the orthonormal basis V[0..m] of every chain one product at a time; every product, inner-product
//    pass and update is ONE batched launch over all chains.  The Gram-Schmidt coefficients stay on the device (the
//    axpy kernel reads them there) and the new vectors are normalised from device-resident norms, so a cycle runs
//    without host synchronisation; coefficients are read back once per cycle.
//  * The projected matrices Hp = V^H H V are assembled per chain from those coefficients (column j: h_ij = c1_i +
//    c2_i, i <= j, and the norm beta_j below the diagonal) and diagonalised on the host.  After a restart the leading
//    block is diag(theta) of the kept Ritz vectors; the couplings to the continued Lanczos vector are measured by the
//    same Gram-Schmidt pass.
//  * Restart: keep the LANCZOS_KEEP lowest Ritz vectors (V Y) plus the residual vector v_m.
//  * A chain that has converged (or whose Krylov space closed) has its Ritz vector copied to theta and is FROZEN: it
//    keeps riding along in the batched launches, its later numbers are ignored.  The call ends when every chain is
//    frozen or max_restarts is reached.
template <typename T>
int lanczos_impl(int batch, int chil, int chir, int d, int w, const void* L, long long sL, const void* W,
                 long long sW, const void* R, long long sR, void* theta, long long s_theta, int krylov,
                 int max_restarts, double tol, double* energy_host, double* resid_host, int* n_matvec_host,
                 void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && chil >= 1 && chir >= 1 && d >= 1 && w >= 1 && krylov >= 2 && max_restarts >= 1,
                 "lanczos_ground: bad arguments");
    const size_t need = lanczos_bytes(batch, chil, chir, d, w, krylov, sizeof(T));
    MPSB_REQUIRE(ws && ws_bytes >= need, "lanczos_ground: workspace %zu B < required %zu B", ws_bytes, need);
    const size_t n = (size_t)chil * d * d * chir;
    const int nchunk = LANCZOS_NCHUNK;
    int m = krylov;
    if ((size_t)m > n) m = (int)n;
    const size_t kk = (size_t)krylov + 2;
    const size_t nb = (size_t)batch;
    unsigned char* p = static_cast<unsigned char*>(ws);
    Heff<T> h;
    int rc = heff_prepare<T>(h, chil, chir, d, w, static_cast<const T*>(L), static_cast<const T*>(W),
                             static_cast<const T*>(R), p, st, batch, sL, sW, sR);
    if (rc) return rc;
    p += heff_bytes(chil, chir, d, w, batch, sizeof(T));
    const size_t vstride = al(n * sizeof(T)) / sizeof(T);
    const long long sV = (long long)(vstride * kk), sVk = (long long)(vstride * LANCZOS_KEEP);
    const long long sPart = (long long)(kk * (nchunk + 1)), sC = (long long)(kk * kk), sN = (long long)kk,
                    sY = (long long)(kk * LANCZOS_KEEP);
    T* V = reinterpret_cast<T*>(p);                 p += nb * al(n * sizeof(T)) * kk;             // [chain][kk][vstride]
    T* Vk = reinterpret_cast<T*>(p);                p += nb * al(n * sizeof(T)) * LANCZOS_KEEP;   // Ritz vectors
    T* part = reinterpret_cast<T*>(p);              p += al(nb * kk * (nchunk + 1) * sizeof(T));   // partials + tickets
    T* C1 = reinterpret_cast<T*>(p);                p += al(nb * kk * kk * sizeof(T));   // pass-1 coefficients, row j
    T* C2 = reinterpret_cast<T*>(p);                p += al(nb * kk * kk * sizeof(T));   // pass-2 coefficients
    T* N2 = reinterpret_cast<T*>(p);                p += al(nb * kk * sizeof(T));        // ||w_j||^2
    T* Yd = reinterpret_cast<T*>(p);                p += al(nb * kk * LANCZOS_KEEP * sizeof(T));    // [chain][keep][kk]
    const size_t nLt = (size_t)w * chil * chil, nR = (size_t)w * chir * chir, nWW = (size_t)w * w * d * d * d * d;
    T* Lt_c = reinterpret_cast<T*>(p);              p += al(nb * nLt * sizeof(T));
    T* R_c = reinterpret_cast<T*>(p);               p += al(nb * nR * sizeof(T));
    T* WW_c = reinterpret_cast<T*>(p);              p += al(nb * nWW * sizeof(T));
    int* idx_d = reinterpret_cast<int*>(p);                                                     // [2][batch]
    T* th = static_cast<T*>(theta);
    // Slots: the batched launches of a cycle run over the chains that are still iterating (`nact` of them, slot ->
    // chain in slot2chain).  Chains that converge are frozen at a restart and the remaining ones move up into the first
    // slots (their bases by gather_rows, their operators re-gathered from the full set), so a coupling sweep whose
    // chains need 40 ... 90 products does not carry the finished ones along.
    int nact = batch;
    std::vector<int> slot2chain(nb);
    for (size_t i = 0; i < nb; ++i) slot2chain[i] = (int)i;
    const Heff<T> hfull = h;
    auto bat = [&](long long sv, long long sw, long long sc) {
        BlasBatch b;
        b.batch = nact; b.sV = sv; b.sw = sw; b.sc = sc; b.spart = sPart;
        return b;
    };
    struct Chain {
        std::vector<double> Hp, Tm, Y;
        std::vector<int> idx;
        double E = 0.0, resid = 0.0, beta_last = 0.0;
        int mm = 0;
        bool frozen = false;
    };
    std::vector<Chain> ch(nb);
    for (auto& c : ch) c.Hp.assign((size_t)m * m, 0.0);
    std::vector<T> h1(nb * kk * kk), h2(nb * kk * kk), hn(nb * kk), hc(nb * kk * LANCZOS_KEEP);
    int n_mv = 0;
    unsigned long long chain_mv = 0ull;          // products actually formed, summed over the chains that took part

    MPSB_CHECK_CUDA(cudaMemsetAsync(part, 0, al(nb * kk * (nchunk + 1) * sizeof(T)), st));
    // normalise the start vectors into V[chain][0]
    rc = zdotc_multi(th, 0, 1, (long long)n, th, N2, part, nchunk, (int)kk, st, bat(s_theta, s_theta, sN));
    if (rc) return rc;
    MPSB_CHECK_CUDA(cudaMemcpy2DAsync(hn.data(), kk * sizeof(T), N2, kk * sizeof(T), sizeof(T), nb,
                                      cudaMemcpyDeviceToHost, st));
    MPSB_CHECK_CUDA(cudaStreamSynchronize(st));
    for (size_t c = 0; c < nb; ++c)
        MPSB_REQUIRE(Sc<T>::re(hn[c * kk]) > 0.0 && std::isfinite(Sc<T>::re(hn[c * kk])), "lanczos_ground: start vector of chain %zu has norm^2 = %g",
                     c, Sc<T>::re(hn[c * kk]));
    if (batch == 1) MPSB_CHECK_CUDA(cudaMemcpyAsync(V, th, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    else MPSB_CHECK_CUDA(cudaMemcpy2DAsync(V, (size_t)sV * sizeof(T), th, (size_t)s_theta * sizeof(T), n * sizeof(T), nb,
                                           cudaMemcpyDeviceToDevice, st));
    rc = znormalise((long long)n, N2, V, st, bat(0, sV, sN));
    if (rc) return rc;

    int j0 = 0;                 // first basis index whose product has not been formed yet
    const bool dbg = getenv("MPSB_LANCZOS_DEBUG") != nullptr;
    double t_enq = 0.0, t_wait = 0.0, t_host = 0.0;
    auto now = [] { return std::chrono::high_resolution_clock::now(); };
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    // Convergence is also polled inside the FIRST cycle (after 4 steps, then every 8: one small download and a host
    // eigen-solve of the projected matrices).  A start vector that is already the answer - finite-DMRG sweeps after
    // the first - then costs a few products instead of a whole 24-step cycle.  Later cycles run to their end:
    // stopping them the moment the tolerance is met was measured to cost more products over an iDMRG run (each step
    // then starts from a slightly less converged prediction).
    static const int poll_env = getenv("MPSB_LANCZOS_POLL") ? atoi(getenv("MPSB_LANCZOS_POLL")) : 1;
    auto small_enough = [&](const Chain& c) { return c.resid <= tol * (std::fabs(c.E) > 1.0 ? std::fabs(c.E) : 1.0); };
    // Ritz data of the bases V[chain][0 .. formed): per live chain T (eigenvalues on the diagonal), Y, idx (ascending),
    // E, resid, mm.
    auto evaluate = [&](int formed) -> int {
        const size_t na = (size_t)nact;
        MPSB_CHECK_CUDA(cudaMemcpy2DAsync(h1.data(), kk * kk * sizeof(T), C1, kk * kk * sizeof(T),
                                          sizeof(T) * kk * formed, na, cudaMemcpyDeviceToHost, st));
        MPSB_CHECK_CUDA(cudaMemcpy2DAsync(h2.data(), kk * kk * sizeof(T), C2, kk * kk * sizeof(T),
                                          sizeof(T) * kk * formed, na, cudaMemcpyDeviceToHost, st));
        MPSB_CHECK_CUDA(cudaMemcpy2DAsync(hn.data(), kk * sizeof(T), N2, kk * sizeof(T), sizeof(T) * formed, na,
                                          cudaMemcpyDeviceToHost, st));
        auto tb = now();
        MPSB_CHECK_CUDA(cudaStreamSynchronize(st));
        auto tc = now();
        t_wait += us(tb, tc);
        auto solve_chain = [&](size_t ci) {             // ci: slot
            Chain& c = ch[slot2chain[ci]];
            if (c.frozen) return;
            const T* a1 = h1.data() + ci * kk * kk;
            const T* a2 = h2.data() + ci * kk * kk;
            const T* an = hn.data() + ci * kk;
            c.mm = formed;
            c.beta_last = 0.0;
            double hscale = 0.0;                      // largest |h_ij| seen: scale of the breakdown test
            for (int j = j0; j < formed; ++j) {
                for (int i = 0; i <= j; ++i) {
                    const double v = Sc<T>::re(a1[(size_t)j * kk + i]) + Sc<T>::re(a2[(size_t)j * kk + i]);
                    c.Hp[(size_t)i * m + j] = v;
                    c.Hp[(size_t)j * m + i] = v;
                    hscale = std::max(hscale, std::fabs(v));
                }
                const double bj = std::sqrt(Sc<T>::re(an[j]) > 0.0 ? Sc<T>::re(an[j]) : 0.0);
                c.beta_last = bj;
                if (j + 1 < m) { c.Hp[(size_t)(j + 1) * m + j] = bj; c.Hp[(size_t)j * m + j + 1] = bj; }
                hscale = std::max(hscale, std::fabs(c.Hp[0]));
                if (!(bj > 1e-14 * (hscale + 1e-300)) || !std::isfinite(bj)) { c.mm = j + 1; break; }
            }
            const int mm = c.mm;
            c.Tm.assign((size_t)mm * mm, 0.0);
            for (int i = 0; i < mm; ++i)
                for (int q = 0; q < mm; ++q) c.Tm[(size_t)i * mm + q] = c.Hp[(size_t)i * m + q];
            sym_eig_host(mm, c.Tm, c.Y);
            c.idx.resize(mm);
            for (int i = 0; i < mm; ++i) c.idx[i] = i;
            std::sort(c.idx.begin(), c.idx.end(), [&](int a, int b) { return c.Tm[(size_t)a * mm + a] < c.Tm[(size_t)b * mm + b]; });
            const int best = c.idx[0];
            c.E = c.Tm[(size_t)best * mm + best];
            c.resid = std::fabs(c.beta_last * c.Y[(size_t)(mm - 1) * mm + best]);
        };
        // the projected eigenproblems of the chains are independent: spread them over a few host threads
        const unsigned nthr = (unsigned)std::max<size_t>(1, std::min<size_t>({na / 8, (size_t)16, (size_t)std::thread::hardware_concurrency()}));
        if (nthr <= 1) {
            for (size_t ci = 0; ci < na; ++ci) solve_chain(ci);
        } else {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nthr; ++t)
                pool.emplace_back([&, t] { for (size_t ci = t; ci < na; ci += nthr) solve_chain(ci); });
            for (auto& th_ : pool) th_.join();
        }
        t_host += us(tc, now());
        return 0;
    };
    bool all_done = false;
    for (int restart = 0; restart < max_restarts && !all_done; ++restart) {
        auto ta = now();
        int formed = m;
        bool evaluated = false;
        MPSB_CHECK_CUDA(cudaMemsetAsync(C1, 0, al(nb * kk * kk * sizeof(T)), st));     // pass 1 fills only part of a row
        for (int j = j0; j < m; ++j) {
            T* vj = V + (size_t)j * vstride;
            T* wv = V + (size_t)(j + 1) * vstride;
            rc = heff_apply(h, vj, wv, st, sV, sV);
            if (rc) return rc;
            ++n_mv;
            chain_mv += (unsigned long long)nact;
            // Gram-Schmidt in two passes.  Pass 1 is the Lanczos recurrence itself: it removes the large components
            // along v_j and v_{j-1} (right after a (re)start: along every basis vector, the kept Ritz vectors all couple
            // to the continued one), so that pass 2 - the full reorthogonalisation - subtracts only rounding-level
            // coefficients and needs no repetition.  Traffic per step: 3 + j vectors instead of 2 (j + 1).
            const int lo = (j == j0) ? 0 : std::max(0, j - 1);
            {
                T* c = C1 + (size_t)j * kk + lo;
                const T* Vlo = V + (size_t)lo * vstride;
                rc = zdotc_multi(Vlo, (long long)vstride, j + 1 - lo, (long long)n, wv, c, part, nchunk, (int)kk, st, bat(sV, sV, sC));
                if (rc) return rc;
                rc = zaxpy_multi(Vlo, (long long)vstride, j + 1 - lo, (long long)n, c, -1.0, wv, st, bat(sV, sV, sC));
                if (rc) return rc;
            }
            {
                T* c = C2 + (size_t)j * kk;
                rc = zdotc_multi(V, (long long)vstride, j + 1, (long long)n, wv, c, part, nchunk, (int)kk, st, bat(sV, sV, sC));
                if (rc) return rc;
                rc = zaxpy_multi(V, (long long)vstride, j + 1, (long long)n, c, -1.0, wv, st, bat(sV, sV, sC));
                if (rc) return rc;
            }
            rc = zdotc_multi(wv, 0, 1, (long long)n, wv, N2 + j, part, nchunk, (int)kk, st, bat(sV, sV, sN));
            if (rc) return rc;
            rc = znormalise((long long)n, N2 + j, wv, st, bat(0, sV, sN));
            if (rc) return rc;
            const int done_now = j + 1 - j0;
            if (poll_env && restart == 0 && j + 1 < m && (done_now == 4 || (done_now > 4 && (done_now - 4) % 8 == 0))) {
                t_enq += us(ta, now());
                rc = evaluate(j + 1);
                if (rc) return rc;
                ta = now();
                bool all = true;
                for (const Chain& c : ch) all = all && (c.frozen || small_enough(c) || c.mm < j + 1);
                if (all) { formed = j + 1; evaluated = true; break; }
            }
        }
        t_enq += us(ta, now());
        if (!evaluated) {
            rc = evaluate(formed);
            if (rc) return rc;
        }
        auto tc = now();
        const bool last = restart + 1 == max_restarts;
        all_done = true;
        const size_t na = (size_t)nact;
        std::vector<char> finish(na, 0);                 // per slot
        for (size_t si = 0; si < na; ++si) {
            Chain& c = ch[slot2chain[si]];
            if (c.frozen) continue;
            if (small_enough(c) || c.mm < formed || formed < m || last) finish[si] = 1;
            else all_done = false;
        }
        const int keep = all_done ? 1 : std::min(LANCZOS_KEEP, m - 1);
        // Ritz vectors Vk[chain][i] = sum_q Y[q][idx[i]] V[chain][q]   (zero coefficients beyond a chain's own mm)
        std::fill(hc.begin(), hc.end(), Sc<T>::zero());
        for (size_t si = 0; si < na; ++si) {
            const Chain& c = ch[slot2chain[si]];
            if (c.frozen) continue;
            const int kc = std::min(keep, c.mm);
            for (int i = 0; i < kc; ++i)
                for (int q = 0; q < c.mm; ++q)
                    hc[si * kk * LANCZOS_KEEP + (size_t)i * kk + q] = Sc<T>::from_re(c.Y[(size_t)q * c.mm + c.idx[i]]);
        }
        MPSB_CHECK_CUDA(cudaMemcpyAsync(Yd, hc.data(), sizeof(T) * na * kk * LANCZOS_KEEP, cudaMemcpyHostToDevice, st));
        MPSB_CHECK_CUDA(cudaMemsetAsync(Vk, 0, na * al(n * sizeof(T)) * LANCZOS_KEEP, st));
        for (int i = 0; i < keep; ++i) {
            rc = zaxpy_multi(V, (long long)vstride, formed, (long long)n, Yd + (size_t)i * kk, 1.0, Vk + (size_t)i * vstride, st,
                             bat(sV, sVk, sY));
            if (rc) return rc;
        }
        std::vector<int> keep_slots, keep_chains;      // slots (chains) that go on iterating, ascending
        for (size_t si = 0; si < na; ++si) {           // freeze the chains that finished in this cycle
            const int ci = slot2chain[si];
            if (!finish[si]) {
                if (!ch[ci].frozen) { keep_slots.push_back((int)si); keep_chains.push_back(ci); }
                continue;
            }
            MPSB_CHECK_CUDA(cudaMemcpyAsync(th + (size_t)ci * (size_t)s_theta, Vk + si * (size_t)sVk, n * sizeof(T),
                                            cudaMemcpyDeviceToDevice, st));
            ch[ci].frozen = true;
        }
        t_host += us(tc, now());
        if (all_done) break;
        // new bases: kept Ritz vectors, then the residual direction v_m (already orthogonal to all of them)
        if (batch == 1) {
            MPSB_CHECK_CUDA(cudaMemcpyAsync(V + (size_t)keep * vstride, V + (size_t)m * vstride, n * sizeof(T),
                                            cudaMemcpyDeviceToDevice, st));
            MPSB_CHECK_CUDA(cudaMemcpyAsync(V, Vk, al(n * sizeof(T)) * keep, cudaMemcpyDeviceToDevice, st));
        } else if ((int)keep_slots.size() == nact) {   // nobody left: same slots
            MPSB_CHECK_CUDA(cudaMemcpy2DAsync(V + (size_t)keep * vstride, (size_t)sV * sizeof(T), V + (size_t)m * vstride,
                                              (size_t)sV * sizeof(T), n * sizeof(T), na, cudaMemcpyDeviceToDevice, st));
            MPSB_CHECK_CUDA(cudaMemcpy2DAsync(V, (size_t)sV * sizeof(T), Vk, (size_t)sVk * sizeof(T),
                                              al(n * sizeof(T)) * keep, na, cudaMemcpyDeviceToDevice, st));
        } else {
            // The survivors move up into slots 0 .. nkeep-1.  Within V only index m of a slot is read and only indices
            // 0 .. keep (< m) of a slot are written, so the in-place gather of the residual vectors is safe.
            const int nkeep = (int)keep_slots.size();
            MPSB_CHECK_CUDA(cudaMemcpyAsync(idx_d, keep_slots.data(), sizeof(int) * nkeep, cudaMemcpyHostToDevice, st));
            MPSB_CHECK_CUDA(cudaMemcpyAsync(idx_d + nb, keep_chains.data(), sizeof(int) * nkeep, cudaMemcpyHostToDevice, st));
            rc = gather_rows<T>(V + (size_t)keep * vstride, sV, V + (size_t)m * vstride, sV, idx_d, (long long)n, nkeep, st);
            if (rc) return rc;
            rc = gather_rows<T>(V, sV, Vk, sVk, idx_d, (long long)(vstride * keep), nkeep, st);
            if (rc) return rc;
            rc = gather_rows<T>(Lt_c, (long long)nLt, hfull.Lt, (long long)nLt, idx_d + nb, (long long)nLt, nkeep, st);
            if (rc) return rc;
            rc = gather_rows<T>(R_c, (long long)nR, hfull.R, hfull.sR, idx_d + nb, (long long)nR, nkeep, st);
            if (rc) return rc;
            h.Lt = Lt_c; h.R = R_c; h.sR = (long long)nR;
            if (hfull.sWW != 0) {
                rc = gather_rows<T>(WW_c, (long long)nWW, hfull.WW, hfull.sWW, idx_d + nb, (long long)nWW, nkeep, st);
                if (rc) return rc;
                h.WW = WW_c;
            }
            for (int i = 0; i < nkeep; ++i) slot2chain[i] = keep_chains[i];
            nact = nkeep;
            h.batch = nact;
        }
        for (Chain& c : ch) {
            if (c.frozen) continue;
            std::vector<double> diag(keep);
            for (int i = 0; i < keep; ++i) diag[i] = c.Tm[(size_t)c.idx[i] * c.mm + c.idx[i]];
            std::fill(c.Hp.begin(), c.Hp.end(), 0.0);
            for (int i = 0; i < keep; ++i) c.Hp[(size_t)i * m + i] = diag[i];
        }
        j0 = keep;
    }
    if (dbg) fprintf(stderr, "[mpsb] lanczos: %d chains, %d matvecs, enqueue %.0f us, wait %.0f us, host ritz %.0f us\n", batch,
                     n_mv, t_enq, t_wait, t_host);
    // returned vectors: lowest Ritz vectors, renormalised; Rayleigh quotients and true residuals with one more product
    // (all chains again, with the operators as given)
    nact = batch;
    h = hfull;
    rc = zdotc_multi(th, 0, 1, (long long)n, th, N2, part, nchunk, (int)kk, st, bat(s_theta, s_theta, sN));
    if (rc) return rc;
    rc = znormalise((long long)n, N2, th, st, bat(0, s_theta, sN));
    if (rc) return rc;
    T* hv = V + (size_t)1 * vstride;
    rc = heff_apply(h, th, hv, st, s_theta, sV);
    if (rc) return rc;
    ++n_mv;
    chain_mv += (unsigned long long)batch;
    g_last_chain_matvecs = chain_mv;
    rc = zdotc_multi(th, 0, 1, (long long)n, hv, C1, part, nchunk, (int)kk, st, bat(s_theta, sV, sC));
    if (rc) return rc;
    rc = zaxpy_multi(th, 0, 1, (long long)n, C1, -1.0, hv, st, bat(s_theta, sV, sC));
    if (rc) return rc;
    rc = zdotc_multi(hv, 0, 1, (long long)n, hv, N2, part, nchunk, (int)kk, st, bat(sV, sV, sN));
    if (rc) return rc;
    MPSB_CHECK_CUDA(cudaMemcpy2DAsync(hc.data(), sizeof(T), C1, kk * kk * sizeof(T), sizeof(T), nb,
                                      cudaMemcpyDeviceToHost, st));
    MPSB_CHECK_CUDA(cudaMemcpy2DAsync(hn.data(), sizeof(T), N2, kk * sizeof(T), sizeof(T), nb, cudaMemcpyDeviceToHost,
                                      st));
    MPSB_CHECK_CUDA(cudaStreamSynchronize(st));
    for (size_t ci = 0; ci < nb; ++ci) {
        if (energy_host) energy_host[ci] = Sc<T>::re(hc[ci]);
        if (resid_host) resid_host[ci] = std::sqrt(Sc<T>::re(hn[ci]) > 0.0 ? Sc<T>::re(hn[ci]) : 0.0);
    }
    if (n_matvec_host) *n_matvec_host = n_mv;
    return 0;
}


// A, B, S, S_prev_inv, theta: contiguous per chain ([batch][...])
template <typename T>
int predict_impl(int batch, int chi, int chip, int d, const void* A, const void* B, const double* S,
                 const double* S_prev_inv, void* theta, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && chi >= 1 && chip >= 1 && d >= 1 && A && B && S && S_prev_inv && theta,
                 "predict_theta: bad arguments");
    const size_t need = 2 * al((size_t)batch * chi * d * chip * sizeof(T));
    MPSB_REQUIRE(ws && ws_bytes >= need, "predict_theta: workspace %zu B < required %zu B", ws_bytes, need);
    T* XL = static_cast<T*>(ws);
    T* XR = reinterpret_cast<T*>(static_cast<unsigned char*>(ws) + need / 2);
    const long long total = (long long)chi * d * chip;
    dim3 grid((unsigned)((total + 255) / 256), batch);
    predict_scale_kernel<T><<<grid, 256, 0, st>>>(chi, chip, d, static_cast<const T*>(B), static_cast<const T*>(A), S,
                                                  S_prev_inv, XL, XR);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    GemmArgsT<T> g;
    g.batch1 = batch; g.batch2 = 1; g.sA1 = total; g.sB1 = total; g.sC1 = (long long)chi * d * d * chi;
    g.sA2 = g.sB2 = g.sC2 = 0; g.accumulate = 0;
    g.A = XL; g.B = XR; g.C = static_cast<T*>(theta);
    g.M = chi * d; g.N = d * chi; g.K = chip;
    g.lda = chip; g.ldb = d * chi; g.ldc = d * chi; g.opA = 0; g.opB = 0;
    return gemm(g, st);
}


size_t env_bytes(int batch, int chi, int chin, int d, int w, size_t esz) {
    const size_t b = (size_t)batch;
    return al(b * w * chi * chi * esz) + al(b * w * chi * d * chin * esz) * 2 + al(b * w * chin * chin * esz);
}

// E [batch][chi][chi][w] (stride sE), UV [batch][...] (sUV), W [w][w][d][d] per chain (sW) or shared (sW = 0),
// E_out [batch][chin][chin][w] (sEo)
template <typename T>
int env_common(int right, int batch, int chi, int chin, int d, int w, const void* E, long long sE, const void* UV,
                      long long sUV, const void* W, long long sW, int swap_conj, void* E_out, long long sEo, void* ws,
                      size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && chi >= 1 && chin >= 1 && d >= 1 && w >= 1, "env: bad dims");
    const size_t need = env_bytes(batch, chi, chin, d, w, sizeof(T));
    MPSB_REQUIRE(ws && ws_bytes >= need, "env: workspace %zu B < required %zu B", ws_bytes, need);
    const size_t b = (size_t)batch;
    unsigned char* p = static_cast<unsigned char*>(ws);
    T* Et = reinterpret_cast<T*>(p); p += al(b * w * chi * chi * sizeof(T));
    T* T1 = reinterpret_cast<T*>(p); p += al(b * w * chi * d * chin * sizeof(T));
    T* T2 = reinterpret_cast<T*>(p); p += al(b * w * chi * d * chin * sizeof(T));
    T* Ot = reinterpret_cast<T*>(p);
    const T* X = static_cast<const T*>(UV);
    const long long sEt = (long long)w * chi * chi, sT = (long long)w * chi * d * chin, sOt = (long long)w * chin * chin;
    // Et[k][i][j] = E[i][j][k]
    int rc = transpose_conj(static_cast<const T*>(E), w, chi * chi, w, Et, chi * chi, 0, st, batch, sE, sEt);
    if (rc) return rc;
    GemmArgsT<T> g;
    g.batch1 = batch; g.batch2 = 1; g.sA1 = sEt; g.sB1 = sUV; g.sC1 = sT; g.sA2 = g.sB2 = g.sC2 = 0; g.accumulate = 0;
    // ket leg (index j of the environment) meets the un-conjugated tensor unless swap_conj
    if (!right) {
        // T1[(k,i)][(m,B)] = sum_j Et[(k,i)][j] U[j][(m,B)]
        g.A = Et; g.B = X; g.C = T1; g.M = w * chi; g.N = d * chin; g.K = chi;
        g.lda = chi; g.ldb = d * chin; g.ldc = d * chin; g.opA = 0; g.opB = swap_conj ? 3 : 0;
    } else {
        // T1[(k,i)][(B,m)] = sum_j Et[(k,i)][j] V[(B,m)][j]
        g.A = Et; g.B = X; g.C = T1; g.M = w * chi; g.N = chin * d; g.K = chi;
        g.lda = chi; g.ldb = chi; g.ldc = chin * d; g.opA = 0; g.opB = swap_conj ? 2 : 1;
    }
    rc = gemm(g, st);
    if (rc) return rc;
    const long long total = (long long)w * chi * d * chin;
    dim3 grid((unsigned)((total + 255) / 256), batch);
    env_w_kernel<T><<<grid, 256, 0, st>>>(right, chi, chin, d, w, static_cast<const T*>(W), sW, T1, T2);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    // bra leg:  Ot[C][A][B] = sum_{K} op(X)[A][K] T2[C][K][B]      (batch level 1: chains, level 2: C)
    g.B = T2; g.C = Ot; g.M = chin; g.N = chin; g.K = chi * d; g.ldb = chin; g.ldc = chin;
    g.batch1 = batch; g.batch2 = w;
    g.sA1 = sUV; g.sB1 = sT; g.sC1 = sOt;
    g.sA2 = 0; g.sB2 = (long long)chi * d * chin; g.sC2 = (long long)chin * chin;
    g.A = X;
    if (!right) { g.lda = chin; g.opA = swap_conj ? 1 : 2; }      // U as [(i,l)][A]
    else        { g.lda = chi * d; g.opA = swap_conj ? 0 : 3; }   // V as [A][(l,i)]
    g.opB = 0;
    rc = gemm(g, st);
    if (rc) return rc;
    // E_out[A][B][C] = Ot[C][A][B]
    return transpose_conj(Ot, chin * chin, w, chin * chin, static_cast<T*>(E_out), w, 0, st, batch, sOt, sEo);
}


// ---- left singular vectors from the right ones (split of large two-site tensors) ----------------------------------
// halve (DMRG/iDMRG.py:17-35) needs U and V.  For rows too long to carry an identity through the Jacobi sweeps the
// second factor used to come from a second SVD (of A^H).  With Vh and S of the first one,
//     Y = A Vh^H diag(1/S)                       (one GEMM; columns orthonormal up to eps * S_0 / S_i)
//     U = two Newton-Schulz steps  Y <- Y (3 I - Y^H Y) / 2  (the defect e becomes ~ e^2, then ~ e^4)
// gives the partner basis of V (the same phases, the same choice inside degenerate multiplets) at the cost of three
// GEMMs.  The caller uses it only while S_k / S_0 is large enough for e << 1 (iDMRG.SPLIT_CUTOFF).
template <typename T>
__global__ void __launch_bounds__(256)
inv_scale_rows_kernel(int rows, int cols, const T* __restrict__ in, const double* __restrict__ sv, T* __restrict__ out,
                      long long s_in, long long s_sv, long long s_out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)rows * cols) return;
    const int i = (int)(e / cols);
    const double v = sv[(size_t)blockIdx.y * s_sv + i];
    out[(size_t)blockIdx.y * s_out + e] = Sc<T>::scale(in[(size_t)blockIdx.y * s_in + e], v > 0.0 ? 1.0 / v : 0.0);
}
template <typename T>
__global__ void __launch_bounds__(256) ns_matrix_kernel(int k, const T* __restrict__ G, T* __restrict__ M) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)k * k) return;
    const size_t off = (size_t)blockIdx.y * k * k + e;
    const int i = (int)(e / k), j = (int)(e - (long long)i * k);
    M[off] = Sc<T>::add(Sc<T>::from_re(i == j ? 1.5 : 0.0), Sc<T>::scale(G[off], -0.5));
}

size_t left_vectors_bytes(int batch, int rows, int cols, int k, size_t esz) {
    const size_t b = (size_t)batch;
    return al(b * k * cols * esz) + al(b * rows * k * esz) + 2 * al(b * k * k * esz);
}

// A [batch][rows][cols] (pitch lda, stride sA), Vh [batch][k][cols], S [batch][.] (stride sS, first k used),
// U_out [batch][rows][k]
template <typename T>
int left_vectors_impl(int batch, int rows, int cols, int k, const void* A, int lda, long long sA, const void* Vh,
                      const double* S, long long sS, void* U_out, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && batch <= 65535 && rows >= 1 && cols >= 1 && k >= 1 && k <= rows && k <= cols && A && Vh && S && U_out,
                 "left_vectors: bad arguments");
    const size_t need = left_vectors_bytes(batch, rows, cols, k, sizeof(T));
    MPSB_REQUIRE(ws && ws_bytes >= need, "left_vectors: workspace %zu B < required %zu B", ws_bytes, need);
    const size_t b = (size_t)batch;
    unsigned char* p = static_cast<unsigned char*>(ws);
    T* Vs = reinterpret_cast<T*>(p); p += al(b * k * cols * sizeof(T));
    T* Y = reinterpret_cast<T*>(p);  p += al(b * rows * k * sizeof(T));
    T* G = reinterpret_cast<T*>(p);  p += al(b * k * k * sizeof(T));
    T* M = reinterpret_cast<T*>(p);
    {
        dim3 grid((unsigned)(((long long)k * cols + 255) / 256), batch);
        inv_scale_rows_kernel<T><<<grid, 256, 0, st>>>(k, cols, static_cast<const T*>(Vh), S, Vs, (long long)k * cols, sS,
                                                       (long long)k * cols);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
    }
    GemmArgsT<T> g;
    g.batch1 = batch; g.batch2 = 1; g.sA2 = g.sB2 = g.sC2 = 0; g.accumulate = 0;
    // Y = A Vs^H
    g.A = static_cast<const T*>(A); g.B = Vs; g.C = Y; g.M = rows; g.N = k; g.K = cols; g.lda = lda; g.ldb = cols; g.ldc = k;
    g.opA = 0; g.opB = 2; g.sA1 = sA; g.sB1 = (long long)k * cols; g.sC1 = (long long)rows * k;
    int rc = gemm(g, st);
    if (rc) return rc;
    // two Newton-Schulz steps  Y <- Y (3 I - Y^H Y) / 2:  a defect e = |Y^H Y - I| becomes ~e^2, then ~e^4
    T* src = Y;
    T* dst = static_cast<T*>(U_out);
    for (int step = 0; step < 2; ++step) {
        g.A = src; g.B = src; g.C = G; g.M = k; g.N = k; g.K = rows; g.lda = k; g.ldb = k; g.ldc = k; g.opA = 2; g.opB = 0;
        g.sA1 = (long long)rows * k; g.sB1 = (long long)rows * k; g.sC1 = (long long)k * k;
        rc = gemm(g, st);
        if (rc) return rc;
        dim3 grid((unsigned)(((long long)k * k + 255) / 256), batch);
        ns_matrix_kernel<T><<<grid, 256, 0, st>>>(k, G, M);
        MPSB_COUNT_LAUNCH();
        MPSB_CHECK_CUDA(cudaGetLastError());
        g.A = src; g.B = M; g.C = dst; g.M = rows; g.N = k; g.K = k; g.lda = k; g.ldb = k; g.ldc = k;
        g.opA = 0; g.opB = 0; g.sA1 = (long long)rows * k; g.sB1 = (long long)k * k; g.sC1 = (long long)rows * k;
        rc = gemm(g, st);
        if (rc) return rc;
        T* t = src; src = dst; dst = t;
    }
    // after two steps the result sits in Y again
    MPSB_CHECK_CUDA(cudaMemcpyAsync(U_out, Y, b * rows * k * sizeof(T), cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // namespace
}  // namespace mpsb

using namespace mpsb;

extern "C" {

// The *_real entry points take tensors of doubles with the same shapes, strides (in doubles) and workspace queries as
// their complex twins (the queries return the complex size, an upper bound for the real call).
size_t mpsb_heff_batched_workspace(int batch, int chil, int chir, int d, int w) {
    return heff_bytes(chil, chir, d, w, batch);
}
size_t mpsb_heff_workspace(int chil, int chir, int d, int w) { return heff_bytes(chil, chir, d, w); }

int mpsb_heff_matvec_batched(int batch, int chil, int chir, int d, int w, const void* L, long long sL, const void* W,
                             long long sW, const void* R, long long sR, const void* theta, long long s_theta, void* out,
                             long long s_out, void* ws, size_t ws_bytes, void* stream) {
    return heff_matvec_impl<cplx>(batch, chil, chir, d, w, L, sL, W, sW, R, sR, theta, s_theta, out, s_out, ws, ws_bytes,
                                  stream);
}
int mpsb_heff_matvec_batched_real(int batch, int chil, int chir, int d, int w, const void* L, long long sL, const void* W,
                                  long long sW, const void* R, long long sR, const void* theta, long long s_theta,
                                  void* out, long long s_out, void* ws, size_t ws_bytes, void* stream) {
    return heff_matvec_impl<double>(batch, chil, chir, d, w, L, sL, W, sW, R, sR, theta, s_theta, out, s_out, ws,
                                    ws_bytes, stream);
}
int mpsb_heff_matvec(int chil, int chir, int d, int w, const void* L, const void* W, const void* R, const void* theta,
                     void* out, void* ws, size_t ws_bytes, void* stream) {
    return mpsb_heff_matvec_batched(1, chil, chir, d, w, L, 0, W, 0, R, 0, theta, 0, out, 0, ws, ws_bytes, stream);
}

unsigned long long mpsb_last_lanczos_chain_matvecs(void) { return g_last_chain_matvecs; }

size_t mpsb_lanczos_batched_workspace(int batch, int chil, int chir, int d, int w, int krylov) {
    return lanczos_bytes(batch, chil, chir, d, w, krylov, sizeof(cplx));
}
size_t mpsb_lanczos_workspace(int chil, int chir, int d, int w, int krylov) {
    return mpsb_lanczos_batched_workspace(1, chil, chir, d, w, krylov);
}
int mpsb_lanczos_ground_batched(int batch, int chil, int chir, int d, int w, const void* L, long long sL, const void* W,
                                long long sW, const void* R, long long sR, void* theta, long long s_theta, int krylov,
                                int max_restarts, double tol, double* energy_host, double* resid_host, int* n_matvec_host,
                                void* ws, size_t ws_bytes, void* stream) {
    return lanczos_impl<cplx>(batch, chil, chir, d, w, L, sL, W, sW, R, sR, theta, s_theta, krylov, max_restarts, tol,
                              energy_host, resid_host, n_matvec_host, ws, ws_bytes, stream);
}
int mpsb_lanczos_ground_batched_real(int batch, int chil, int chir, int d, int w, const void* L, long long sL,
                                     const void* W, long long sW, const void* R, long long sR, void* theta,
                                     long long s_theta, int krylov, int max_restarts, double tol, double* energy_host,
                                     double* resid_host, int* n_matvec_host, void* ws, size_t ws_bytes, void* stream) {
    return lanczos_impl<double>(batch, chil, chir, d, w, L, sL, W, sW, R, sR, theta, s_theta, krylov, max_restarts, tol,
                                energy_host, resid_host, n_matvec_host, ws, ws_bytes, stream);
}
int mpsb_lanczos_ground(int chil, int chir, int d, int w, const void* L, const void* W, const void* R, void* theta,
                        int krylov, int max_restarts, double tol, double* energy_host, double* resid_host,
                        int* n_matvec_host, void* ws, size_t ws_bytes, void* stream) {
    const long long n = (long long)chil * d * d * chir;
    return mpsb_lanczos_ground_batched(1, chil, chir, d, w, L, 0, W, 0, R, 0, theta, n, krylov, max_restarts, tol,
                                       energy_host, resid_host, n_matvec_host, ws, ws_bytes, stream);
}

size_t mpsb_predict_batched_workspace(int batch, int chi, int chip, int d) {
    return 2 * al((size_t)batch * chi * d * chip * sizeof(cplx));
}
size_t mpsb_predict_workspace(int chi, int chip, int d) { return mpsb_predict_batched_workspace(1, chi, chip, d); }
int mpsb_predict_theta_batched(int batch, int chi, int chip, int d, const void* A, const void* B, const double* S,
                               const double* S_prev_inv, void* theta, void* ws, size_t ws_bytes, void* stream) {
    return predict_impl<cplx>(batch, chi, chip, d, A, B, S, S_prev_inv, theta, ws, ws_bytes, stream);
}
int mpsb_predict_theta_batched_real(int batch, int chi, int chip, int d, const void* A, const void* B, const double* S,
                                    const double* S_prev_inv, void* theta, void* ws, size_t ws_bytes, void* stream) {
    return predict_impl<double>(batch, chi, chip, d, A, B, S, S_prev_inv, theta, ws, ws_bytes, stream);
}
int mpsb_predict_theta(int chi, int chip, int d, const void* A, const void* B, const double* S, const double* S_prev_inv,
                       void* theta, void* ws, size_t ws_bytes, void* stream) {
    return mpsb_predict_theta_batched(1, chi, chip, d, A, B, S, S_prev_inv, theta, ws, ws_bytes, stream);
}

size_t mpsb_env_batched_workspace(int batch, int chi, int chin, int d, int w) {
    return env_bytes(batch, chi, chin, d, w, sizeof(cplx));
}
size_t mpsb_env_workspace(int chi, int chin, int d, int w) { return mpsb_env_batched_workspace(1, chi, chin, d, w); }
int mpsb_env_left(int chi, int chin, int d, int w, const void* L, const void* U, const void* W, int swap_conj,
                  void* L_out, void* ws, size_t ws_bytes, void* stream) {
    return env_common<cplx>(0, 1, chi, chin, d, w, L, 0, U, 0, W, 0, swap_conj, L_out, 0, ws, ws_bytes, stream);
}
int mpsb_env_right(int chi, int chin, int d, int w, const void* R, const void* V, const void* W, int swap_conj,
                   void* R_out, void* ws, size_t ws_bytes, void* stream) {
    return env_common<cplx>(1, 1, chi, chin, d, w, R, 0, V, 0, W, 0, swap_conj, R_out, 0, ws, ws_bytes, stream);
}
int mpsb_env_left_batched(int batch, int chi, int chin, int d, int w, const void* L, long long sL, const void* U, long long sU,
                          const void* W, long long sW, int swap_conj, void* L_out, long long sLo, void* ws, size_t ws_bytes,
                          void* stream) {
    return env_common<cplx>(0, batch, chi, chin, d, w, L, sL, U, sU, W, sW, swap_conj, L_out, sLo, ws, ws_bytes, stream);
}
int mpsb_env_right_batched(int batch, int chi, int chin, int d, int w, const void* R, long long sR, const void* V, long long sV,
                           const void* W, long long sW, int swap_conj, void* R_out, long long sRo, void* ws, size_t ws_bytes,
                           void* stream) {
    return env_common<cplx>(1, batch, chi, chin, d, w, R, sR, V, sV, W, sW, swap_conj, R_out, sRo, ws, ws_bytes, stream);
}
int mpsb_env_left_batched_real(int batch, int chi, int chin, int d, int w, const void* L, long long sL, const void* U,
                               long long sU, const void* W, long long sW, void* L_out, long long sLo, void* ws,
                               size_t ws_bytes, void* stream) {
    return env_common<double>(0, batch, chi, chin, d, w, L, sL, U, sU, W, sW, 0, L_out, sLo, ws, ws_bytes, stream);
}
int mpsb_env_right_batched_real(int batch, int chi, int chin, int d, int w, const void* R, long long sR, const void* V,
                                long long sV, const void* W, long long sW, void* R_out, long long sRo, void* ws,
                                size_t ws_bytes, void* stream) {
    return env_common<double>(1, batch, chi, chin, d, w, R, sR, V, sV, W, sW, 0, R_out, sRo, ws, ws_bytes, stream);
}

size_t mpsb_left_vectors_workspace(int batch, int rows, int cols, int k) {
    return left_vectors_bytes(batch, rows, cols, k, sizeof(cplx));
}
int mpsb_left_vectors(int batch, int rows, int cols, int k, const void* A, int lda, long long sA, const void* Vh,
                      const double* S, long long sS, void* U_out, int real, void* ws, size_t ws_bytes, void* stream) {
    if (real) return left_vectors_impl<double>(batch, rows, cols, k, A, lda, sA, Vh, S, sS, U_out, ws, ws_bytes, stream);
    return left_vectors_impl<cplx>(batch, rows, cols, k, A, lda, sA, Vh, S, sS, U_out, ws, ws_bytes, stream);
}

}  // extern "C"
