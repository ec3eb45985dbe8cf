// Layout and BLAS-1 helpers around the GEMM / Jacobi kernels.  All HBM-bound, written for coalesced
// 16-byte accesses; none of them has a counterpart in the reference, which leaves index shuffling
// to np.einsum (DMRG/MPS.py:257-284, DMRG/iDMRG.py:40-60) and vector algebra to ARPACK.
#include "common.cuh"
#include "mpsb_internal.h"

namespace mpsb {

namespace {

__device__ __forceinline__ cplx fetch(const MatSrc& m, int b, int i, int j) {
    cplx v = m.src[(long long)b * m.bs + (long long)i * m.rs + (long long)j * m.cs];
    if (m.conj) v.y = -v.y;
    if (m.scale) {
        const double s = m.scale[(long long)b * m.scale_bs + (j / m.sdiv) % m.smod];
        v.x *= s;
        v.y *= s;
    }
    return v;
}

__global__ void __launch_bounds__(256)
pack_kernel(cplx* __restrict__ X, int n_pad, int ld, int rows, int ldot, int ltot, MatSrc dot, MatSrc aug,
            int identity) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_pad * ld) return;
    const int b = blockIdx.y;
    const int j = (int)(e % ld), i = (int)(e / ld);
    cplx v = make_double2(0.0, 0.0);
    if (i < rows) {
        if (j < ldot) v = fetch(dot, b, i, j);
        else if (j < ltot) {
            if (aug.src) v = fetch(aug, b, i, j - ldot);
            else if (identity && (j - ldot) == i) v = make_double2(1.0, 0.0);
        }
    }
    X[(size_t)b * n_pad * ld + e] = v;
}

// Real path of the SVD (mpsb_svd_trunc_real): the real matrix enters the complex work matrix with a zero imaginary
// part (the QR kernels then stay real, every product with an exact zero being zero) ...
__global__ void __launch_bounds__(256)
pack_real_kernel(cplx* __restrict__ X, int n_pad, int ld, int rows, int cols, int ldot, const double* __restrict__ A,
                 int lda, long long sA, int identity) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_pad * ld) return;
    const int b = blockIdx.y;
    const int j = (int)(e % ld), i = (int)(e / ld);
    double v = 0.0;
    if (i < rows) {
        if (j < cols) v = A[(size_t)b * sA + (size_t)i * lda + j];
        else if (identity && j == ldot + i) v = 1.0;
    }
    X[(size_t)b * n_pad * ld + e] = make_double2(v, 0.0);
}
// ... and after the QR adjacent columns are packed into one complex element each for the Jacobi sweeps, which halves
// the row length (and lets rotate_pair & co. run real rotations on the packed rows).
__global__ void __launch_bounds__(256)
compact_pairs_kernel(const cplx* __restrict__ X, int n_pad, int ld, int rows, cplx* __restrict__ Xp, int n_pad_p, int ld_p,
                     int ltot_p) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_pad_p * ld_p) return;
    const int b = blockIdx.y;
    const int j = (int)(e % ld_p), i = (int)(e / ld_p);
    cplx v = make_double2(0.0, 0.0);
    if (i < rows && j < ltot_p) {
        const double4 two = *reinterpret_cast<const double4*>(X + (size_t)b * n_pad * ld + (size_t)i * ld + 2 * j);
        v = make_double2(two.x, two.z);
    }
    Xp[(size_t)b * n_pad_p * ld_p + e] = v;
}

// 32x32 tile transpose through shared memory (padded against bank conflicts).
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ in, int ldi, int rows, int cols, T* __restrict__ out, int ldo,
                 int conj, long long s_in, long long s_out) {
    __shared__ T tile[32][33];
    in += (size_t)blockIdx.z * s_in;
    out += (size_t)blockIdx.z * s_out;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        if (i < rows && j < cols) tile[r][tx] = in[(size_t)i * ldi + j];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, i = i0 + tx;
        if (i < rows && j < cols) {
            T v = tile[tx][r];
            if (conj) v = Sc<T>::conj(v);
            out[(size_t)j * ldo + i] = v;
        }
    }
}

__device__ __forceinline__ double warp_sum_t(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ cplx warp_sum_t(cplx v) { return make_double2(warp_sum_t(v.x), warp_sum_t(v.y)); }

constexpr int DOT_THREADS = 256;
// y[j] = sum_i conj(V[j][i]) w[i]: every block reduces one chunk of one vector into part[j][chunk]; the block that
// draws the last ticket of vector j adds the partials in a fixed order (deterministic) and re-arms the ticket.
// `part` holds m * nchunk partial sums followed by m ticket counters (zero on entry, zero again on exit).
template <typename T>
__global__ void __launch_bounds__(DOT_THREADS)
dot_fused_kernel(const T* __restrict__ V, long long ldv, long long n, const T* __restrict__ w,
                 T* __restrict__ part, int nchunk, int m_cap, T* __restrict__ y, BlasBatch bb) {
    __shared__ T sred[DOT_THREADS / 32];
    __shared__ int is_last;
    const int j = blockIdx.y, ch = blockIdx.x;
    V += (size_t)blockIdx.z * bb.sV; w += (size_t)blockIdx.z * bb.sw;
    part += (size_t)blockIdx.z * bb.spart; y += (size_t)blockIdx.z * bb.sc;
    unsigned int* tickets = reinterpret_cast<unsigned int*>(part + (size_t)m_cap * nchunk);
    const long long per = (n + nchunk - 1) / nchunk;
    const long long lo = (long long)ch * per, hi = lo + per < n ? lo + per : n;
    const T* v = V + (size_t)j * ldv;
    T acc = Sc<T>::zero();
    for (long long i = lo + threadIdx.x; i < hi; i += DOT_THREADS) Sc<T>::macc(acc, v[i], w[i]);
    acc = warp_sum_t(acc);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T a = Sc<T>::zero();
        for (int q = 0; q < DOT_THREADS / 32; ++q) a = Sc<T>::add(a, sred[q]);
        part[(size_t)j * nchunk + ch] = a;
        __threadfence();
        is_last = (atomicAdd(&tickets[j], 1u) == (unsigned)(nchunk - 1));
    }
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        T a = Sc<T>::zero();
        for (int q = threadIdx.x; q < nchunk; q += 32) a = Sc<T>::add(a, __ldcg(&part[(size_t)j * nchunk + q]));
        a = warp_sum_t(a);
        if (threadIdx.x == 0) {
            y[j] = a;
            tickets[j] = 0u;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
axpy_multi_kernel(const T* __restrict__ V, long long ldv, int m, long long n, const T* __restrict__ c,
                  double sign, T* __restrict__ w, BlasBatch bb) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V += (size_t)blockIdx.y * bb.sV; w += (size_t)blockIdx.y * bb.sw; c += (size_t)blockIdx.y * bb.sc;
    T acc = w[i];
    for (int j = 0; j < m; ++j) Sc<T>::mac(acc, Sc<T>::scale(c[j], sign), V[(size_t)j * ldv + i]);
    w[i] = acc;
}

// w *= 1/sqrt(re(nrm2[0])); a vanishing norm (Lanczos breakdown) zeroes the vector instead of producing NaNs
template <typename T>
__global__ void __launch_bounds__(256) normalise_kernel(long long n, const T* __restrict__ nrm2, T* __restrict__ w,
                                                        BlasBatch bb) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    nrm2 += (size_t)blockIdx.y * bb.sc; w += (size_t)blockIdx.y * bb.sw;
    const double v2 = Sc<T>::re(nrm2[0]);
    const double sc = v2 > 0.0 ? rsqrt(v2) : 0.0;
    w[i] = Sc<T>::scale(w[i], sc);
}

__global__ void __launch_bounds__(256) dscal_kernel(long long n, double alpha, cplx* __restrict__ w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx v = w[i];
    v.x *= alpha;
    v.y *= alpha;
    w[i] = v;
}

}  // namespace

int pack_work_matrix(cplx* X, int n_pad, int ld, int rows, int ldot, int ltot, const MatSrc& dot, const MatSrc& aug,
                     int identity, cudaStream_t st, int batch) {
    const long long total = (long long)n_pad * ld;
    MPSB_REQUIRE(batch >= 1 && batch <= 65535, "pack: batch %d exceeds grid.y", batch);
    dim3 grid((unsigned)((total + 255) / 256), batch);
    pack_kernel<<<grid, 256, 0, st>>>(X, n_pad, ld, rows, ldot, ltot, dot, aug, identity);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int pack_real_matrix(cplx* X, int n_pad, int ld, int rows, int cols, int ldot, const double* A, int lda, long long sA,
                     int identity, cudaStream_t st, int batch) {
    const long long total = (long long)n_pad * ld;
    MPSB_REQUIRE(batch >= 1 && batch <= 65535, "pack: batch %d exceeds grid.y", batch);
    dim3 grid((unsigned)((total + 255) / 256), batch);
    pack_real_kernel<<<grid, 256, 0, st>>>(X, n_pad, ld, rows, cols, ldot, A, lda, sA, identity);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int compact_real_pairs(const cplx* X, int n_pad, int ld, int rows, cplx* Xp, int n_pad_p, int ld_p, int ltot_p,
                       cudaStream_t st, int batch) {
    const long long total = (long long)n_pad_p * ld_p;
    MPSB_REQUIRE(batch >= 1 && batch <= 65535 && 2 * ltot_p <= ld, "compact: bad batch %d / widths %d, %d", batch, ltot_p, ld);
    dim3 grid((unsigned)((total + 255) / 256), batch);
    compact_pairs_kernel<<<grid, 256, 0, st>>>(X, n_pad, ld, rows, Xp, n_pad_p, ld_p, ltot_p);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int transpose_t(const T* in, int ldi, int rows, int cols, T* out, int ldo, int conj, cudaStream_t st,
                       int batch, long long s_in, long long s_out) {
    if (rows == 0 || cols == 0) return 0;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
    MPSB_REQUIRE(grid.y <= 65535 && batch >= 1 && batch <= 65535, "transpose: %d rows / batch %d exceed the grid", rows, batch);
    transpose_kernel<T><<<grid, 256, 0, st>>>(in, ldi, rows, cols, out, ldo, conj, s_in, s_out);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int transpose_conj(const cplx* in, int ldi, int rows, int cols, cplx* out, int ldo, int conj, cudaStream_t st,
                   int batch, long long s_in, long long s_out) {
    return transpose_t(in, ldi, rows, cols, out, ldo, conj, st, batch, s_in, s_out);
}
int transpose_conj(const double* in, int ldi, int rows, int cols, double* out, int ldo, int conj, cudaStream_t st,
                   int batch, long long s_in, long long s_out) {
    return transpose_t(in, ldi, rows, cols, out, ldo, conj, st, batch, s_in, s_out);
}

// out[i][j] = s[i / group] * in[i][j]: the Lambda_l scaling of a (chi_l * group) x cols matrix view (DMRG/MPS.py:446).
__global__ void __launch_bounds__(256)
scale_rows_kernel(int rows, int cols, const cplx* __restrict__ in, int ldi, const double* __restrict__ s, int group,
                  cplx* __restrict__ out, int ldo) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)rows * cols) return;
    const int i = (int)(e / cols), j = (int)(e - (long long)i * cols);
    const double f = s[i / group];
    const cplx x = in[(size_t)i * ldi + j];
    out[(size_t)i * ldo + j] = make_double2(f * x.x, f * x.y);
}

int scale_rows(int rows, int cols, const cplx* in, int ldi, const double* s, int group, cplx* out, int ldo,
               cudaStream_t st) {
    MPSB_REQUIRE(rows >= 0 && cols >= 0 && group >= 1, "scale_rows: bad dims");
    const long long total = (long long)rows * cols;
    if (total == 0) return 0;
    scale_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rows, cols, in, ldi, s, group, out, ldo);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int perm_last_to_first(const cplx* in, int n_ab, int n_c, cplx* out, cudaStream_t st) {
    return transpose_conj(in, n_c, n_ab, n_c, out, n_ab, 0, st);
}
int perm_first_to_last(const cplx* in, int n_c, int n_ab, cplx* out, cudaStream_t st) {
    return transpose_conj(in, n_ab, n_c, n_ab, out, n_c, 0, st);
}

template <typename T>
static int dotc_t(const T* V, long long ldv, int m, long long n, const T* w, T* y, T* part, int nchunk, int m_cap,
                  cudaStream_t st, BlasBatch bb) {
    if (m == 0) return 0;
    MPSB_REQUIRE(m <= 65535 && m <= m_cap && nchunk >= 1 && bb.batch >= 1 && bb.batch <= 65535,
                 "zdotc_multi: bad m=%d (cap %d) nchunk=%d batch=%d", m, m_cap, nchunk, bb.batch);
    dim3 grid(nchunk, m, bb.batch);
    dot_fused_kernel<T><<<grid, DOT_THREADS, 0, st>>>(V, ldv, n, w, part, nchunk, m_cap, y, bb);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int zdotc_multi(const cplx* V, long long ldv, int m, long long n, const cplx* w, cplx* y, cplx* part, int nchunk,
                int m_cap, cudaStream_t st, BlasBatch bb) {
    return dotc_t(V, ldv, m, n, w, y, part, nchunk, m_cap, st, bb);
}
int zdotc_multi(const double* V, long long ldv, int m, long long n, const double* w, double* y, double* part, int nchunk,
                int m_cap, cudaStream_t st, BlasBatch bb) {
    return dotc_t(V, ldv, m, n, w, y, part, nchunk, m_cap, st, bb);
}

template <typename T>
static int axpy_t(const T* V, long long ldv, int m, long long n, const T* c, double sign, T* w, cudaStream_t st,
                  BlasBatch bb) {
    if (m == 0 || n == 0) return 0;
    dim3 grid((unsigned)((n + 255) / 256), bb.batch);
    axpy_multi_kernel<T><<<grid, 256, 0, st>>>(V, ldv, m, n, c, sign, w, bb);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int zaxpy_multi(const cplx* V, long long ldv, int m, long long n, const cplx* c, double sign, cplx* w,
                cudaStream_t st, BlasBatch bb) {
    return axpy_t(V, ldv, m, n, c, sign, w, st, bb);
}
int zaxpy_multi(const double* V, long long ldv, int m, long long n, const double* c, double sign, double* w,
                cudaStream_t st, BlasBatch bb) {
    return axpy_t(V, ldv, m, n, c, sign, w, st, bb);
}

template <typename T>
static int normalise_t(long long n, const T* nrm2, T* w, cudaStream_t st, BlasBatch bb) {
    if (n == 0) return 0;
    dim3 grid((unsigned)((n + 255) / 256), bb.batch);
    normalise_kernel<T><<<grid, 256, 0, st>>>(n, nrm2, w, bb);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int znormalise(long long n, const cplx* nrm2, cplx* w, cudaStream_t st, BlasBatch bb) { return normalise_t(n, nrm2, w, st, bb); }
int znormalise(long long n, const double* nrm2, double* w, cudaStream_t st, BlasBatch bb) { return normalise_t(n, nrm2, w, st, bb); }

int zdscal(long long n, double alpha, cplx* w, cudaStream_t st) {
    if (n == 0) return 0;
    dscal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, alpha, w);
    MPSB_COUNT_LAUNCH();
    MPSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mpsb
