// extern "C" surface of libmps_b200 (declared in include/mpsb.h) and the launch sequences behind the
// MPS entry points: update_double / update_single / svd_truncate / ortho_*_site / dot / corr of
// DMRG/MPS.py.  The iDMRG entry points live in idmrg.cu.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <map>
#include <utility>
#include "../../include/mpsb.h"
#include "common.cuh"
#include "mpsb_internal.h"

namespace mpsb {

unsigned long long g_launch_count = 0;
static thread_local char g_error[1024] = "";
static thread_local int g_last_sweeps = 0;

Profile g_prof = {};
void prof_mark(int idx, cudaStream_t st) {
    if (!g_prof.enabled) return;
    if (!g_prof.have_events) {
        for (int i = 0; i < 8; ++i) cudaEventCreate(&g_prof.ev[i]);
        g_prof.have_events = 1;
    }
    cudaEventRecord(g_prof.ev[idx], st);
    g_prof.recorded |= 1u << idx;
}

int ensure_dynamic_smem(const void* func, size_t bytes) {
    static thread_local std::map<std::pair<const void*, int>, size_t> granted;
    int dev = 0;
    MPSB_CHECK_CUDA(cudaGetDevice(&dev));
    size_t& have = granted[std::make_pair(func, dev)];
    if (bytes > have) {
        MPSB_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return 0;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

namespace {

inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

// Bump allocator over the caller's workspace.
struct Arena {
    unsigned char* base;
    size_t size, used;
    Arena(void* p, size_t n) : base(static_cast<unsigned char*>(p)), size(n), used(0) {}
    void* take(size_t bytes) {
        void* r = base + used;
        used += al(bytes);
        return r;
    }
};

struct TebdLayout {
    SvdPlan plan;
    size_t theta0, thetaB, X, state, total;
};
int tebd_layout(int batch, int chil, int chi, int chir, int d, TebdLayout* L) {
    (void)chi;
    const int n = chil * d, l = d * chir;
    int rc = svd_make_plan(batch, n, l, l, &L->plan);
    if (rc) return rc;
    const size_t mat = (size_t)n * l * sizeof(cplx);
    L->theta0 = al(mat * batch);
    L->thetaB = al(mat * batch);
    L->X = al((size_t)L->plan.n_pad * L->plan.ld * sizeof(cplx) * batch);
    L->state = al(svd_state_bytes(L->plan));
    L->total = L->theta0 + L->thetaB + L->X + L->state;
    return 0;
}

struct SvdLayout {
    SvdPlan plan;
    size_t X, state, total;
};
int svd_layout(int batch, int rows, int cols, int naug, SvdLayout* L) {
    int rc = svd_make_plan(batch, rows, cols, cols + naug, &L->plan);
    if (rc) return rc;
    L->X = al((size_t)L->plan.n_pad * L->plan.ld * sizeof(cplx) * batch);
    L->state = al(svd_state_bytes(L->plan));
    L->total = L->X + L->state;
    return 0;
}

}  // namespace
}  // namespace mpsb

using namespace mpsb;

extern "C" {

int mpsb_version(void) { return 100; }
const char* mpsb_last_error(void) { return g_error; }
unsigned long long mpsb_launch_count(void) { return g_launch_count; }
int mpsb_last_svd_sweeps(void) { return g_last_sweeps; }
int mpsb_debug_sweep_schedule(int nblk, int quad, int* out_host, int cap) { return sweep_schedule(nblk, quad, out_host, cap); }

void mpsb_profile_enable(int on) {
    g_prof.enabled = on;
    g_prof.recorded = 0u;
}
int mpsb_profile_read(double* phase_ms, unsigned long long* rotations, unsigned long long* problem_sweeps) {
    MPSB_REQUIRE(g_prof.enabled && g_prof.have_events, "profile_read: profiling was not enabled for the last call");
    // Calls other than the two-site update only pass some of the marks (an SVD alone: 3 and 4); phases without both
    // ends read as -1.
    for (int i = 0; i < 6; ++i) {
        phase_ms[i] = -1.0;
        if ((g_prof.recorded >> i & 3u) != 3u) continue;
        float ms = 0.f;
        MPSB_CHECK_CUDA(cudaEventSynchronize(g_prof.ev[i + 1]));
        MPSB_CHECK_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]));
        phase_ms[i] = ms;
    }
    if (rotations) *rotations = g_prof.rotations;
    if (problem_sweeps) *problem_sweeps = g_prof.problem_sweeps;
    return 0;
}

int mpsb_zgemm(int opA, int opB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C,
               int ldc, int batch1, int batch2, long long sA1, long long sA2, long long sB1, long long sB2,
               long long sC1, long long sC2, int accumulate, void* stream) {
    GemmArgs p;
    p.A = static_cast<const cplx*>(A);
    p.B = static_cast<const cplx*>(B);
    p.C = static_cast<cplx*>(C);
    p.M = M; p.N = N; p.K = K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.opA = opA; p.opB = opB;
    p.batch1 = batch1; p.batch2 = batch2;
    p.sA1 = sA1; p.sA2 = sA2; p.sB1 = sB1; p.sB2 = sB2; p.sC1 = sC1; p.sC2 = sC2;
    p.accumulate = accumulate;
    return zgemm(p, static_cast<cudaStream_t>(stream));
}

int mpsb_dgemm(int opA, int opB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C,
               int ldc, int batch1, int batch2, long long sA1, long long sA2, long long sB1, long long sB2,
               long long sC1, long long sC2, int accumulate, void* stream) {
    GemmArgsT<double> p;
    p.A = static_cast<const double*>(A);
    p.B = static_cast<const double*>(B);
    p.C = static_cast<double*>(C);
    p.M = M; p.N = N; p.K = K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.opA = opA; p.opB = opB;
    p.batch1 = batch1; p.batch2 = batch2;
    p.sA1 = sA1; p.sA2 = sA2; p.sB1 = sB1; p.sB2 = sB2; p.sC1 = sC1; p.sC2 = sC2;
    p.accumulate = accumulate;
    return dgemm(p, static_cast<cudaStream_t>(stream));
}

// ---- svd_truncate / halve ---------------------------------------------------------------------------
size_t mpsb_svd_trunc_workspace(int batch, int rows, int cols, int want_u) {
    SvdLayout L;
    if (svd_layout(batch, rows, cols, want_u ? rows : 0, &L)) return 0;
    return L.total;
}

int mpsb_svd_trunc(int batch, int rows, int cols, const void* A, int lda, long long sA, int trun, double err,
                   int normalise, int k_cap, void* Vh, void* Uh, double* S, double* S_raw, int* rank, void* ws,
                   size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && rows >= 1 && cols >= 1 && k_cap >= 1 && trun >= 0, "svd_trunc: bad arguments");
    SvdLayout L;
    int rc = svd_layout(batch, rows, cols, Uh ? rows : 0, &L);
    if (rc) return rc;
    MPSB_REQUIRE(ws && ws_bytes >= L.total, "svd_trunc: workspace %zu B < required %zu B", ws_bytes, L.total);
    Arena ar(ws, ws_bytes);
    cplx* X = static_cast<cplx*>(ar.take(L.X));
    void* state = ar.take(L.state);
    const SvdPlan& pl = L.plan;
    {
        MatSrc dot = {static_cast<const cplx*>(A), lda, 1, 0, nullptr, 1, 1};
        dot.bs = sA;
        MatSrc aug = {nullptr, 0, 0, 0, nullptr, 1, 1};
        rc = pack_work_matrix(X, pl.n_pad, pl.ld, rows, pl.ldot, pl.ltot, dot, aug, Uh ? 1 : 0, st, batch);
        if (rc) return rc;
    }
    rc = svd_orthogonalise(pl, X, state, st, &g_last_sweeps);
    if (rc) return rc;
    SvdOutputs o;
    memset(&o, 0, sizeof(o));
    o.vh = static_cast<cplx*>(Vh); o.vh_stride = (long long)k_cap * cols; o.vh_ld = cols;
    o.aug = static_cast<cplx*>(Uh); o.aug_stride = (long long)k_cap * rows; o.aug_ld = rows; o.aug_conj = 0;
    o.s = S; o.s_stride = k_cap;
    o.s_raw = S_raw; o.n_raw = rows < cols ? rows : cols; o.s_raw_stride = o.n_raw;
    o.rank = rank; o.k_cap = k_cap; o.trun = trun; o.err = err; o.normalise = normalise;
    return svd_finalise(pl, X, state, o, st);
}

// Real matrices (the real iDMRG path): QR on the unpacked work matrix (zero imaginary parts stay zero), then the
// Jacobi sweeps on the copy with adjacent columns packed pairwise into complex elements - half the row length, real
// rotations - and a finalise that writes arrays of doubles.
namespace {
struct SvdRealLayout {
    SvdPlan qr, jac;           // unpacked (QR) and packed (sweeps) plans
    int ldot_c, ltot_c;        // unpacked widths, rounded up to even
    size_t X, Xp, state, total;
};
int svd_real_layout(int batch, int rows, int cols, int want_u, SvdRealLayout* L) {
    L->ldot_c = round_up(cols, 2);
    L->ltot_c = L->ldot_c + (want_u ? round_up(rows, 2) : 0);
    int rc = svd_make_plan(batch, rows, L->ldot_c, L->ltot_c, &L->qr);
    if (rc) return rc;
    rc = svd_make_plan(batch, rows, L->ldot_c / 2, L->ltot_c / 2, &L->jac);
    if (rc) return rc;
    L->qr.phase = 1;
    L->jac.phase = 2;
    L->jac.real = 1;
    L->jac.qr_grid = 0;
    L->X = al((size_t)L->qr.n_pad * L->qr.ld * sizeof(cplx) * batch);
    L->Xp = al((size_t)L->jac.n_pad * L->jac.ld * sizeof(cplx) * batch);
    L->state = al(std::max(svd_state_bytes(L->qr), svd_state_bytes(L->jac)));
    L->total = L->X + L->Xp + L->state;
    return 0;
}
}  // namespace

size_t mpsb_svd_trunc_real_workspace(int batch, int rows, int cols, int want_u) {
    SvdRealLayout L;
    if (svd_real_layout(batch, rows, cols, want_u, &L)) return 0;
    return L.total;
}

int mpsb_svd_trunc_real(int batch, int rows, int cols, const void* A, int lda, long long sA, int trun, double err,
                        int normalise, int k_cap, void* Vh, void* Uh, double* S, double* S_raw, int* rank, void* ws,
                        size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && rows >= 1 && cols >= 1 && k_cap >= 1 && trun >= 0, "svd_trunc_real: bad arguments");
    SvdRealLayout L;
    int rc = svd_real_layout(batch, rows, cols, Uh ? 1 : 0, &L);
    if (rc) return rc;
    MPSB_REQUIRE(ws && ws_bytes >= L.total, "svd_trunc_real: workspace %zu B < required %zu B", ws_bytes, L.total);
    Arena ar(ws, ws_bytes);
    cplx* X = static_cast<cplx*>(ar.take(L.X));
    cplx* Xp = static_cast<cplx*>(ar.take(L.Xp));
    void* state = ar.take(L.state);
    rc = pack_real_matrix(X, L.qr.n_pad, L.qr.ld, rows, cols, L.ldot_c, static_cast<const double*>(A), lda, sA, Uh ? 1 : 0, st,
                          batch);
    if (rc) return rc;
    rc = svd_orthogonalise(L.qr, X, state, st, nullptr);
    if (rc) return rc;
    rc = compact_real_pairs(X, L.qr.n_pad, L.qr.ld, rows, Xp, L.jac.n_pad, L.jac.ld, L.jac.ltot, st, batch);
    if (rc) return rc;
    rc = svd_orthogonalise(L.jac, Xp, state, st, &g_last_sweeps);
    if (rc) return rc;
    SvdOutputs o;
    memset(&o, 0, sizeof(o));
    o.real = 1; o.real_cols = cols; o.real_naug = Uh ? rows : 0;
    o.vh = static_cast<cplx*>(Vh); o.vh_stride = (long long)k_cap * cols; o.vh_ld = cols;
    o.aug = static_cast<cplx*>(Uh); o.aug_stride = (long long)k_cap * rows; o.aug_ld = rows; o.aug_conj = 0;
    o.s = S; o.s_stride = k_cap;
    o.s_raw = S_raw; o.n_raw = rows < cols ? rows : cols; o.s_raw_stride = o.n_raw;
    o.rank = rank; o.k_cap = k_cap; o.trun = trun; o.err = err; o.normalise = normalise;
    return svd_finalise(L.jac, Xp, state, o, st);
}

// ---- update_double -------------------------------------------------------------------------------------
size_t mpsb_tebd_two_site_workspace(int batch, int chil, int chi, int chir, int d) {
    TebdLayout L;
    if (tebd_layout(batch, chil, chi, chir, d, &L)) return 0;
    return L.total;
}

int mpsb_tebd_two_site(int batch, int chil, int chi, int chir, int d, const void* Bi, long long sBi, const void* Bj,
                       long long sBj, const double* Sl, long long sSl, const void* U, long long sU, int trun,
                       double err, int k_cap, void* Bi_out, long long sBio, void* Bj_out, long long sBjo,
                       double* Sr_out, long long sSr, int* rank, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && chil >= 1 && chi >= 1 && chir >= 1 && d >= 1 && k_cap >= 1 && trun >= 0,
                 "tebd_two_site: bad dims batch=%d chil=%d chi=%d chir=%d d=%d k_cap=%d", batch, chil, chi, chir, d, k_cap);
    TebdLayout L;
    int rc = tebd_layout(batch, chil, chi, chir, d, &L);
    if (rc) return rc;
    MPSB_REQUIRE(ws && ws_bytes >= L.total, "tebd_two_site: workspace %zu B < required %zu B", ws_bytes, L.total);
    Arena ar(ws, ws_bytes);
    cplx* theta0 = static_cast<cplx*>(ar.take(L.theta0));
    cplx* thetaB = static_cast<cplx*>(ar.take(L.thetaB));
    cplx* X = static_cast<cplx*>(ar.take(L.X));
    void* state = ar.take(L.state);
    const SvdPlan& pl = L.plan;
    const int n = chil * d, l = d * chir;
    const long long sTheta = (long long)n * l, sX = (long long)pl.n_pad * pl.ld;

    // theta0[(l,c),(j,k)] = sum_r Bi[(l,c), r] Bj[r, (j,k)]        (DMRG/MPS.py:445, first two operands)
    prof_mark(0, st);
    GemmArgs g;
    g.A = static_cast<const cplx*>(Bi); g.B = static_cast<const cplx*>(Bj); g.C = theta0;
    g.M = n; g.N = l; g.K = chi; g.lda = chi; g.ldb = l; g.ldc = l; g.opA = 0; g.opB = 0;
    g.batch1 = batch; g.batch2 = 1; g.sA1 = sBi; g.sB1 = sBj; g.sC1 = sTheta; g.sA2 = g.sB2 = g.sC2 = 0;
    g.accumulate = 0;
    rc = zgemm(g, st);
    if (rc) return rc;
    prof_mark(1, st);
    if (pl.n_pad != n || pl.ld != l) MPSB_CHECK_CUDA(cudaMemsetAsync(X, 0, sizeof(cplx) * sX * batch, st));
    // thetaB = U . theta0 ; X = Sl * thetaB                        (:445 third operand, :446)
    rc = gate_scale(batch, chil, chir, d, theta0, sTheta, static_cast<const cplx*>(U), sU, Sl, sSl, thetaB, sTheta, X,
                    sX, pl.ld, st);
    if (rc) return rc;
    prof_mark(2, st);
    // (s, Vh) = svd_truncate(X) -> Bj' = Vh, Sr = s, rank = k       (:448-452, :455)
    rc = svd_orthogonalise(pl, X, state, st, &g_last_sweeps);
    if (rc) return rc;
    SvdOutputs o;
    memset(&o, 0, sizeof(o));
    o.vh = static_cast<cplx*>(Bj_out); o.vh_stride = sBjo; o.vh_ld = l;
    o.s = Sr_out; o.s_stride = sSr;
    o.rank = rank; o.k_cap = k_cap; o.trun = trun; o.err = err; o.normalise = 1;
    rc = svd_finalise(pl, X, state, o, st);
    if (rc) return rc;
    prof_mark(5, st);
    // Bi'[(l,a), m] = sum_{(b,k)} thetaB[(l,a),(b,k)] conj(Vh[m,(b,k)])   (:454)
    g.A = thetaB; g.B = static_cast<const cplx*>(Bj_out); g.C = static_cast<cplx*>(Bi_out);
    g.M = n; g.N = k_cap; g.K = l; g.lda = l; g.ldb = l; g.ldc = k_cap; g.opA = 0; g.opB = 2;
    g.sA1 = sTheta; g.sB1 = sBjo; g.sC1 = sBio;
    rc = zgemm(g, st);
    prof_mark(6, st);
    return rc;
}

int mpsb_transpose(int rows, int cols, const void* in, int ldi, void* out, int ldo, int conj, void* stream) {
    return transpose_conj(static_cast<const cplx*>(in), ldi, rows, cols, static_cast<cplx*>(out), ldo, conj,
                          static_cast<cudaStream_t>(stream));
}

int mpsb_transpose_batched(int batch, int rows, int cols, const void* in, int ldi, long long s_in, void* out, int ldo,
                           long long s_out, int conj, void* stream) {
    return transpose_conj(static_cast<const cplx*>(in), ldi, rows, cols, static_cast<cplx*>(out), ldo, conj,
                          static_cast<cudaStream_t>(stream), batch, s_in, s_out);
}

int mpsb_transpose_batched_real(int batch, int rows, int cols, const void* in, int ldi, long long s_in, void* out, int ldo,
                                long long s_out, void* stream) {
    return transpose_conj(static_cast<const double*>(in), ldi, rows, cols, static_cast<double*>(out), ldo, 0,
                          static_cast<cudaStream_t>(stream), batch, s_in, s_out);
}

int mpsb_scale_rows(int rows, int cols, const void* in, int ldi, const double* s, int group, void* out, int ldo,
                    void* stream) {
    MPSB_REQUIRE(in && s && out, "scale_rows: null buffer");
    return scale_rows(rows, cols, static_cast<const cplx*>(in), ldi, s, group, static_cast<cplx*>(out), ldo,
                      static_cast<cudaStream_t>(stream));
}

int mpsb_one_site(int batch, int chil, int chir, int d, const void* M, long long sM, const void* U, long long sU,
                  void* M_out, long long sMo, void* stream) {
    return one_site(batch, chil, chir, d, static_cast<const cplx*>(M), sM, static_cast<const cplx*>(U), sU,
                    static_cast<cplx*>(M_out), sMo, static_cast<cudaStream_t>(stream));
}

int mpsb_mpo_site(int batch, int chil, int chir, int d, int wl, int wr, const void* M, long long sM, const void* W,
                  long long sW, void* out, long long sOut, void* stream) {
    MPSB_REQUIRE(M && W && out && M != out, "mpo_site: null or aliased buffers");
    return mpo_site(batch, chil, chir, d, wl, wr, static_cast<const cplx*>(M), sM, static_cast<const cplx*>(W), sW,
                    static_cast<cplx*>(out), sOut, static_cast<cudaStream_t>(stream));
}

// ---- canon: gauge steps ------------------------------------------------------------------------------------
// Left step: work matrix X = [ (Sl M_i)^H | M_next ], rows = right bond r (chir of them).  After the row
// orthogonalisation  X = [ diag(s) U^H | Vh M_next ]  (DMRG/MPS.py:257-264).
// Right step: X = [ M_i Sr | M_prev^H ], rows = left bond l; afterwards X = [ diag(s) Vh | (M_prev U)^H ]
// (DMRG/MPS.py:277-284).
size_t mpsb_gauge_batched_workspace(int batch, int chil, int chir, int d, int nother, int k_cap) {
    SvdLayout L, R;
    if (svd_layout(batch, chir, chil * d, nother, &L)) return 0;
    if (svd_layout(batch, chil, d * chir, nother, &R)) return 0;
    const size_t left = L.total + al((size_t)batch * k_cap * chil * d * sizeof(cplx));
    const size_t right = R.total + al((size_t)batch * k_cap * (nother > 0 ? nother : 1) * sizeof(cplx));
    return left > right ? left : right;
}
size_t mpsb_gauge_workspace(int chil, int chir, int d, int nother, int k_cap) {
    return mpsb_gauge_batched_workspace(1, chil, chir, d, nother, k_cap);
}

int mpsb_gauge_left_batched(int batch, int chil, int chir, int d, const void* M, long long sM, const double* Sl,
                            long long sSl, const void* M_next, long long sMn, int nnext, int trun, double err, int k_cap,
                            void* M_out, long long sMo, double* Sr_out, long long sSr, void* M_next_out, long long sMno,
                            int* rank, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && chil >= 1 && chir >= 1 && d >= 1 && nnext >= 0 && k_cap >= 1, "gauge_left: bad dims");
    const int rows = chir, ldot = chil * d;
    SvdLayout L;
    int rc = svd_layout(batch, rows, ldot, nnext, &L);
    if (rc) return rc;
    const size_t tmp_bytes = al((size_t)batch * k_cap * ldot * sizeof(cplx));
    MPSB_REQUIRE(ws && ws_bytes >= L.total + tmp_bytes, "gauge_left: workspace %zu B < required %zu B", ws_bytes,
                 L.total + tmp_bytes);
    Arena ar(ws, ws_bytes);
    cplx* X = static_cast<cplx*>(ar.take(L.X));
    void* state = ar.take(L.state);
    cplx* tmp = static_cast<cplx*>(ar.take(tmp_bytes));
    const SvdPlan& pl = L.plan;
    // X[r][(l,c)] = conj(Sl[l] M[l,c,r]);   X[r][ldot + q] = M_next[r][q]
    MatSrc dot = {static_cast<const cplx*>(M), 1, chir, 1, Sl, d, chil};
    dot.bs = sM; dot.scale_bs = sSl;
    MatSrc aug = {static_cast<const cplx*>(M_next), nnext, 1, 0, nullptr, 1, 1};
    aug.bs = sMn;
    rc = pack_work_matrix(X, pl.n_pad, pl.ld, rows, pl.ldot, pl.ltot, dot, aug, 0, st, batch);
    if (rc) return rc;
    rc = svd_orthogonalise(pl, X, state, st, &g_last_sweeps);
    if (rc) return rc;
    SvdOutputs o;
    memset(&o, 0, sizeof(o));
    o.vh = tmp; o.vh_stride = (long long)k_cap * ldot; o.vh_ld = ldot;        // rows i: conj(U[:, i])
    o.aug = nnext ? static_cast<cplx*>(M_next_out) : nullptr; o.aug_stride = sMno; o.aug_ld = nnext; o.aug_conj = 0;
    o.s = Sr_out; o.s_stride = sSr; o.rank = rank; o.k_cap = k_cap; o.trun = trun; o.err = err; o.normalise = 1;
    rc = svd_finalise(pl, X, state, o, st);
    if (rc) return rc;
    // M_out[(l,c)][i] = conj(tmp[i][(l,c)])
    return transpose_conj(tmp, ldot, k_cap, ldot, static_cast<cplx*>(M_out), k_cap, 1, st, batch, (long long)k_cap * ldot,
                          sMo);
}

int mpsb_gauge_right_batched(int batch, int chil, int chir, int d, const void* M, long long sM, const double* Sr,
                             long long sSr, const void* M_prev, long long sMp, int nprev, int trun, double err, int k_cap,
                             void* M_out, long long sMo, double* Sl_out, long long sSl, void* M_prev_out, long long sMpo,
                             int* rank, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1 && chil >= 1 && chir >= 1 && d >= 1 && nprev >= 0 && k_cap >= 1, "gauge_right: bad dims");
    const int rows = chil, ldot = d * chir;
    SvdLayout L;
    int rc = svd_layout(batch, rows, ldot, nprev, &L);
    if (rc) return rc;
    const size_t tmp_bytes = al((size_t)batch * k_cap * (nprev > 0 ? nprev : 1) * sizeof(cplx));
    MPSB_REQUIRE(ws && ws_bytes >= L.total + tmp_bytes, "gauge_right: workspace %zu B < required %zu B", ws_bytes,
                 L.total + tmp_bytes);
    Arena ar(ws, ws_bytes);
    cplx* X = static_cast<cplx*>(ar.take(L.X));
    void* state = ar.take(L.state);
    cplx* tmp = static_cast<cplx*>(ar.take(tmp_bytes));
    const SvdPlan& pl = L.plan;
    // X[l][(c,r)] = M[l,c,r] Sr[r];   X[l][ldot + a] = conj(M_prev[a][l])
    MatSrc dot = {static_cast<const cplx*>(M), ldot, 1, 0, Sr, 1, chir};
    dot.bs = sM; dot.scale_bs = sSr;
    MatSrc aug = {static_cast<const cplx*>(M_prev), 1, chil, 1, nullptr, 1, 1};
    aug.bs = sMp;
    rc = pack_work_matrix(X, pl.n_pad, pl.ld, rows, pl.ldot, pl.ltot, dot, aug, 0, st, batch);
    if (rc) return rc;
    rc = svd_orthogonalise(pl, X, state, st, &g_last_sweeps);
    if (rc) return rc;
    SvdOutputs o;
    memset(&o, 0, sizeof(o));
    o.vh = static_cast<cplx*>(M_out); o.vh_stride = sMo; o.vh_ld = ldot;
    o.aug = nprev ? tmp : nullptr; o.aug_stride = (long long)k_cap * nprev; o.aug_ld = nprev; o.aug_conj = 1;   // tmp[i][a] = (M_prev U)[a][i]
    o.s = Sl_out; o.s_stride = sSl; o.rank = rank; o.k_cap = k_cap; o.trun = trun; o.err = err; o.normalise = 1;
    rc = svd_finalise(pl, X, state, o, st);
    if (rc) return rc;
    if (nprev == 0) return 0;
    return transpose_conj(tmp, nprev, k_cap, nprev, static_cast<cplx*>(M_prev_out), k_cap, 0, st, batch,
                          (long long)k_cap * nprev, sMpo);
}

int mpsb_gauge_left(int chil, int chir, int d, const void* M, const double* Sl, const void* M_next, int nnext,
                    int trun, double err, int k_cap, void* M_out, double* Sr_out, void* M_next_out, int* rank,
                    void* ws, size_t ws_bytes, void* stream) {
    return mpsb_gauge_left_batched(1, chil, chir, d, M, 0, Sl, 0, M_next, 0, nnext, trun, err, k_cap, M_out, 0, Sr_out, 0,
                                   M_next_out, 0, rank, ws, ws_bytes, stream);
}

int mpsb_gauge_right(int chil, int chir, int d, const void* M, const double* Sr, const void* M_prev, int nprev,
                     int trun, double err, int k_cap, void* M_out, double* Sl_out, void* M_prev_out, int* rank,
                     void* ws, size_t ws_bytes, void* stream) {
    return mpsb_gauge_right_batched(1, chil, chir, d, M, 0, Sr, 0, M_prev, 0, nprev, trun, err, k_cap, M_out, 0, Sl_out, 0,
                                    M_prev_out, 0, rank, ws, ws_bytes, stream);
}

// ---- dot / corr: transfer-matrix step -----------------------------------------------------------------------
size_t mpsb_transfer_batched_workspace(int batch, int ka, int kb, int d, int oa, int rb) {
    (void)kb; (void)oa;
    return al((size_t)batch * ka * d * rb * sizeof(cplx)) * 2;
}
size_t mpsb_transfer_workspace(int ka, int kb, int d, int oa, int rb) {
    return mpsb_transfer_batched_workspace(1, ka, kb, d, oa, rb);
}

int mpsb_transfer_step_batched(int batch, int ka, int kb, int d, int oa, int rb, const void* E, long long sE, const void* A,
                               long long sA, const void* B, long long sB, const void* Op, long long sOp, void* E_out,
                               long long sEo, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MPSB_REQUIRE(batch >= 1, "transfer_step: bad batch");
    const size_t need = mpsb_transfer_batched_workspace(batch, ka, kb, d, oa, rb);
    MPSB_REQUIRE(ws && ws_bytes >= need, "transfer_step: workspace %zu B < required %zu B", ws_bytes, need);
    Arena ar(ws, ws_bytes);
    const long long sT = (long long)ka * d * rb;
    cplx* T = static_cast<cplx*>(ar.take((size_t)batch * sT * sizeof(cplx)));
    cplx* T2 = static_cast<cplx*>(ar.take((size_t)batch * sT * sizeof(cplx)));
    // T[k][(j,r)] = sum_l E[k][l] B[l][(j,r)]
    GemmArgs g;
    g.A = static_cast<const cplx*>(E); g.B = static_cast<const cplx*>(B); g.C = T;
    g.M = ka; g.N = d * rb; g.K = kb; g.lda = kb; g.ldb = d * rb; g.ldc = d * rb; g.opA = 0; g.opB = 0;
    g.batch1 = batch; g.batch2 = 1; g.sA1 = sE; g.sB1 = sB; g.sC1 = sT; g.sA2 = g.sB2 = g.sC2 = 0; g.accumulate = 0;
    int rc = zgemm(g, st);
    if (rc) return rc;
    const cplx* Tuse = T;
    if (Op) {   // T2[k][i][r] = sum_j Op[i][j] T[k][j][r]
        if (d <= 16) {
            rc = one_site(batch, ka, rb, d, T, sT, static_cast<const cplx*>(Op), sOp, T2, sT, st);
        } else {   // fused multi-site operators (measure on n sites: d^n x d^n): one small GEMM per left index
            GemmArgs h;
            h.A = static_cast<const cplx*>(Op); h.B = T; h.C = T2;
            h.M = d; h.N = rb; h.K = d; h.lda = d; h.ldb = rb; h.ldc = rb; h.opA = 0; h.opB = 0;
            h.batch1 = batch; h.batch2 = ka; h.sA1 = sOp; h.sB1 = sT; h.sC1 = sT;
            h.sA2 = 0; h.sB2 = (long long)d * rb; h.sC2 = (long long)d * rb; h.accumulate = 0;
            rc = zgemm(h, st);
        }
        if (rc) return rc;
        Tuse = T2;
    }
    // E'[o][r] = sum_{(k,i)} conj(A[(k,i)][o]) T[(k,i)][r]
    g.A = static_cast<const cplx*>(A); g.B = Tuse; g.C = static_cast<cplx*>(E_out);
    g.M = oa; g.N = rb; g.K = ka * d; g.lda = oa; g.ldb = rb; g.ldc = rb; g.opA = 2; g.opB = 0;
    g.sA1 = sA; g.sB1 = sT; g.sC1 = sEo;
    return zgemm(g, st);
}

int mpsb_transfer_step(int ka, int kb, int d, int oa, int rb, const void* E, const void* A, const void* B,
                       const void* Op, void* E_out, void* ws, size_t ws_bytes, void* stream) {
    return mpsb_transfer_step_batched(1, ka, kb, d, oa, rb, E, 0, A, 0, B, 0, Op, 0, E_out, 0, ws, ws_bytes, stream);
}

}  // extern "C"
