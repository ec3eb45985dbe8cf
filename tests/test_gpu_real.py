"""GPU: the real-arithmetic path (reference: MPO.tomatrix(dtype='double'), DMRG/MPO.py:68-76 - real L, W, R make the
whole growth step of DMRG/iDMRG.py:49-61 real).  DGEMM kernel against numpy; real matvec / Lanczos / environments /
prediction against the complex path and the oracle; energies of the two paths agree to 1e-12."""
import numpy as np
import pytest

from oracle import idmrg_oracle as io_

pytestmark = pytest.mark.gpu

OPS = {0: lambda a: a, 1: lambda a: a.T, 2: lambda a: a.T, 3: lambda a: a}


def _dmatmul(lib, A, B, opA, opB):
    M, K = A.shape if opA in (0, 3) else A.shape[::-1]
    N = B.shape[1] if opB in (0, 3) else B.shape[0]
    C = lib.empty((M, N), "f64")
    lib.dgemm(lib.to_device(A, np.float64), lib.to_device(B, np.float64), C, M, N, K, A.shape[1], B.shape[1], N, opA, opB)
    return lib.to_host(C)


@pytest.mark.parametrize("opA", [0, 1, 2, 3])
@pytest.mark.parametrize("opB", [0, 1])
@pytest.mark.parametrize("shape", [(1, 1, 1), (5, 7, 3), (6, 8, 4), (64, 64, 16), (70, 130, 34), (72, 130, 33), (256, 256, 128),
                                   (384, 512, 128), (1024, 640, 100)])
def test_dgemm_ops(lib, opA, opB, shape):
    """All op combinations; even shapes ride the TMA kernels (64- and 128-tiles), odd leading dimensions the SIMT
    fallback; K tails and M / N tails are zero-filled by the TMA unit."""
    M, N, K = shape
    rs = np.random.RandomState(M * 1000 + N * 10 + K + opA * 4 + opB)
    A = rs.randn(*((M, K) if opA in (0, 3) else (K, M)))
    B = rs.randn(*((K, N) if opB in (0, 3) else (N, K)))
    ref = OPS[opA](A) @ OPS[opB](B)
    np.testing.assert_allclose(_dmatmul(lib, A, B, opA, opB), ref, rtol=0, atol=1e-13 * K * 4)


def test_dgemm_batched_strided_accumulate(lib):
    rs = np.random.RandomState(3)
    for (b1, b2, M, N, K) in ((3, 2, 20, 24, 10), (40, 4, 128, 136, 48), (2, 3, 9, 7, 5)):
        A = rs.randn(b1, b2, M, K)
        B = rs.randn(b2, K, N)                 # shared over b1 (stride 0)
        C0 = rs.randn(b1, b2, M, N)
        Cd = lib.to_device(C0, np.float64)
        lib.dgemm(lib.to_device(A, np.float64), lib.to_device(B, np.float64), Cd, M, N, K, K, N, N, batch1=b1, batch2=b2,
                  sA=(b2 * M * K, M * K), sB=(0, K * N), sC=(b2 * M * N, M * N), accumulate=True)
        ref = C0 + np.einsum('xymk,ykn->xymn', A, B)
        np.testing.assert_allclose(lib.to_host(Cd), ref, atol=1e-12)


def test_real_heff_matvec_and_env(lib):
    from dmrg_b200 import iDMRG
    rs = np.random.RandomState(5)
    for xl, xr, d, w in ((1, 1, 2, 3), (7, 5, 3, 4), (32, 32, 2, 5), (40, 24, 2, 3), (3, 9, 3, 11)):
        L, R, W, th = rs.randn(xl, xl, w), rs.randn(xr, xr, w), rs.randn(w, w, d, d), rs.randn(xl, d, d, xr)
        W[0, 1:] = 0
        ref = io_.heff_matvec(L, W, R, th)
        out = iDMRG.heff_matvec(L, W, R, th)
        assert out.dtype == np.float64                       # took the real kernels
        np.testing.assert_allclose(out, ref.real, atol=1e-11 * np.abs(ref).max())
        chin = 6
        U, V = rs.randn(xl, d, chin), rs.randn(chin, d, xr)
        Ln = iDMRG._env(True, lib.to_device(L, np.float64), lib.to_device(U, np.float64), lib.to_device(W, np.float64), False)
        Rn = iDMRG._env(False, lib.to_device(R, np.float64), lib.to_device(V, np.float64), lib.to_device(W, np.float64), False)
        assert not Ln.is_complex() and not Rn.is_complex()
        np.testing.assert_allclose(lib.to_host(Ln), io_.env_left(L, U, W, False).real, atol=1e-11 * np.abs(L).max() * xl)
        np.testing.assert_allclose(lib.to_host(Rn), io_.env_right(R, V, W, False).real, atol=1e-11 * np.abs(R).max() * xr)


def _svd_real(lib, A, trunc, want_u, normalise=0, err=0.0):
    """mpsb_svd_trunc_real on a batch [n, rows, cols] of real matrices: (Uh | None, S_raw, S, Vh, rank) as numpy."""
    l = lib.load()
    n, rows, cols = A.shape
    k = max(1, min(rows, cols, trunc))
    Ad = lib.to_device(A, np.float64)
    vh, uh = lib.empty((n, k, cols), "f64"), (lib.empty((n, k, rows), "f64") if want_u else None)
    s, sraw, rank = lib.empty((n, k), "f64"), lib.empty((n, min(rows, cols)), "f64"), lib.empty((n,), "i32")
    ws, wsn = lib.workspace(l.mpsb_svd_trunc_real_workspace(n, rows, cols, int(want_u)))
    lib.check(l.mpsb_svd_trunc_real(n, rows, cols, lib.ptr(Ad), cols, rows * cols, trunc, err, normalise, k, lib.ptr(vh),
                                    lib.ptr(uh), lib.ptr(s), lib.ptr(sraw), lib.ptr(rank), ws, wsn, lib.stream_ptr()),
              "mpsb_svd_trunc_real")
    return (lib.to_host(uh) if want_u else None), lib.to_host(sraw), lib.to_host(s), lib.to_host(vh), lib.to_host(rank)


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 2, 3), (2, 5, 4), (3, 7, 5), (4, 16, 16), (3, 33, 17), (2, 17, 40), (5, 64, 64),
                                   (3, 100, 100), (2, 128, 256), (6, 256, 256), (1, 320, 320), (1, 512, 512),
                                   (1, 1024, 1024)])
@pytest.mark.parametrize("want_u", [False, True])
def test_real_svd_matches_lapack(lib, shape, want_u):
    """Real SVD (QR unpacked, Jacobi sweeps on pair-packed rows) against LAPACK: singular values 1e-11 relative,
    Vh (and U) orthonormal to 1e-12, A V = U S; even and odd shapes, all kernel families (resident, cross / self,
    Gram, tiled Gram)."""
    n, rows, cols = shape
    if want_u and rows >= 1024:
        pytest.skip("carried identity at 1024 rows: covered by the split path")
    rs = np.random.RandomState(rows * 7 + cols + n)
    A = rs.randn(n, rows, cols) * np.exp(-0.03 * np.arange(cols))[None, None, :]      # decaying column scales
    trunc = min(rows, cols, 48)
    uh, sraw, s, vh, rank = _svd_real(lib, A, trunc, want_u)
    for b in range(n):
        s_ref = np.linalg.svd(A[b], compute_uv=False)
        np.testing.assert_allclose(sraw[b], s_ref, rtol=1e-11, atol=1e-13 * s_ref[0])
        k = int(rank[b])
        assert k == trunc
        V = vh[b, :k]
        np.testing.assert_allclose(V @ V.T, np.eye(k), atol=1e-12)
        if want_u:
            U = uh[b, :k].T
            np.testing.assert_allclose(U.T @ U, np.eye(k), atol=1e-12)
            np.testing.assert_allclose(A[b] @ V.T, U * sraw[b, :k], atol=1e-11 * s_ref[0])
        else:                                               # the projector onto the kept right singular space
            _, _, vr = np.linalg.svd(A[b])
            gap_ok = k == min(rows, cols) or s_ref[k - 1] - s_ref[k] > 1e-6 * s_ref[0]
            if gap_ok:
                np.testing.assert_allclose(V.T @ V, vr[:k].T @ vr[:k], atol=1e-8)


def test_real_svd_truncation_rule(lib):
    """normalise = 1: the reference's `truncate` (DMRG/MPS.py:33-38) on the real path."""
    from oracle.mps_oracle import svd_truncate as oracle_svd_truncate
    rs = np.random.RandomState(11)
    A = rs.randn(1, 24, 30) @ np.diag(np.exp(-1.5 * np.arange(30)))
    _, _, s, vh, rank = _svd_real(lib, A, 10, False, normalise=1, err=1e-9)
    uo, so, vo = oracle_svd_truncate(A[0], 10, 1e-9)
    assert int(rank[0]) == len(so)
    np.testing.assert_allclose(s[0, :len(so)], so, rtol=1e-10)


def test_real_halve_returns_real_factors(lib):
    from dmrg_b200 import iDMRG
    rs = np.random.RandomState(8)
    for sh in ((6, 2, 2, 5), (32, 2, 2, 32), (160, 2, 2, 160)):
        psi = rs.randn(*sh)
        U, S, V = iDMRG.halve(psi.ravel(), sh, trunc=12)
        assert U.dtype == np.float64 and V.dtype == np.float64
        Uo, So, Vo = io_.halve(psi, sh, 12)
        np.testing.assert_allclose(S, So, rtol=1e-11)
        np.testing.assert_allclose(np.tensordot(U, U, axes=([2], [2])), np.tensordot(Uo, Uo.conj(), axes=([2], [2])).real,
                                   atol=1e-9)
        th = np.einsum('abk,k,kcd->abcd', U, S[:12], V)
        thk = np.einsum('abk,k,kcd->abcd', Uo, So[:12], Vo)
        np.testing.assert_allclose(th, thk.real, atol=1e-10 * So[0])


@pytest.mark.parametrize("model", ["tfim", "heisenberg"])
def test_real_and_complex_paths_give_the_same_energies(lib, golden, model):
    """energies identical to the complex path to 1e-12 (relative)."""
    from dmrg_b200 import iDMRG
    W = io_.tfim_mpo(g=1.0) if model == "tfim" else golden("idmrg_energies.npz")["heis_W"]
    w = W.shape[0]
    trunc, steps = 20, 9
    runs = {}
    for real in (True, False):
        iDMRG.REAL_PATH = real
        try:
            ch = iDMRG.Chain(W, trunc=trunc)
            assert ch.real == real
            e = [ch.step() for _ in range(steps)]
            assert ch.L.is_complex() != real
            runs[real] = (np.array(e), ch.S.copy(), list(ch.n_matvec))
        finally:
            iDMRG.REAL_PATH = True
    er, ec = runs[True][0], runs[False][0]
    np.testing.assert_allclose(er, ec, rtol=1e-12)
    if model == "tfim":                                     # non-degenerate spectrum: Schmidt values comparable too
        np.testing.assert_allclose(runs[True][1][:10], runs[False][1][:10], rtol=1e-8)


def test_real_next_LR_returns_real_arrays_and_matches_oracle(lib):
    from dmrg_b200 import iDMRG
    W = io_.tfim_mpo(g=0.7).real
    L, R = (a.real for a in io_.boundary(3))
    Lo, Ro = io_.boundary(3)
    for i in range(8):
        L, R, E = iDMRG.next_LR(L, R, W, trunc=16)
        Lo, Ro, Eo = io_.next_LR(Lo, Ro, W, trunc=16)
        assert L.dtype == np.float64 and R.dtype == np.float64
        assert abs(E - Eo) < 1e-9 * abs(Eo)


def test_real_chain_batch_matches_complex(lib):
    from dmrg_b200 import iDMRG
    gs = np.linspace(0.5, 1.5, 4)
    Ws = np.stack([io_.tfim_mpo(g=g) for g in gs])
    out = {}
    for real in (True, False):
        iDMRG.REAL_PATH = real
        try:
            run = iDMRG.ChainBatch(Ws, trunc=12)
            assert run.real == real
            out[real] = np.array([run.step() for _ in range(9)])
            assert run.L.is_complex() != real
        finally:
            iDMRG.REAL_PATH = True
    np.testing.assert_allclose(out[True], out[False], rtol=1e-12)


@pytest.mark.parametrize("real", [True, False])
def test_large_split_left_vectors_from_right_ones(lib, real):
    """`_halve_split` (rows too long for a carried identity): U = A V^H S^-1 + one Newton-Schulz step gives orthonormal
    Schmidt partners of V - also when `trunc` cuts a degenerate multiplet, where U S V must still be a best rank-k
    approximation (||A - U S V||_F^2 = sum of the discarded S^2) - and falls back to a second SVD when the kept
    spectrum reaches below SPLIT_CUTOFF."""
    from dmrg_b200 import iDMRG
    rs = np.random.RandomState(31 + real)
    n, k = 288, 48
    def unitary(m):
        z = rs.randn(m, m) if real else rs.randn(m, m) + 1j * rs.randn(m, m)
        return np.linalg.qr(z)[0]
    for floor, fast in ((1e-3, True), (1e-20, True), (1e-40, False)):    # S_k / S_0 = 0.32, 5.3e-4, 2.8e-7
        sv = np.exp(np.linspace(0, np.log(floor), n))
        sv[k - 2:k + 2] = sv[k]                              # a four-fold multiplet cut in the middle
        A = (unitary(n) * sv) @ unitary(n)
        sh = (n // 2, 2, 2, n // 2)
        Ad = lib.to_device(A, np.float64 if real else np.complex128)
        U, S, V = iDMRG._halve_split(Ad.reshape(sh), sh, k)
        took_fast = bool(sv[k - 1] > iDMRG.SPLIT_CUTOFF * sv[0])
        assert took_fast == fast
        U, S, V = lib.to_host(U).reshape(n, k), lib.to_host(S), lib.to_host(V).reshape(k, n)
        np.testing.assert_allclose(S, np.sort(sv)[::-1], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(U.conj().T @ U, np.eye(k), atol=1e-12)
        np.testing.assert_allclose(V @ V.conj().T, np.eye(k), atol=1e-12)
        if fast:
            err2 = np.linalg.norm(A - (U * S[:k]) @ V) ** 2
            np.testing.assert_allclose(err2, np.sum(np.sort(sv)[::-1][k:] ** 2), rtol=1e-6, atol=1e-26)
