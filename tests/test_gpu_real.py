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


def test_complex_svd_of_real_matrix_stays_real(lib):
    """The split of the real path runs the complex Jacobi kernels on data with a zero imaginary part; every product
    with an exact zero is zero, so the factors come back exactly real."""
    from dmrg_b200 import iDMRG
    rs = np.random.RandomState(8)
    for sh in ((6, 2, 2, 5), (32, 2, 2, 32), (100, 2, 2, 100), (160, 2, 2, 160)):
        psi = rs.randn(*sh)
        split = iDMRG.halve if sh[0] * sh[1] <= 256 else iDMRG._halve_split
        U, S, V = split(iDMRG._as_complex(lib.to_device(psi, np.float64)), sh, 24)
        assert float(U.imag.abs().max()) == 0.0 and float(V.imag.abs().max()) == 0.0
        s_ref = np.linalg.svd(psi.reshape(sh[0] * sh[1], -1), compute_uv=False)
        np.testing.assert_allclose(lib.to_host(S)[:24], s_ref[:24], rtol=1e-11)


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
