"""GPU: each C-ABI entry point against numpy / the oracle on seeded inputs (run with -m gpu)."""
import os

import numpy as np
import pytest
import scipy.linalg as la

from oracle.mps_oracle import OracleMPS, svd_truncate as oracle_svd_truncate

pytestmark = pytest.mark.gpu


def crand(rs, *shape):
    return rs.randn(*shape) + 1j * rs.randn(*shape)


OPS = {0: lambda a: a, 1: lambda a: a.T, 2: lambda a: a.conj().T, 3: lambda a: a.conj()}


@pytest.mark.parametrize("opA", [0, 1, 2, 3])
@pytest.mark.parametrize("opB", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(1, 1, 1), (5, 7, 3), (64, 64, 16), (70, 130, 33), (256, 256, 128)])
def test_zgemm_ops(lib, opA, opB, shape):
    M, N, K = shape
    rs = np.random.RandomState(M * 1000 + N * 10 + K + opA * 4 + opB)
    A = crand(rs, *((M, K) if opA in (0, 3) else (K, M)))
    B = crand(rs, *((K, N) if opB in (0, 3) else (N, K)))
    C = lib.matmul(lib.to_device(A), lib.to_device(B), opA, opB)
    ref = OPS[opA](A) @ OPS[opB](B)
    np.testing.assert_allclose(lib.to_host(C), ref, rtol=0, atol=1e-13 * K * 4)


def test_zgemm_batched_strided_accumulate(lib):
    rs = np.random.RandomState(3)
    b1, b2, M, N, K = 3, 2, 20, 24, 9
    A = crand(rs, b1, b2, M, K)
    B = crand(rs, b2, K, N)                 # shared over b1 (stride 0)
    C0 = crand(rs, b1, b2, M, N)
    Cd = lib.to_device(C0)
    lib.zgemm(lib.to_device(A), lib.to_device(B), Cd, M, N, K, K, N, N, batch1=b1, batch2=b2, sA=(b2 * M * K, M * K),
              sB=(0, K * N), sC=(b2 * M * N, M * N), accumulate=True)
    ref = C0 + np.einsum('xymk,ykn->xymn', A, B)
    np.testing.assert_allclose(lib.to_host(Cd), ref, atol=1e-12)


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (3, 5), (6, 6), (17, 9), (32, 32), (40, 64), (64, 40), (128, 128),
                                   (256, 256)])
def test_svd_trunc_matches_lapack(lib, shape):
    from dmrg_b200.MPS import svd_truncate
    rows, cols = shape
    rs = np.random.RandomState(rows * 31 + cols)
    m = crand(rs, rows, cols)
    u, s, vh = svd_truncate(m, trun=10 ** 6, err=1e-13)
    uo, so, vo = oracle_svd_truncate(m, trun=10 ** 6, err=1e-13)
    assert s.shape == so.shape
    np.testing.assert_allclose(s, so, rtol=1e-11, atol=1e-15)
    k = len(s)
    s_raw = la.svd(m, compute_uv=False)[:k]
    np.testing.assert_allclose((u * s_raw) @ vh, m, atol=1e-12 * s_raw[0])          # full rank kept -> exact reconstruction
    np.testing.assert_allclose(u.conj().T @ u, np.eye(k), atol=1e-12)
    np.testing.assert_allclose(vh @ vh.conj().T, np.eye(k), atol=1e-12)


@pytest.mark.parametrize("kind", ["decay", "lowrank", "zero", "rowgraded", "degenerate"])
def test_svd_trunc_hard_spectra(lib, kind):
    from dmrg_b200.MPS import svd_truncate
    n = 96
    rs = np.random.RandomState(17)
    qu, _ = np.linalg.qr(crand(rs, n, n))
    qv, _ = np.linalg.qr(crand(rs, n, n))
    if kind == "decay":            # exponentially decaying Schmidt spectrum over 17 decades
        m = (qu * np.exp(-np.arange(n) * 40.0 / n)) @ qv
    elif kind == "lowrank":
        m = crand(rs, n, n // 3) @ crand(rs, n // 3, n)
    elif kind == "zero":
        m = np.zeros((n, n), dtype=complex)
    elif kind == "rowgraded":
        m = np.exp(-np.arange(n) * 30.0 / n)[:, None] * crand(rs, n, n)
    else:                          # AKLT-like exactly degenerate pairs
        m = (qu * np.repeat(np.exp(-np.arange(n // 2) * 0.5), 2)) @ qv
    trun, err = 64, 1e-10
    u, s, vh = svd_truncate(m, trun=trun, err=err)
    uo, so, vo = oracle_svd_truncate(m, trun=trun, err=err)
    assert len(s) == len(so)
    if len(s) == 0:
        return
    np.testing.assert_allclose(s, so, rtol=1e-10, atol=1e-14)
    # projector onto the kept right singular space is gauge invariant when the cut is not inside a multiplet
    if kind != "degenerate":
        np.testing.assert_allclose(vh.conj().T @ vh, vo.conj().T @ vo, atol=1e-6)
    np.testing.assert_allclose(vh @ vh.conj().T, np.eye(len(s)), atol=1e-11)
    assert lib.load().mpsb_last_svd_sweeps() <= 16


def test_one_site_gate(lib):
    rs = np.random.RandomState(2)
    for d in (2, 3, 5):
        M = crand(rs, 3, 7, d, 11)
        U = crand(rs, 3, d, d)
        Md = lib.to_device(M)
        Ud = lib.to_device(U)
        lib.check(lib.load().mpsb_one_site(3, 7, 11, d, lib.ptr(Md), 7 * d * 11, lib.ptr(Ud), d * d,
                                           lib.ptr(Md), 7 * d * 11, lib.stream_ptr()), "one_site")
        np.testing.assert_allclose(lib.to_host(Md), np.einsum('blcr,bdc->bldr', M, U), atol=1e-13)


def _tebd_call(lib, Bi, Bj, Sl, U, trun, err, kcap):
    """Batched mpsb_tebd_two_site on host arrays [b, ...]; returns host outputs."""
    L = lib.load()
    b, xl, d, x = Bi.shape
    xr = Bj.shape[3]
    per_chain_gate = (U.ndim == 5)
    bio, bjo = lib.empty((b, xl, d, kcap)), lib.empty((b, kcap, d, xr))
    sro, rank = lib.empty((b, kcap), "f64"), lib.empty((b,), "i32")
    ws, wsn = lib.workspace(L.mpsb_tebd_two_site_workspace(b, xl, x, xr, d))
    # keep every device input alive until the call has been issued (temporaries would alias in the caching allocator)
    Bi_d, Bj_d, Sl_d, U_d = lib.to_device(Bi), lib.to_device(Bj), lib.to_device(Sl, np.float64), lib.to_device(U)
    lib.check(L.mpsb_tebd_two_site(b, xl, x, xr, d, lib.ptr(Bi_d), xl * d * x, lib.ptr(Bj_d), x * d * xr, lib.ptr(Sl_d), xl,
                                   lib.ptr(U_d), d ** 4 if per_chain_gate else 0, trun, err, kcap, lib.ptr(bio),
                                   xl * d * kcap, lib.ptr(bjo), kcap * d * xr, lib.ptr(sro), kcap, lib.ptr(rank), ws, wsn,
                                   lib.stream_ptr()), "tebd_two_site")
    return lib.to_host(bio), lib.to_host(bjo), lib.to_host(sro), lib.to_host(rank)


def _oracle_update(Bi, Bj, Sl, U, trun, err):
    o = OracleMPS([Bi, Bj], trun=trun, err=err)
    o.s[0] = Sl
    o.update_double(U, 0)
    return o.M[0], o.M[1], o.s[1]


@pytest.mark.parametrize("dims", [(1, 1, 1, 2), (2, 4, 2, 2), (8, 16, 8, 2), (5, 7, 6, 3), (32, 32, 32, 2), (16, 16, 16, 3)])
def test_tebd_two_site_batch_vs_oracle(lib, dims):
    xl, x, xr, d = dims
    rs = np.random.RandomState(sum(dims))
    b = 5
    Bi, Bj = crand(rs, b, xl, d, x), crand(rs, b, x, d, xr)
    Sl = np.abs(rs.randn(b, xl)) + 0.1
    U = np.stack([la.expm(-0.1j * (h + h.conj().T)).reshape([d] * 4) for h in crand(rs, b, d * d, d * d)])
    trun, err = 24, 1e-10
    kcap = max(1, min(trun, d * xl, d * xr))
    bio, bjo, sro, rank = _tebd_call(lib, Bi, Bj, Sl, U, trun, err, kcap)
    for c in range(b):
        Mi, Mj, s = _oracle_update(Bi[c], Bj[c], Sl[c], U[c], trun, err)
        k = len(s)
        assert rank[c] == k
        np.testing.assert_allclose(sro[c, :k], s, rtol=1e-10, atol=1e-15)
        assert np.all(sro[c, k:] == 0) and np.all(bio[c, :, :, k:] == 0) and np.all(bjo[c, k:] == 0)
        th_g = np.tensordot(bio[c, :, :, :k], bjo[c, :k], axes=([2], [0]))
        th_o = np.tensordot(Mi, Mj, axes=([2], [0]))
        np.testing.assert_allclose(th_g, th_o, atol=1e-9 * max(1.0, np.abs(th_o).max()))


def test_tebd_result_independent_of_zero_padding(lib):
    """Same bond embedded in a zero-padded chi = 24 buffer gives the same Schmidt values and two-site tensor."""
    rs = np.random.RandomState(9)
    xl, x, xr, d, P = 6, 9, 7, 2, 24
    Bi, Bj, Sl = crand(rs, 1, xl, d, x), crand(rs, 1, x, d, xr), np.abs(rs.randn(1, xl)) + 0.1
    h = crand(rs, 4, 4)
    U = la.expm(-0.2j * (h + h.conj().T)).reshape(2, 2, 2, 2)
    trun, err = 24, 1e-12
    a = _tebd_call(lib, Bi, Bj, Sl, U, trun, err, min(trun, d * xl, d * xr))
    Bip, Bjp, Slp = np.zeros((1, P, d, P), complex), np.zeros((1, P, d, P), complex), np.zeros((1, P))
    Bip[0, :xl, :, :x], Bjp[0, :x, :, :xr], Slp[0, :xl] = Bi[0], Bj[0], Sl[0]
    p = _tebd_call(lib, Bip, Bjp, Slp, U, trun, err, P)
    k = int(a[3][0])
    assert int(p[3][0]) == k
    np.testing.assert_allclose(p[2][0, :k], a[2][0, :k], rtol=1e-11)
    th_a = np.tensordot(a[0][0, :, :, :k], a[1][0, :k], axes=([2], [0]))
    th_p = np.tensordot(p[0][0, :xl, :, :k], p[1][0, :k, :, :xr], axes=([2], [0]))
    np.testing.assert_allclose(th_p, th_a, atol=1e-11)
    assert np.all(p[0][0, xl:] == 0) and np.all(p[1][0, :, :, xr:] == 0)


def test_chi128_bench_unit_golden(lib, golden):
    """The benchmark's unit of work at full size: chi = 128 bond of the seed-1000 chain (golden from the reference)."""
    g = golden("chi128_bond7_seed1000.npz")
    L, d = 16, 2
    chis = [min(2 ** i, 2 ** (L - i), 128) for i in range(L + 1)]
    rs = np.random.RandomState(1000)
    Ms = [crand(rs, chis[i], d, chis[i + 1]) for i in range(L)]
    A = crand(rs, 4, 4)
    U = la.expm(-1j * 0.05 * (A + A.conj().T)).reshape([d] * 4)
    o = OracleMPS(Ms, trun=128, err=1e-14)
    o.canon()
    bio, bjo, sro, rank = _tebd_call(lib, o.M[7][None], o.M[8][None], o.s[7][None], U, 128, 1e-14, 128)
    assert int(rank[0]) == int(g["k"])
    np.testing.assert_allclose(sro[0], g["Sr"], rtol=1e-10, atol=1e-16)
    th = np.tensordot(bio[0], bjo[0], axes=([2], [0]))
    assert abs(la.norm(th) - float(g["theta_fro"])) < 1e-10
    np.testing.assert_allclose(th[::17, :, :, ::19], g["theta_probe"], atol=1e-11)
    p = sro[0] ** 2
    assert abs(-np.sum(p * np.log(p)) + np.sum(g["Sr"] ** 2 * np.log(g["Sr"] ** 2))) < 1e-10     # entanglement entropy


def test_transfer_step(lib):
    from dmrg_b200.MPS import MPS
    rs = np.random.RandomState(4)
    ka, kb, d, oa, rb = 5, 6, 3, 7, 4
    E, A, B, Op = crand(rs, ka, kb), crand(rs, ka, d, oa), crand(rs, kb, d, rb), crand(rs, d, d)
    out = MPS._transfer(lib.to_device(E), lib.to_device(A), lib.to_device(B), Op)
    np.testing.assert_allclose(lib.to_host(out), np.einsum('kl,kio,ij,ljr->or', E, A.conj(), Op, B), atol=1e-12)
    out = MPS._transfer(lib.to_device(E), lib.to_device(A), lib.to_device(B), None)
    np.testing.assert_allclose(lib.to_host(out), np.einsum('kl,kio,lir->or', E, A.conj(), B), atol=1e-12)


def test_svd_fuzz_shapes(lib):
    """40 random shapes (1..90 rows/cols, tall, wide, tiny) through every Jacobi kernel variant and padding path."""
    from dmrg_b200.MPS import svd_truncate
    rs = np.random.RandomState(2024)
    for trial in range(40):
        rows, cols = int(rs.randint(1, 91)), int(rs.randint(1, 91))
        m = crand(rs, rows, cols)
        if trial % 5 == 0:                          # rank deficient
            r = max(1, min(rows, cols) // 2)
            m = crand(rs, rows, r) @ crand(rs, r, cols)
        trun = int(rs.choice([3, 17, 1000]))
        err = float(rs.choice([1e-13, 1e-6]))
        u, s, vh = svd_truncate(m, trun=trun, err=err)
        uo, so, vo = oracle_svd_truncate(m, trun=trun, err=err)
        assert len(s) == len(so), (rows, cols, trun, err, len(s), len(so))
        if len(s) == 0:
            continue
        np.testing.assert_allclose(s, so, rtol=1e-10, atol=1e-14, err_msg=str((rows, cols, trun, err)))
        k = len(s)
        s_raw = la.svd(m, compute_uv=False)[:k]
        np.testing.assert_allclose(u.conj().T @ u, np.eye(k), atol=1e-11)
        np.testing.assert_allclose(vh @ vh.conj().T, np.eye(k), atol=1e-11)
        # u s vh is the best rank-k approximation: same as LAPACK's up to the gauge of degenerate values
        np.testing.assert_allclose((u * s_raw) @ vh, (uo * s_raw) @ vo, atol=1e-9 * s_raw[0])


def test_tebd_fuzz_shapes(lib):
    """Random (chi_l, chi, chi_r, d) two-site updates, batch 3, against the oracle."""
    rs = np.random.RandomState(77)
    for trial in range(16):
        d = int(rs.choice([2, 2, 3, 4]))
        xl, x, xr = (int(v) for v in rs.randint(1, 24, size=3))
        b = 3
        Bi, Bj = crand(rs, b, xl, d, x), crand(rs, b, x, d, xr)
        Sl = np.abs(rs.randn(b, xl)) + 0.05
        h = crand(rs, d * d, d * d)
        U = la.expm(-0.15j * (h + h.conj().T)).reshape([d] * 4)
        trun, err = int(rs.choice([4, 12, 64])), 1e-11
        kcap = max(1, min(trun, d * xl, d * xr))
        bio, bjo, sro, rank = _tebd_call(lib, Bi, Bj, Sl, U, trun, err, kcap)
        for c in range(b):
            Mi, Mj, s = _oracle_update(Bi[c], Bj[c], Sl[c], U, trun, err)
            k = len(s)
            assert rank[c] == k, (xl, x, xr, d, trun)
            np.testing.assert_allclose(sro[c, :k], s, rtol=1e-9, atol=1e-14)
            th_g = np.tensordot(bio[c, :, :, :k], bjo[c, :k], axes=([2], [0]))
            th_o = np.tensordot(Mi, Mj, axes=([2], [0]))
            np.testing.assert_allclose(th_g, th_o, atol=1e-8 * max(1.0, np.abs(th_o).max()))


def test_large_local_dimension_d10(lib):
    """Boson-chain sized local space (transmit/bhopping.py uses d = 10): two-site update and one-site gate at d = 10."""
    rs = np.random.RandomState(10)
    d, xl, x, xr, b = 10, 4, 6, 5, 2
    Bi, Bj = crand(rs, b, xl, d, x), crand(rs, b, x, d, xr)
    Sl = np.abs(rs.randn(b, xl)) + 0.1
    h = crand(rs, d * d, d * d)
    U = la.expm(-0.05j * (h + h.conj().T)).reshape([d] * 4)
    trun, err = d, 1e-10                                  # BMPS uses trun = dim (transmit/bhopping.py:37)
    kcap = max(1, min(trun, d * xl, d * xr))
    bio, bjo, sro, rank = _tebd_call(lib, Bi, Bj, Sl, U, trun, err, kcap)
    for c in range(b):
        Mi, Mj, s = _oracle_update(Bi[c], Bj[c], Sl[c], U, trun, err)
        k = len(s)
        assert rank[c] == k
        np.testing.assert_allclose(sro[c, :k], s, rtol=1e-10)
        th_g = np.tensordot(bio[c, :, :, :k], bjo[c, :k], axes=([2], [0]))
        th_o = np.tensordot(Mi, Mj, axes=([2], [0]))
        np.testing.assert_allclose(th_g, th_o, atol=1e-9 * np.abs(th_o).max())


@pytest.mark.parametrize("shape", [(200, 300), (330, 170), (520, 520), (24, 1500), (700, 40)])
def test_svd_large_single_problem(lib, shape):
    """One big problem: cooperative multi-CTA QR (min(rows, cols) >= 160) and, once the carried identity makes the rows
    longer than shared memory holds (520 + 520 columns, 1500 columns), the column-tiled Gram Jacobi kernel."""
    from dmrg_b200.MPS import svd_truncate
    rows, cols = shape
    rs = np.random.RandomState(rows + 7 * cols)
    k0 = min(rows, cols)
    qu, _ = np.linalg.qr(crand(rs, rows, k0))
    qv, _ = np.linalg.qr(crand(rs, cols, k0))
    sv = np.exp(-np.arange(k0) * 20.0 / k0)                     # Schmidt-like decay over 9 decades
    sv[k0 // 2] = sv[k0 // 2 + 1]                               # one exactly degenerate pair
    m = (qu * sv) @ qv.conj().T
    u, s, vh = svd_truncate(m, trun=10 ** 6, err=1e-7)
    uo, so, vo = oracle_svd_truncate(m, trun=10 ** 6, err=1e-7)
    assert len(s) == len(so)
    k = len(s)
    np.testing.assert_allclose(s, so, rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(u.conj().T @ u, np.eye(k), atol=1e-11)
    np.testing.assert_allclose(vh @ vh.conj().T, np.eye(k), atol=1e-11)
    s_raw = la.svd(m, compute_uv=False)[:k]
    np.testing.assert_allclose((u * s_raw) @ vh, (uo * s_raw) @ vo, atol=1e-9)
    assert lib.load().mpsb_last_svd_sweeps() <= 20


def test_svd_grid_qr_small_batch_and_zero_columns(lib):
    """Batch of 3 problems sharing the GPU in the cooperative QR; one has zero columns (reflector skipped), one is
    all zeros."""
    rs = np.random.RandomState(99)
    n, L, nb = 180, 220, 3
    a = crand(rs, nb, n, L)
    a[1][:, ::7] = 0.0
    a[2][:] = 0.0
    Lb = lib.load()
    ad = lib.to_device(a)
    k = min(n, L)
    vh, s, rank = lib.empty((nb, k, L)), lib.empty((nb, k), "f64"), lib.empty((nb,), "i32")
    sraw = lib.empty((nb, k), "f64")
    ws, wsn = lib.workspace(Lb.mpsb_svd_trunc_workspace(nb, n, L, 0))
    lib.check(Lb.mpsb_svd_trunc(nb, n, L, lib.ptr(ad), L, n * L, k, 1e-12, 1, k, lib.ptr(vh), None, lib.ptr(s),
                                lib.ptr(sraw), lib.ptr(rank), ws, wsn, lib.stream_ptr()), "mpsb_svd_trunc")
    ranks = lib.to_host(rank)
    assert ranks[2] == 0
    for b in range(2):
        ref = la.svd(a[b], compute_uv=False)
        kk = int(np.sum(ref / ref[0] > 1e-12))
        assert ranks[b] == kk
        np.testing.assert_allclose(lib.to_host(sraw)[b][:kk], ref[:kk], rtol=1e-10, atol=1e-13 * ref[0])
        v = lib.to_host(vh)[b][:kk]
        np.testing.assert_allclose(v @ v.conj().T, np.eye(kk), atol=1e-11)
        np.testing.assert_allclose(la.norm(a[b] @ v.conj().T, axis=0), ref[:kk], rtol=1e-9, atol=1e-12 * ref[0])


def test_svd_resident_small_batches(lib):
    """Tiny matrices in small batches run all sweeps inside one launch (matrix resident in shared memory)."""
    rs = np.random.RandomState(5)
    Lb = lib.load()
    for n, L, nb in [(2, 2, 1), (7, 5, 4), (16, 16, 30), (33, 40, 9), (48, 64, 5), (20, 250, 3)]:
        a = crand(rs, nb, n, L) * np.exp(-0.2 * np.arange(L))[None, None, :]
        ad = lib.to_device(a)
        k = min(n, L)
        vh, s, rank = lib.empty((nb, k, L)), lib.empty((nb, k), "f64"), lib.empty((nb,), "i32")
        sraw = lib.empty((nb, k), "f64")
        ws, wsn = lib.workspace(Lb.mpsb_svd_trunc_workspace(nb, n, L, 0))
        n0 = Lb.mpsb_launch_count()
        lib.check(Lb.mpsb_svd_trunc(nb, n, L, lib.ptr(ad), L, n * L, k, 0.0, 1, k, lib.ptr(vh), None, lib.ptr(s),
                                    lib.ptr(sraw), lib.ptr(rank), ws, wsn, lib.stream_ptr()), "mpsb_svd_trunc")
        assert Lb.mpsb_launch_count() - n0 <= 6 + nb, "expected the single-launch resident path (+ one pack per problem)"
        for b in range(nb):
            ref = la.svd(a[b], compute_uv=False)
            np.testing.assert_allclose(lib.to_host(sraw)[b], ref, rtol=1e-10, atol=1e-14 * ref[0])
            v = lib.to_host(vh)[b]
            np.testing.assert_allclose(v @ v.conj().T, np.eye(k), atol=1e-11)


def test_svd_rows_at_underflow_scale(lib):
    """Rows 60 to 250 decades below the leading ones (products of noise-level Schmidt values in an iDMRG run on a
    rank-deficient state) must neither spin the relative convergence test for ever nor come back as unnormalised
    junk: live rows stay orthonormal, rows below 1e-62 of the scale come back as zero vectors."""
    from dmrg_b200.MPS import _device_svd
    rs = np.random.RandomState(11)
    n = 96
    expo = np.concatenate([[0, 0, 3, 3, 20, 40, 55], [70, 100, 130, 160, 200, 250], np.full(n - 13 - 20, 300.0)])
    scale = np.concatenate([10.0 ** (-expo), np.zeros(20)])
    m = scale[:, None] * crand(rs, n, n)
    md = lib.to_device(m)
    for want_u in (False, True):
        u, s_all, vh, k = _device_svd(md, n, n, 64, 0.0, want_u=want_u, normalise=0)
        assert lib.load().mpsb_last_svd_sweeps() < 25
        s = lib.to_host(s_all)
        v = lib.to_host(vh)
        live = s[:k] > 1e-60
        assert live.sum() == 7
        g = v[live] @ v[live].conj().T
        np.testing.assert_allclose(g, np.eye(7), atol=1e-11)
        ref = la.svd(m, compute_uv=False)
        np.testing.assert_allclose(s[:4], ref[:4], rtol=1e-10)
        dead = s[:k] < 1e-63
        assert dead.sum() >= 50 and np.all(v[dead] == 0)
        if want_u:
            uu = lib.to_host(u)
            np.testing.assert_allclose(uu[:, live].conj().T @ uu[:, live], np.eye(7), atol=1e-11)
            np.testing.assert_allclose((uu[:, live] * s[:k][live]) @ v[live], m, atol=1e-12)


@pytest.mark.skipif(bool(os.environ.get("MPSB_JACOBI_QUAD_MIN_CTAS")), reason="already inside the child run")
def test_super_block_ordering_forced_on_small_batches(lib):
    """jacobi_quad_kernel (modulus ordering on 16-row super blocks, four cross tasks per CTA) only serves calls with
    >= 1776 CTAs per step.  The library reads its knobs once per process, so a child process with
    MPSB_JACOBI_QUAD_MIN_CTAS=1 re-runs the SVD and two-site tests of this file with that path taken wherever the
    modulus ordering applies (complex and real, full and ragged row lengths, stamps, de Rijk swaps)."""
    import subprocess
    import sys
    env = dict(os.environ, MPSB_JACOBI_QUAD="2", MPSB_JACOBI_QUAD_MIN_CTAS="1")        # 2: every shape, not only 256 complex columns
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "svd or tebd or chi128",
                        os.path.join(root, "tests", "test_gpu_kernels.py"), os.path.join(root, "tests", "test_gpu_real.py")],
                       env=env, cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
