"""GPU: the reference's own MPS tests (DMRG/MPS_test.py) ported verbatim onto dmrg_b200.MPS, plus the golden
vectors generated from the reference and oracle comparisons at larger bond dimension."""
import numpy as np
import pytest
import scipy.linalg as la

from oracle.mps_oracle import OracleMPS

pytestmark = pytest.mark.gpu


def test_canon_reference_case(lib, golden):                    # DMRG/MPS_test.py:10-28
    from dmrg_b200.MPS import MPS
    s = MPS((1, 2, 2))
    s.M[0][0, :, :] = np.array([[1, 2 + 1j], [9j, 6j]])
    s.M[1][:, :, 0] = np.array([[- 1j, 2 - 3j], [0.3, 4]])
    s.M[1][:, :, 1] = np.array([[5 - 1j, 2 - 3j], [0.3, 4]])
    s.canon()
    assert abs(s.dot(s) - 1) < 1e-12
    for i in s.s:
        assert abs(la.norm(i) - 1) < 1e-12
    for i in range(s.L):
        S = s.block_single(i)
        TSS = np.einsum("ijk, ijn->kn", S.conj(), S)
        SST = np.einsum("ijk, ljk->il", S, S.conj())
        assert la.norm(TSS - np.diag(s.Sr[i]) ** 2) < 1e-12
        assert la.norm(SST - np.diag(s.Sl[i]) ** 2) < 1e-12
    s.testCanon(err=1e-12)
    g = golden("canon_testcanon.npz")
    for b in range(3):
        np.testing.assert_allclose(np.asarray(s.s[b], dtype=float), g[f"s{b}"], rtol=1e-10, atol=1e-14)


def test_two_body_reference_case(lib, golden):                 # DMRG/MPS_test.py:30-50
    from dmrg_b200.MPS import MPS
    from dmrg_b200.spin import sigma
    s = MPS.naive([1 + 3j, 1], [1 - 8j, 1 - 2j])
    eig_states = [[0, 0], [0, 1], [1, 0], [1, 1]]
    eig_vals = np.array([1, -1, -1, 1])
    s.canon()
    n = 20
    H = np.kron(sigma[3], sigma[3])
    U = la.expm(np.pi / n * 1j * H).reshape([2, 2, 2, 2])
    phi = np.exp(np.pi / n * 1j * eig_vals)
    L0 = np.array([s[eig_states[j]] for j in range(4)])
    H0 = s.corr((0, sigma[3]), (1, sigma[3]))
    g = golden("two_body.npz")
    for i in range(n):
        s.update_double(U, 0)
        assert abs(H0 - s.corr((0, sigma[3]), (1, sigma[3]))) < 1e-10
        assert abs(H0 - s.measure(0, H)) < 1e-10
        L = np.array([s[eig_states[j]] for j in range(4)])
        L0 *= phi
        assert la.norm(L - L0) < 1e-10
        np.testing.assert_allclose(L, g["amps"][i], atol=1e-11)
    assert abs(H0 - g["zz"]) < 1e-11


def test_aklt_reference_case(lib, golden, capsys):             # DMRG/MPS_test.py:52-64
    from dmrg_b200 import AKLT as ak
    N = 11
    s = ak.AKLT_State(N)
    E = ak.Energy(s)
    assert abs(E / (N - 1) + 2 / 3) < 1e-12
    s.canon()
    n = 5
    l1 = ak.evolve(s, n=n, time=1, k=10)
    l2 = np.exp(-1j * E * (np.arange(n) + 1))
    np.testing.assert_allclose(l1, l2, rtol=1e-7)
    ak.correlator(s)
    g = golden("aklt_N11.npz")
    assert abs(E - g["energy"]) < 1e-11
    for b in range(N + 1):
        np.testing.assert_allclose(np.asarray(s.s[b], dtype=float), g[f"s_{b}"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(ak.Z(s), g["sz"], atol=1e-11)
    np.testing.assert_allclose(l1[:3], g["overlaps"], atol=1e-9)
    np.testing.assert_allclose([s.corr(*ak.string_op(2, j)) for j in range(3, 9)], g["string"], atol=1e-11)


def test_update_double_golden_every_bond(lib, golden):
    from dmrg_b200.MPS import MPS
    g = golden("update_double_L6.npz")
    L = 6
    for i in range(L - 1):
        s = MPS([g[f"M_in_{k}"] for k in range(L)], 2, trun=int(g["trun"]), err=float(g["err"]))
        for b in range(L + 1):
            s.s[b] = g[f"s_in_{b}"]
        s.update_double(g["U"], i)
        assert int(s.xr[i]) == int(g[f"k_out_{i}"])
        np.testing.assert_allclose(s.Sr[i], g[f"Sr_out_{i}"], rtol=1e-10, atol=1e-15)
        th = np.tensordot(s.M[i], s.M[i + 1], axes=([2], [0]))
        np.testing.assert_allclose(th, g[f"theta_out_{i}"], atol=1e-11)
        s.verify_shape()


def test_tfim_quench_golden_and_ed(lib, golden):
    """iTEBD_double (batched half sweeps) on an L = 10 quench: <Z_i> and all Schmidt spectra vs the reference."""
    from dmrg_b200.MPS import MPS
    from dmrg_b200.spin import sigma
    g = golden("tfim_quench_L10.npz")
    L = 10
    s = MPS.naive([[1, 0]] * L)
    s.trun, s.err = 32, 1e-12
    s.canon()
    s.iTEBD_double(g["Hb"], float(g["T"]), int(g["n"]))
    s.canon()
    z = np.array([s.corr((i, sigma[3])) for i in range(L)])
    np.testing.assert_allclose(z, g["z"], rtol=1e-9, atol=1e-10)
    assert np.max(np.abs(z - g["z_ed"])) < 2e-4
    for b in range(L + 1):
        a, r = np.asarray(s.s[b], dtype=float), g[f"s_{b}"]
        assert len(a) == len(r)
        np.testing.assert_allclose(a, r, rtol=1e-8, atol=1e-12)
    ent = lambda v: -np.sum(v ** 2 * np.log(v ** 2))
    assert abs(ent(np.asarray(s.s[5], dtype=float)) - ent(g["s_5"])) < 1e-10


def test_canon_random_chain_vs_oracle(lib):
    """canon() at bond dimension up to 32 (d = 2) and a d = 3 chain: Schmidt spectra, overlaps, observables."""
    from dmrg_b200.MPS import MPS
    for d, chis, seed in ((2, [1, 2, 4, 8, 16, 32, 32, 16, 8, 4, 2, 1], 21), (3, [1, 3, 9, 12, 9, 3, 1], 22)):
        rs = np.random.RandomState(seed)
        Ms = [rs.randn(chis[i], d, chis[i + 1]) + 1j * rs.randn(chis[i], d, chis[i + 1]) for i in range(len(chis) - 1)]
        o = OracleMPS([m.copy() for m in Ms], trun=32, err=1e-12)
        o.canon()
        s = MPS([m.copy() for m in Ms], d, trun=32, err=1e-12)
        s.canon()
        s.testCanon(err=1e-11)
        for b in range(len(chis)):
            np.testing.assert_allclose(np.asarray(s.s[b], dtype=float), o.s[b], rtol=1e-10, atol=1e-14)
        op = rs.randn(d, d) + 1j * rs.randn(d, d)
        op = op + op.conj().T
        site = len(Ms) // 2
        assert abs(s.corr((site, op)) - o.corr((site, op))) < 1e-11
        assert abs(s.corr((1, op), (site, op)) - o.corr((1, op), (site, op))) < 1e-11
        op2 = np.kron(op, op.T)
        assert abs(s.measure(site, op2) - o.measure(site, op2)) < 1e-11
        # overlap with a second state
        Ms2 = [m + 0.1 * (rs.randn(*m.shape) + 1j * rs.randn(*m.shape)) for m in Ms]
        o2 = OracleMPS(Ms2, trun=32, err=1e-12); o2.canon()
        s2 = MPS(Ms2, d, trun=32, err=1e-12); s2.canon()
        assert abs(abs(s.dot(s2)) - abs(o.dot(o2))) < 1e-11


def test_update_single_and_copy(lib):
    from dmrg_b200.MPS import MPS
    from dmrg_b200 import spin
    rs = np.random.RandomState(8)
    Ms = [rs.randn(1, 2, 2) + 0j, rs.randn(2, 2, 2) + 0j, rs.randn(2, 2, 1) + 0j]
    s = MPS(Ms, 2)
    s.canon()
    p = s.copy()
    U = la.expm(0.3j * spin.sigma[1])
    p.update_single(U, 1)
    o = OracleMPS(Ms, trun=20, err=1e-6); o.canon()
    q = o.copy(); q.update_single(U, 1)
    assert abs(abs(s.dot(p)) - abs(o.dot(q))) < 1e-12
    assert abs(s.dot(s) - 1) < 1e-12                       # the copy did not alias the original


def test_chain_batch_matches_per_chain_mps(lib):
    """ChainBatch (padded, in place, one gate per chain) == independent MPS objects, after a short TFIM sweep."""
    from dmrg_b200.MPS import MPS
    from dmrg_b200.batched import ChainBatch
    from dmrg_b200.spin import sigma
    L, n, cap = 8, 3, 16
    Z, X, I2 = sigma[3], sigma[1], np.eye(2)
    gs = [0.5, 1.0, 1.5]
    Hs = np.stack([-np.kron(Z, Z) - 0.5 * g * (np.kron(X, I2) + np.kron(I2, X)) for g in gs])
    cb = ChainBatch.product([[1, 0]] * L, n, cap, trun=cap, err=1e-12)
    cb.tebd(Hs, 0.6, 6, order=2)
    ent = cb.entropies()
    for c, g in enumerate(gs):
        s = MPS.naive([[1, 0]] * L)
        s.trun, s.err = cap, 1e-12
        s.canon()
        s.iTEBD_double(Hs[c], 0.6, 6)
        Ms, ss = cb.chain(c)
        for b in range(L + 1):
            a = np.asarray(s.s[b], dtype=float)
            assert len(ss[b]) == len(a)
            np.testing.assert_allclose(ss[b], a, rtol=1e-9, atol=1e-13)
        assert abs(ent[c, L // 2] + np.sum(np.asarray(s.s[L // 2]) ** 2 * np.log(np.asarray(s.s[L // 2]) ** 2))) < 1e-10


def test_tebd_chi64_vs_oracle(lib):
    """A random canonical chain at bond dimension 64: one full second-order Trotter step (all bonds, batched half
    sweeps) against the oracle; Schmidt spectra, entropies and <sigma> compared after canon()."""
    from dmrg_b200.MPS import MPS
    from dmrg_b200.spin import sigma
    rs = np.random.RandomState(33)
    chis = [1, 2, 4, 8, 16, 32, 64, 64, 64, 32, 16, 8, 4, 2, 1]
    Ms = [rs.randn(chis[i], 2, chis[i + 1]) + 1j * rs.randn(chis[i], 2, chis[i + 1]) for i in range(len(chis) - 1)]
    Z, X, I2 = sigma[3], sigma[1], np.eye(2)
    Hb = -np.kron(Z, Z) - 0.5 * 1.3 * (np.kron(X, I2) + np.kron(I2, X))
    o = OracleMPS([m.copy() for m in Ms], trun=64, err=1e-12)
    o.canon()
    s = MPS([m.copy() for m in Ms], 2, trun=64, err=1e-12)
    s.canon()
    o.tebd(Hb, 0.1, 2)
    s.iTEBD_double(Hb, 0.1, 2)
    o.canon()
    s.canon()
    for b in range(len(chis)):
        a, r = np.asarray(s.s[b], dtype=float), o.s[b]
        assert len(a) == len(r)
        np.testing.assert_allclose(a, r, rtol=1e-9, atol=1e-13)
        assert abs(np.sum(a ** 2 * np.log(a ** 2)) - np.sum(r ** 2 * np.log(r ** 2))) < 1e-10
    for i in (0, 5, 7, 13):
        for op in (Z, X):
            assert abs(s.corr((i, op)) - o.corr((i, op))) < 1e-10


def test_fourth_order_trotter_beats_second_order(lib):
    """ST4 task line through MPS.iTEBD_double(order=4) and ChainBatch.tebd(order=4): L = 6 TFIM quench vs exact
    evolution; the 4th-order error is far below the 2nd-order one at the same step count."""
    from dmrg_b200.MPS import MPS
    from dmrg_b200.batched import ChainBatch
    from dmrg_b200.spin import sigma
    from dmrg_b200 import Ising
    L, T, n = 6, 0.8, 4
    Z, X, I2 = sigma[3], sigma[1], np.eye(2)
    Hb = -np.kron(Z, Z) - 0.5 * (np.kron(X, I2) + np.kron(I2, X))
    g = np.ones(L); g[0] = g[-1] = 0.5
    Hfull = Ising.Hamilton_trans(L, g=g, J=1)['H']
    psi0 = np.zeros(2 ** L); psi0[0] = 1
    psi = la.expm(-1j * T * Hfull) @ psi0
    z_ed = np.array([(psi.conj() @ np.kron(np.kron(np.eye(2 ** i), Z), np.eye(2 ** (L - i - 1))) @ psi).real for i in range(L)])
    errs = {}
    for order in (2, 4):
        s = MPS.naive([[1, 0]] * L)
        s.trun, s.err = 8, 1e-14
        s.canon()
        s.iTEBD_double(Hb, T, n, order=order)
        s.canon()
        errs[order] = np.max(np.abs(np.array([s.corr((i, Z)) for i in range(L)]) - z_ed))
    assert errs[4] < errs[2] / 20 and errs[4] < 2e-4
    cb = ChainBatch.product([[1, 0]] * L, 2, 8, trun=8, err=1e-14)
    cb.tebd(Hb, T, n, order=4)
    Ms, ss = cb.chain(1)
    s4 = MPS([m.copy() for m in Ms], 2, trun=8, err=1e-14)
    for b in range(L + 1):
        s4.s[b] = ss[b]
    s4.canon()
    assert np.max(np.abs(np.array([s4.corr((i, Z)) for i in range(L)]) - z_ed)) < 2e-4


def test_host_bond_updater_matches_direct_call(lib):
    """Pipelined host-buffer path (HostBondUpdater) == one direct batched call on device-resident inputs."""
    import torch
    from dmrg_b200.batched import HostBondUpdater
    rs = np.random.RandomState(12)
    n, xl, x, xr, d = 7, 6, 8, 5, 2
    Bi = rs.randn(n, xl, d, x) + 1j * rs.randn(n, xl, d, x)
    Bj = rs.randn(n, x, d, xr) + 1j * rs.randn(n, x, d, xr)
    Sl = np.abs(rs.randn(n, xl)) + 0.1
    U = np.stack([la.expm(-0.1j * (h + h.conj().T)).reshape([d] * 4) for h in rs.randn(n, 4, 4) + 1j * rs.randn(n, 4, 4)])
    trun, err = 8, 1e-12
    upd = HostBondUpdater(n, xl, x, xr, d, trun, err)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    k = upd.kcap
    Bio, Bjo = torch.empty((n, xl, d, k), dtype=torch.complex128).pin_memory(), torch.empty((n, k, d, xr), dtype=torch.complex128).pin_memory()
    Sro, rk = torch.empty((n, k), dtype=torch.float64).pin_memory(), torch.empty((n,), dtype=torch.int32).pin_memory()
    ins = (pin(Bi), pin(Bj), pin(Sl), pin(U))
    upd(*ins, Bio, Bjo, Sro, rk)
    torch.cuda.synchronize()
    first = (Bio.clone(), Sro.clone())
    upd.stage(*ins)                       # streamed form: two batches in flight
    upd.stage(*ins)
    upd.run(Bio, Bjo, Sro, rk)
    upd.run(Bio, Bjo, Sro, rk)
    upd.finish()
    torch.cuda.synchronize()
    assert torch.equal(first[0], Bio) and torch.equal(first[1], Sro)
    for c in range(n):
        o = OracleMPS([Bi[c], Bj[c]], trun=trun, err=err)
        o.s[0] = Sl[c]
        o.update_double(U[c], 0)
        kk = len(o.s[1])
        assert int(rk[c]) == kk
        np.testing.assert_allclose(Sro[c, :kk].numpy(), o.s[1], rtol=1e-10)
        th_g = np.tensordot(Bio[c, :, :, :kk].numpy(), Bjo[c, :kk].numpy(), axes=([2], [0]))
        th_o = np.tensordot(o.M[0], o.M[1], axes=([2], [0]))
        np.testing.assert_allclose(th_g, th_o, atol=1e-10 * max(1.0, np.abs(th_o).max()))


def test_non_unitary_update_branch(lib):
    """`unitary=False` (imaginary-time gates, DMRG/MPS.py:424-425, 456-457 — present but untested in the reference):
    a left-to-right sweep of exp(-tau H) gates followed by canon(), GPU vs oracle."""
    from dmrg_b200.MPS import MPS
    from dmrg_b200.spin import sigma
    L = 6
    Z, X, I2 = sigma[3], sigma[1], np.eye(2)
    Hb = -np.kron(Z, Z) - 0.5 * 1.1 * (np.kron(X, I2) + np.kron(I2, X))
    G = la.expm(-0.2 * Hb).reshape(2, 2, 2, 2)
    rs = np.random.RandomState(3)
    chis = [1, 2, 4, 4, 4, 2, 1]
    Ms = [rs.randn(chis[i], 2, chis[i + 1]) + 1j * rs.randn(chis[i], 2, chis[i + 1]) for i in range(L)]
    s = MPS([m.copy() for m in Ms], 2, trun=8, err=1e-12)
    o = OracleMPS([m.copy() for m in Ms], trun=8, err=1e-12)
    s.canon(); o.canon()
    for sweep in range(3):
        for i in range(L - 1):
            s.update_double(G, i, unitary=False)
            o.update_double(G, i, unitary=False)
        s.canon(); o.canon()
    for b in range(L + 1):
        np.testing.assert_allclose(np.asarray(s.s[b], dtype=float), o.s[b], rtol=1e-9, atol=1e-13)
    e_gpu = sum(s.measure(i, Hb) for i in range(L - 1))
    e_ora = sum(o.measure(i, Hb) for i in range(L - 1))
    assert abs(e_gpu - e_ora) < 1e-10


def test_update_k_multi_site_gates(lib):
    """`update_k` (NotImplementedError in the reference, DMRG/MPS.py:459-461) against the oracle's
    restatement (itself checked against dense state vectors in tests/test_oracle.py) - Schmidt values on every bond,
    overlap with the oracle state, and k = 2 identical to update_double."""
    from dmrg_b200.MPS import MPS
    rs = np.random.RandomState(4)
    L, d = 6, 2
    chis = [1, 2, 4, 8, 4, 2, 1]
    tens = [rs.randn(chis[j], d, chis[j + 1]) + 1j * rs.randn(chis[j], d, chis[j + 1]) for j in range(L)]
    for k, i in [(2, 1), (3, 0), (3, 2), (4, 1), (1, 3)]:
        D = d ** k
        A = rs.randn(D, D) + 1j * rs.randn(D, D)
        U = la.expm(-0.3j * (A + A.conj().T)).reshape([d] * (2 * k))
        g = MPS([t.copy() for t in tens], trun=64, err=1e-14)
        g.canon()
        o = OracleMPS([t.copy() for t in tens], trun=64, err=1e-14)
        o.canon()
        g.update_k(U, i, k)
        if k == 1:
            o.update_single(U, i)
        else:
            o.update_k(U, i, k)
        assert [int(v) for v in g.x] == [int(v) for v in o.x]
        for b in range(L + 1):
            np.testing.assert_allclose(np.asarray(g.s[b], dtype=float), o.s[b], rtol=1e-10, atol=1e-13)
        ref = MPS([np.array(m) for m in o.M[:L]], trun=64, err=1e-14)      # oracle state as a device MPS (B form, s[0] = 1)
        ref.s = np.array([np.asarray(v) for v in o.s] , dtype=object)
        assert abs(abs(g.dot(ref)) - 1) < 1e-11
        if k == 2:
            h = MPS([t.copy() for t in tens], trun=64, err=1e-14)
            h.canon()
            h.update_double(U, i)
            for b in range(L + 1):
                np.testing.assert_allclose(np.asarray(g.s[b], dtype=float), np.asarray(h.s[b], dtype=float), rtol=1e-10, atol=1e-13)
    # d = 3, three sites, truncation active
    rs = np.random.RandomState(8)
    chis = [1, 3, 6, 6, 3, 1]
    tens = [rs.randn(chis[j], 3, chis[j + 1]) + 1j * rs.randn(chis[j], 3, chis[j + 1]) for j in range(5)]
    A = rs.randn(27, 27) + 1j * rs.randn(27, 27)
    U = la.expm(-0.2j * (A + A.conj().T))
    g = MPS([t.copy() for t in tens], dim=3, trun=5, err=1e-8)
    o = OracleMPS([t.copy() for t in tens], trun=5, err=1e-8)
    g.canon(); o.canon()
    g.update_k(U, 1, 3); o.update_k(U, 1, 3)
    assert [int(v) for v in g.x] == [int(v) for v in o.x]
    for b in range(6):
        np.testing.assert_allclose(np.asarray(g.s[b], dtype=float), o.s[b], rtol=1e-10, atol=1e-13)


def test_scale_rows_kernel(lib):
    rs = np.random.RandomState(1)
    a = rs.randn(12, 7) + 1j * rs.randn(12, 7)
    s = np.abs(rs.randn(4)) + 0.1
    out = lib.empty((12, 7))
    ad, sd = lib.to_device(a), lib.to_device(s, np.float64)
    lib.check(lib.load().mpsb_scale_rows(12, 7, lib.ptr(ad), 7, lib.ptr(sd), 3, lib.ptr(out), 7, lib.stream_ptr()), "scale")
    np.testing.assert_allclose(lib.to_host(out), np.repeat(s, 3)[:, None] * a, rtol=1e-15)


def test_chain_batch_canon_and_observables_match_per_chain_mps(lib):
    """ChainBatch.canon / corr / measure (batched gauge and transfer entry points) against one `MPS` per chain:
    random NON-canonical chains of different bond dimensions, padded to a common capacity.  Tolerance 1e-10 on Schmidt
    values, <Z>, <ZZ>, a two-site and a per-chain operator (DMRG/MPS.py:246-302, 358-410)."""
    from dmrg_b200.MPS import MPS
    from dmrg_b200.batched import ChainBatch
    from dmrg_b200.spin import sigma
    n, Lc, cap = 6, 8, 16
    Z, X = sigma[3], sigma[1]
    H2 = -np.kron(Z, Z) - 0.7 * (np.kron(X, np.eye(2)) + np.kron(np.eye(2), X))
    chains, singles = [], []
    for c in range(n):
        rs = np.random.RandomState(900 + c)
        top = 4 + 2 * c
        x = [min(2 ** i, 2 ** (Lc - i), top) for i in range(Lc + 1)]
        Ms = [rs.randn(x[i], 2, x[i + 1]) + 1j * rs.randn(x[i], 2, x[i + 1]) for i in range(Lc)]
        chains.append((Ms, [np.ones(1)] + [np.ones(x[b]) for b in range(1, Lc + 1)]))
        m = MPS([t.copy() for t in Ms], 2, trun=cap, err=1e-12)
        m.canon()
        singles.append(m)
    batch = ChainBatch.from_chains(chains, cap=cap, trun=cap, err=1e-12)
    batch.canon()
    ranks = batch.ranks()
    for c in range(n):
        _, ss = batch.chain(c)
        for b in range(Lc + 1):
            ref = np.asarray(singles[c].s[b], dtype=float)
            assert ranks[c, b] == len(ref)
            np.testing.assert_allclose(ss[b], ref, rtol=1e-10, atol=1e-13)
    z3 = batch.corr((3, Z))
    zz = batch.corr((2, Z), (5, Z))
    e2 = batch.measure(3, H2)
    ops = np.stack([np.cos(0.3 * c) * Z + np.sin(0.3 * c) * X for c in range(n)])
    per = batch.corr((4, ops))
    for c in range(n):
        assert abs(z3[c] - singles[c].corr((3, Z))) < 1e-10
        assert abs(zz[c] - singles[c].corr((2, Z), (5, Z))) < 1e-10
        assert abs(e2[c] - singles[c].measure(3, H2)) < 1e-10
        assert abs(per[c] - singles[c].corr((4, ops[c]))) < 1e-10
    # canon() after a truncating Trotter step restores unit norm: <1> = 1 for every chain
    batch.tebd(H2, 0.1, 1)
    batch.canon()
    np.testing.assert_allclose(batch.corr((0, np.eye(2))), np.ones(n), atol=1e-12)


def test_two_streams_do_not_share_scratch(lib):
    """Calls issued on two CUDA streams draw their scratch from separate workspaces: two batches
    updated concurrently on two streams equal the same batches updated one after the other."""
    import torch
    from dmrg_b200.batched import ChainBatch
    from dmrg_b200.spin import sigma
    Z, X = sigma[3], sigma[1]
    H2 = -np.kron(Z, Z) - 0.5 * (np.kron(X, np.eye(2)) + np.kron(np.eye(2), X))
    U = lib.to_device(la.expm(-0.2j * H2).reshape(2, 2, 2, 2))

    def make(seed):
        return ChainBatch.from_chains([_random_canonical_chain(seed + c, 8, 16) for c in range(8)], cap=16, trun=16, err=1e-12)

    serial = [make(50), make(70)]
    for b in serial:
        for i in range(0, 7, 2):
            b.apply_bond(U, i)
    torch.cuda.synchronize()
    conc = [make(50), make(70)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for i in range(0, 7, 2):
        for b, st in zip(conc, streams):
            with torch.cuda.stream(st):
                b.apply_bond(U, i)
    torch.cuda.synchronize()
    for a, b in zip(serial, conc):
        np.testing.assert_array_equal(lib.to_host(a.S), lib.to_host(b.S))
        np.testing.assert_array_equal(lib.to_host(a.M), lib.to_host(b.M))


def _random_canonical_chain(seed, Lc, chi):
    rs = np.random.RandomState(seed)
    x = [min(2 ** i, 2 ** (Lc - i), chi) for i in range(Lc + 1)]
    o = OracleMPS([rs.randn(x[i], 2, x[i + 1]) + 1j * rs.randn(x[i], 2, x[i + 1]) for i in range(Lc)], trun=chi, err=1e-12)
    o.canon()
    return [m.copy() for m in o.M[:Lc]], [np.array(v) for v in o.s]


def test_update_single_large_local_dimension(lib):
    """`update_single` for d > 16 (the reference accepts any d, DMRG/MPS.py:413-425): the GEMM path of mpsb_one_site."""
    from dmrg_b200.MPS import MPS
    rs = np.random.RandomState(12)
    d = 20
    Ms = [rs.randn(1, d, 3) + 1j * rs.randn(1, d, 3), rs.randn(3, d, 1) + 1j * rs.randn(3, d, 1)]
    A = rs.randn(d, d) + 1j * rs.randn(d, d)
    U = la.expm(0.1j * (A + A.conj().T))
    s = MPS([m.copy() for m in Ms], d, trun=8, err=1e-12)
    s.update_single(U, 0)
    np.testing.assert_allclose(np.asarray(s.M[0]), np.einsum('lcr, dc->ldr', Ms[0], U), atol=1e-12)
