"""CPU: the oracle against the golden vectors generated from the unmodified reference
(oracle/validate_against_reference.py) and against exact results."""
import numpy as np
import scipy.linalg as la

from oracle import idmrg_oracle as io_
from oracle.mps_oracle import OracleMPS, svd_truncate, truncate


def test_truncate_rule():
    s = np.array([2.0, 1.0, 1e-3, 1e-9])
    kept, k = truncate(s, trun=20, err=1e-6)
    assert k == 3 and abs(la.norm(kept) - 1) < 1e-15
    kept, k = truncate(s, trun=2, err=1e-6)
    assert k == 2
    kept, k = truncate(np.zeros(3), trun=5, err=1e-6)     # all-zero matrix: reference yields k = 0 (NaN > err is False)
    assert k == 0


def test_svd_truncate_reconstructs():
    rs = np.random.RandomState(0)
    m = rs.randn(12, 9) + 1j * rs.randn(12, 9)
    u, s, vh = svd_truncate(m, trun=20, err=1e-12)
    s_raw = la.svd(m, compute_uv=False)
    np.testing.assert_allclose(s, s_raw / la.norm(s_raw), rtol=1e-13)
    np.testing.assert_allclose((u * s_raw) @ vh, m, atol=1e-13)


def test_canon_testcanon_golden(golden):
    g = golden("canon_testcanon.npz")
    o = OracleMPS([g["M0"], g["M1"]], trun=20, err=1e-6)
    o.canon()
    for b in range(3):
        np.testing.assert_allclose(o.s[b], g[f"s{b}"], atol=1e-14)
    assert abs(o.dot(o) - g["norm"]) < 1e-14
    for i in range(o.L):                                   # both orthonormality relations, DMRG/MPS.py:495-508
        S = o.block_single(i)
        assert la.norm(np.tensordot(S.conj(), S, axes=([0, 1], [0, 1])) - np.diag(o.s[i + 1]) ** 2) < 1e-13
        assert la.norm(np.tensordot(S, S.conj(), axes=([1, 2], [1, 2])) - np.diag(o.s[i]) ** 2) < 1e-13


def test_update_double_golden(golden):
    g = golden("update_double_L6.npz")
    L = 6
    for i in range(L - 1):
        o = OracleMPS([g[f"M_in_{k}"] for k in range(L)], trun=int(g["trun"]), err=float(g["err"]))
        o.s = [g[f"s_in_{b}"] for b in range(L + 1)]
        o.update_double(g["U"], i)
        assert o.x[i + 1] == int(g[f"k_out_{i}"])
        np.testing.assert_allclose(o.s[i + 1], g[f"Sr_out_{i}"], rtol=1e-10, atol=1e-15)
        th = np.tensordot(o.M[i], o.M[i + 1], axes=([2], [0]))
        np.testing.assert_allclose(th, g[f"theta_out_{i}"], atol=1e-12)


def test_two_body_phases(golden):
    g = golden("two_body.npz")
    Z = np.diag([1.0, -1.0])
    o = OracleMPS.product([[1 + 3j, 1], [1 - 8j, 1 - 2j]], trun=20, err=1e-6)
    o.canon()
    U = la.expm(np.pi / 20 * 1j * np.kron(Z, Z)).reshape([2] * 4)
    for step in range(20):
        o.update_double(U, 0)
        amps = np.array([OracleMPS.product(e).dot(o) for e in ([0, 0], [0, 1], [1, 0], [1, 1])])
        np.testing.assert_allclose(amps, g["amps"][step], atol=1e-13)
    assert abs(o.corr((0, Z), (1, Z)) - g["zz"]) < 1e-13


def test_aklt_golden(golden):
    g = golden("aklt_N11.npz")
    N = 11
    o = OracleMPS([g[f"raw_{i}"] for i in range(N)], trun=20, err=1e-6)
    o.canon()
    for b in range(N + 1):
        np.testing.assert_allclose(o.s[b], g[f"s_{b}"], atol=1e-13)
    E = sum(o.measure(i, g["H"]) for i in range(N - 1)).real
    assert abs(E / (N - 1) + 2 / 3) < 1e-13 and abs(E - g["energy"]) < 1e-12
    sz = np.diag([1.0, 0.0, -1.0])
    np.testing.assert_allclose([o.corr((i, sz)) for i in range(N)], g["sz"], atol=1e-12)
    p = o.copy()
    for k in range(3):
        p.tebd(g["H"], 1, 10)
        ov = o.dot(p)
        assert abs(ov - g["overlaps"][k]) < 1e-10
        assert abs(ov - np.exp(-1j * E * (k + 1))) < 1e-6          # DMRG/MPS_test.py:62-63


def test_tfim_quench_golden(golden):
    g = golden("tfim_quench_L10.npz")
    L = 10
    Z = np.diag([1.0, -1.0])
    o = OracleMPS.product([[1, 0]] * L, trun=32, err=1e-12)
    o.canon()
    o.tebd(g["Hb"], float(g["T"]), int(g["n"]))
    o.canon()
    z = np.array([o.corr((i, Z)) for i in range(L)])
    np.testing.assert_allclose(z, g["z"], atol=1e-10)
    assert np.max(np.abs(z - g["z_ed"])) < 2e-4                     # Trotter error at n = 40
    for b in range(L + 1):
        np.testing.assert_allclose(o.s[b], g[f"s_{b}"], atol=1e-10)


def test_heff_and_env_golden(golden):
    g = golden("heff_small.npz")
    np.testing.assert_allclose(io_.heff_matvec(g["L"], g["W"], g["R"], g["theta"]), g["out"], atol=1e-12)
    np.testing.assert_allclose(io_.env_left(g["L"], g["U"], g["W"], swap_conj=True), g["Lnext_ref"], atol=1e-12)
    np.testing.assert_allclose(io_.env_right(g["R"], g["V"], g["W"], swap_conj=True), g["Rnext_ref"], atol=1e-12)


def test_idmrg_tfim_energies_golden(golden):
    g = golden("idmrg_energies.npz")
    W = g["W"]
    for trunc, steps in ((10, 20), (16, 12), (32, 9)):       # trunc = 32: BASELINE config 1
        L, R = io_.boundary(3)
        e = []
        for i in range(steps):
            L, R, val = io_.next_LR(L, R, W, trunc=trunc)
            e.append(val / (2 * i + 2))
        np.testing.assert_allclose(e, g[f"tfim_e_trunc{trunc}"][:steps], rtol=2e-10)
    assert abs(g["tfim_e_trunc10"][0] + 1.1180339887499) < 1e-12
    assert abs(g["tfim_e_trunc10"][39] + 1.2687174555821) < 1e-9
    assert abs(g["tfim_e_trunc32"][8] + 1.253444929157) < 1e-11         # SURVEY App. C, trunc = 32 series


def _bmps_cases(golden):
    import json
    g = golden("bmps.npz")
    return g, json.loads(str(g["cases"]))


def test_bmps_oracle_golden(golden):
    """the oracle's bosonic hopping chain against the reference's `transmit/bhopping.py` run."""
    from oracle.bmps_oracle import OracleBMPS
    g, cases = _bmps_cases(golden)
    for name, c in cases.items():
        o = OracleBMPS(c["dim"], c["l"], c["para"], c["pack"])
        np.testing.assert_allclose(o.occupations(), g[name + "_occ0"], atol=1e-13)
        o.evolve(c["t"], c["n"], c["tail"])
        assert list(o.x) == list(g[name + "_x_raw"])
        o.canon()
        assert list(o.x) == list(g[name + "_x"])
        np.testing.assert_allclose(o.occupations(), g[name + "_occ"], atol=1e-12)
        for b in range(o.L + 1):
            np.testing.assert_allclose(o.s[b], g[name + "_s%d" % b], atol=1e-12)
        nn = np.kron(o.N, o.N)
        np.testing.assert_allclose([o.measure(i, nn) for i in range(o.l - 1)], g[name + "_nn"], atol=1e-12)
        np.testing.assert_allclose(o.c_n, g[name + "_c_n"], atol=1e-13)


def test_bmps_free_hopping_keeps_coherent_product():
    """Known answer independent of the reference: under pure hopping a product of coherent states stays one, with
    amplitudes c(t) = exp(-i g A t) c(0) (A = open-chain adjacency) -> <N_i> = |c_i(t)|^2 up to Fock-cutoff error."""
    from oracle.bmps_oracle import OracleBMPS
    para = {"omega0": 0, "g": 1.0, "u": 0, "h": 0, "K": 0}
    pack = {"dk": 0.6, "center": 3, "k_c": np.pi / 2, "trun": False, "alpha": 0.6}
    o = OracleBMPS(9, 6, para, pack)
    c0 = o.c_n.copy()
    o.evolve(0.4, 20)
    o.canon()
    A = np.diag(np.ones(5), 1) + np.diag(np.ones(5), -1)
    ct = la.expm(-1j * 0.4 * A) @ c0
    np.testing.assert_allclose(o.occupations()[:6], abs(ct) ** 2, atol=2e-5)     # Trotter error O(dt^2) + cutoff


def test_oracle_update_k_against_dense_state():
    """k-site gates: k = 2 is update_double; k = 3, 4 reproduce the exactly evolved state vector."""
    rs = np.random.RandomState(4)
    L, d = 6, 2
    chis = [1, 2, 4, 8, 4, 2, 1]
    tens = [rs.randn(chis[j], d, chis[j + 1]) + 1j * rs.randn(chis[j], d, chis[j + 1]) for j in range(L)]
    base = OracleMPS(tens, trun=64, err=1e-14)
    base.canon()
    for k, i in [(2, 1), (3, 0), (3, 2), (4, 1)]:
        D = d ** k
        A = rs.randn(D, D) + 1j * rs.randn(D, D)
        U = la.expm(-0.3j * (A + A.conj().T))
        a = base.copy()
        psi0 = a.to_dense()
        a.update_k(U.reshape([d] * (2 * k)), i, k)
        full = np.kron(np.kron(np.eye(d ** i), U), np.eye(d ** (L - i - k)))
        np.testing.assert_allclose(a.to_dense(), full @ psi0, atol=1e-13)
        for b in range(L + 1):
            assert abs(la.norm(a.s[b]) - 1) < 1e-13
        if k == 2:
            b2 = base.copy()
            b2.update_double(U.reshape([d] * 4), i)
            for b in range(L + 1):
                np.testing.assert_allclose(a.s[b], b2.s[b], atol=1e-14)
