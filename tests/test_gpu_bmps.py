"""GPU: SURVEY 8(f) row 1 — the reference's large-local-dimension caller `transmit/bhopping.py::BMPS` on the CUDA
path, against the goldens generated from the unmodified reference and against the oracle on the same inputs."""
import json

import numpy as np
import pytest
import scipy.linalg as la

from oracle.bmps_oracle import OracleBMPS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 1, 3, 1, 1), (4, 5, 2, 3, 3), (7, 3, 10, 10, 1), (1, 6, 6, 1, 6), (16, 16, 4, 5, 5)])
def test_mpo_site_kernel(lib, shape):
    """mpsb_mpo_site against the reference's einsum (transmit/bhopping.py:97-100), batched with shared and own W."""
    xl, xr, d, wl, wr = shape
    rs = np.random.RandomState(sum(shape))
    nb = 3
    M = rs.randn(nb, xl, d, xr) + 1j * rs.randn(nb, xl, d, xr)
    W = rs.randn(nb, wl, wr, d, d) + 1j * rs.randn(nb, wl, wr, d, d)
    L = lib.load()
    Md, Wd = lib.to_device(M), lib.to_device(W)
    for shared in (False, True):
        out = lib.empty((nb, xl * wl, d, xr * wr))
        lib.check(L.mpsb_mpo_site(nb, xl, xr, d, wl, wr, lib.ptr(Md), xl * d * xr, lib.ptr(Wd), 0 if shared else wl * wr * d * d,
                                  lib.ptr(out), xl * wl * d * xr * wr, lib.stream_ptr()), "mpsb_mpo_site")
        got = lib.to_host(out)
        for b in range(nb):
            ref = np.einsum("ijk, lmjo->ilokm", M[b], W[0 if shared else b], optimize=False)
            ref = ref.reshape(xl * wl, d, xr * wr)
            np.testing.assert_allclose(got[b], ref, rtol=0, atol=1e-13 * max(1.0, abs(ref).max()))


def test_mpo_site_rejects_aliasing(lib):
    L = lib.load()
    t = lib.empty((2, 2, 2))
    w = lib.empty((1, 1, 2, 2))
    assert L.mpsb_mpo_site(1, 2, 2, 2, 1, 1, lib.ptr(t), 0, lib.ptr(w), 0, lib.ptr(t), 0, lib.stream_ptr()) != 0
    assert b"mpo_site" in L.mpsb_last_error()


def _cases(golden):
    g = golden("bmps.npz")
    return g, json.loads(str(g["cases"]))


@pytest.mark.parametrize("name", ["d6", "d10", "d5_notail"])
def test_bmps_matches_reference_golden(lib, golden, name):
    from dmrg_b200.transmit.bhopping import BMPS
    g, cases = _cases(golden)
    c = cases[name]
    psi = BMPS(c["dim"], c["l"], c["para"], c["pack"])
    assert psi.L == c["l"] + 1 and psi.l == c["l"] and psi.trun == c["dim"] and psi.err == 1e-6
    occ0 = np.array([psi.measure(i, psi.N) for i in range(psi.L)])
    np.testing.assert_allclose(occ0, g[name + "_occ0"], rtol=1e-10, atol=1e-13)
    psi.evolve(c["t"], c["n"], c["tail"])
    assert [int(v) for v in psi.x] == [int(v) for v in g[name + "_x_raw"]]
    psi.canon()
    assert [int(v) for v in psi.x] == [int(v) for v in g[name + "_x"]]
    psi.testCanon(err=1e-9)                 # truncated at err = 1e-6: canonical up to the discarded weight
    occ = np.array([psi.measure(i, psi.N) for i in range(psi.L)])
    np.testing.assert_allclose(occ, g[name + "_occ"], rtol=1e-10, atol=1e-12)
    for b in range(psi.L + 1):
        # LAPACK resolves singular values to eps * s_max, the Jacobi path to eps * s_i: compare at eps * s_max level
        np.testing.assert_allclose(np.asarray(psi.s[b], dtype=float), g[name + "_s%d" % b], rtol=1e-10, atol=1e-13)
    nn = np.kron(psi.N, psi.N)
    np.testing.assert_allclose([psi.measure(i, nn) for i in range(psi.l - 1)], g[name + "_nn"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(psi.c_n, g[name + "_c_n"], atol=1e-13)
    assert abs(psi.time - c["t"]) < 1e-15


def test_bmps_matches_oracle_fresh_inputs(lib):
    """Not in the goldens: another local dimension, packet and couplings, evolved in two periods with `evolve_measure`."""
    from dmrg_b200.transmit.bhopping import BMPS
    para = {"omega0": 0.1, "g": 0.9, "u": 1.1, "h": 0.3, "K": 0.4}
    pack = {"dk": 0.7, "center": 2, "k_c": np.pi / 2, "trun": True, "alpha": 1.0}
    psi = BMPS(7, 5, para, pack)
    orc = OracleBMPS(7, 5, para, pack)
    rows = psi.evolve_measure(0.25, n=2, k=3)
    want = [orc.occupations()[:5]]
    for _ in range(2):
        orc.evolve(0.25, 3)
        orc.canon()
        want.append(orc.occupations()[:5])
    np.testing.assert_allclose(rows, np.array(want), rtol=1e-10, atol=1e-12)
    assert [int(v) for v in psi.x] == [int(v) for v in orc.x]
    for b in range(psi.L + 1):
        np.testing.assert_allclose(np.asarray(psi.s[b], dtype=float), orc.s[b], rtol=1e-10, atol=1e-13)
    a, b = psi.copy(), psi.copy()
    assert abs(a.dot(b) - 1) < 1e-9


def test_bmps_wigner_and_overlap(lib, golden):
    """`wigner` (a product of per-site operators through `corr`, transmit/bhopping.py:192-197) on the initial product
    state has a closed form: prod_i <c_i| W(a_i) |c_i>."""
    from dmrg_b200.transmit import cat
    from dmrg_b200.transmit.bhopping import BMPS
    para = {"omega0": 0, "g": 1.0, "u": 0, "h": 0, "K": 0}
    pack = {"dk": 0.5, "center": 2, "k_c": -np.pi / 2, "trun": False, "alpha": 0.8}
    psi = BMPS(8, 5, para, pack)
    amps = 0.3 * psi.c_n / la.norm(psi.c_n)
    want = np.prod([cat.expect(cat.wigner_matrix(a, 8), cat.coherent(c, 8)) for a, c in zip(amps, psi.c_n)])
    assert abs(psi.wigner(0.3) - want) < 1e-10
    g, cases = _cases(golden)
    c = cases["d6"]
    phi = BMPS(c["dim"], c["l"], c["para"], c["pack"])
    phi.evolve(c["t"], c["n"], c["tail"])
    phi.canon()
    assert abs(phi.wigner(0.3) - float(g["d6_wigner"])) < 1e-9     # operator entries themselves are ~1e-10 (cat.displace)


def test_bmps_free_hopping_known_answer(lib):
    """Independent of reference and oracle: free hopping moves coherent amplitudes by exp(-i g A t)."""
    from dmrg_b200.transmit.bhopping import BMPS
    para = {"omega0": 0, "g": 1.0, "u": 0, "h": 0, "K": 0}
    pack = {"dk": 0.6, "center": 3, "k_c": np.pi / 2, "trun": False, "alpha": 0.6}
    psi = BMPS(9, 6, para, pack)
    c0 = psi.c_n.copy()
    psi.evolve(0.4, 20)
    psi.canon()
    A = np.diag(np.ones(5), 1) + np.diag(np.ones(5), -1)
    ct = la.expm(-1j * 0.4 * A) @ c0
    occ = np.array([psi.measure(i, psi.N) for i in range(psi.l)])
    np.testing.assert_allclose(occ, abs(ct) ** 2, atol=2e-5)
