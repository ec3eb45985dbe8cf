"""CPU: host-side logic (task lines, model builders, MPO algebra) and the C-ABI library surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from dmrg_b200 import ST, Ising, spin
from dmrg_b200.MPO import MPO, two_site_mpo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

log = ''


def A(k):
    def _A():
        global log
        log += "A({})".format(k)
    return _A


def B(k):
    def _B():
        global log
        log += "B({})".format(k)
    return _B


def test_ST2():                                         # DMRG/ST_test.py:20-26
    global log
    log = ''
    ST.ST2((A, B), 5)
    assert log == 'A(0.1)B(0.2)A(0.2)B(0.2)A(0.2)B(0.2)A(0.2)B(0.2)A(0.2)B(0.2)A(0.1)'


def test_ST1():                                         # DMRG/ST_test.py:28-33
    global log
    log = ''
    ST.ST1((A, B), 5)
    assert log == 'A(0.2)B(0.2)A(0.2)B(0.2)A(0.2)B(0.2)A(0.2)B(0.2)A(0.2)B(0.2)'


def test_streamline_merges_and_keeps_order():
    assert ST.streamline([(1, 0.5), (1, 0.75), (2, 0.4), (1, 0.1)]) == [(1, 1.25), (2, 0.4), (1, 0.1)]
    assert ST.streamline([]) == []


def test_ST4_is_fourth_order():
    """exp(-i(A+B)) from the 4th-order line: error falls ~16x when n doubles, and beats ST2 at equal n."""
    import scipy.linalg as la
    rs = np.random.RandomState(1)
    def herm():
        m = rs.randn(4, 4) + 1j * rs.randn(4, 4)
        return m + m.conj().T
    Ha, Hb = herm(), herm()
    exact = la.expm(-1j * (Ha + Hb))

    def run(tasks):
        U = np.eye(4, dtype=complex)
        for kind, span in ST.streamline(tasks):
            U = la.expm(-1j * (Ha, Hb)[kind] * span) @ U
        return la.norm(U - exact)
    e4 = [run(ST.ST4_tasks(2, n)) for n in (4, 8)]
    e2 = run(ST.ST2_tasks(2, 8))
    assert e4[0] / e4[1] > 12 and e4[1] < e2 / 10
    assert abs(sum(s for _, s in ST.ST4_tasks(2, 3)) - 2.0) < 1e-12       # each operator runs for total time 1


def test_pauli():                                       # DMRG/test_spin.py:6-17
    np.testing.assert_allclose(spin.sigma[0], np.eye(2))
    s = spin.sigma[1:]
    for i in range(3):
        for j in range(3):
            np.testing.assert_allclose(s[i] @ s[j] + s[j] @ s[i], 2 * ((i == j) * np.eye(2)))
    for i in range(3):
        np.testing.assert_allclose(2j * s[i - 2], s[i - 1] @ s[i] - s[i] @ s[i - 1])


@pytest.mark.parametrize("s", [0.5, 1, 1.5, 2])
def test_spin_algebra(s):
    X, Y, Z = spin.X(s), spin.Y(s), spin.Z(s)
    np.testing.assert_allclose(X @ Y - Y @ X, 1j * Z, atol=1e-14)
    np.testing.assert_allclose(X @ X + Y @ Y + Z @ Z, s * (s + 1) * np.eye(spin.dof(s)), atol=1e-14)
    np.testing.assert_allclose(spin.R(s) @ Z @ spin.R(s), -Z, atol=1e-14)


def test_tfim_mpo_matches_reference_layout():
    W = (-MPO('sx') - MPO('sz') ** 2).tomatrix()
    X, Z = spin.sigma[1], spin.sigma[3]
    assert W.shape == (3, 3, 2, 2) and W.dtype == complex
    np.testing.assert_allclose(W[0, 0], np.eye(2)); np.testing.assert_allclose(W[2, 2], np.eye(2))
    np.testing.assert_allclose(W[1, 0], -Z); np.testing.assert_allclose(W[2, 1], Z); np.testing.assert_allclose(W[2, 0], -X)
    assert np.count_nonzero(np.abs(W).sum(axis=(2, 3))) == 5          # lower triangular, 5 non-zero blocks


def _mpo_to_dense(W, n):
    w, d = W.shape[0], W.shape[2]
    T = W[w - 1]                                          # row selected by the left boundary
    for _ in range(n - 1):
        T = np.einsum('jAB,jkab->kAaBb', T, W).reshape(w, T.shape[1] * d, T.shape[2] * d)
    return T[0]


def test_mpo_dense_equals_ising_builder():
    n = 5
    W = (-0.7 * MPO('sx') - MPO('sz') ** 2).tomatrix()
    np.testing.assert_allclose(_mpo_to_dense(W, n), Ising.Hamilton_trans(n, g=0.7, J=1)['H'], atol=1e-13)
    Wh = (MPO('sx') ** 2 + MPO('sy') ** 2 + MPO('sz') ** 2).tomatrix()
    H = -Ising.Hamilton_XZ(n, delta=1, g=0, h=0)['H']
    np.testing.assert_allclose(_mpo_to_dense(Wh, n), H, atol=1e-13)


def test_two_site_mpo_aklt():
    ss = sum(np.kron(a, a) for a in spin.S(1)[1:])
    H2 = (ss + ss @ ss / 3).real
    W = two_site_mpo(H2, 3)
    n = 3
    dense = sum(np.kron(np.kron(np.eye(3 ** i), H2), np.eye(3 ** (n - 2 - i))) for i in range(n - 1))
    np.testing.assert_allclose(_mpo_to_dense(W, n), dense, atol=1e-13)


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "mpsb.h")).read()
    declared = set(re.findall(r"\b(mpsb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(so, name), f"{name} declared in include/mpsb.h but not exported"
    assert set(lib.SIGNATURES) == declared, "ctypes table and header disagree"
    assert lib.load().mpsb_version() >= 100


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dmrg_b200.MPS import MPS
    with pytest.raises(lib.MpsbError):
        MPS((1, 2, 1))


# ---- transmit/cat helpers (transmit/cat_test.py ported) ---------------------------------------------------
def test_cat_exp_an():                                   # transmit/cat_test.py:6-18
    import scipy.linalg as la
    from dmrg_b200.transmit import cat
    for n in [1, 2, 4, 8]:
        for r in np.linspace(0, 1, 7):
            for theta in np.linspace(0, 2 * np.pi, 13):
                alpha = r * np.exp(1j * theta)
                np.testing.assert_allclose(cat.exp_an(alpha, n), la.expm(alpha * cat.a_n(n)), atol=1e-13)
                np.testing.assert_allclose(cat.exp_ap(alpha, n), la.expm(alpha * cat.a_p(n)), atol=1e-13)


def test_cat_coherent_wigner_displace():                 # transmit/cat_test.py:20-47
    from dmrg_b200.transmit import cat
    s1 = cat.coherent(1, 20)
    s30 = cat.coherent(1, 30)
    for r in np.linspace(0, 2, 11):
        for theta in np.linspace(0, 2 * np.pi, 7):
            z = r * np.exp(1j * theta)
            assert abs(cat.husimi(z, s1) - np.exp(-abs(z - 1) ** 2)) < 1e-6, "Q function test fail"
            assert abs(cat.wigner(z, s30) - np.exp(-2 * abs(z - 1) ** 2)) < 1e-6, "Wigner function test fail"
            s2 = cat.displace(z - 1, 30) @ s30
            assert abs(cat.husimi(z, s2) - 1) < 1e-6, "Displacement function test fail"
    rho = np.outer(s1, s1.conj())
    W = cat.wigner_matrix(0.3 - 0.2j, 20)
    assert abs(cat.expect(W, rho) - cat.expect(W, s1)) < 1e-12
    assert cat.fock(3, 7)[3] == 1 and cat.fock(3, 7).sum() == 1
    np.testing.assert_allclose(abs(cat.kerr(0.3, 2, 10)), abs(cat.coherent(2, 10)))


def test_bond_grouping_for_batched_half_sweeps():
    """Host bookkeeping of MPS._update_bonds: which bonds share one batched mpsb_tebd_two_site call."""
    from dmrg_b200.MPS import MPS
    # saturated L = 128, chi = 128 chain: the edge bonds (1-2-4, 4-8-16, ...) ride along with the bulk -> one call
    L = 128
    x = np.array([min(2 ** i, 2 ** (L - i), 128) for i in range(L + 1)])
    for parity in (0, 1):
        bonds = list(range(parity, L - 1, 2))
        g = MPS._group_bonds(x, bonds, 2)
        assert list(g) == [(128, 128, 128)] and sorted(g[(128, 128, 128)]) == bonds
    # d = 10 boson chain (transmit/bhopping.py): all bonds small -> one call padded to the largest bond triple
    x = np.array([1, 5, 10, 10, 10, 10, 8, 5, 4, 3, 1, 1])
    g = MPS._group_bonds(x, [0, 2, 4, 6, 8], 10)
    assert list(g) == [(10, 10, 10)] and g[(10, 10, 10)] == [0, 2, 4, 6, 8]
    # every bond fits its class; a large class with few members is not merged away
    x = np.array([4] * 30 + [100, 128, 128, 128, 128])
    bonds = list(range(0, 32, 2))
    g = MPS._group_bonds(x, bonds, 2)
    for (PL, PX, PR), idx in g.items():
        for i in idx:
            assert x[i] <= PL and x[i + 1] <= PX and x[i + 2] <= PR
    assert sorted(i for idx in g.values() for i in idx) == bonds
    assert any(k[1] == 4 for k in g) and any(k[1] == 128 for k in g)
    # a single class stays as it is
    assert MPS._group_bonds(np.array([8, 8, 8, 8, 8]), [0, 2], 2) == {(8, 8, 8): [0, 2]}


@pytest.mark.parametrize("nblk,quad", [(4, 0), (6, 0), (8, 0), (32, 0), (8, 1), (12, 1), (16, 1), (32, 1), (64, 1)])
def test_jacobi_sweep_schedule_is_a_complete_ordering(lib, nblk, quad):
    """One sweep of the block ordering, listed on the host by the helpers the kernels and launchers use
    (mpsb_debug_sweep_schedule; quad = 1: the 16-row super-block form of jacobi_quad_kernel): every pair of 8-row blocks
    meets exactly once, every block meets itself exactly once, and the launches of one step - main stream and side
    stream run concurrently - never touch a block twice except in the side stream's own serial order."""
    import ctypes
    cap = 4 * nblk * nblk
    buf = (ctypes.c_int * (4 * cap))()
    n = lib.load().mpsb_debug_sweep_schedule(nblk, quad, buf, cap)
    assert n > 0
    tasks = [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]
    pairs = [(i, j) for _, _, i, j in tasks]
    assert all(0 <= i <= j < nblk for i, j in pairs)
    assert sorted(pairs) == sorted((i, j) for i in range(nblk) for j in range(i, nblk))       # each exactly once
    steps = nblk // 2 if quad else nblk
    assert {t[0] for t in tasks} == set(range(steps))
    for k in range(steps):
        main = [(i, j) for s, st, i, j in tasks if s == k and st == 0]
        side = [(i, j) for s, st, i, j in tasks if s == k and st == 1]
        side_blocks = {b for p in side for b in p}
        if quad:
            # a CTA of the main launch owns the two super blocks of its four tasks; CTAs are disjoint
            owners = {}
            for i, j in main:
                cta = (i // 2, j // 2)
                for b in (i, j):
                    assert owners.setdefault(b, cta) == cta
        else:
            blocks = [b for p in main for b in p]
            assert len(blocks) == len(set(blocks))
        assert not side_blocks & {b for p in main for b in p}
        assert bool(side) == (k % 2 == 0)
