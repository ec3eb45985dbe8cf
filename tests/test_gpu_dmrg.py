"""GPU: finite-system two-site DMRG against exact diagonalisation through the reference's own
dense builders (`DMRG/Ising.py:29-83`, mirrored in dmrg_b200/Ising.py)."""
import numpy as np
import pytest
import scipy.linalg as la

pytestmark = pytest.mark.gpu


def test_finite_dmrg_tfim_vs_ed(lib):
    from dmrg_b200 import DMRG, Ising
    from dmrg_b200.MPO import MPO
    from dmrg_b200.spin import sigma
    n = 10
    W = (-MPO('sx') - MPO('sz') ** 2).tomatrix()                    # H = -sum X - sum Z Z, the reference's driver MPO
    H = Ising.Hamilton_trans(n, g=1, J=1)['H']
    evals, evecs = la.eigh(H)
    ch = DMRG.FiniteChain(W, n, trunc=32)
    E = ch.run(sweeps=4)
    assert abs(E - evals[0]) < 1e-9 * abs(evals[0])
    assert ch.energies[-1] <= ch.energies[0] + 1e-12                # variational: sweeps never raise the energy
    assert max(ch.discarded[-(n - 1):]) < 1e-20                     # chi = 32 is exact for 10 spins
    # observables of the converged state through the MPS class
    mps = ch.to_mps()
    psi = evecs[:, 0]
    for i in (0, 4, 8):
        zz = np.kron(np.kron(np.eye(2 ** i), np.kron(sigma[3], sigma[3])), np.eye(2 ** (n - i - 2)))
        assert abs(mps.measure(i, np.kron(sigma[3], sigma[3])) - (psi.conj() @ zz @ psi).real) < 1e-8
    # mid-chain entanglement entropy against the exact Schmidt decomposition
    sv = la.svd(psi.reshape(2 ** 5, 2 ** 5), compute_uv=False)
    p = sv ** 2
    p = p[p > 1e-30]
    assert abs(ch.entropy(5) - float(-np.sum(p * np.log(p)))) < 1e-8
    # late sweeps restart from the current tensors: a handful of H_eff products per bond
    assert np.mean(ch.n_matvec[-(n - 1):]) < 0.5 * np.mean(ch.n_matvec[:n - 1]), (ch.n_matvec[:n - 1], ch.n_matvec[-(n - 1):])


def test_finite_dmrg_truncated_heisenberg(lib):
    """Real-form Heisenberg MPO (SURVEY 8c caveat 2), L = 12, chi = 16 < exact: energy within the discarded weight."""
    from dmrg_b200 import DMRG, spin
    from dmrg_b200.Ising import nearest
    from dmrg_b200.MPO import MPO
    from dmrg_b200.spin import sigma
    n = 12
    W = (2 * MPO('P', op=spin.P()) * MPO('N', op=spin.N()) + 2 * MPO('N', op=spin.N()) * MPO('P', op=spin.P())
         + MPO('sz') ** 2).tomatrix()
    H = sum(nearest(n, sigma[a], sigma[a], sparse=True) for a in (1, 2, 3))
    import scipy.sparse.linalg as sla
    e0 = sla.eigsh(H.real.astype(float) if abs(H.imag).max() == 0 else H, k=1, which='SA')[0][0]
    ch = DMRG.FiniteChain(W, n, trunc=16)
    E = ch.run(sweeps=5)
    worst = max(ch.discarded[-(n - 1):])
    assert E >= e0 - 1e-9                                           # variational
    assert E - e0 < 50 * worst * abs(e0) + 1e-7, (E, e0, worst)
    assert worst < 1e-5
