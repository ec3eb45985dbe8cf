"""CPU oracle for the iDMRG growth step — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Matrix-free numpy restatement of `DMRG/iDMRG.py:38-61` of the reference.  The reference builds the
dense effective Hamiltonian (chi^2 d^2)^2 (`Hamiltonian`, :38-40) and hands it to ARPACK (:57); that
needs 268 MB at chi = 32 and 68.7 GB at chi = 128, so beyond chi ~ 32 the only CPU check possible
is this index-for-index restatement with the same contraction evaluated as four tensordots.
`oracle/validate_against_reference.py` pins it against the unmodified reference where the reference
can run (dense H_eff entries, `next_LR` energies for TFIM at trunc = 10 and 16) and against exact
diagonalisation through `DMRG/Ising.py`.

Index conventions: L[A,a,i] (bra, ket, mpo), R[D,d,k], W[i,j,B,b] (mpo-row, mpo-col, bra, ket),
theta[a,b,c,d].
"""
import numpy as np
import scipy.linalg as la
import scipy.sparse.linalg as sla


def heff_matvec(L, W, R, theta):
    """out[A,B,C,D] = sum L[A,a,i] W[i,j,B,b] W[j,k,C,c] R[D,d,k] theta[a,b,c,d]  (DMRG/iDMRG.py:40)."""
    t = np.tensordot(L, theta, axes=([1], [0]))                 # A i b c d
    t = np.tensordot(W, t, axes=([0, 3], [1, 2]))               # j B A c d
    t = np.tensordot(W, t, axes=([0, 3], [0, 3]))               # k C B A d
    t = np.tensordot(t, R, axes=([0, 4], [2, 1]))               # C B A D
    return t.transpose(2, 1, 0, 3)


def heff_dense(L, W, R):
    """Dense H_eff exactly as the reference builds it (DMRG/iDMRG.py:38-46); small chi only."""
    A = np.einsum('Aai, ijBb, jkCc, Ddk->ABCDabcd', L, W, W, R, optimize=True)
    n = int(round(np.sqrt(A.size)))
    return A.reshape(n, n)


def halve(psi, shape, trunc):
    """DMRG/iDMRG.py:17-35 without the print: full S (unnormalised), U[:, :k], V[:k], k = min(trunc, S.size)."""
    xl, d1, d2, xr = shape
    U, S, V = la.svd(np.asarray(psi).reshape(xl * d1, d2 * xr), full_matrices=False)
    k = min(trunc, S.size)
    return U[:, :k].reshape(xl, d1, k), S, V[:k].reshape(k, d2, xr)


def env_left(L, U, W, swap_conj=False):
    """DMRG/iDMRG.py:59, 'ijk, ilA, jmB, kClm->ABC'.  swap_conj=True is the reference as written (conjugate on the ket
    leg); False is the mathematically correct projection.  Identical for real U.
    Evaluated as three BLAS tensordots (np.einsum's own path takes minutes at chi = 256)."""
    bra, ket = (U, U.conj()) if swap_conj else (U.conj(), U)
    t = np.tensordot(L, bra, axes=([0], [0]))                   # j k l A
    t = np.tensordot(t, W, axes=([1, 2], [0, 2]))               # j A C m
    t = np.tensordot(t, ket, axes=([0, 3], [0, 1]))             # A C B
    return t.transpose(0, 2, 1)


def env_right(R, V, W, swap_conj=False):
    """DMRG/iDMRG.py:60, 'ijk, Ali, Bmj, Cklm->ABC'."""
    bra, ket = (V, V.conj()) if swap_conj else (V.conj(), V)
    t = np.tensordot(R, bra, axes=([0], [2]))                   # j k A l
    t = np.tensordot(t, W, axes=([1, 3], [1, 2]))               # j A C m
    t = np.tensordot(t, ket, axes=([0, 3], [2, 1]))             # A C B
    return t.transpose(0, 2, 1)


def ground_state(L, W, R, v0=None, tol=0):
    """Lowest eigenpair of H_eff; ARPACK through a LinearOperator (same solver as DMRG/iDMRG.py:57)."""
    xl, xr, d = L.shape[0], R.shape[0], W.shape[2]
    shape = (xl, d, d, xr)
    n = int(np.prod(shape))
    if n <= 64:   # ARPACK needs k < n - 1; tiny first steps go through LAPACK
        w, v = la.eigh(heff_dense(L, W, R))
        return w[0], v[:, 0].reshape(shape)
    dtype = np.result_type(L.dtype, W.dtype, R.dtype, np.float64)
    op = sla.LinearOperator((n, n), dtype=dtype,
                            matvec=lambda x: heff_matvec(L, W, R, x.reshape(shape)).ravel())
    w, v = sla.eigsh(op, k=1, which='SA', v0=v0, tol=tol)
    return w[0], v[:, 0].reshape(shape)


def next_LR(L, R, W, trunc=10, swap_conj=False, return_state=False):
    """One growth step, DMRG/iDMRG.py:49-61: returns (L', R', E_total[, S, theta])."""
    E, theta = ground_state(L, W, R)
    U, S, V = halve(theta, theta.shape, trunc)
    Ln, Rn = env_left(L, U, W, swap_conj), env_right(R, V, W, swap_conj)
    if return_state:
        return Ln, Rn, E, S, theta
    return Ln, Rn, E


def entropy_from_S(S, k=None):
    """Entanglement entropy (nats) of the top-k renormalised singular values."""
    s = np.asarray(S, dtype=float)[:k]
    p = s ** 2 / np.sum(s ** 2)
    p = p[p > 0]
    return float(-np.sum(p * np.log(p)))


def tfim_mpo(g=1.0, J=1.0, dtype=complex):
    """W for H = -g sum X - J sum ZZ in the reference's lower-triangular convention
    (what `(-g*MPO('sx') - J*MPO('sz')**2).tomatrix()` yields, DMRG/iDMRG.py:69)."""
    X = np.array([[0, 1], [1, 0]], dtype=dtype)
    Z = np.array([[1, 0], [0, -1]], dtype=dtype)
    W = np.zeros((3, 3, 2, 2), dtype=dtype)
    W[0, 0] = np.eye(2)
    W[2, 2] = np.eye(2)
    W[1, 0] = -J * Z
    W[2, 1] = Z
    W[2, 0] = -g * X
    return W


def boundary(w, dtype=complex):
    """L selects the last MPO row, R the first column (DMRG/iDMRG.py:70-71)."""
    L = np.zeros((1, 1, w), dtype=dtype)
    R = np.zeros((1, 1, w), dtype=dtype)
    L[0, 0, w - 1] = 1
    R[0, 0, 0] = 1
    return L, R
