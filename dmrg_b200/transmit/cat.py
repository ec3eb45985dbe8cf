"""Single-mode bosonic helpers with the names of the reference's `transmit/cat.py:13-114` (operators, coherent
states, displacement and Wigner operators in a Fock space cut at n levels).  Host-side numpy: these build the d x d
operators and d-vectors that `bhopping.BMPS` feeds to the GPU kernels; nothing here is on the hot path."""
import math

import numpy as np
import scipy.linalg as la


def a_n(n):
    """Annihilation operator (transmit/cat.py:13-15)."""
    return np.diag(np.sqrt(np.arange(1, n)), k=1)


def a_p(n):
    """Creation operator (transmit/cat.py:17-19)."""
    return np.diag(np.sqrt(np.arange(1, n)), k=-1)


def N(n):
    """Diagonal of the number operator (transmit/cat.py:21-23)."""
    return np.arange(n)


def P(n):
    """Diagonal of the parity operator (transmit/cat.py:25-27)."""
    return (-1) ** np.arange(n)


def exp_ap(alpha, n=20):
    """exp(alpha a^+) in the truncated Fock space (transmit/cat.py:43-50): lower triangular with
    <i|.|j> = alpha^(i-j) / (i-j)! * sqrt(i! / j!).  The factorial ratios are formed in exact integer arithmetic
    so that each entry is good to a few ulp (products of these matrices cancel heavily, see `displace`)."""
    fact = [math.factorial(k) for k in range(n)]
    pw = np.power(complex(alpha), np.arange(n))
    A = np.zeros((n, n), dtype=complex)
    for i in range(n):
        for j in range(i + 1):
            A[i, j] = pw[i - j] * (math.sqrt(fact[i] // fact[j]) / fact[i - j])
    return A


def exp_an(alpha, n=20):
    """exp(alpha a) (transmit/cat.py:39-41)."""
    return exp_ap(np.conj(alpha), n).conj().T


def coherent(alpha, n=50):
    """Normalised coherent state |alpha> cut at n levels (transmit/cat.py:52-59)."""
    amp = np.array([complex(alpha) ** k / math.sqrt(math.factorial(k)) for k in range(n)])
    return amp / la.norm(amp)


def displace(alpha, n=20):
    """D(alpha) = exp(alpha a^+) exp(-conj(alpha) a) exp(-|alpha|^2/2) (transmit/cat.py:61-63)."""
    return exp_ap(alpha, n) @ exp_an(-np.conj(alpha), n) * np.exp(-abs(alpha) ** 2 / 2)


def wigner_matrix(alpha, n=20):
    """Parity times D(-2 alpha) (transmit/cat.py:65-67)."""
    return P(n)[:, None] * displace(-2 * alpha, n)


def kerr(k, alpha=2, n=20):
    """Coherent state after a Kerr phase exp(i pi k N^2) (transmit/cat.py:91-94)."""
    return np.exp(1j * np.pi * N(n) ** 2 * k) * coherent(alpha, n)


def fock(k, n=50):
    """Number state |k> (transmit/cat.py:96-99)."""
    v = np.zeros([n])
    v[k] = 1
    return v


def expect(W, state):
    """<W> of a state vector or of a density matrix (transmit/cat.py:101-105)."""
    state = np.asarray(state)
    if state.ndim == 1:
        return np.vdot(state, W @ state).real
    return np.sum(W * state.T).real


def wigner(alpha, psi):
    """Wigner function of psi at alpha, up to the reference's normalisation (transmit/cat.py:107-110)."""
    return expect(wigner_matrix(alpha, psi.shape[0]), psi)


def husimi(alpha, psi):
    """Husimi Q function |<alpha|psi>|^2 (transmit/cat.py:112-113)."""
    return abs(np.vdot(coherent(alpha, len(psi)), psi)) ** 2
