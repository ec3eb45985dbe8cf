"""Dense 2^n x 2^n spin-chain Hamiltonians with the names of the reference's `DMRG/Ising.py`.

Exact-diagonalisation cross-checks and model-builder API only; nothing here is accelerated
(SURVEY §2.1 row 6).  Sign conventions are the reference's.
"""
from functools import reduce

import numpy as np
import scipy.sparse as sps

from .spin import sigma


def nearest(n, *ops, coef=1, sparse=False):
    """sum_i coef[i] * ops[0]_i ops[1]_{i+1} ... on an open chain of n sites (`DMRG/Ising.py:29-56`)."""
    k = len(ops)
    one = np.eye(ops[0].shape[0])
    weight = coef * np.ones(n)
    # sparse: the first factor is wrapped as a sparse MATRIX, which keeps scipy.sparse.kron on the spmatrix interface
    # the reference relies on (and silences scipy's "switching to the sparse array interface" deprecation)
    kron = (lambda a, b: sps.kron(sps.coo_matrix(a), b)) if sparse else np.kron
    total = 0
    for start in range(n - k + 1):
        factors = [one] * start + list(ops) + [one] * (n - k - start)
        total = total + weight[start] * reduce(kron, factors)
    return total


def Hamilton_trans(n, g=0, J=1):
    """Transverse-field Ising, H = -sum J Z Z - sum g X (`DMRG/Ising.py:59-83`)."""
    H = -nearest(n, sigma[3], sigma[3], coef=J).astype('double')
    H = H - nearest(n, sigma[1], coef=g)
    return {'H': H, 'H_template': r'$H=-\sum_i JZ_iZ_{i+1} - \sum_i gX_i$', 'n': n, 'J': J, 'g': g}


def Hamilton_XX(n, delta=1 / 2, g=1):
    """H = -sum (Z Z + delta X X) - sum g X (`DMRG/Ising.py:86-111`)."""
    H = -nearest(n, sigma[1], coef=g).astype('double')
    H = H - nearest(n, sigma[3], sigma[3])
    H = H - nearest(n, sigma[1], sigma[1], coef=delta)
    return {'H': H, 'H_template': r'$H=-\sum (Z_iZ_{i+1}+\Delta X_iX_{i+1})-\sum g_iX_i$', 'n': n, 'delta': delta,
            'g': g}


def Hamilton_XZ(n, delta=1 / 2, g=1, h=0.1):
    """XXZ chain with fields, H = -sum (XX + YY + delta ZZ) + sum (g X + h Z) (`DMRG/Ising.py:114-123`)."""
    H = -nearest(n, sigma[1], sigma[1]).astype('complex128')
    H = H - nearest(n, sigma[2], sigma[2])
    H = H - nearest(n, sigma[3], sigma[3], coef=delta)
    H = H + nearest(n, sigma[1], coef=g)
    H = H + nearest(n, sigma[3], coef=h)
    return {'H': H, 'H_template': r'$H=-\sum (X_iX_j+Y_iY_j+\Delta Z_iZ_j)+\sum (gX_i+hZ_i)$', 'n': n,
            'delta': delta, 'g': g, 'h': h}


def Hamilton_TL(n, J=1, g=0.945, h=0.8090):
    """Ising chain in a tilted field, H = -sum J Z Z + sum (g X + h Z) (`DMRG/Ising.py:126-152`)."""
    H = -nearest(n, sigma[3], sigma[3], coef=J).astype('double')
    H = H + nearest(n, sigma[1], coef=g)
    H = H + nearest(n, sigma[3], coef=h)
    return {'H': H, 'H_template': r'$H=-\sum J Z_iZ_{i+1}+\sum (gX_i+hZ_i)$', 'n': n, 'J': J, 'g': g, 'h': h}
