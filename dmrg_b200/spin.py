"""Spin-s operators and Pauli matrices with the names of the reference's `DMRG/spin.py`.

Model-builder API only (tiny host-side matrices that end up as gate / MPO inputs of the kernels).
Basis order is m = s, s-1, ..., -s as in the reference (`DMRG/spin.py:51`).
"""
import numpy as np


def dof(s):
    """Hilbert-space dimension 2s+1 of a spin s (`DMRG/spin.py:11-24`)."""
    return round(2 * s + 1)


def I(s=1 / 2):
    """Identity (`DMRG/spin.py:27-38`)."""
    return np.eye(dof(s))


def R(s=1 / 2):
    """Spin reversal: the anti-diagonal permutation (`DMRG/spin.py:41-45`)."""
    return np.fliplr(np.eye(dof(s)))


def Z(s=1 / 2):
    """S_z = diag(s, s-1, ..., -s) (`DMRG/spin.py:48-59`)."""
    n = dof(s)
    return np.diag([s - k for k in range(n)]).astype(float)


def P(s=1 / 2):
    """S_+ : <m+1|S_+|m> = sqrt(s(s+1) - m(m+1)) on the super-diagonal (`DMRG/spin.py:62-77`)."""
    n = dof(s)
    m = np.array([s - k for k in range(1, n)], dtype=float)      # the lower m of each raised pair
    return np.diag(np.sqrt(s * (s + 1) - m * (m + 1)), k=1)


def N(s=1 / 2):
    """S_- = S_+^T (`DMRG/spin.py:80-91`)."""
    return P(s).T.copy()


def X(s=1 / 2):
    """S_x = (S_+ + S_-)/2 (`DMRG/spin.py:94-105`)."""
    return 0.5 * (P(s) + N(s))


def Y(s=1 / 2):
    """S_y = (S_+ - S_-)/(2i) (`DMRG/spin.py:108-119`)."""
    return (P(s) - N(s)) / 2j


def S(s=1 / 2):
    """(1, S_x, S_y, S_z) (`DMRG/spin.py:122-140`)."""
    return I(s), X(s), Y(s), Z(s)


sigma = [I(), 2 * X(), 2 * Y(), 2 * Z()]
"""Pauli matrices (sigma_0, sigma_x, sigma_y, sigma_z) (`DMRG/spin.py:143`)."""
