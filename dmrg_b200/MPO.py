"""Symbolic matrix-product-operator algebra with the interface of the reference's `DMRG/MPO.py`.

An `MPO` is a lower-triangular operator-valued matrix of size (dof+2)^2 whose first column / last
row carry the finished / unfinished terms of a translation-invariant Hamiltonian:

    MPO('sz')            one-site term          [[1, 0], [sz, 1]]
    a + b                sum of Hamiltonians    (`DMRG/MPO.py:81-97`)
    a * b                a on site i times b on site i+1   (`DMRG/MPO.py:107-117, 135-138`)
    a ** n               a on n consecutive sites          (`DMRG/MPO.py:140-144`)
    c * a, -a, a - b     scalar multiples; the scalar lands on the entries of column 0 (`:119-125`)

`tomatrix()` returns the dense W[w, w, d, d] (mpo-row, mpo-col, bra, ket) that the iDMRG kernels
consume (`DMRG/MPO.py:68-76`).  Entries are addressed by the parallel lists `x` (row), `y` (column),
`op`, `name` exactly like the reference so that its users keep working.  Host-side only.
"""
import copy

import numpy as np

from .spin import sigma

op2 = dict(zip(('s0', 'sx', 'sy', 'sz'), sigma))


class MPO:
    width = 4
    lazy = True

    def __init__(self, name='s0', op=None, dim=2):
        self.x, self.y = [1], [0]
        self.op, self.name = [op], [name]
        self.dof = 0
        self.dim = dim
        self.N = 1

    # ---- evaluation ------------------------------------------------------------------------------
    def eval_op(self):
        """Resolve operators that were given by name only (Pauli namespace `op2`, `DMRG/MPO.py:44-47`)."""
        for k, o in enumerate(self.op):
            if o is None:
                self.op[k] = np.array(eval(self.name[k], dict(op2)))

    def tomatrix(self, dtype='complex', names=op2):
        w = self.dof + 2
        W = np.zeros((w, w, self.dim, self.dim), dtype=dtype)
        W[0, 0] = W[w - 1, w - 1] = np.eye(self.dim)
        self.eval_op()
        for r, c, o in zip(self.x, self.y, self.op):
            W[r, c] += o
        return W

    def tomatrix_name(self):
        w = self.dof + 2
        T = np.zeros((w, w), dtype='U10')
        T[0, 0] = T[w - 1, w - 1] = '1'
        for k in range(1, self.dof + 1):
            T[k, k] = '↰'
        for r, c, label in zip(self.x, self.y, self.name):
            signed = label.startswith('+') or label.startswith('-')
            T[r, c] += ('+' + label) if (T[r, c] and not signed) else label
        return T

    def __str__(self):
        return '\n'.join(''.join('{:^{}}'.format(cell, MPO.width) for cell in row) for row in self.tomatrix_name())

    __repr__ = __str__

    def copy(self):
        return copy.deepcopy(self)

    # ---- algebra -----------------------------------------------------------------------------------
    def __iadd__(self, rhs):
        last = self.dof + rhs.dof + 1                       # index of the shared bottom row
        self.x = [r if r <= self.dof else last for r in self.x] + [r + self.dof for r in rhs.x]
        self.y = list(self.y) + [c + self.dof if c else 0 for c in rhs.y]
        self.op = self.op + rhs.op
        self.name = self.name + rhs.name
        self.dof += rhs.dof
        self.N += rhs.N
        return self

    def __add__(self, rhs):
        out = self.copy()
        out += rhs
        return out

    def __rmul__(self, lhs):
        out = self.copy()
        out.eval_op()
        out.op = [o * lhs if c == 0 else o for o, c in zip(out.op, out.y)]
        return out

    def __neg__(self):
        return (-1) * self

    def __sub__(self, rhs):
        out = self.copy()
        out += (-1) * rhs
        return out

    def __imul__(self, rhs):
        assert(self.dim == rhs.dim)
        shift = rhs.dof + 1
        self.x = [r + shift for r in self.x] + list(rhs.x)
        self.y = [c + shift for c in self.y] + list(rhs.y)
        self.op = self.op + rhs.op
        self.name = self.name + rhs.name
        self.dof += shift
        self.N += rhs.N
        return self

    def __mul__(self, rhs):
        out = self.copy()
        out *= rhs
        return out

    def __pow__(self, n):
        out = self.copy()
        for _ in range(n - 1):
            out *= self
        return out


def two_site_mpo(H2, dim, tol=1e-12):
    """MPO of sum_i H2_{i,i+1} for an arbitrary two-site operator H2 (d^2 x d^2, kron ordering):
    operator-Schmidt decomposition H2 = sum_k A_k (x) B_k, then W[k+1, 0] = B_k, W[w-1, k+1] = A_k.
    New (the reference has no such builder; BASELINE config 2 needs an AKLT MPO — SURVEY §8d)."""
    d = dim
    T = np.asarray(H2).reshape(d, d, d, d).transpose(0, 2, 1, 3).reshape(d * d, d * d)   # (a a'),(b b')
    u, s, vh = np.linalg.svd(T)
    keep = int(np.sum(s > tol * s[0]))
    w = keep + 2
    real = np.isrealobj(H2) or np.allclose(np.asarray(H2).imag, 0)
    W = np.zeros((w, w, d, d), dtype=float if real else complex)
    W[0, 0] = W[w - 1, w - 1] = np.eye(d)
    for k in range(keep):
        A = (u[:, k] * np.sqrt(s[k])).reshape(d, d)
        B = (vh[k] * np.sqrt(s[k])).reshape(d, d)
        W[w - 1, k + 1] = A.real if real else A
        W[k + 1, 0] = B.real if real else B
    return W
