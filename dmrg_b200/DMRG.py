"""Finite-system two-site DMRG on the GPU kernels of the iDMRG step (new).

The reference stops at the infinite-size growth step (`DMRG/iDMRG.py:49-61`); its `DMRG-notes.ipynb` (cell 15) only
sketches the finite-size sweep.  Everything a sweep needs is already behind the C ABI: the matrix-free H_eff and the
Lanczos solver (`mpsb_lanczos_ground`), the split (`mpsb_svd_trunc`, through `iDMRG.halve`), the environment growth
(`mpsb_env_left/right`) and the Lambda scaling (`mpsb_scale_rows`).  This module only orders those calls:

    right sweep, bond (j, j+1), j = 0 .. L-2:   theta0 = C_j . B_{j+1}   ->  ground state  ->  U S V
                                                 A_j = U,  L_{j+1} = env(L_j, U),  C_{j+1} = S V
    left sweep,  bond (j, j+1), j = L-2 .. 0:   theta0 = A_j . C_{j+1}   ->  ground state  ->  U S V
                                                 B_{j+1} = V,  R_{j+1} = env(R_{j+2}, V),  C_j = U S

Index conventions are those of `iDMRG` (L[A,a,i], R[D,d,k], M[i,j,B,b]); the MPO tensor is the same on every site
and lower triangular, with the open-boundary vectors of the reference's driver (`DMRG/iDMRG.py:69-70`).  Every
Lanczos call starts from the current two-site tensor, so late sweeps need only a handful of H_eff products per bond.
"""
import numpy as np

from . import _lib
from . import iDMRG
from .MPS import MPS


def _scale_rows(T, s):
    """diag(s) . T for a device tensor T[k, ...] (s on the host)."""
    lib = _lib.load()
    k = int(T.shape[0])
    cols = T.numel() // k
    out = _lib.empty(tuple(T.shape))
    sd = _lib.to_device(np.asarray(s, dtype=float), np.float64)
    _lib.check(lib.mpsb_scale_rows(k, cols, _lib.ptr(T), cols, _lib.ptr(sd), 1, _lib.ptr(out), cols, _lib.stream_ptr()),
               "mpsb_scale_rows")
    return out


class FiniteChain:
    """Ground state of an open chain of `L` sites with the uniform MPO tensor `M[w, w, d, d]`, bond dimension `trunc`."""

    def __init__(self, M, L, trunc=10, seed=0, swap_conj=False):
        assert L >= 2
        M = np.asarray(M)
        self.w, self.d = int(M.shape[0]), int(M.shape[2])
        self.L, self.trunc, self.swap_conj = int(L), int(trunc), swap_conj
        self.M = _lib.to_device(M)
        d, w = self.d, self.w
        rs = np.random.RandomState(seed)
        x = [min(d ** j, d ** (L - j), 2) for j in range(L + 1)]          # small random start, right-canonical
        start = MPS([rs.randn(x[j], d, x[j + 1]) + 1j * rs.randn(x[j], d, x[j + 1]) for j in range(L)], dim=d,
                    trun=max(2, trunc), err=0.0)
        start.canon()
        self.B = [t.clone() for t in start._dev[:L]]                       # right-canonical tensors
        self.A = [None] * L                                                # left-canonical tensors (after a right sweep)
        self.S = [np.ones(1)] + [np.asarray(start.s[j + 1], dtype=float) for j in range(L)]
        L0 = np.zeros((1, 1, w), dtype=complex)
        L0[0, 0, w - 1] = 1
        R0 = np.zeros((1, 1, w), dtype=complex)
        R0[0, 0, 0] = 1
        self.Lenv = [None] * (L + 1)
        self.Renv = [None] * (L + 1)
        self.Lenv[0] = _lib.to_device(L0)
        self.Renv[L] = _lib.to_device(R0)
        for j in range(L - 1, 0, -1):
            self.Renv[j] = iDMRG._env(False, self.Renv[j + 1], self.B[j], self.M, swap_conj)
        self.C = self.B[0]                                                 # centre tensor, on site 0 (Lambda_0 = 1)
        self.centre = 0
        self.energies, self.n_matvec, self.discarded = [], [], []

    def _solve(self, j, theta0):
        E, theta = iDMRG.ground_state(self.Lenv[j], self.M, self.Renv[j + 2], theta0=theta0)
        self.n_matvec.append(iDMRG.last_info["n_matvec"])
        sh = tuple(int(v) for v in theta.shape)
        U, S_all, V = iDMRG._halve_public(theta, sh, self.trunc)
        k = int(U.shape[2])
        s_all = _lib.to_host(S_all)
        self.discarded.append(float(np.sum(s_all[k:] ** 2) / np.sum(s_all ** 2)))
        s = s_all[:k] / np.linalg.norm(s_all[:k])
        return E, U.contiguous(), s, V.contiguous()

    def sweep_right(self):
        assert self.centre == 0
        for j in range(self.L - 1):
            Cm = self.C.reshape(-1, self.C.shape[2])
            nxt = self.B[j + 1]
            theta0 = _lib.matmul(Cm, nxt.reshape(nxt.shape[0], -1)).reshape(self.C.shape[0], self.d, self.d, nxt.shape[2])
            E, U, s, V = self._solve(j, theta0)
            self.A[j] = U
            self.B[j + 1] = V
            self.S[j + 1] = s
            self.Lenv[j + 1] = iDMRG._env(True, self.Lenv[j], U, self.M, self.swap_conj)
            self.C = _scale_rows(V, s)
        self.centre = self.L - 1
        self.energies.append(E)
        return E

    def sweep_left(self):
        assert self.centre == self.L - 1
        for j in range(self.L - 2, -1, -1):
            prv = self.A[j]
            theta0 = _lib.matmul(prv.reshape(-1, prv.shape[2]), self.C.reshape(self.C.shape[0], -1))
            theta0 = theta0.reshape(prv.shape[0], self.d, self.d, self.C.shape[2])
            E, U, s, V = self._solve(j, theta0)
            self.B[j + 1] = V
            self.S[j + 1] = s
            self.Renv[j + 1] = iDMRG._env(False, self.Renv[j + 2], V, self.M, self.swap_conj)
            xl, d, k = (int(v) for v in U.shape)
            Ut = _lib.conj_transpose(U.reshape(xl * d, k), conj=False)       # [k, xl d]
            self.C = _lib.conj_transpose(_scale_rows(Ut, s), conj=False).reshape(xl, d, k)
        self.centre = 0
        self.energies.append(E)
        return E

    def run(self, sweeps=4, tol=1e-10):
        """Full sweeps (right + left) until the energy moves by less than tol * |E|; returns the energy."""
        last = None
        for _ in range(sweeps):
            self.sweep_right()
            E = self.sweep_left()
            if last is not None and abs(E - last) <= tol * max(1.0, abs(E)):
                break
            last = E
        return self.energies[-1]

    def to_mps(self):
        """The state as a canonical `MPS` (centre must be on site 0, i.e. after a left sweep)."""
        assert self.centre == 0
        tens = [_lib.to_host(self.C)] + [_lib.to_host(b) for b in self.B[1:]]
        mps = MPS(tens, dim=self.d, trun=max(self.trunc, 2), err=0.0)
        mps.canon()
        return mps

    def entropy(self, bond):
        p = np.asarray(self.S[bond]) ** 2
        p = p[p > 0]
        return float(-np.sum(p * np.log(p)))
