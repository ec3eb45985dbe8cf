"""CPU oracle for the bosonic hopping chain `BMPS` — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's `transmit/bhopping.py` (`BMPS`, :32-237) on top of `OracleMPS`: the
large-local-dimension caller of the two-site update.  Only `tests/` may import this module.

Pinning: `oracle/validate_against_reference.py` runs the unmodified reference class (imported from
`/root/reference` with `scipy.misc.imresize`, `matplotlib` and `joblib.Memory(cachedir=...)` stubbed — the three
imports of `transmit/` that no longer exist in this image; none is used by the code under test) next to this module
and writes `tests/golden/bmps_*.npz`; `tests/test_oracle.py` re-checks against those vectors.
"""
import numpy as np
import scipy.linalg as la

from .mps_oracle import OracleMPS


def coherent(alpha, n):
    """Coherent state in a Fock space cut at n levels (transmit/cat.py:52-59)."""
    v = np.ones(n, dtype=complex)
    for k in range(1, n):
        v[k] = v[k - 1] * alpha / np.sqrt(k)
    return v / la.norm(v)


def st2_schedule(n_ops, n):
    """Merged second-order Suzuki-Trotter task line (DMRG/ST.py:37-51, 77-93): per slice the operators run
    0..n_ops-1 and back with weight 1/(2n); equal neighbours are fused."""
    line = []
    for _ in range(n):
        for op in list(range(n_ops)) + list(reversed(range(n_ops))):
            if line and line[-1][0] == op:
                line[-1][1] += 0.5 / n
            else:
                line.append([op, 0.5 / n])
    return [(op, w) for op, w in line]


class OracleBMPS(OracleMPS):
    def __init__(self, dim, l, para, pack):
        vac = np.zeros((1, dim, 1), dtype=complex)
        vac[0, 0, 0] = 1
        super().__init__([vac.copy() for _ in range(l + 1)], trun=dim, err=1e-6)     # :34-37, MPS default err
        self.l = l
        self.set_model(**para)
        self.wavepacket(**pack)
        self.canon()

    def set_model(self, omega0=0, g=1, u=2, h=1, K=0):
        """transmit/bhopping.py:52-74."""
        d = self.dim
        self.omega0, self.g, self.u, self.h, self.K = omega0, g, u, h, K
        a = np.diag(np.sqrt(np.arange(1, d)), k=1)
        self.N = np.diag(np.arange(d)).astype(float)
        self.Htc = h * (a + a.T)
        hop = np.kron(a.T, a)
        self.H1 = g * (hop + hop.T)
        self.O = np.zeros((d, d, d, d))
        for i in range(d):
            for j in range(i, d):
                self.O[i, j, j - i, j - i] = 1.0

    def mpo(self, dt):
        """transmit/bhopping.py:76-92."""
        tail = np.stack([la.expm(-1j * m * dt * self.Htc) for m in range(self.dim)])[:, None]
        return [self.O[:1]] + [self.O] * (self.L - 2) + [tail]

    def update_MPO(self, mpo):
        """transmit/bhopping.py:94-104."""
        for i in range(self.L):
            t = np.tensordot(self.M[i], mpo[i], axes=([1], [2]))          # i k l m o
            t = t.transpose(0, 2, 4, 1, 3)                                  # i l o k m
            sh = t.shape
            self.M[i] = t.reshape(sh[0] * sh[1], sh[2], sh[3] * sh[4])
            self.x[i] = sh[0] * sh[1]
            self.x[i + 1] = sh[3] * sh[4]
        self.canon()

    def evolve(self, t, n=100, enable_tail=True):
        """transmit/bhopping.py:106-156."""
        d = self.dim
        kinds = []
        if self.g:
            kinds += ["even", "odd"]
        if self.omega0 or self.K:
            kinds.append("single")
        if self.u:
            kinds.append("tail")
        if self.h and enable_tail:
            kinds.append("coupling")
        for op, w in st2_schedule(len(kinds), n):
            kind = kinds[op]
            if kind in ("even", "odd"):
                U = la.expm(-1j * self.H1 * w * t).reshape([d] * 4)
                for i in range(0 if kind == "even" else 1, self.l - 1, 2):
                    self.update_double(U, i)
            elif kind == "single":
                U1 = la.expm(-1j * (self.omega0 * self.N - (self.K / 2) * self.N @ self.N) * w * t)
                for i in range(self.l):
                    self.update_single(U1, i)
            elif kind == "tail":
                self.update_single(la.expm(-1j * self.u * self.N * w * t), self.l)
            else:
                self.update_MPO(self.mpo(w * t))
        self.time += t
        self.c_n = self.packet_amplitudes(self.time)

    def packet_amplitudes(self, t=0.0, linear=False):
        """transmit/bhopping.py:199-225."""
        dk, center, k_c = self.pack
        fold = lambda v: (v + np.pi) % (2 * np.pi) - np.pi
        k = 2 * np.pi * np.arange(self.l) / self.l
        f_k = np.exp(-(fold(k - k_c) / dk) ** 2 / 2)
        if linear:
            w_k = self.omega0 + 2 * self.g * (np.cos(k_c) - np.sin(k_c) * fold(k - k_c))
        else:
            w_k = self.omega0 + 2 * self.g * np.cos(k)
        x = np.arange(self.l) - center
        return np.exp(1j * (x[:, None] * k[None, :] - t * w_k[None, :])) @ f_k

    def wavepacket(self, dk, center, k_c=np.pi / 2, alpha=1, t=0, trun=True, linear=False):
        """transmit/bhopping.py:227-237 (+ init_wave :223-225)."""
        self.pack = (dk, center, k_c)
        self.time = t
        c = self.packet_amplitudes(t, linear)
        if trun:
            lb, rb = center - self.l // 2, center + self.l // 2
            if lb >= 0:
                c[:lb + 1] = 0
            if rb <= self.l - 1:
                c[rb:] = 0
        self.c_n = c / la.norm(c) * alpha
        for i in range(self.l):
            self.M[i][0, :, 0] = coherent(self.c_n[i], self.dim)

    def occupations(self):
        return np.array([self.measure(i, self.N) for i in range(self.L)])
