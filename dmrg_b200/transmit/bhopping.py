"""TEBD for a chain of bosonic modes with nearest-neighbour hopping, Kerr terms and a tail mode coupled through
a matrix product operator — the reference's `transmit/bhopping.py` (`BMPS`, :32-237) on top of the GPU `MPS`.

It is the reference's own large-local-dimension caller of the two-site update (d = 10 in its `__main__` and in
`qcircuit-prog/naive.py:33`; `trun = dim`): every gate is a d^2 x d^2 matrix, every SVD is (chi d) x (d chi).
Same constructor arguments, attributes and method names; what changes is where the work runs:

* `evolve` issues all even (odd) bonds of a half sweep as one batched `mpsb_tebd_two_site` call and uploads each
  gate once per Trotter coefficient instead of once per bond;
* `update_MPO` (:94-104) runs `mpsb_mpo_site` per site instead of a five-index `einsum`, then `canon()`.

Chain layout (reference :32-70): `l` cavity modes on sites 0..l-1 plus one tail mode on site l, so `L = l + 1`.
"""
import numpy as np
import scipy.linalg as la

from .. import _lib
from ..MPS import MPS
from ..ST import ST2
from .cat import coherent, wigner_matrix


def dist_circ(t):
    """Angle folded into [-pi, pi) (transmit/bhopping.py:28-30)."""
    return (np.asarray(t) + np.pi) % (2 * np.pi) - np.pi


class BMPS(MPS):
    def __init__(self, dim, l, para, pack):
        self.dim = dim
        vac = [self.zero().reshape([1, dim, 1]) for _ in range(l + 1)]
        super().__init__(vac, dim, trun=self.dim)
        self._init_H(**para)
        self.wavepacket(**pack)
        self.canon()

    # ---- local operators (transmit/bhopping.py:41-74) ----------------------------------------------------------
    def zero(self):
        """Vacuum amplitude vector."""
        v = np.zeros(self.dim)
        v[0] = 1
        return v

    def p(self, i):
        """Projector on level i; the zero matrix for i < 0 (transmit/bhopping.py:46-50)."""
        v = np.zeros(self.dim)
        if i >= 0:
            v[i] = 1
        return np.diag(v)

    def _init_H(self, omega0=0, g=1, u=2, h=1, K=0):
        d = self.dim
        self.omega0, self.K, self.g, self.u, self.h = omega0, K, g, u, h
        self.a = np.diag(np.sqrt(np.arange(1, d)), k=1)
        self.a_p = self.a.T.conj()
        self.N = np.diag(np.arange(d))
        self.l = self.L - 1
        self.Ht = np.diag(np.arange(d))
        # O[i, j] = projector on level j - i: the bulk MPO tensor that carries the photon number to the tail.
        self.O = np.zeros((d,) * 4)
        for i in range(d):
            for j in range(i, d):
                self.O[i, j, j - i, j - i] = 1
        self.Htc = h * (self.a_p + self.a)
        hop = np.kron(self.a_p, self.a)
        self.H1 = g * (hop + hop.T.conj())

    def Bm_single(self, mt):
        return la.expm(-1j * mt * self.Htc)

    def Bm(self, dt):
        """Tail MPO tensor: displacement-like kick exp(-i m dt Htc) selected by the incoming bond m (:76-80)."""
        return np.stack([self.Bm_single(m * dt) for m in range(self.dim)])[:, np.newaxis, :, :]

    def MPO(self, dt):
        mpo = [self.O for _ in range(self.L)]
        mpo[0] = self.O[:1]
        mpo[-1] = self.Bm(dt)
        return mpo

    # ---- MPO application (transmit/bhopping.py:94-104) ------------------------------------------------------
    def update_MPO(self, mpo):
        assert len(mpo) == self.L, "Incompatible MPO"
        self._flush()
        lib = _lib.load()
        d = self.dim
        for i in range(self.L):
            W = np.asarray(mpo[i])
            wl, wr = int(W.shape[0]), int(W.shape[1])
            assert W.shape[2:] == (d, d), "Incompatible MPO"
            t = self._dev[i]
            xl, _, xr = (int(v) for v in t.shape)
            out = _lib.empty((xl * wl, d, xr * wr))
            Wd = _lib.to_device(W)
            _lib.check(lib.mpsb_mpo_site(1, xl, xr, d, wl, wr, _lib.ptr(t), 0, _lib.ptr(Wd), 0, _lib.ptr(out), 0,
                                         _lib.stream_ptr()), "mpsb_mpo_site")
            self._dev[i] = out
            self.xl[i] = xl * wl
            self.xr[i] = xr * wr
        self.canon()

    # ---- evolution (transmit/bhopping.py:106-156) -----------------------------------------------------------
    def evolve(self, t, n=100, enable_tail=True):
        d = self.dim

        def bonds(parity):
            def worker(k):
                U = _lib.to_device(la.expm(-1j * self.H1 * k * t).reshape([d] * 4))
                sites = list(range(parity, self.l - 1, 2))
                return lambda: self._update_bonds(U, sites) if sites else None
            return worker

        def single(k):
            U1 = _lib.to_device(la.expm(-1j * (self.omega0 * self.N - (self.K / 2) * self.N ** 2) * k * t))
            return lambda: self._apply_one_site(U1, range(self.l))

        def tail(k):
            U2 = _lib.to_device(la.expm(-1j * self.u * self.Ht * k * t))
            return lambda: self._apply_one_site(U2, [self.l])

        def coupling(k):
            mpo = self.MPO(k * t)
            return lambda: self.update_MPO(mpo)

        tasks = ()
        if self.g:
            tasks += (bonds(0), bonds(1))
        if self.omega0 or self.K:
            tasks += (single,)
        if self.u:
            tasks += (tail,)
        if self.h and enable_tail:
            tasks += (coupling,)
        ST2(tasks, n)
        self.time += t
        self.c_n = self.wavepacket_mode(self.time, linear=False)

    def evolve_measure(self, time, n=1, k=10):
        """<N_i> on the cavity sites before and after each of n evolution periods (:158-174)."""
        rows = [[self.measure(i, self.N) for i in range(self.l)]]
        for step in range(n):
            print("{}/{}".format(step, n))
            self.evolve(time, k)
            self.canon()
            rows.append([self.measure(i, self.N) for i in range(self.l)])
        return np.array(rows)

    def exact_evolve(self, time, n, linear=False):
        """Occupations of the exactly propagated free wave packet (:176-184)."""
        rows = []
        for step in range(n + 1):
            self.wavepacket(*self.pack, t=time * step, trun=False, linear=linear)
            rows.append([self.measure(i, self.N) for i in range(self.l)])
        return np.array(rows)

    def alpha(self, alpha, linear=False):
        self.wavepacket(*self.pack, alpha=alpha, t=self.time, trun=False, linear=linear)

    def husimi(self, alpha, linear=False):
        pass

    def wigner(self, alpha, linear=False, threshold=0):
        """Product of single-mode Wigner operators along the current packet shape (:192-197)."""
        amps = alpha * self.c_n / la.norm(self.c_n)
        ops = [(i, wigner_matrix(a, self.dim)) for i, a in enumerate(amps)]
        print(len(ops))
        return self.corr(*ops)

    # ---- wave packets (transmit/bhopping.py:199-237) --------------------------------------------------------
    def omega(self, k, k0):
        return self.omega0 + 2 * self.g * np.cos(k)

    def linear_omega(self, k, k0):
        return self.omega0 + 2 * self.g * (np.cos(k0) - np.sin(k0) * dist_circ(k - k0))

    def wavepacket_mode(self, t=0, linear=False):
        """Site amplitudes of a Gaussian packet in momentum space, propagated freely to time t."""
        dk, center, k_c = self.pack
        x = np.arange(self.l) - center
        k = 2 * np.pi * np.arange(self.l) / self.l
        weight = np.exp(-(dist_circ(k - k_c) / dk) ** 2 / 2)
        disp = self.linear_omega if linear else self.omega
        phase = np.outer(x, k) - t * disp(k, k_c)[np.newaxis, :]
        return np.exp(1j * phase) @ weight

    def init_wave(self):
        for i in range(self.l):
            site = self.M[i]
            site[0, :, 0] = coherent(self.c_n[i], self.dim)

    def wavepacket(self, dk, center, k_c=np.pi / 2, alpha=1, t=0, trun=True, linear=False):
        self.pack = (dk, center, k_c)
        self.time = t
        self.c_n = self.wavepacket_mode(t, linear)
        if trun:
            lb, rb = center - self.l // 2, center + self.l // 2
            if lb >= 0:
                self.c_n[:lb + 1] = 0
            if rb <= self.l - 1:
                self.c_n[rb:] = 0
        self.c_n = self.c_n / la.norm(self.c_n) * alpha
        self.init_wave()
