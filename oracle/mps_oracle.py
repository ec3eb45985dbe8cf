"""CPU oracle for the MPS two-site update path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's algorithms (peijunz-archive/DMRG, `DMRG/MPS.py`), written
with BLAS-backed `tensordot`/`matmul` instead of the reference's loop-evaluated `np.einsum` so it
finishes in seconds at chi = 128.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py` may import this module; the product (`dmrg_b200/`) never
does and fails loudly without its CUDA library.

Pinning: `oracle/validate_against_reference.py` runs this module side by side with the unmodified
reference imported from `/root/reference` (update_double, canon, dot, corr, measure, iTEBD_double,
AKLT flow) and writes the golden vectors under `tests/golden/`; `tests/test_oracle.py` re-checks the
oracle against those vectors on every run.  The arithmetic itself lives in numpy/scipy (LAPACK
zgesdd through `scipy.linalg.svd`, `scipy.linalg.expm`); versions used for the goldens are recorded
in `tests/golden/MANIFEST.json`.

State layout mirrors the reference (`DMRG/MPS.py:97-127`): `M` is a list of L+2 site tensors
[chi_l, d, chi_r] (the last two are the dummy boundary tensors), `s` a list of L+1 Schmidt vectors
(`s[i]` sits on the bond to the left of site i), B-form after `canon`.
"""
import math

import numpy as np
import scipy.linalg as la


def truncate(s, trun, err):
    """Reference `truncate` (DMRG/MPS.py:14-38): relative threshold on s/s[0], cap at trun, renormalise."""
    s = np.array(s, dtype=float)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = s / s[0]
    k = min(int(np.sum(np.abs(rel) > err)), trun)
    kept = rel[:k]
    kept = kept / la.norm(kept) if k else kept
    return kept, k


def svd_truncate(m, trun=20, err=1e-10):
    """Reference `svd_truncate` (DMRG/MPS.py:41-45)."""
    u, s, vh = la.svd(m, full_matrices=False)
    s, k = truncate(s, trun, err)
    return u[:, :k], s, vh[:k]


class OracleMPS:
    """Functional twin of the reference `MPS` class restricted to the hot path."""

    def __init__(self, tensors, trun=20, err=1e-6):
        tensors = [np.array(t, dtype=complex) for t in tensors]
        for a, b in zip(tensors[:-1], tensors[1:]):
            assert a.shape[2] == b.shape[0], "Incompatible matrices!"
        self.L = len(tensors)
        self.dim = tensors[0].shape[1]
        self.x = np.array([t.shape[0] for t in tensors] + [tensors[-1].shape[2]], dtype=np.int64)
        # dummy boundary tensors, DMRG/MPS.py:124-125
        self.M = tensors + [np.eye(self.x[-1], dtype=complex)[:, :, None], np.eye(self.x[0], dtype=complex)[None]]
        self.s = [np.ones(self.x[0])] + [np.zeros(0) for _ in range(self.L)]
        self.trun, self.err = trun, err
        self.canonical = False

    # -- constructors ------------------------------------------------------------------------------
    @staticmethod
    def product(local, **kw):
        """Reference `MPS.naive` (DMRG/MPS.py:173-188); an int i means the vector (i, 1-i)."""
        vecs = [(v, 1 - v) if isinstance(v, (int, np.integer)) else v for v in local]
        return OracleMPS([np.asarray(v, dtype=complex)[None, :, None] for v in vecs], **kw)

    def copy(self):
        import copy
        return copy.deepcopy(self)

    # -- gauge -------------------------------------------------------------------------------------
    def ortho_left_site(self, i):
        """DMRG/MPS.py:246-264."""
        xl, d, xr = self.M[i].shape
        m = (self.s[i][:, None, None] * self.M[i]).reshape(xl * d, xr)
        u, s, vh = svd_truncate(m, self.trun, self.err)
        self.M[i] = u.reshape(xl, d, -1)
        self.x[i + 1] = len(s)
        self.s[i + 1] = s
        nxt = self.M[i + 1]
        self.M[i + 1] = (vh @ nxt.reshape(nxt.shape[0], -1)).reshape(len(s), *nxt.shape[1:])

    def ortho_right_site(self, i):
        """DMRG/MPS.py:266-284."""
        xl, d, xr = self.M[i].shape
        m = (self.M[i] * self.s[i + 1][None, None, :]).reshape(xl, d * xr)
        u, s, vh = svd_truncate(m, self.trun, self.err)
        self.M[i] = vh.reshape(-1, d, xr)
        self.x[i] = len(s)
        self.s[i] = s
        prv = self.M[i - 1]
        self.M[i - 1] = (prv.reshape(-1, prv.shape[-1]) @ u).reshape(*prv.shape[:-1], len(s))

    def unify_end(self):
        """DMRG/MPS.py:286-293."""
        if self.x[0] == 1:
            self.M[0] = self.M[0] / self.M[-1][0, 0, 0]
            self.M[-1][0, 0, 0] = 1
        if self.x[self.L] == 1:
            self.M[self.L - 1] = self.M[self.L - 1] / self.M[self.L][0, 0, 0]
            self.M[self.L][0, 0, 0] = 1

    def canon(self):
        """DMRG/MPS.py:295-302."""
        for i in range(self.L):
            self.ortho_left_site(i)
        for i in reversed(range(self.L)):
            self.ortho_right_site(i)
        self.unify_end()
        self.canonical = True

    # -- updates -----------------------------------------------------------------------------------
    def update_single(self, U, i):
        """DMRG/MPS.py:423."""
        self.M[i] = np.tensordot(self.M[i], U, axes=([1], [1])).transpose(0, 2, 1)

    def update_double(self, U, i, unitary=True):
        """DMRG/MPS.py:445-457."""
        Bi, Bj = self.M[i], self.M[i + 1]
        xl, d, _ = Bi.shape
        xr = Bj.shape[2]
        theta0 = np.tensordot(Bi, Bj, axes=([2], [0]))                      # l c j k
        thetaB = np.tensordot(theta0, U, axes=([1, 2], [2, 3]))            # l k a b
        thetaB = thetaB.transpose(0, 2, 3, 1)                               # l a b k
        m = (self.s[i][:, None, None, None] * thetaB).reshape(xl * d, d * xr)
        _, s, vh = svd_truncate(m, self.trun, self.err)
        k = len(s)
        self.x[i + 1] = k
        self.s[i + 1] = s
        self.M[i] = (thetaB.reshape(xl * d, d * xr) @ vh.conj().T).reshape(xl, d, k)
        self.M[i + 1] = vh.reshape(k, d, xr)
        if not unitary:
            self.ortho_left_site(i + 1)

    def update_k(self, U, i, k):
        """k-site gate (declared at DMRG/MPS.py:459-461, NotImplemented there): the scheme of `update_double`
        (:445-455) applied k-1 times from the right.  For k = 2 this IS update_double."""
        d = self.dim
        D = d ** k
        theta = self.M[i]
        for j in range(1, k):
            theta = np.tensordot(theta, self.M[i + j], axes=([theta.ndim - 1], [0]))
        xl, xr = theta.shape[0], theta.shape[-1]
        T = np.einsum('ac,lcr->lar', np.asarray(U).reshape(D, D), theta.reshape(xl, D, xr))
        cur = xr
        for j in range(k - 1, 0, -1):
            Tm = T.reshape(xl * d ** j, d * cur)
            m = (np.repeat(self.s[i], d ** j)[:, None]) * Tm
            _, s, vh = svd_truncate(m, self.trun, self.err)
            kk = len(s)
            self.M[i + j] = vh.reshape(kk, d, cur)
            self.x[i + j] = kk
            self.s[i + j] = s
            T = Tm @ vh.conj().T
            cur = kk
        self.M[i] = T.reshape(xl, d, cur)

    def to_dense(self):
        """Full state vector psi[c_0 .. c_{L-1}] of a chain with trivial outer bonds (test helper)."""
        psi = self.s[0][:, None, None] * self.M[0]
        for j in range(1, self.L):
            psi = np.tensordot(psi, self.M[j], axes=([psi.ndim - 1], [0]))
        return psi.reshape(-1)

    def half_sweep(self, U, parity):
        """Body of `_even_update` / `_odd_update`, DMRG/MPS.py:479-480, 487-488."""
        for i in range(parity, self.L - 1, 2):
            self.update_double(U, i)

    def tebd(self, H, t, n):
        """`iTEBD_double` with the second-order schedule A(t/2n) [B(t/n) A(t/n)]^(n-1) B(t/n) A(t/2n)
        that `ST.ST2` + `streamline` produce (DMRG/MPS.py:463-490, DMRG/ST.py:37-117, DMRG/ST_test.py:24)."""
        d = self.dim
        Uh = la.expm(-1j * H * (0.5 / n) * t).reshape([d] * 4)
        U1 = la.expm(-1j * H * (1.0 / n) * t).reshape([d] * 4)
        seq = [(0, Uh)]
        for _ in range(n - 1):
            seq += [(1, U1), (0, U1)]
        seq += [(1, U1), (0, Uh)]
        for parity, U in seq:
            self.half_sweep(U, parity)

    # -- observables -------------------------------------------------------------------------------
    def block_single(self, i):
        return self.s[i][:, None, None] * self.M[i]

    def block(self, start, n):
        """DMRG/MPS.py:312-329."""
        s = self.block_single(start)
        for i in range(start + 1, start + n):
            s = np.tensordot(s, self.M[i], axes=([2], [0]))
            s = s.reshape(s.shape[0], -1, s.shape[-1])
        return s

    def dot(self, rhs):
        """<self|rhs>, DMRG/MPS.py:339-355."""
        a, b = self.block_single(0), rhs.block_single(0)
        E = np.tensordot(a.conj(), b, axes=([0, 1], [0, 1]))
        for i in range(1, self.L):
            T = np.tensordot(E, rhs.M[i], axes=([1], [0]))                  # k n r
            E = np.tensordot(self.M[i].conj(), T, axes=([0, 1], [0, 1]))    # o r
        return np.trace(E)

    def corr(self, *ops):
        """DMRG/MPS.py:358-389."""
        assert self.canonical
        D = dict(ops)
        if not D:
            return 1
        start, end = min(D), max(D) + 1
        E = np.diag(self.s[start] ** 2).astype(complex)
        for i in range(start, end):
            b = self.M[i]
            T = np.tensordot(E, b, axes=([1], [0]))                          # k j r
            if i in D:
                T = np.tensordot(D[i], T, axes=([1], [1])).transpose(1, 0, 2)  # k i r
            E = np.tensordot(b.conj(), T, axes=([0, 1], [0, 1]))
        return np.trace(E).real

    def measure(self, start, op, Hermitian=True):
        """DMRG/MPS.py:391-410 (note: contracts op[b,e] with psi[b] psi*[e], i.e. <op^T>)."""
        n = round(math.log(op.size, self.dim) / 2)
        s = self.block(start, n)
        rho = np.tensordot(s, s.conj(), axes=([0, 2], [0, 2]))               # b e
        ret = np.sum(rho * op)
        return ret.real if Hermitian else ret

    def entropy(self, bond):
        """Von Neumann entropy (nats) of the Schmidt spectrum on `bond`."""
        p = np.asarray(self.s[bond]) ** 2
        p = p[p > 0]
        return float(-np.sum(p * np.log(p)))
