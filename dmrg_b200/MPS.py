"""Matrix product state with the reference's interface (peijunz-archive/DMRG, `DMRG/MPS.py`) whose
numerics run in libmps_b200's sm_100a kernels.

Same names, argument meaning, in-place mutation and error behaviour as the reference class
(`DMRG/MPS.py:48-508`): B-form + Lambda storage, `M` holds L+2 site tensors (two dummy boundary
tensors at the end, :124-125), `s` the L+1 Schmidt vectors, `x` the bond dimensions.

Where the data lives: site tensors are device-resident complex128 buffers.  `mps.M[i]` hands out a
host copy (numpy) and remembers that it did; whatever the caller wrote into it is uploaded again
before the next kernel runs, so reference-style code such as `s.M[0][0, :, :] = ...`
(`DMRG/MPS_test.py:13`) works unchanged.  Schmidt vectors are tiny and stay on the host.
"""
import copy as _copy
import math

import numpy as np
import scipy.linalg as la

from . import ST
from . import _lib


def truncate(s, trun, err):
    """Reference `truncate` (DMRG/MPS.py:14-38): in place `s /= s[0]`, keep `min(#(|s| > err), trun)`
    values, renormalise.  Host-side: it acts on a vector of at most chi*d doubles."""
    s /= s[0]
    ind = min(int(np.sum(np.abs(s) > err)), trun)
    s = s[:ind]
    s /= la.norm(s)
    return s, ind


def svd_truncate(m, trun=20, err=1e-10):
    """Reference `svd_truncate` (DMRG/MPS.py:41-45) on the GPU: returns (u[:, :k], s, vh[:k]) as numpy
    arrays; singular vectors carry an arbitrary phase per singular value, as LAPACK's do."""
    m = np.asarray(m)
    rows, cols = m.shape
    u, s, vh, k = _device_svd(_lib.to_device(m), rows, cols, trun, err, want_u=True)
    return _lib.to_host(u), _lib.to_host(s), _lib.to_host(vh)


def _device_svd(a_dev, rows, cols, trun, err, want_u, normalise=1):
    """mpsb_svd_trunc on one device matrix; returns device tensors cut to the rank (one sync).  A real (float64)
    matrix goes through mpsb_svd_trunc_real and comes back as real factors."""
    lib = _lib.load()
    real = not a_dev.is_complex()
    dt = "f64" if real else "c128"
    kcap = max(1, min(rows, cols, max(int(trun), 1)))
    vh = _lib.empty((kcap, cols), dt)
    uh = _lib.empty((kcap, rows), dt) if want_u else None
    s = _lib.empty((kcap,), "f64")
    sraw = _lib.empty((min(rows, cols),), "f64")
    rank = _lib.empty((1,), "i32")
    wsq, fn = (lib.mpsb_svd_trunc_real_workspace, lib.mpsb_svd_trunc_real) if real else (lib.mpsb_svd_trunc_workspace,
                                                                                         lib.mpsb_svd_trunc)
    ws, wsn = _lib.workspace(wsq(1, rows, cols, int(want_u)))
    _lib.check(fn(1, rows, cols, _lib.ptr(a_dev), cols, 0, int(trun), float(err), int(normalise), kcap,
                  _lib.ptr(vh), _lib.ptr(uh), _lib.ptr(s), _lib.ptr(sraw), _lib.ptr(rank), ws, wsn,
                  _lib.stream_ptr()), "mpsb_svd_trunc")
    k = int(rank.item())
    _lib.warn_if_unconverged("svd_truncate")
    u = None
    if want_u:
        u = _lib.conj_transpose(uh[:k]) if k else _lib.empty((rows, 0), dt)
    if normalise:
        return u, s[:k], vh[:k], k
    return u, sraw, vh[:k], k


class _SiteList:
    """`mps.M`: list-like view of the device-resident site tensors (see module docstring)."""

    def __init__(self, owner):
        self._o = owner

    def _norm(self, i):
        n = len(self._o._dev)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("site index out of range")
        return i

    def __len__(self):
        return len(self._o._dev)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        i = self._norm(i)
        h = self._o._host.get(i)
        if h is None:
            h = _lib.to_host(self._o._dev[i])
            self._o._host[i] = h
        return h

    def __setitem__(self, i, value):
        i = self._norm(i)
        self._o._host[i] = np.array(value, dtype=complex)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def append(self, value):
        self._o._dev.append(_lib.to_device(value))


class MPS:
    """Matrix product state, B form (`M[i]` right-canonical, includes the right Lambda); see the
    reference docstring `DMRG/MPS.py:49-102` for the picture."""

    def __init__(self, arg=(1, 1), dim=2, trun=20, err=1e-6):
        self._dev = []
        self._host = {}
        if isinstance(arg[0], np.ndarray):
            self.from_M(arg)
        else:
            self.from_chi(arg, dim)
        # dummy boundary tensors (DMRG/MPS.py:124-125)
        self._dev.append(_lib.to_device(np.eye(self.x[-1])[:, :, np.newaxis]))
        self._dev.append(_lib.to_device(np.eye(self.x[0])[np.newaxis]))
        self.s = np.zeros_like(self.x, dtype='object')
        self.s[0] = np.ones(self.x[0])
        self.canonical = False
        self.trun = trun
        self.err = err

    # ---- construction -----------------------------------------------------------------------------
    def from_chi(self, x, dim):
        """Zero tensors with bond dimensions x[0..L] (DMRG/MPS.py:129-144)."""
        self.dim = dim
        self.x = np.array(x, dtype='i8')
        self.L = len(x) - 1
        self._dev = [_lib.zeros(tuple(int(v) for v in self.shape(i))) for i in range(self.L)]
        self._host = {}

    def from_M(self, M):
        """Adopt a list of site tensors T[left, phys, right] (DMRG/MPS.py:146-171)."""
        lefts = [m.shape[0] for m in M]
        rights = [m.shape[-1] for m in M]
        phys = [m.shape[1] for m in M]
        assert lefts[1:] == rights[:-1], "Incompatible matrices!"
        for p in phys:
            assert p == phys[0], "Incompatible shape!"
        self.dim = phys[0]
        self.x = np.array(lefts + [rights[-1]], dtype='i8')
        self.L = len(M)
        self._dev = [_lib.to_device(m) for m in M]
        self._host = {}

    @staticmethod
    def naive(*local):
        """Product state from per-site amplitude vectors; an int i stands for (i, 1-i) (DMRG/MPS.py:173-188)."""
        if len(local) == 1:
            local = local[0]
        if isinstance(local[0], int):
            local = [(i, 1 - i) for i in local]
        local = np.array(local)[:, np.newaxis, :, np.newaxis]
        return MPS(local)

    def __getitem__(self, ind) -> float:
        """Amplitude <product state(ind)|self> (DMRG/MPS.py:190-195)."""
        return self.naive(ind).dot(self)

    # ---- bookkeeping --------------------------------------------------------------------------------
    @property
    def M(self):
        return _SiteList(self)

    @M.setter
    def M(self, value):
        self._dev = [_lib.to_device(v) for v in value]
        self._host = {}

    def _flush(self):
        """Upload every site tensor that was handed out to (and possibly modified by) the caller."""
        if self._host:
            for i, h in self._host.items():
                self._dev[i] = _lib.to_device(h)
            self._host = {}

    def _schmidt(self, b):
        """Device copy of the Schmidt vector on bond b (the integer 0 before canon(), DMRG/MPS.py:126)."""
        v = np.asarray(self.s[b], dtype=float)
        if v.ndim == 0:
            v = np.full(int(self.x[b]), float(v))
        return _lib.to_device(v, np.float64)

    def shape(self, i):
        return self.xl[i], self.dim, self.xr[i]

    def graph(self):
        top = '-- v --'.join('{:^3}'.format(int(i)) for i in self.x)
        bars = ''.join('|' if ch == 'v' else ' ' for ch in top)
        dims = bars.replace('  |  ', '{:^5}'.format(self.dim))
        return '\n'.join((top, bars, dims))

    __repr__ = graph
    __str__ = graph

    @property
    def Sl(self):
        return self.s[:-1]

    @property
    def Sr(self):
        return self.s[1:]

    @property
    def xl(self):
        return self.x[:-1]

    @property
    def xr(self):
        return self.x[1:]

    def B(self, i):
        return self.M[i]

    def block_single(self, i):
        """Lambda_l . B_i on the host (DMRG/MPS.py:308-310)."""
        return self.Sl[i][:, np.newaxis, np.newaxis] * self.M[i]

    def block(self, start, n):
        """Sites [start, start+n) contracted into (chi_l, d^n, chi_r), left Lambda included (DMRG/MPS.py:312-329)."""
        return self.Sl[start][:, np.newaxis, np.newaxis] * _lib.to_host(self._block_dev(start, n))

    def _block_dev(self, start, n):
        self._flush()
        blk = self._dev[start]
        for i in range(start + 1, start + n):
            nxt = self._dev[i]
            xl, dd, x = blk.shape
            blk = _lib.matmul(blk.view(xl * dd, x), nxt.view(x, -1)).view(xl, dd * self.dim, nxt.shape[2])
        return blk

    def verify_shape(self):
        for i in range(self.L):
            expt = tuple(int(v) for v in self.shape(i))
            got = tuple(self._host[i].shape) if i in self._host else tuple(self._dev[i].shape)
            assert expt == got, 'Shape error at {}, expected {}, got {}'.format(i, expt, got)

    def copy(self):
        self._flush()
        new = _copy.copy(self)
        new._dev = [t.clone() for t in self._dev]
        new._host = {}
        new.s = _copy.deepcopy(self.s)
        new.x = self.x.copy()
        return new

    # ---- gauge ------------------------------------------------------------------------------------------
    def ortho_left_site(self, i):
        """Left orthogonalisation of site i (DMRG/MPS.py:246-264), one fused launch sequence."""
        self._flush()
        lib = _lib.load()
        xl, d, xr = (int(v) for v in self._dev[i].shape)
        nxt = self._dev[i + 1]
        nnext = int(nxt.numel() // nxt.shape[0])
        kcap = max(1, min(xl * d, xr, max(int(self.trun), 1)))
        m_out = _lib.empty((xl, d, kcap))
        nxt_out = _lib.empty((kcap, nnext))
        s_out = _lib.empty((kcap,), "f64")
        rank = _lib.empty((1,), "i32")
        ws, wsn = _lib.workspace(lib.mpsb_gauge_workspace(xl, xr, d, nnext, kcap))
        sl = self._schmidt(i)
        _lib.check(lib.mpsb_gauge_left(xl, xr, d, _lib.ptr(self._dev[i]), _lib.ptr(sl), _lib.ptr(nxt), nnext,
                                       int(self.trun), float(self.err), kcap, _lib.ptr(m_out), _lib.ptr(s_out),
                                       _lib.ptr(nxt_out), _lib.ptr(rank), ws, wsn, _lib.stream_ptr()), "mpsb_gauge_left")
        k = int(rank.item())
        self._dev[i] = m_out if k == kcap else m_out[:, :, :k].contiguous()
        self._dev[i + 1] = (nxt_out if k == kcap else nxt_out[:k].contiguous()).view(k, *nxt.shape[1:])
        self.xr[i] = k
        self.Sr[i] = _lib.to_host(s_out[:k])

    def ortho_right_site(self, i):
        """Right orthogonalisation of site i (DMRG/MPS.py:266-284)."""
        self._flush()
        lib = _lib.load()
        xl, d, xr = (int(v) for v in self._dev[i].shape)
        prv = self._dev[i - 1]
        nprev = int(prv.numel() // prv.shape[-1])
        kcap = max(1, min(xl, d * xr, max(int(self.trun), 1)))
        m_out = _lib.empty((kcap, d, xr))
        prv_out = _lib.empty((nprev, kcap))
        s_out = _lib.empty((kcap,), "f64")
        rank = _lib.empty((1,), "i32")
        ws, wsn = _lib.workspace(lib.mpsb_gauge_workspace(xl, xr, d, nprev, kcap))
        sr = self._schmidt(i + 1)
        _lib.check(lib.mpsb_gauge_right(xl, xr, d, _lib.ptr(self._dev[i]), _lib.ptr(sr), _lib.ptr(prv),
                                        nprev, int(self.trun), float(self.err), kcap, _lib.ptr(m_out), _lib.ptr(s_out),
                                        _lib.ptr(prv_out), _lib.ptr(rank), ws, wsn, _lib.stream_ptr()), "mpsb_gauge_right")
        k = int(rank.item())
        self._dev[i] = m_out if k == kcap else m_out[:k].contiguous()
        self._dev[i - 1] = (prv_out if k == kcap else prv_out[:, :k].contiguous()).view(*prv.shape[:-1], k)
        self.xl[i] = k
        self.Sl[i] = _lib.to_host(s_out[:k])

    def _scale_site(self, i, factor):
        """M[i] *= factor through the one-site kernel (U = factor * identity)."""
        t = self._dev[i]
        xl, d, xr = (int(v) for v in t.shape)
        U = _lib.to_device(np.eye(d) * factor)
        out = t if d <= 16 else _lib.empty(tuple(t.shape))       # the in-place kernel covers d <= 16, larger d go through GEMM
        _lib.check(_lib.load().mpsb_one_site(1, xl, xr, d, _lib.ptr(t), 0, _lib.ptr(U), 0, _lib.ptr(out), 0,
                                             _lib.stream_ptr()), "mpsb_one_site")
        self._dev[i] = out

    def unify_end(self):
        """Divide the phases parked in the dummy boundary tensors back in (DMRG/MPS.py:286-293)."""
        self._flush()
        if self.xl[0] == 1:
            ph = complex(self._dev[-1].view(-1)[0].item())
            self._scale_site(0, 1.0 / ph)
            self._dev[-1] = _lib.to_device(np.ones((1, 1, 1)))
        if self.xr[-1] == 1:
            ph = complex(self._dev[self.L].view(-1)[0].item())
            self._scale_site(self.L - 1, 1.0 / ph)
            self._dev[self.L] = _lib.to_device(np.ones((1, 1, 1)))

    def canon(self):
        """Left sweep, right sweep, strip boundary phases (DMRG/MPS.py:295-302)."""
        for i in range(self.L):
            self.ortho_left_site(i)
        for i in reversed(range(self.L)):
            self.ortho_right_site(i)
        self.unify_end()
        self.canonical = True

    # ---- observables ----------------------------------------------------------------------------------------
    @staticmethod
    def _transfer(E, A, B, op):
        lib = _lib.load()
        ka, d, oa = (int(v) for v in A.shape)
        kb, _, rb = (int(v) for v in B.shape)
        out = _lib.empty((oa, rb))
        ws, wsn = _lib.workspace(lib.mpsb_transfer_workspace(ka, kb, d, oa, rb))
        op_dev = None if op is None else _lib.to_device(op)
        _lib.check(lib.mpsb_transfer_step(ka, kb, d, oa, rb, _lib.ptr(E), _lib.ptr(A), _lib.ptr(B), _lib.ptr(op_dev),
                                          _lib.ptr(out), ws, wsn, _lib.stream_ptr()), "mpsb_transfer_step")
        return out

    def dot(self, rhs):
        """<self|rhs> by transfer matrices (DMRG/MPS.py:339-355)."""
        assert self.dim == rhs.dim, "Shape conflict between states"
        self._flush()
        rhs._flush()
        E = _lib.to_device(np.diag(np.asarray(self.Sl[0], dtype=float) * np.asarray(rhs.Sl[0], dtype=float)))
        for i in range(self.L):
            E = self._transfer(E, self._dev[i], rhs._dev[i], None)
        return np.trace(_lib.to_host(E))

    def corr(self, *ops):
        """<prod_i op_i> for (site, operator) pairs given in site order (DMRG/MPS.py:358-389)."""
        assert(self.canonical)
        D = dict(ops)
        if len(D.keys()) == 0:
            return 1
        self._flush()
        start, end = min(D.keys()), max(D.keys()) + 1
        E = _lib.to_device(np.diag(np.asarray(self.Sl[start], dtype=float) ** 2))
        for i in range(start, end):
            E = self._transfer(E, self._dev[i], self._dev[i], D.get(i))
        return np.trace(_lib.to_host(E)).real

    def measure(self, start, op, Hermitian=True):
        """Expectation value of an operator acting on n = log_d(sqrt(op.size)) sites (DMRG/MPS.py:391-410).
        Like the reference it contracts op[b, e] with psi[b] psi*[e], i.e. returns <op^T>."""
        n = round(math.log(op.size, self.dim) / 2)
        blk = self._block_dev(start, n)
        E = _lib.to_device(np.diag(np.asarray(self.Sl[start], dtype=float) ** 2))
        E = self._transfer(E, blk, blk, np.asarray(op).T)
        ret = np.trace(_lib.to_host(E))
        return ret.real if Hermitian else ret

    # ---- updates -----------------------------------------------------------------------------------------------
    def update_single(self, U, i, unitary=True):
        """M[i][l,e,r] = sum_c U[e,c] M[i][l,c,r] (DMRG/MPS.py:413-425)."""
        self._apply_one_site(_lib.to_device(U), [i])
        if not unitary:
            self.ortho_left_site(i)

    def _apply_one_site(self, U_dev, sites):
        """The same one-site gate (already on the device) on several sites: one upload, one launch per site."""
        self._flush()
        lib = _lib.load()
        for i in sites:
            t = self._dev[i]
            xl, d, xr = (int(v) for v in t.shape)
            out = t if d <= 16 else _lib.empty(tuple(t.shape))
            _lib.check(lib.mpsb_one_site(1, xl, xr, d, _lib.ptr(t), 0, _lib.ptr(U_dev), 0, _lib.ptr(out), 0,
                                         _lib.stream_ptr()), "mpsb_one_site")
            self._dev[i] = out

    def update_double(self, U, i, unitary=True):
        """Two-site gate on sites (i, i+1) with SVD truncation (DMRG/MPS.py:428-457)."""
        self._update_bonds(_lib.to_device(np.asarray(U)), [i])
        if not unitary:
            self.ortho_left_site(i + 1)

    @staticmethod
    def _bucket(v):
        """Bond dimensions are padded up to a few size classes so that bonds of similar size share one launch
        sequence (zero padding is exact: it only adds zero singular values, which `truncate` drops)."""
        if v <= 4:
            return v
        step = 8 if v <= 64 else 32
        return (v + step - 1) // step * step

    @classmethod
    def _group_bonds(cls, x, bonds, d):
        """Partition disjoint bonds into batched calls: {(PL, PX, PR): [bond, ...]} with every bond's dimensions
        (x[i], x[i+1], x[i+2]) <= its class.  Pure bookkeeping (no device work)."""
        groups = {}
        for i in bonds:
            key = (cls._bucket(int(x[i])), cls._bucket(int(x[i + 1])), cls._bucket(int(x[i + 2])))
            groups.setdefault(key, []).append(i)
        if len(groups) > 1:
            # Small bonds are bound by launch latency, not by work: one call padded to the largest size class beats
            # one call per class (a d = 10 chain with bonds 1,5,10,...,10,8,5,4,3,1 has five classes per half sweep).
            top = tuple(max(int(x[i + j]) for i in bonds) for j in range(3))
            if top[0] * d <= 128 and d * top[2] <= 128:
                groups = {top: list(bonds)}
            else:
                # The few bonds near the ends of a saturated chain (1-2-4, 4-8-16, ...) would each be a call of their
                # own, every one paying the full single-problem latency; padded into the smallest class that covers
                # them they ride along with the bulk for a few per cent more work.
                for key in sorted(groups, key=lambda g: g[0] * g[1] * g[2]):
                    if len(groups[key]) >= 8:
                        continue
                    hosts = [g for g in groups if g != key and all(g[j] >= key[j] for j in range(3))]
                    if hosts:
                        host = min(hosts, key=lambda g: g[0] * g[1] * g[2])
                        groups[host] = groups[host] + groups.pop(key)
        return groups

    def _update_bonds(self, U_dev, bonds):
        """Apply the same gate to disjoint bonds; bonds in the same padded size class share one batched call."""
        self._flush()
        lib = _lib.load()
        d = self.dim
        groups = self._group_bonds(self.x, bonds, d)
        pending = []
        for (PL, PX, PR), idx in groups.items():
            nb = len(idx)
            kcap = max(1, min(PL * d, d * PR, max(int(self.trun), 1)))
            exact = all((int(self.x[i]), int(self.x[i + 1]), int(self.x[i + 2])) == (PL, PX, PR) for i in idx)
            if nb == 1 and exact:
                bi, bj = self._dev[idx[0]], self._dev[idx[0] + 1]
                sl = self._schmidt(idx[0])
            else:
                # one gather per operand; only the bonds smaller than the class are zero-padded first
                t = _lib.torch()

                def fit(src, shape):
                    if tuple(src.shape) == shape:
                        return src
                    out = _lib.zeros(shape)
                    out[:src.shape[0], :, :src.shape[2]].copy_(src)
                    return out
                bi = t.stack([fit(self._dev[i], (PL, d, PX)) for i in idx])
                bj = t.stack([fit(self._dev[i + 1], (PX, d, PR)) for i in idx])
                slh = np.zeros((nb, PL))
                for b, i in enumerate(idx):
                    xl = int(self.x[i])
                    slh[b, :xl] = np.asarray(self.s[i], dtype=float) * np.ones(xl)
                sl = _lib.to_device(slh, np.float64)
            bio = _lib.empty((nb, PL, d, kcap))
            bjo = _lib.empty((nb, kcap, d, PR))
            sro = _lib.empty((nb, kcap), "f64")
            rank = _lib.empty((nb,), "i32")
            ws, wsn = _lib.workspace(lib.mpsb_tebd_two_site_workspace(nb, PL, PX, PR, d))
            _lib.check(lib.mpsb_tebd_two_site(nb, PL, PX, PR, d, _lib.ptr(bi), PL * d * PX, _lib.ptr(bj), PX * d * PR,
                                              _lib.ptr(sl), PL, _lib.ptr(U_dev), 0, int(self.trun), float(self.err), kcap,
                                              _lib.ptr(bio), PL * d * kcap, _lib.ptr(bjo), kcap * d * PR, _lib.ptr(sro),
                                              kcap, _lib.ptr(rank), ws, wsn, _lib.stream_ptr()), "mpsb_tebd_two_site")
            pending.append((idx, kcap, bio, bjo, sro, rank, (PL, PR)))
        for idx, kcap, bio, bjo, sro, rank, (PL, PR) in pending:
            ranks = _lib.to_host(rank)
            _lib.warn_if_unconverged("update_double")
            sr = _lib.to_host(sro)
            for b, i in enumerate(idx):
                k = int(ranks[b])
                xl, xr = int(self.x[i]), int(self.x[i + 2])
                whole = (k == kcap and xl == PL and xr == PR)
                self._dev[i] = bio[b] if whole else bio[b, :xl, :, :k].contiguous()
                self._dev[i + 1] = bjo[b] if whole else bjo[b, :k, :, :xr].contiguous()
                self.xr[i] = k
                self.Sr[i] = sr[b, :k].copy()

    def update_k(self, U, i, k, unitary=True):
        """Gate on the k neighbouring sites i .. i+k-1 (declared in the reference, DMRG/MPS.py:459-461, where it raises
        NotImplementedError).  U has 2k legs of size d, outputs first - the convention of `update_double`
        (U[a,b,c,j] = <a b|U|c j>, :445) - or is the d^k x d^k matrix.  The k-site tensor is formed by GEMMs, the
        gate acts on the fused physical leg, and sites are split off from the right exactly as `update_double` does
        it (:446-455): SVD of Lambda_l * theta, B' = Vh, the rest = theta . Vh^H, k-1 times."""
        d = self.dim
        if k == 1:
            return self.update_single(np.asarray(U).reshape(d, d), i, unitary)
        assert k >= 2 and i + k <= self.L, "sites out of range"
        lib = _lib.load()
        D = d ** k
        Um = _lib.to_device(np.asarray(U).reshape(D, D))
        blk = self._block_dev(i, k)                                   # B_i ... B_{i+k-1} as [xl, d^k, xr]
        xl, xr = int(blk.shape[0]), int(blk.shape[2])
        T = _lib.empty((xl, D, xr))
        _lib.zgemm(Um, blk, T, D, xr, D, D, xr, xr, batch1=xl, sB=(D * xr, 0), sC=(D * xr, 0))
        sl = self._schmidt(i)
        cur = xr
        for j in range(k - 1, 0, -1):                                 # split site i+j off the right end
            rows, cols = xl * d ** j, d * cur
            Tm = T.reshape(rows, cols)
            m = _lib.empty((rows, cols))
            _lib.check(lib.mpsb_scale_rows(rows, cols, _lib.ptr(Tm), cols, _lib.ptr(sl), d ** j, _lib.ptr(m), cols,
                                           _lib.stream_ptr()), "mpsb_scale_rows")
            _, s, vh, kk = _device_svd(m, rows, cols, self.trun, self.err, want_u=False)
            self._dev[i + j] = vh.reshape(kk, d, cur).contiguous()
            self.x[i + j] = kk
            self.s[i + j] = _lib.to_host(s)
            T = _lib.matmul(Tm, vh.contiguous(), opB=_lib.OP_C)       # theta . Vh^H: [xl * d^j, kk]
            cur = kk
        self._dev[i] = T.reshape(xl, d, cur)
        if not unitary:
            self.ortho_left_site(i + k - 1)

    def iTEBD_double(self, H, t, n, order=2):
        """Second-order Suzuki-Trotter evolution exp(-i H t) with n slices for a nearest-neighbour two-site
        Hamiltonian H (DMRG/MPS.py:463-490).  All even (odd) bonds of a half sweep are independent and go to the
        GPU as one batch.  `order` (new, default = the reference's 2) selects the ST1 / ST2 / ST4 task line."""
        def even_update(k):
            U = _lib.to_device(la.expm(-1j * H * k * t).reshape([self.dim] * 4))

            def _even_update():
                self._update_bonds(U, list(range(0, self.L - 1, 2)))
            return _even_update

        def odd_update(k):
            U = _lib.to_device(la.expm(-1j * H * k * t).reshape([self.dim] * 4))

            def _odd_update():
                self._update_bonds(U, list(range(1, self.L - 1, 2)))
            return _odd_update
        {1: ST.ST1, 2: ST.ST2, 4: ST.ST4}[order]((even_update, odd_update), n)

    def testCanon(self, err=1e-13):
        """Assert both orthonormality relations of the canonical form (DMRG/MPS.py:495-508)."""
        for i in range(self.L):
            S = self.block_single(i)
            TSS = np.tensordot(S.conj(), S, axes=([0, 1], [0, 1]))
            SST = np.tensordot(S, S.conj(), axes=([1, 2], [1, 2]))
            assert la.norm(TSS - np.diag(self.Sr[i]) ** 2) < err
            assert la.norm(SST - np.diag(self.Sl[i]) ** 2) < err
