"""Many independent chains on one GPU (coupling sweeps, Trotter ensembles) — BASELINE config 4.

The reference has no batching: a parameter sweep is a Python loop over `MPS` objects (or a PBS job per
parameter, `ETH-prog/submit_job.py:3-11`).  Here `n` chains of equal length share one padded device
buffer and every two-site update of a half sweep (`DMRG/MPS.py:479-480, 487-488`) is issued once for
all chains:

    M[site][chain][P][d][P]   complex128, zero outside the chain's current bond dimensions
    S[bond][chain][P]         float64 Schmidt values, zero beyond the rank
    rank[bond][chain]         int32

P is the bond capacity (>= trun).  Zero padding is exact: padded rows/columns only add zero singular
values, which the reference's truncation rule (`DMRG/MPS.py:33-38`) drops, so results equal those of
unpadded per-chain updates.  Chains are sharded over GPUs by index (one process per GPU); the only
collective is the gather of per-chain observables (`gather_observables`).
"""
import numpy as np
import scipy.linalg as la

from . import ST
from . import _lib


class ChainBatch:
    def __init__(self, n, L, dim=2, cap=32, trun=None, err=1e-6):
        self.n, self.L, self.dim, self.P = int(n), int(L), int(dim), int(cap)
        self.trun = int(trun if trun is not None else cap)
        assert self.trun <= self.P, "trun exceeds the bond capacity"
        self.err = float(err)
        P, d = self.P, self.dim
        self.M = _lib.zeros((self.L, self.n, P, d, P))
        self.S = _lib.zeros((self.L + 1, self.n, P), "f64")
        self.rank = _lib.zeros((self.L + 1, self.n), "i32")
        self.canonical = False

    @classmethod
    def from_chains(cls, chains, cap, trun=None, err=1e-6):
        """chains: list of (site tensors, Schmidt vectors) in B form, e.g. taken from canonical MPS objects."""
        n, L = len(chains), len(chains[0][0])
        d = chains[0][0][0].shape[1]
        self = cls(n, L, d, cap, trun, err)
        P = self.P
        Mh = np.zeros((L, n, P, d, P), dtype=complex)
        Sh = np.zeros((L + 1, n, P))
        Rh = np.zeros((L + 1, n), dtype=np.int32)
        for c, (Ms, ss) in enumerate(chains):
            for i, m in enumerate(Ms):
                xl, _, xr = m.shape
                Mh[i, c, :xl, :, :xr] = m
            for b, s in enumerate(ss):
                s = np.asarray(s, dtype=float)
                Sh[b, c, :len(s)] = s
                Rh[b, c] = len(s)
        self.M, self.S, self.rank = _lib.to_device(Mh), _lib.to_device(Sh, np.float64), _lib.to_device(Rh, np.int32)
        self.canonical = True          # by contract of this constructor (call canon() otherwise)
        return self

    @classmethod
    def product(cls, local, n, cap, trun=None, err=1e-6):
        """n copies of the product state with per-site amplitude vectors `local` (cf. MPS.naive)."""
        Ms = [np.asarray(v, dtype=complex)[None, :, None] for v in local]
        ss = [np.ones(1) for _ in range(len(local) + 1)]
        return cls.from_chains([(Ms, ss)] * n, cap, trun, err)

    def chain(self, c):
        """Host copy of chain c cut to its bond dimensions: (site tensors, Schmidt vectors)."""
        r = _lib.to_host(self.rank[:, c])
        Ms = [_lib.to_host(self.M[i, c])[:r[i], :, :r[i + 1]] for i in range(self.L)]
        ss = [_lib.to_host(self.S[b, c])[:r[b]] for b in range(self.L + 1)]
        return Ms, ss

    # ---- updates -------------------------------------------------------------------------------------------
    def apply_bond(self, U, i):
        """Two-site gate on sites (i, i+1) of every chain; U is a device tensor [d,d,d,d] (shared) or
        [n,d,d,d,d] (one gate per chain).  In place."""
        lib = _lib.load()
        P, d = self.P, self.dim
        site = P * d * P
        sU = d ** 4 if U.dim() == 5 else 0
        ws, wsn = _lib.workspace(lib.mpsb_tebd_two_site_workspace(self.n, P, P, P, d))
        _lib.check(lib.mpsb_tebd_two_site(self.n, P, P, P, d, _lib.ptr(self.M[i]), site, _lib.ptr(self.M[i + 1]), site,
                                          _lib.ptr(self.S[i]), P, _lib.ptr(U), sU, self.trun, self.err, P,
                                          _lib.ptr(self.M[i]), site, _lib.ptr(self.M[i + 1]), site, _lib.ptr(self.S[i + 1]),
                                          P, _lib.ptr(self.rank[i + 1]), ws, wsn, _lib.stream_ptr()), "mpsb_tebd_two_site")

    def half_sweep(self, U, parity):
        for i in range(parity, self.L - 1, 2):
            self.apply_bond(U, i)

    def tebd(self, H, t, nsteps, order=2):
        """exp(-i H t) with `nsteps` Trotter slices; H is one two-site Hamiltonian [d^2, d^2] or one per chain
        [n, d^2, d^2].  Same even/odd schedule as MPS.iTEBD_double (`DMRG/MPS.py:463-490`); order 1, 2 or 4."""
        H = np.asarray(H)
        d = self.dim

        def gates(k):
            if H.ndim == 3:
                return _lib.to_device(np.stack([la.expm(-1j * h * k * t).reshape([d] * 4) for h in H]))
            return _lib.to_device(la.expm(-1j * H * k * t).reshape([d] * 4))

        def worker(parity):
            def make(k):
                U = gates(k)
                return lambda: self.half_sweep(U, parity)
            return make
        {1: ST.ST1, 2: ST.ST2, 4: ST.ST4}[order]((worker(0), worker(1)), nsteps)

    # ---- gauge ------------------------------------------------------------------------------------------------
    def canon(self):
        """Canonical B form of every chain: left sweep, right sweep, boundary phases divided back in - the reference's
        `MPS.canon` (DMRG/MPS.py:295-302: `ortho_left_site` :246-264 for every site, `ortho_right_site` :266-284 back,
        `unify_end` :286-293) for all n chains at once: 2 L batched gauge calls, no per-site host synchronisation (the
        ranks stay on the device).  Chains must be closed (bond dimension 1 at both ends), as `ChainBatch` builds them."""
        lib = _lib.load()
        n, P, d, L = self.n, self.P, self.dim, self.L
        site = P * d * P
        right = _lib.zeros((n, P, 1))          # the reference's dummy boundary tensors (DMRG/MPS.py:124-125), padded
        left = _lib.zeros((n, 1, P))
        right[:, 0, 0] = 1.0
        left[:, 0, 0] = 1.0
        ws, wsn = _lib.workspace(lib.mpsb_gauge_batched_workspace(n, P, P, d, d * P, P))
        for i in range(L):
            nxt, nn, sn = (self.M[i + 1], d * P, site) if i + 1 < L else (right, 1, P)
            _lib.check(lib.mpsb_gauge_left_batched(n, P, P, d, _lib.ptr(self.M[i]), site, _lib.ptr(self.S[i]), P,
                                                   _lib.ptr(nxt), sn, nn, self.trun, self.err, P, _lib.ptr(self.M[i]),
                                                   site, _lib.ptr(self.S[i + 1]), P, _lib.ptr(nxt), sn,
                                                   _lib.ptr(self.rank[i + 1]), ws, wsn, _lib.stream_ptr()),
                       "mpsb_gauge_left_batched")
        for i in reversed(range(L)):
            prv, npv, sp = (self.M[i - 1], P * d, site) if i > 0 else (left, 1, P)
            _lib.check(lib.mpsb_gauge_right_batched(n, P, P, d, _lib.ptr(self.M[i]), site, _lib.ptr(self.S[i + 1]), P,
                                                    _lib.ptr(prv), sp, npv, self.trun, self.err, P, _lib.ptr(self.M[i]),
                                                    site, _lib.ptr(self.S[i]), P, _lib.ptr(prv), sp,
                                                    _lib.ptr(self.rank[i]), ws, wsn, _lib.stream_ptr()),
                       "mpsb_gauge_right_batched")
        # unify_end: the phases parked in the boundary tensors go back into the end sites (one gate per chain)
        for tensor, i in ((left[:, 0, 0], 0), (right[:, 0, 0], L - 1)):
            ph = _lib.to_host(tensor.contiguous())
            U = _lib.to_device(np.eye(d)[None, :, :] / ph[:, None, None])
            _lib.check(lib.mpsb_one_site(n, P, P, d, _lib.ptr(self.M[i]), site, _lib.ptr(U), d * d, _lib.ptr(self.M[i]),
                                         site, _lib.stream_ptr()), "mpsb_one_site")
        self.canonical = True

    # ---- observables ---------------------------------------------------------------------------------------
    def _transfer(self, E, A, B, op):
        """One batched transfer-matrix step (DMRG/MPS.py:351-355, 384-388 for every chain): E [n, ka, kb],
        A [n, ka, d, oa], B [n, kb, d, rb], op None | [d, d] (shared) | [n, d, d] (one per chain) -> [n, oa, rb]."""
        lib = _lib.load()
        n, ka, d, oa = (int(v) for v in A.shape)
        kb, rb = int(B.shape[1]), int(B.shape[3])
        out = _lib.empty((n, oa, rb))
        ws, wsn = _lib.workspace(lib.mpsb_transfer_batched_workspace(n, ka, kb, d, oa, rb))
        op_dev, s_op = None, 0
        if op is not None:
            op = np.asarray(op, dtype=complex)
            op_dev, s_op = _lib.to_device(op), (d * d if op.ndim == 3 else 0)
        _lib.check(lib.mpsb_transfer_step_batched(n, ka, kb, d, oa, rb, _lib.ptr(E), ka * kb, _lib.ptr(A), ka * d * oa,
                                                  _lib.ptr(B), kb * d * rb, _lib.ptr(op_dev), s_op, _lib.ptr(out), oa * rb,
                                                  ws, wsn, _lib.stream_ptr()), "mpsb_transfer_step_batched")
        return out

    def _schmidt_weights(self, bond):
        """diag(S[bond]^2) per chain as a device tensor [n, P, P] (the reference starts `corr` from
        np.diag(Sl**2), DMRG/MPS.py:380)."""
        s2 = _lib.to_device(_lib.to_host(self.S[bond]) ** 2)
        E = _lib.zeros((self.n, self.P, self.P))
        E.diagonal(dim1=1, dim2=2).copy_(s2)
        return E

    @staticmethod
    def _traces(E):
        return _lib.to_host(E.diagonal(dim1=1, dim2=2).contiguous()).sum(axis=1)

    def corr(self, *ops):
        """<prod_i op_i> of every chain for (site, operator) pairs in site order: array [n] (DMRG/MPS.py:358-389 per
        chain).  An operator is [d, d] (the same for all chains) or [n, d, d]."""
        assert getattr(self, "canonical", False), "canon() first"
        D = dict(ops)
        if not D:
            return np.ones(self.n)
        start, end = min(D), max(D) + 1
        E = self._schmidt_weights(start)
        for i in range(start, end):
            E = self._transfer(E, self.M[i], self.M[i], D.get(i))
        return self._traces(E).real

    def _block(self, start, nsites):
        """Product of `nsites` consecutive site tensors of every chain: [n, P, d^nsites, P] (DMRG/MPS.py:323-337)."""
        n, P, d = self.n, self.P, self.dim
        blk, dd = self.M[start], d
        for i in range(start + 1, start + nsites):
            out = _lib.empty((n, P, dd * d, P))
            _lib.zgemm(blk, self.M[i], out, P * dd, d * P, P, P, d * P, d * P, batch1=n, sA=(P * dd * P, 0),
                       sB=(P * d * P, 0), sC=(P * dd * d * P, 0))
            blk, dd = out, dd * d
        return blk

    def measure(self, start, op, Hermitian=True):
        """Expectation value of an operator on n_sites = log_d(sqrt(op.size)) consecutive sites for every chain: array
        [n] (DMRG/MPS.py:391-410 per chain; like the reference it returns <op^T>).  `op` is
        [d^k, d^k] or one per chain [n, d^k, d^k]."""
        assert getattr(self, "canonical", False), "canon() first"
        op = np.asarray(op)
        per_chain = op.ndim == 3
        size = op[0].size if per_chain else op.size
        nsites = round(np.log(size) / np.log(self.dim) / 2)
        blk = self._block(start, nsites)
        opT = np.swapaxes(op, -1, -2)
        E = self._transfer(self._schmidt_weights(start), blk, blk, np.ascontiguousarray(opT))
        ret = self._traces(E)
        return ret.real if Hermitian else ret

    def entropies(self):
        """Von Neumann entropy (nats) of every bond of every chain: array [n, L+1]."""
        p = _lib.to_host(self.S) ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(p > 0, p * np.log(p), 0.0)
        return -t.sum(axis=2).T

    def ranks(self):
        return _lib.to_host(self.rank).T


class HostBondUpdater:
    """Two-site updates on HOST-resident (pinned) batches: the public entry point for callers whose tensors live
    in host memory (the reference keeps everything in numpy).  Device staging is double-buffered and all copies run
    on a private copy stream, so a caller that streams batches through

        upd = HostBondUpdater(n, chi_l, chi, chi_r, d, trun, err)
        upd.stage(Bi, Bj, Sl, U)                      # host -> device of batch 0 (asynchronous)
        for each batch k:
            upd.stage(<inputs of batch k+1>)          # overlaps the kernels of batch k
            upd.run(Bi_out, Bj_out, Sr_out, rank_out) # kernels of batch k, then device -> host (asynchronous)
        upd.finish()                                  # outputs of the last batch are in the host buffers

    hides the PCIe time behind the SVD.  `upd(Bi, Bj, Sl, U, Bi_out, Bj_out, Sr_out, rank_out)` is the plain
    synchronous form (stage + run + finish).  All host tensors are pinned torch CPU tensors, batch first."""

    def __init__(self, n, xl, x, xr, d, trun, err):
        t = _lib.torch()
        self.n, self.xl, self.x, self.xr, self.d = n, xl, x, xr, d
        self.trun, self.err = int(trun), float(err)
        self.kcap = max(1, min(self.trun, d * xl, d * xr))
        self.copy_stream = t.cuda.Stream()
        self.bufs = []
        for _ in range(2):
            self.bufs.append(dict(Bi=_lib.empty((n, xl, d, x)), Bj=_lib.empty((n, x, d, xr)), Sl=_lib.empty((n, xl), "f64"),
                                  U=_lib.empty((n, d, d, d, d)), Bio=_lib.empty((n, xl, d, self.kcap)),
                                  Bjo=_lib.empty((n, self.kcap, d, xr)), Sro=_lib.empty((n, self.kcap), "f64"),
                                  rk=_lib.empty((n,), "i32"), up=t.cuda.Event(), computed=t.cuda.Event(),
                                  down=t.cuda.Event()))
        self.ws_bytes = int(_lib.load().mpsb_tebd_two_site_workspace(n, xl, x, xr, d))
        self.staged = []          # buffer indices uploaded and not yet run
        self.next_buf = 0
        main = t.cuda.current_stream()
        for b in self.bufs:
            b["computed"].record(main)
            b["down"].record(main)

    def stage(self, Bi, Bj, Sl, U):
        t = _lib.torch()
        assert len(self.staged) < 2, "two batches are already staged"
        k = self.next_buf
        self.next_buf ^= 1
        buf = self.bufs[k]
        with t.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(buf["computed"])       # the kernels that read these inputs have finished
            buf["Bi"].copy_(Bi, non_blocking=True)
            buf["Bj"].copy_(Bj, non_blocking=True)
            buf["Sl"].copy_(Sl, non_blocking=True)
            buf["U"].copy_(U, non_blocking=True)
            buf["up"].record(self.copy_stream)
        self.staged.append(k)

    def run(self, Bi_out, Bj_out, Sr_out, rank_out):
        t = _lib.torch()
        lib = _lib.load()
        assert self.staged, "nothing staged"
        buf = self.bufs[self.staged.pop(0)]
        xl, x, xr, d, kcap, n = self.xl, self.x, self.xr, self.d, self.kcap, self.n
        main = t.cuda.current_stream()
        ws, wsn = _lib.workspace(self.ws_bytes)
        main.wait_event(buf["up"])
        main.wait_event(buf["down"])                             # previous outputs of this buffer have left the device
        _lib.check(lib.mpsb_tebd_two_site(n, xl, x, xr, d, _lib.ptr(buf["Bi"]), xl * d * x, _lib.ptr(buf["Bj"]), x * d * xr,
                                          _lib.ptr(buf["Sl"]), xl, _lib.ptr(buf["U"]), d ** 4, self.trun, self.err, kcap,
                                          _lib.ptr(buf["Bio"]), xl * d * kcap, _lib.ptr(buf["Bjo"]), kcap * d * xr,
                                          _lib.ptr(buf["Sro"]), kcap, _lib.ptr(buf["rk"]), ws, wsn, _lib.stream_ptr()),
                   "mpsb_tebd_two_site")
        buf["computed"].record(main)
        with t.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(buf["computed"])
            Bi_out.copy_(buf["Bio"], non_blocking=True)
            Bj_out.copy_(buf["Bjo"], non_blocking=True)
            Sr_out.copy_(buf["Sro"], non_blocking=True)
            rank_out.copy_(buf["rk"], non_blocking=True)
            buf["down"].record(self.copy_stream)

    def finish(self):
        _lib.torch().cuda.current_stream().wait_stream(self.copy_stream)

    def __call__(self, Bi, Bj, Sl, U, Bi_out, Bj_out, Sr_out, rank_out):
        self.stage(Bi, Bj, Sl, U)
        self.run(Bi_out, Bj_out, Sr_out, rank_out)
        self.finish()


def shard(n_total, rank, world):
    """Contiguous block of chain indices owned by `rank`."""
    per, rem = divmod(n_total, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def gather_observables(local, group=None):
    """All-gather a [n_local, n_obs] float64 tensor of per-chain observables over the process group
    (NCCL on GPUs, gloo in the CPU tests).  The one collective of the whole path; shards may be ragged."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c.item()) for c in counts]
    width = max(counts)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
