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

    # ---- observables ---------------------------------------------------------------------------------------
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
