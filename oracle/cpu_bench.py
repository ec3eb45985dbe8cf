"""CPU timing leg of bench.py — TEST/BENCH INFRASTRUCTURE (imports the oracle; never used by the product).

Times `update_double` (DMRG/MPS.py:428-457) on the same synthetic bond shape the GPU arm uses, one worker process
per host core with OMP_NUM_THREADS=1, in two flavours:
  kind "reference": the UNMODIFIED reference class `DMRG.MPS.MPS`, imported from oracle/_ref/DMRG - a git-ignored
                    copy of /root/reference/DMRG that `__graft_entry__.build()` stages where the reference tree is
                    visible (this container); it travels to the GPU box like the built .so files do;
  kind "port":      the oracle's restatement (numpy tensordot + LAPACK zgesdd instead of the reference's 3-operand
                    einsum, ~7x faster) - the fallback when oracle/_ref is absent, and reported beside the reference.
"""
import os
import time

import numpy as np


def make_bond(seed, chi, d):
    """Synthetic saturated bond: right-isometric B tensors (QR of complex Gaussians), Schmidt vector from the
    normalised spectrum of a Gaussian matrix, GUE gate exp(-0.05i H).  Same recipe as bench.py's GPU inputs."""
    import scipy.linalg as la
    rs = np.random.RandomState(seed)

    def iso():
        q, _ = np.linalg.qr(rs.randn(d * chi, chi) + 1j * rs.randn(d * chi, chi))
        return np.ascontiguousarray(q.conj().T).reshape(chi, d, chi)
    Bi, Bj = iso(), iso()
    s = np.linalg.svd(rs.randn(chi, chi) + 1j * rs.randn(chi, chi), compute_uv=False)
    Sl = s / np.linalg.norm(s)
    A = rs.randn(d * d, d * d) + 1j * rs.randn(d * d, d * d)
    U = la.expm(-0.05j * (A + A.conj().T)).reshape([d] * 4)
    return Bi, Bj, Sl, U


def worker(args):
    seed, n_updates, chi, d, trun, err = args
    from oracle.mps_oracle import OracleMPS
    Bi, Bj, Sl, U = make_bond(seed, chi, d)
    o = OracleMPS([Bi, Bj], trun=trun, err=err)
    o.s[0] = Sl
    o.update_double(U, 0)                      # warm-up (BLAS/LAPACK initialisation)
    t0 = time.perf_counter()
    for _ in range(n_updates):
        o = OracleMPS([Bi, Bj], trun=trun, err=err)
        o.s[0] = Sl
        o.update_double(U, 0)
    return time.perf_counter() - t0, int(o.x[1])


REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "DMRG", "MPS.py"))


def worker_reference(args):
    """The unmodified reference: DMRG.MPS.MPS([Bi, Bj], trun=, err=).update_double(U, 0)."""
    import sys
    seed, n_updates, chi, d, trun, err = args
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from DMRG.MPS import MPS as RefMPS
    Bi, Bj, Sl, U = make_bond(seed, chi, d)

    def fresh():
        o = RefMPS([Bi, Bj], trun=trun, err=err)
        o.s[0] = Sl
        return o
    fresh().update_double(U, 0)                # warm-up
    t0 = time.perf_counter()
    for _ in range(n_updates):
        o = fresh()
        o.update_double(U, 0)
    return time.perf_counter() - t0, int(o.x[1])


def run(n_per_worker, chi, d, trun, err, cores=None, kind="port"):
    """Returns (updates_per_second over all workers, cores, total updates, wall seconds).  The rate is total updates
    over the busiest worker's own loop time (pool start-up and the warm-up update are not counted)."""
    import multiprocessing as mp
    fn = worker_reference if kind == "reference" else worker
    cores = cores or len(os.sched_getaffinity(0))
    old = os.environ.get("OMP_NUM_THREADS")
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(cores) as pool:
            t0 = time.perf_counter()
            res = pool.map(fn, [(5000 + w, n_per_worker, chi, d, trun, err) for w in range(cores)])
            wall = time.perf_counter() - t0
    finally:
        if old is None:
            os.environ.pop("OMP_NUM_THREADS", None)
        else:
            os.environ["OMP_NUM_THREADS"] = old
    busiest = max(r[0] for r in res)
    total = n_per_worker * cores
    return total / busiest, cores, total, wall
