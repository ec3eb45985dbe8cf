"""Latency/throughput of mpsb_svd_trunc on small matrices; run once per setting of MPSB_SVD_RESIDENT_BATCH."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmrg_b200 import _lib

lib = _lib.load()
rs = np.random.RandomState(0)
print("MPSB_SVD_RESIDENT_BATCH =", os.environ.get("MPSB_SVD_RESIDENT_BATCH", "(default)"))
for n, L in [(16, 16), (32, 32), (64, 64), (100, 100), (100, 200), (50, 250)]:
    for batch in [1, 16, 148, 296, 1024, 4096]:
        if n * L * batch > 5e7:
            continue
        a = rs.randn(batch, n, L) + 1j * rs.randn(batch, n, L)
        a *= np.exp(-0.3 * np.arange(L))[None, None, :]                     # decaying spectrum like a Schmidt matrix
        ad = _lib.to_device(a)
        k = min(n, L)
        vh, s, rank = _lib.empty((batch, k, L)), _lib.empty((batch, k), "f64"), _lib.empty((batch,), "i32")
        ws, wsn = _lib.workspace(lib.mpsb_svd_trunc_workspace(batch, n, L, 0))

        def call():
            _lib.check(lib.mpsb_svd_trunc(batch, n, L, _lib.ptr(ad), L, n * L, k, 0.0, 1, k, _lib.ptr(vh), None,
                                          _lib.ptr(s), None, _lib.ptr(rank), ws, wsn, _lib.stream_ptr()), "svd")
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        sv = _lib.to_host(s)[0]
        ref = np.linalg.svd(a[0], compute_uv=False)
        ref = ref / np.linalg.norm(ref)
        err = np.max(np.abs(sv - ref) / ref[0])
        print("n=%3d L=%3d batch=%4d  %8.3f ms/call  %8.1f us/problem  sweeps=%d  err=%.1e" %
              (n, L, batch, dt * 1e3, dt * 1e6 / batch, lib.mpsb_last_svd_sweeps(), err))
