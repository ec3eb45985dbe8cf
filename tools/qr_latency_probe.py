"""QR phase of mpsb_svd_trunc for small batches: run once with MPSB_SVD_QR_LATENCY_BATCH=0 (blocked DMMA kernel always)
and once with the default to place the cross-over (development aid)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmrg_b200 import _lib as L
lib = L.load()
lib.mpsb_profile_enable(1)
rs = np.random.RandomState(0)
for n in (64, 128, 256):
    for batch in (1, 4, 16, 32, 64, 148):
        A = L.to_device(rs.randn(batch, n, n) + 1j * rs.randn(batch, n, n))
        k = n // 2
        vh, s, sraw, rank = L.empty((batch, k, n)), L.empty((batch, k), "f64"), L.empty((batch, n), "f64"), L.empty((batch,), "i32")
        ws, wsn = L.workspace(lib.mpsb_svd_trunc_workspace(batch, n, n, 0))
        qr = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.mpsb_svd_trunc(batch, n, n, L.ptr(A), n, n * n, k, 1e-14, 1, k, L.ptr(vh), None, L.ptr(s), L.ptr(sraw),
                                       L.ptr(rank), ws, wsn, L.stream_ptr()), "svd")
            e1.record(); torch.cuda.synchronize()
            ph = (ctypes.c_double * 6)(); r_, s_ = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
            lib.mpsb_profile_read(ph, ctypes.byref(r_), ctypes.byref(s_))
            qr.append((e0.elapsed_time(e1), ph[3]))
        tot, jac = min(q[0] for q in qr[1:]), min(q[1] for q in qr[1:])
        print(f"n={n} batch={batch}: whole call {tot:.3f} ms  jacobi {jac:.3f} ms  -> pack + QR + finalise {tot - jac:.3f} ms", flush=True)
