"""Phase breakdown (ms) of one mpsb_tebd_two_site call at small bond dimension / large local dimension."""
import ctypes
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
for chi, d, batch in [(10, 10, 1), (10, 10, 13), (32, 2, 1), (32, 2, 64), (64, 2, 1), (64, 3, 1), (128, 2, 1), (128, 2, 64), (128, 2, 148), (128, 2, 160)]:
    bi = _lib.to_device(rs.randn(batch, chi, d, chi) + 1j * rs.randn(batch, chi, d, chi))
    bj = _lib.to_device(rs.randn(batch, chi, d, chi) + 1j * rs.randn(batch, chi, d, chi))
    sl = _lib.to_device(np.abs(rs.randn(batch, chi)) + 0.1, np.float64)
    A = rs.randn(d * d, d * d) + 1j * rs.randn(d * d, d * d)
    import scipy.linalg as la
    U = _lib.to_device(la.expm(-0.05j * (A + A.conj().T)).reshape([d] * 4))
    kcap = chi
    bio, bjo = _lib.empty((batch, chi, d, kcap)), _lib.empty((batch, kcap, d, chi))
    sro, rank = _lib.empty((batch, kcap), "f64"), _lib.empty((batch,), "i32")
    ws, wsn = _lib.workspace(lib.mpsb_tebd_two_site_workspace(batch, chi, chi, chi, d))

    def call():
        _lib.check(lib.mpsb_tebd_two_site(batch, chi, chi, chi, d, _lib.ptr(bi), chi * d * chi, _lib.ptr(bj), chi * d * chi,
                                          _lib.ptr(sl), chi, _lib.ptr(U), 0, chi, 1e-12, kcap, _lib.ptr(bio), chi * d * kcap,
                                          _lib.ptr(bjo), kcap * d * chi, _lib.ptr(sro), kcap, _lib.ptr(rank), ws, wsn,
                                          _lib.stream_ptr()), "tebd")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    n0 = lib.mpsb_launch_count()
    t0 = time.perf_counter()
    for _ in range(10):
        call()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    nl = (lib.mpsb_launch_count() - n0) / 10
    lib.mpsb_profile_enable(1)
    call()
    torch.cuda.synchronize()
    ph = (ctypes.c_double * 6)()
    rot, sw = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    lib.mpsb_profile_read(ph, ctypes.byref(rot), ctypes.byref(sw))
    lib.mpsb_profile_enable(0)
    print("chi=%3d d=%2d batch=%3d: %.3f ms/call, %5.0f launches | gemm %.3f gate %.3f qr %.3f jacobi %.3f fin %.3f gemm %.3f | sweeps %d"
          % (chi, d, batch, dt * 1e3, nl, *list(ph), lib.mpsb_last_svd_sweeps()))
