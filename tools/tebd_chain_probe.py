"""Where a Trotter step of ONE saturated chain spends its time (bench.py's tebd_chain leg with the library's phase
profile and the sweep trace switched on): python tools/tebd_chain_probe.py"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g

g.build()
from dmrg_b200 import _lib as L
from dmrg_b200.MPS import MPS
from dmrg_b200.spin import sigma

lib = L.load()
rs = np.random.RandomState(7)
Lc, CHI, D = 128, 128, 2
x = [min(2 ** i, 2 ** (Lc - i), CHI) for i in range(Lc + 1)]
s = MPS([rs.randn(x[i], D, x[i + 1]) + 1j * rs.randn(x[i], D, x[i + 1]) for i in range(Lc)], trun=CHI, err=1e-14)
s.canon()
Z, X, I2 = sigma[3], sigma[1], np.eye(2)
Hb = -np.kron(Z, Z) - 0.5 * (np.kron(X, I2) + np.kron(I2, X))
s.iTEBD_double(Hb, 0.05, 1)
lib.mpsb_profile_enable(1)
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = L.launch_count()
    s.iTEBD_double(Hb, 0.05, 1)
    torch.cuda.synchronize(); el = time.perf_counter() - t0
    ph = (ctypes.c_double * 6)()
    r_, s_ = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    lib.mpsb_profile_read(ph, ctypes.byref(r_), ctypes.byref(s_))
    print(f"Trotter step {k}: {el * 1e3:.1f} ms for 3 half sweeps, {L.launch_count() - l0} launches; last bulk call: phases (ms) "
          f"theta {ph[0]:.2f} gate {ph[1]:.2f} qr {ph[2]:.2f} jacobi {ph[3]:.2f} finalise {ph[4]:.2f} backproj {ph[5]:.2f}, "
          f"sweeps of slowest problem {lib.mpsb_last_svd_sweeps()}, rotations {r_.value}, problem-sweeps {s_.value}", flush=True)
