"""Timing of the single-chain API (canon, update_double, corr) vs the CPU oracle (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.linalg as la
from dmrg_b200.MPS import MPS
from dmrg_b200 import _lib
from oracle.mps_oracle import OracleMPS
L, d = 16, 2
chis = [min(2 ** i, 2 ** (L - i), 128) for i in range(L + 1)]
rs = np.random.RandomState(1000)
Ms = [rs.randn(chis[i], d, chis[i + 1]) + 1j * rs.randn(chis[i], d, chis[i + 1]) for i in range(L)]
A = rs.randn(4, 4) + 1j * rs.randn(4, 4); U = la.expm(-0.05j * (A + A.conj().T)).reshape(2, 2, 2, 2)
def T(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
def gpu_canon():
    s = MPS([m.copy() for m in Ms], d, trun=128, err=1e-14); s.canon(); return s
def cpu_canon():
    o = OracleMPS([m.copy() for m in Ms], trun=128, err=1e-14); o.canon(); return o
l0 = _lib.launch_count(); s = gpu_canon(); nl = _lib.launch_count() - l0
print(f"canon L=16 chi<=128: GPU {T(gpu_canon):.1f} ms ({nl} launches), CPU oracle {T(cpu_canon, 1):.1f} ms")
o = cpu_canon()
print(f"update_double chi=128 bond: GPU {T(lambda: s.copy().update_double(U, 7)):.1f} ms (incl. copy), CPU oracle {T(lambda: o.copy().update_double(U, 7), 1):.1f} ms")
Z = np.diag([1.0, -1.0])
print(f"corr 2-point: GPU {T(lambda: s.corr((2, Z), (12, Z))):.2f} ms, CPU {T(lambda: o.corr((2, Z), (12, Z)), 1):.2f} ms")
print(f"dot: GPU {T(lambda: s.dot(s)):.2f} ms, CPU {T(lambda: o.dot(o), 1):.2f} ms")
