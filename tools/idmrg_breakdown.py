"""Where does one iDMRG growth step spend its time?  (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmrg_b200 import iDMRG, _lib
from dmrg_b200.MPO import MPO
chi = int(sys.argv[1])
if len(sys.argv) > 2 and sys.argv[2] == "heis":
    from dmrg_b200 import spin
    W = (2 * MPO('P', op=spin.P()) * MPO('N', op=spin.N()) + 2 * MPO('N', op=spin.N()) * MPO('P', op=spin.P()) + MPO('sz') ** 2).tomatrix()
else:
    W = (-MPO('sx') - MPO('sz') ** 2).tomatrix()
w = W.shape[0]
L0 = np.zeros((1, 1, w)); L0[0, 0, w - 1] = 1
R0 = np.zeros((1, 1, w)); R0[0, 0, 0] = 1
real = 'complex' not in sys.argv
Ld, Rd, Wd = iDMRG._to_device_auto(real, L0, R0, W)
for i in range(max(14, int(np.log2(chi)) + 3)):
    Ld, Rd, E, S = iDMRG.next_LR_device(Ld, Rd, Wd, trunc=chi)
def T(f, reps=2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, out
t_gs, (E, theta) = T(lambda: iDMRG.ground_state(Ld, Wd, Rd, seed=0))
sh = tuple(theta.shape)
t_sp, (U, S, V) = T(lambda: iDMRG._halve_split(theta, sh, chi))
t_au = float("nan")
if chi <= 256:
    t_au, _ = T(lambda: iDMRG.halve(theta, sh, chi))
lib0 = _lib.load()
lib0.mpsb_profile_enable(1)
iDMRG._halve_split(theta, sh, chi)
import ctypes
ph = (ctypes.c_double * 6)(); rot = ctypes.c_ulonglong(); swp = ctypes.c_ulonglong()
lib0.mpsb_profile_read(ph, ctypes.byref(rot), ctypes.byref(swp))
lib0.mpsb_profile_enable(0)
print("last SVD of the split: jacobi %.1f ms, rotations %d, sweeps %d" % (ph[3], rot.value, swp.value))
t_env, _ = T(lambda: (iDMRG._env(True, Ld, U.contiguous(), Wd, False), iDMRG._env(False, Rd, V.contiguous(), Wd, False)))
lib = _lib.load()
print(f"chi={chi}: lanczos {t_gs:.1f} ms ({iDMRG.last_info['n_matvec']} matvecs), halve split {t_sp:.1f} ms, halve augmented {t_au:.1f} ms "
      f"(sweeps {lib.mpsb_last_svd_sweeps()}), env {t_env:.2f} ms")
