"""Where does one iDMRG growth step spend its time?  (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmrg_b200 import iDMRG, _lib
from dmrg_b200.MPO import MPO
chi = int(sys.argv[1])
W = (-MPO('sx') - MPO('sz') ** 2).tomatrix()
Ld, Rd, Wd = _lib.to_device(np.array([[[0, 0, 1]]])), _lib.to_device(np.array([[[1, 0, 0]]])), _lib.to_device(W)
for i in range(14):
    Ld, Rd, E, S = iDMRG.next_LR_device(Ld, Rd, Wd, trunc=chi)
def T(f, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, out
t_gs, (E, theta) = T(lambda: iDMRG.ground_state(Ld, Wd, Rd, seed=0))
sh = tuple(theta.shape)
t_sp, (U, S, V) = T(lambda: iDMRG._halve_split(theta, sh, chi))
t_au, _ = T(lambda: iDMRG.halve(theta, sh, chi))
t_env, _ = T(lambda: (iDMRG._env(True, Ld, U.contiguous(), Wd, False), iDMRG._env(False, Rd, V.contiguous(), Wd, False)))
lib = _lib.load()
print(f"chi={chi}: lanczos {t_gs:.1f} ms ({iDMRG.last_info['n_matvec']} matvecs), halve split {t_sp:.1f} ms, halve augmented {t_au:.1f} ms "
      f"(sweeps {lib.mpsb_last_svd_sweeps()}), env {t_env:.2f} ms")
