"""Where does a batched iDMRG growth step (ChainBatch.step) spend its time?  (development aid)
python tools/idmrg_batch_breakdown.py [n_chains] [chi] [complex]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmrg_b200 import iDMRG, _lib
from dmrg_b200.MPO import MPO
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iDMRG.REAL_PATH = "complex" not in sys.argv
gs = np.linspace(0.2, 2.0, n)
Ws = np.stack([(-g * MPO('sx') - MPO('sz') ** 2).tomatrix() for g in gs])
run = iDMRG.ChainBatch(Ws, trunc=chi)
for i in range(int(np.log2(chi)) + 3):
    run.step()
def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = f(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, out
os.environ["MPSB_LANCZOS_DEBUG"] = "1"
t_pred, guess = T(run._prediction)
t_gs, (E, theta) = T(lambda: iDMRG.ground_state_batched(run.L, run.W, run.R, seed=1, theta0=guess))
t_sv, (U, S, V) = T(lambda: iDMRG.halve_batched(theta, chi))
t_env, _ = T(lambda: (iDMRG._env_batched(True, run.L, U.contiguous(), run.W, False), iDMRG._env_batched(False, run.R, V.contiguous(), run.W, False)))
t_all, _ = T(run.step)
print(f"n={n} chi={chi} real={run.real}: predict {t_pred:.1f}  lanczos {t_gs:.1f} ({iDMRG.last_info['n_matvec']} matvecs)  "
      f"halve {t_sv:.1f}  env {t_env:.1f}  | whole step {t_all:.1f} ms")
