"""End-to-end TEBD probe (BASELINE config 3 shape): TFIM quench of |up...up>, L sites, trun chi, n Trotter steps of
dt through MPS.iTEBD_double; prints wall time per Trotter step, bond dimensions and the mid-chain entropy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dmrg_b200.MPS import MPS
from dmrg_b200.spin import sigma
from dmrg_b200 import _lib

L, chi, nsteps, dt = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
order = int(sys.argv[5]) if len(sys.argv) > 5 else 2
Z, X, I2 = sigma[3], sigma[1], np.eye(2)
Hb = -np.kron(Z, Z) - 0.5 * (np.kron(X, I2) + np.kron(I2, X))
s = MPS.naive([[1, 0]] * L)
s.trun, s.err = chi, 1e-10
s.canon()
block = max(1, nsteps // 10)
done = 0
while done < nsteps:
    k = min(block, nsteps - done)
    torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = _lib.launch_count()
    s.iTEBD_double(Hb, k * dt, k, order=order)
    torch.cuda.synchronize(); el = time.perf_counter() - t0
    done += k
    sm = np.asarray(s.s[L // 2], dtype=float)
    ent = -np.sum(sm ** 2 * np.log(sm ** 2))
    nb = (2 * k + 1) * (L - 1) / 2
    print(f"t={done*dt:.2f} steps {done}: {el/k*1e3:.1f} ms/Trotter step, {nb/el:.0f} bond updates/s, max chi {int(s.x.max())}, "
          f"S_mid {ent:.6f}, launches/step {(_lib.launch_count()-l0)/k:.0f}", flush=True)
