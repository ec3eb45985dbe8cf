import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dmrg_b200 import _lib
from dmrg_b200.MPS import _device_svd
rs = np.random.RandomState(11)
n = 96
def crand(*s): return rs.randn(*s) + 1j * rs.randn(*s)
G = crand(n, n)
for name, expo in [("to55", [0, 0, 3, 3, 20, 40, 55]), ("to100", [0, 0, 3, 3, 20, 40, 55, 70, 100]), ("to130", [0, 0, 3, 3, 20, 40, 55, 70, 100, 130]),
                   ("to160", [0, 0, 3, 3, 20, 40, 55, 70, 100, 130, 160]), ("to250", [0, 0, 3, 3, 20, 40, 55, 70, 100, 130, 160, 200, 250])]:
    scale = np.zeros(n); scale[:len(expo)] = 10.0 ** (-np.array(expo, dtype=float))
    m = scale[:, None] * G
    u, s_all, vh, k = _device_svd(_lib.to_device(m), n, n, 64, 0.0, want_u=False, normalise=0)
    print(name, "sweeps", _lib.load().mpsb_last_svd_sweeps(), _lib.to_host(s_all)[:len(expo)])
