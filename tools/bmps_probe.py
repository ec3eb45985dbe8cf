"""Times the reference's own d = 10 workload (qcircuit-prog/naive.py:27-43: BMPS(10, L, H, pack), evolve(T/N, 60))
on the GPU path and on the CPU oracle.  Usage: python tools/bmps_probe.py [--cpu]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

H = {"omega0": 0, "g": 1, "u": 0, "h": 0, "K": 1}
pack = {"dk": 0.25, "center": 10, "k_c": -np.pi / 2, "trun": True, "alpha": 2}
T = np.pi
L = 2 * pack["center"] + int(2 * T * H["g"]) + 1
N = 2


def run(cls, sync=lambda: None):
    psi = cls(10, L, H, pack)
    occ = []
    sync()
    t0 = time.perf_counter()
    for _ in range(N):
        psi.evolve(T / N, 60)
        psi.canon()
        occ.append([psi.measure(i, psi.N) for i in range(psi.l)])
    sync()
    return time.perf_counter() - t0, np.array(occ), [int(v) for v in psi.x]


if __name__ == "__main__":
    if "--cpu" in sys.argv:
        from oracle.bmps_oracle import OracleBMPS
        dt, occ, x = run(OracleBMPS)
        print("cpu oracle: %.2f s for %d x evolve(T/%d, 60), L=%d, bonds %s" % (dt, N, N, L, x))
        np.save("/tmp/bmps_probe_cpu.npy", occ)
    else:
        import torch
        from dmrg_b200 import _lib
        from dmrg_b200.transmit.bhopping import BMPS
        run(BMPS, torch.cuda.synchronize)
        n0 = _lib.load().mpsb_launch_count()
        dt, occ, x = run(BMPS, torch.cuda.synchronize)
        print("gpu: %.2f s for %d x evolve(T/%d, 60), L=%d, bonds %s, launches %d" % (dt, N, N, L, x, _lib.load().mpsb_launch_count() - n0))
        if "--gpu-only" in sys.argv:
            sys.exit(0)
        from oracle.bmps_oracle import OracleBMPS
        dtc, occ_c, xc = run(OracleBMPS)
        print("cpu oracle (1 process): %.2f s, bonds %s; max |d<N>| = %.2e" % (dtc, xc, abs(occ - occ_c).max()))
