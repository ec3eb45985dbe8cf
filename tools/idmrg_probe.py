"""Timing probe for the iDMRG growth step on the GPU (development aid): TFIM / AKLT / Heisenberg at several chi."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dmrg_b200 import iDMRG, _lib, spin
from dmrg_b200.MPO import MPO, two_site_mpo

def run(name, W, chi, steps, exact=None):
    w = W.shape[0]
    L = np.zeros((1, 1, w), dtype=complex); R = np.zeros((1, 1, w), dtype=complex)
    L[0, 0, w - 1] = 1; R[0, 0, 0] = 1
    Ld, Rd, Wd = _lib.to_device(L), _lib.to_device(R), _lib.to_device(W)
    Eprev = 0.0
    for i in range(steps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Ld, Rd, E, S = iDMRG.next_LR_device(Ld, Rd, Wd, trunc=chi)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ebond = (E - Eprev) / 2; Eprev = E
        if i >= steps - 3 or i % 10 == 0:
            print(f"{name} chi={chi} step {i}: bond dim {Ld.shape[0]} E/site {E/(2*i+2):.12f} e_bond {ebond:.12f} "
                  f"matvecs {iDMRG.last_info['n_matvec']} resid {iDMRG.last_info['residual']:.1e} time {dt*1e3:.1f} ms", flush=True)
    if exact is not None:
        print(f"   exact e_inf = {exact:.12f}")
    return _lib.to_host(S)

which = sys.argv[1]
if which == "tfim":
    W = (-MPO('sx') - MPO('sz') ** 2).tomatrix()
    for chi in (32, 128):
        run("TFIM", W, chi, int(sys.argv[2]), -4 / np.pi)
elif which == "aklt":
    ss = sum(np.kron(a, a) for a in spin.S(1)[1:]); H2 = (ss + ss @ ss / 3).real
    S = run("AKLT", two_site_mpo(H2, 3).astype(complex), 64, int(sys.argv[2]), -2 / 3)
    p = S[:8] ** 2 / np.sum(S ** 2)
    print("   top Schmidt weights:", p)
elif which == "heis":
    Wh = (2 * MPO('P', op=spin.P()) * MPO('N', op=spin.N()) + 2 * MPO('N', op=spin.N()) * MPO('P', op=spin.P()) + MPO('sz') ** 2).tomatrix()
    for chi in [int(c) for c in sys.argv[3:]]:
        run("Heisenberg", Wh, chi, int(sys.argv[2]), 4 * (0.25 - np.log(2)))
elif which == "chain":                       # python tools/idmrg_probe.py chain <tfim|heis> <steps> <chi> [nopredict]
    model, steps, chi = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    if model == "heis":
        W = (2 * MPO('P', op=spin.P()) * MPO('N', op=spin.N()) + 2 * MPO('N', op=spin.N()) * MPO('P', op=spin.P()) + MPO('sz') ** 2).tomatrix()
    else:
        W = (-MPO('sx') - MPO('sz') ** 2).tomatrix()
    ch = iDMRG.Chain(W, trunc=chi, predict="nopredict" not in sys.argv)
    for i in range(steps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ch.step()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if i >= steps - 4 or i % 8 == 0:
            fid = ch.fidelity[-1] if ch.fidelity else float('nan')
            print(f"{model} chi={chi} step {i}: bond {ch.A.shape[2]} e_bond {ch.energy_per_site():.12f} matvecs {ch.n_matvec[-1]} "
                  f"1-F {1 - fid:.2e} time {dt*1e3:.1f} ms", flush=True)
