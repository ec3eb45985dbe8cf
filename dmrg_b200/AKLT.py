"""AKLT model helpers with the names of the reference's `DMRG/AKLT.py` — the strongest known-answer
test of the MPS path: exact chi = 2 ground state, E = -2/3 per bond, string order.

Everything numerical goes through `MPS` (GPU kernels); this file only assembles operators.
"""
from functools import reduce

import numpy as np
import scipy.linalg as la

from . import spin
from .MPS import MPS

# two-site AKLT Hamiltonian  S.S + (S.S)^2 / 3  for spin 1 (`DMRG/AKLT.py:12-13`)
ss = sum(np.kron(a, a) for a in spin.S(1)[1:])
H = ss + ss @ ss / 3

sz1 = spin.Z(1)


def AKLT_State(n=30):
    """Exact AKLT matrix product state on n sites with the boundary spins fixed (`DMRG/AKLT.py:16-26`)."""
    site = 2 * np.stack([np.sqrt(2 / 3) * spin.P(), -np.sqrt(1 / 3) * spin.Z() * 2, -np.sqrt(2 / 3) * spin.N()], axis=1)
    tensors = [site.copy() for _ in range(n)]
    tensors[0] = tensors[0][:1]
    tensors[-1] = tensors[-1][:, :, 1:]
    s = MPS(tensors, 3)
    s.canon()
    return s


def Energy(s):
    """Sum of bond energies <H_{i,i+1}> (`DMRG/AKLT.py:29-30`)."""
    return sum(s.measure(i, H) for i in range(s.L - 1)).real


def zz_op(k, i):
    """Operator list for <Sz_k Sz_i> (`DMRG/AKLT.py:36-37`)."""
    return ((k, sz1), (i, sz1))


def string_op(l, r):
    """Operator list of the string order parameter Sz_l exp(i pi sum Sz) Sz_r (`DMRG/AKLT.py:40-42`)."""
    flip = la.expm(np.pi * 1j * sz1)
    return ((l, sz1),) + tuple((i, flip) for i in range(l + 1, r)) + ((r, sz1),)


def oneop(ops):
    """Fuse a contiguous operator list into (first site, kron of the operators) (`DMRG/AKLT.py:45-49`)."""
    return ops[0][0], reduce(np.kron, [o[1] for o in ops])


def evolve(s, time=0.1, n=5, k=40):
    """Evolve a copy of s under H for n periods of `time` (k Trotter slices each) and return the overlaps
    <s|p(t)>; asserts |<s|p>| stays constant, i.e. that s is an eigenstate (`DMRG/AKLT.py:52-63`)."""
    p = s.copy()
    overlaps = []
    start = np.abs(s.dot(p))
    for i in range(n):
        p.iTEBD_double(H, time, k)
        s.canon()
        ov = s.dot(p)
        print('Time {:.3f}, Overlap {:.5f}*exp({:.5f}j)'.format((i + 1) * time, np.abs(ov), np.angle(ov)))
        assert abs(start - np.abs(ov)) < 1e-6, "Not eigenstate?"
        overlaps.append(ov)
    return overlaps


def Z(s):
    """<Sz_i> profile (`DMRG/AKLT.py:66-67`)."""
    return np.array([s.corr((i, spin.Z(1))) for i in range(s.L)]).real


def Ztj(s, t, n=50):
    """<Sz_i>(t_j) for increasing times t (`DMRG/AKLT.py:70-79`)."""
    out = []
    for dt in np.diff([0, *t]):
        print(dt)
        s.iTEBD_double(H, dt, n)
        out.append(Z(s))
    return np.array(out).real


def correlator(s):
    """Print ZZ and string correlators and check corr == measure on short strings (`DMRG/AKLT.py:82-93`)."""
    for i in range(0, s.L - 2):
        zcorr = np.array([s.corr(*zz_op(i, j)) for j in range(i + 1, s.L)]).real
        print('Z correlation starting at {}:\n{}\n>>> Ratio: {}'.format(i, zcorr, zcorr[1:] / zcorr[:-1]))
        string = np.array([s.corr(*string_op(i, j)) for j in range(i + 1, s.L)]).real
        print('String Corr: \n{}'.format(string))
        for j in range(i + 1, min(i + 6, s.L - 1)):
            ops = string_op(i, j)
            assert np.abs(s.corr(*ops) - s.measure(*oneop(ops))) < 1e-13
