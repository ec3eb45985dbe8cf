"""Pins the oracle to the UNMODIFIED reference and writes the golden vectors under tests/golden/.

Run in the build container only (`/root/reference` does not exist on the GPU box):

    python oracle/validate_against_reference.py

For each case the same inputs go through the reference (`DMRG.MPS.MPS`, `DMRG.iDMRG.next_LR`,
`DMRG.AKLT`, imported from /root/reference) and through `oracle/`; the script asserts agreement on
gauge-invariant quantities and stores inputs + reference outputs as .npz fixtures.  TEST
INFRASTRUCTURE — nothing in dmrg_b200/ imports it.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import scipy
import scipy.linalg as la

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from DMRG import AKLT as ref_aklt          # noqa: E402
from DMRG import iDMRG as ref_idmrg        # noqa: E402
from DMRG import Ising as ref_ising        # noqa: E402
from DMRG.MPO import MPO as RefMPO         # noqa: E402
from DMRG.MPS import MPS as RefMPS         # noqa: E402
from DMRG.spin import sigma                # noqa: E402

from oracle import idmrg_oracle as io_     # noqa: E402
from oracle.mps_oracle import OracleMPS    # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def random_mps(seed, chis, d=2):
    rs = np.random.RandomState(seed)
    return [rs.randn(chis[i], d, chis[i + 1]) + 1j * rs.randn(chis[i], d, chis[i + 1]) for i in range(len(chis) - 1)]


def gue_gate(rs, d=2, dt=0.05):
    A = rs.randn(d * d, d * d) + 1j * rs.randn(d * d, d * d)
    H = A + A.conj().T
    return la.expm(-1j * dt * H).reshape([d] * 4), H


def close(a, b, tol, what):
    err = np.max(np.abs(np.asarray(a) - np.asarray(b))) if np.size(a) else 0.0
    assert err <= tol, f"{what}: max abs diff {err:.3e} > {tol:.1e}"
    return err


def case_update_double():
    """Random canonical L=6 chain, one gate on every bond (both paths start from the reference's canon)."""
    chis = [1, 2, 4, 8, 4, 2, 1]
    Ms = random_mps(7, chis)
    ref = RefMPS([m.copy() for m in Ms], 2, trun=8, err=1e-12)
    ref.canon()
    ora = OracleMPS([m.copy() for m in Ms], trun=8, err=1e-12)
    ora.canon()
    for b in range(len(chis)):
        close(ref.s[b], ora.s[b], 1e-13, f"canon s[{b}]")
    rs = np.random.RandomState(11)
    U, _ = gue_gate(rs)
    out = {"U": U, "trun": 8, "err": 1e-12}
    for i in range(ref.L):
        out[f"M_in_{i}"] = ref.M[i].copy()
    for b in range(ref.L + 1):
        out[f"s_in_{b}"] = np.asarray(ref.s[b], dtype=float)
    for i in range(ref.L - 1):
        r2, o2 = ref.copy(), ora.copy()
        r2.update_double(U, i)
        o2.update_double(U, i)
        close(r2.Sr[i], o2.s[i + 1], 1e-13, f"update_double Sr bond {i}")
        th_r = np.tensordot(r2.M[i], r2.M[i + 1], axes=([2], [0]))
        th_o = np.tensordot(o2.M[i], o2.M[i + 1], axes=([2], [0]))
        close(th_r, th_o, 1e-12, f"update_double theta bond {i}")
        out[f"Sr_out_{i}"] = np.asarray(r2.Sr[i], dtype=float)
        out[f"theta_out_{i}"] = th_r
        out[f"k_out_{i}"] = int(r2.xr[i])
    np.savez(os.path.join(GOLD, "update_double_L6.npz"), **out)
    print("update_double L=6: oracle == reference")


def case_canon_testcanon():
    """The reference's own testCanon input (DMRG/MPS_test.py:10-16)."""
    ref = RefMPS((1, 2, 2))
    ref.M[0][0, :, :] = np.array([[1, 2 + 1j], [9j, 6j]])
    ref.M[1][:, :, 0] = np.array([[-1j, 2 - 3j], [0.3, 4]])
    ref.M[1][:, :, 1] = np.array([[5 - 1j, 2 - 3j], [0.3, 4]])
    Min = [ref.M[0].copy(), ref.M[1].copy()]
    ref.canon()
    ora = OracleMPS(Min, trun=20, err=1e-6)
    ora.canon()
    for b in range(3):
        close(ref.s[b], ora.s[b], 1e-14, f"testCanon s[{b}]")
    close(ref.dot(ref), ora.dot(ora), 1e-14, "testCanon norm")
    np.savez(os.path.join(GOLD, "canon_testcanon.npz"), M0=Min[0], M1=Min[1], s0=ref.s[0], s1=ref.s[1], s2=ref.s[2],
             norm=ref.dot(ref))
    print("testCanon input: oracle == reference")


def case_two_body():
    """DMRG/MPS_test.py:30-50: ZZ rotation phases."""
    ref = RefMPS.naive([1 + 3j, 1], [1 - 8j, 1 - 2j])
    ref.canon()
    ora = OracleMPS.product([[1 + 3j, 1], [1 - 8j, 1 - 2j]], trun=20, err=1e-6)
    ora.canon()
    n = 20
    H = np.kron(sigma[3], sigma[3])
    U = la.expm(np.pi / n * 1j * H).reshape([2] * 4)
    amps = []
    for _ in range(n):
        ref.update_double(U, 0)
        ora.update_double(U, 0)
        a_r = np.array([ref[e] for e in ([0, 0], [0, 1], [1, 0], [1, 1])])
        a_o = np.array([OracleMPS.product(e).dot(ora) for e in ([0, 0], [0, 1], [1, 0], [1, 1])])
        close(a_r, a_o, 1e-13, "two-body amplitudes")
        amps.append(a_r)
    zz = ref.corr((0, sigma[3]), (1, sigma[3]))
    close(zz, ora.corr((0, sigma[3]), (1, sigma[3])), 1e-13, "two-body <ZZ>")
    np.savez(os.path.join(GOLD, "two_body.npz"), amps=np.array(amps), zz=zz)
    print("testTwoBody flow: oracle == reference")


def case_aklt():
    """DMRG/MPS_test.py:52-64 / DMRG/AKLT.py: exact AKLT state, energy, TEBD phases, <Sz>, correlators."""
    N = 11
    ref = quiet(ref_aklt.AKLT_State, N)
    Ms = [m.copy() for m in ref.M[:N]]
    # rebuild the un-canonicalised input exactly as AKLT_State does (DMRG/AKLT.py:16-24)
    from DMRG import spin
    A = 2 * np.array([np.sqrt(2 / 3) * spin.P(), -np.sqrt(1 / 3) * spin.Z() * 2, -np.sqrt(2 / 3) * spin.N()]).transpose([1, 0, 2])
    raw = [A.copy() for _ in range(N)]
    raw[0] = raw[0][:1]
    raw[-1] = raw[-1][:, :, 1:]
    ora = OracleMPS(raw, trun=20, err=1e-6)
    ora.canon()
    for b in range(N + 1):
        close(ref.s[b], ora.s[b], 1e-13, f"AKLT s[{b}]")
    E_ref = ref_aklt.Energy(ref)
    E_ora = sum(ora.measure(i, ref_aklt.H) for i in range(N - 1)).real
    close(E_ref, E_ora, 1e-12, "AKLT energy")
    sz_ref = quiet(ref_aklt.Z, ref)
    sz_ora = np.array([ora.corr((i, spin.Z(1))) for i in range(N)])
    close(sz_ref, sz_ora, 1e-12, "AKLT <Sz>")
    ov_ref = quiet(ref_aklt.evolve, ref, n=3, time=1, k=10)
    p = ora.copy()
    ov_ora = []
    for _ in range(3):
        p.tebd(ref_aklt.H, 1, 10)
        ov_ora.append(ora.dot(p))
    close(ov_ref, ov_ora, 1e-10, "AKLT TEBD overlaps")
    string = [ref.corr(*ref_aklt.string_op(2, j)) for j in range(3, 9)]
    zzc = [ref.corr(*ref_aklt.zz_op(2, j)) for j in range(3, 9)]
    close(string, [ora.corr(*ref_aklt.string_op(2, j)) for j in range(3, 9)], 1e-12, "AKLT string order")
    np.savez(os.path.join(GOLD, "aklt_N11.npz"), energy=E_ref, sz=sz_ref, overlaps=np.array(ov_ref),
             string=np.array(string), zz=np.array(zzc), H=ref_aklt.H,
             **{f"s_{b}": np.asarray(ref.s[b], dtype=float) for b in range(N + 1)},
             **{f"raw_{i}": raw[i] for i in range(N)})
    del Ms
    print("AKLT N=11 flow: oracle == reference  (E/bond = %.16f)" % (E_ref / (N - 1)))


def case_tfim_quench():
    """iTEBD_double quench of |up...up> under TFIM, L=10 (SURVEY App. C last item) + ED cross-check."""
    L, n, T = 10, 40, 1.0
    Z, X, I2 = sigma[3], sigma[1], np.eye(2)
    Hb = -np.kron(Z, Z) - 0.5 * (np.kron(X, I2) + np.kron(I2, X))
    ref = RefMPS.naive([[1, 0]] * L)
    ref.trun, ref.err = 32, 1e-12
    ref.canon()
    ref.iTEBD_double(Hb, T, n)
    ref.canon()
    ora = OracleMPS.product([[1, 0]] * L, trun=32, err=1e-12)
    ora.canon()
    ora.tebd(Hb, T, n)
    ora.canon()
    z_ref = np.array([ref.corr((i, Z)) for i in range(L)])
    z_ora = np.array([ora.corr((i, Z)) for i in range(L)])
    close(z_ref, z_ora, 1e-11, "quench <Z_i>")
    for b in range(L + 1):
        close(ref.s[b], ora.s[b], 1e-11, f"quench s[{b}]")
    # ED with halved end fields (what the two-site split evolves)
    g = np.ones(L); g[0] = g[-1] = 0.5
    Hfull = ref_ising.Hamilton_trans(L, g=g, J=1)['H']
    psi0 = np.zeros(2 ** L); psi0[0] = 1
    psi = la.expm(-1j * T * Hfull) @ psi0
    z_ed = []
    for i in range(L):
        op = np.kron(np.kron(np.eye(2 ** i), Z), np.eye(2 ** (L - i - 1)))
        z_ed.append((psi.conj() @ op @ psi).real)
    print("TFIM quench L=10 n=40: oracle == reference; max |<Z> - ED| = %.2e (Trotter error)" % np.max(np.abs(z_ref - z_ed)))
    np.savez(os.path.join(GOLD, "tfim_quench_L10.npz"), Hb=Hb, T=T, n=n, z=z_ref, z_ed=np.array(z_ed),
             **{f"s_{b}": np.asarray(ref.s[b], dtype=float) for b in range(L + 1)})


def case_chi128_bond():
    """The benchmark unit of work: saturated chi=128 bond of an L=16 random chain, chain seed 1000."""
    L, d = 16, 2
    chis = [min(2 ** i, 2 ** (L - i), 128) for i in range(L + 1)]
    rs = np.random.RandomState(1000)
    Ms = [rs.randn(chis[i], d, chis[i + 1]) + 1j * rs.randn(chis[i], d, chis[i + 1]) for i in range(L)]
    U, _ = gue_gate(rs)
    ref = RefMPS([m.copy() for m in Ms], d, trun=128, err=1e-14)
    ref.canon()
    ora = OracleMPS([m.copy() for m in Ms], trun=128, err=1e-14)
    ora.canon()
    close(ref.s[8], ora.s[8], 1e-13, "chi128 canon s[8]")
    i = 7
    ora.update_double(U, i)
    ref.update_double(U, i)       # ~0.2 s in the reference (naive einsum)
    close(ref.Sr[i], ora.s[i + 1], 1e-13, "chi128 update_double Sr")
    th_r = np.tensordot(ref.M[i], ref.M[i + 1], axes=([2], [0]))
    th_o = np.tensordot(ora.M[i], ora.M[i + 1], axes=([2], [0]))
    close(th_r, th_o, 1e-11, "chi128 update_double theta")
    np.savez(os.path.join(GOLD, "chi128_bond7_seed1000.npz"), Sr=np.asarray(ref.Sr[i], dtype=float), k=int(ref.xr[i]),
             s_in=np.asarray(ref.s[8], dtype=float), theta_fro=la.norm(th_r), theta_probe=th_r[::17, :, :, ::19])
    print("chi=128 bench bond: oracle == reference (k = %d)" % ref.xr[i])


def case_idmrg():
    """iDMRG: dense H_eff vs matrix-free product; TFIM energies at trunc 10 (40 steps) and 16 (12 steps);
    Heisenberg with the real-form MPO."""
    W = (-RefMPO('sx') - RefMPO('sz') ** 2).tomatrix()
    close(W, io_.tfim_mpo(), 0, "TFIM MPO")
    rs = np.random.RandomState(3)
    Lr = rs.randn(5, 5, 3) + 1j * rs.randn(5, 5, 3)
    Rr = rs.randn(4, 4, 3) + 1j * rs.randn(4, 4, 3)
    th = rs.randn(5, 2, 2, 4) + 1j * rs.randn(5, 2, 2, 4)
    dense = ref_idmrg.matrixify(ref_idmrg.Hamiltonian(Lr, W, Rr))
    close(dense @ th.ravel(), io_.heff_matvec(Lr, W, Rr, th).ravel(), 1e-12, "H_eff matvec vs dense reference")
    Un = rs.randn(5, 2, 3) + 1j * rs.randn(5, 2, 3)
    Vn = rs.randn(3, 2, 4) + 1j * rs.randn(3, 2, 4)
    close(np.einsum('ijk, ilA, jmB, kClm', Lr, Un, Un.conj(), W), io_.env_left(Lr, Un, W, swap_conj=True), 1e-12, "env_left as written")
    close(np.einsum('ijk, Ali, Bmj, Cklm', Rr, Vn, Vn.conj(), W), io_.env_right(Rr, Vn, W, swap_conj=True), 1e-12, "env_right as written")
    np.savez(os.path.join(GOLD, "heff_small.npz"), L=Lr, R=Rr, W=W, theta=th, out=(dense @ th.ravel()).reshape(th.shape),
             U=Un, V=Vn, Lnext_ref=np.einsum('ijk, ilA, jmB, kClm', Lr, Un, Un.conj(), W),
             Rnext_ref=np.einsum('ijk, Ali, Bmj, Cklm', Rr, Vn, Vn.conj(), W))
    out = {"W": W}
    orig_halve = ref_idmrg.halve
    for trunc, steps in ((10, 40), (16, 12), (32, 9)):     # trunc = 32 is BASELINE config 1
        ref_idmrg.halve = lambda p, sh, trunc=trunc: orig_halve(p, sh, trunc)
        L, R = np.array([[[0, 0, 1]]]), np.array([[[1, 0, 0]]])
        Lo, Ro = io_.boundary(3)
        e_ref, e_ora = [], []
        for i in range(steps):
            L, R, val = quiet(ref_idmrg.next_LR, L, R, W)
            Lo, Ro, vo = io_.next_LR(Lo, Ro, W, trunc=trunc, swap_conj=False)
            e_ref.append(val / (2 * i + 2))
            e_ora.append(vo / (2 * i + 2))
        err = close(e_ref, e_ora, 2e-10, f"TFIM iDMRG energies trunc={trunc}")
        out[f"tfim_e_trunc{trunc}"] = np.array(e_ref)
        print(f"TFIM iDMRG trunc={trunc}: oracle == reference over {steps} steps (max diff {err:.1e})")
    ref_idmrg.halve = orig_halve
    # Heisenberg, real-form MPO, float64: reference is deterministic and ED-consistent here
    from DMRG.spin import P, N
    Wh = (2 * RefMPO('P', op=P()) * RefMPO('N', op=N()) + 2 * RefMPO('N', op=N()) * RefMPO('P', op=P())
          + RefMPO('sz') ** 2).tomatrix(dtype='double')
    ref_idmrg.halve = lambda p, sh, trunc=10: orig_halve(p, sh, 10)
    L, R = np.zeros((1, 1, 5)), np.zeros((1, 1, 5))
    L[0, 0, 4] = 1; R[0, 0, 0] = 1
    Lo, Ro = L.copy(), R.copy()
    e_ref, e_ora = [], []
    for i in range(7):
        L, R, val = quiet(ref_idmrg.next_LR, L, R, Wh)
        Lo, Ro, vo = io_.next_LR(Lo, Ro, Wh, trunc=10)
        e_ref.append(val); e_ora.append(vo)
    ref_idmrg.halve = orig_halve
    print("Heisenberg per-step |oracle - reference|:", np.abs(np.array(e_ref) - np.array(e_ora)))
    # trunc = 10 cuts through a degenerate Schmidt multiplet from n = 10 on, where the kept subspace (and
    # hence the energy at the 1e-8 level) depends on LAPACK's arbitrary basis inside the multiplet
    close(e_ref[:4], e_ora[:4], 1e-11, "Heisenberg iDMRG total energies (before truncation)")
    err = close(e_ref, e_ora, 1e-6, "Heisenberg iDMRG total energies")
    print("Heisenberg iDMRG (real MPO) trunc=10: oracle == reference (max diff %.1e): %s" % (err, np.round(e_ref, 6)))
    out["heis_W"] = Wh
    out["heis_e_total"] = np.array(e_ref)
    np.savez(os.path.join(GOLD, "idmrg_energies.npz"), **out)


def import_reference_transmit():
    """`transmit/` imports three things this image no longer has (`scipy.misc.imresize`, `matplotlib`, and
    `joblib.Memory(cachedir=...)`); none is used by `BMPS`.  Stub exactly those and import the package unmodified."""
    import types

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Memory:
        def __init__(self, *a, **k):
            pass

        def cache(self, f):
            return f
    stub("joblib", Memory=_Memory)
    stub("matplotlib")
    stub("matplotlib.pyplot")
    stub("matplotlib.colors", Normalize=object)
    scipy.misc = stub("scipy.misc", imresize=None, imsave=None)
    from transmit import bhopping, cat
    return bhopping, cat


BMPS_CASES = {
    "d6": dict(dim=6, l=6, para={"omega0": 0.3, "g": 1.0, "u": 0.7, "h": 0.4, "K": 0.5},
               pack={"dk": 0.5, "center": 3, "k_c": -np.pi / 2, "trun": True, "alpha": 1.2}, t=0.3, n=4, tail=True),
    "d10": dict(dim=10, l=5, para={"omega0": 0, "g": 1.0, "u": 0, "h": 0, "K": 1.0},
                pack={"dk": 0.5, "center": 2, "k_c": -np.pi / 2, "trun": True, "alpha": 1.5}, t=0.4, n=5, tail=True),
    "d5_notail": dict(dim=5, l=7, para={"omega0": 0.2, "g": 0.8, "u": 0.5, "h": 0.6, "K": 0.0},
                      pack={"dk": 0.4, "center": 3, "k_c": np.pi / 3, "trun": False, "alpha": 0.9}, t=0.5, n=3,
                      tail=False),
}


def case_bmps():
    """the reference's d >> 2 caller (transmit/bhopping.py), MPO application included."""
    bh, cat = import_reference_transmit()
    from oracle.bmps_oracle import OracleBMPS, coherent
    for al in (0.0, 0.4 - 0.3j, 1.7j):
        close(cat.coherent(al, 12), coherent(al, 12), 1e-15, "coherent state")
    out = {}
    for name, c in BMPS_CASES.items():
        ref = quiet(bh.BMPS, c["dim"], c["l"], c["para"], c["pack"])
        orc = OracleBMPS(c["dim"], c["l"], c["para"], c["pack"])
        occ0 = np.array([ref.measure(i, ref.N) for i in range(ref.L)])
        close(occ0, orc.occupations(), 1e-13, name + " initial occupations")
        quiet(ref.evolve, c["t"], c["n"], c["tail"])
        orc.evolve(c["t"], c["n"], c["tail"])
        x_raw = np.array(ref.x)
        assert list(ref.x) == list(orc.x), name + ": bond dimensions after evolve"
        ref.canon()
        orc.canon()
        assert list(ref.x) == list(orc.x), name + ": bond dimensions after canon"
        occ = np.array([ref.measure(i, ref.N) for i in range(ref.L)])
        close(occ, orc.occupations(), 1e-12, name + " occupations")
        for b in range(ref.L + 1):
            close(ref.s[b], orc.s[b], 1e-12, name + " Schmidt values bond %d" % b)
        nn = np.kron(ref.N, ref.N)
        corr = np.array([ref.measure(i, nn) for i in range(ref.l - 1)])
        close(corr, [orc.measure(i, nn) for i in range(ref.l - 1)], 1e-12, name + " <N N>")
        wig = quiet(ref.wigner, 0.3)
        out[name + "_occ0"], out[name + "_occ"], out[name + "_nn"] = occ0, occ, corr
        out[name + "_x_raw"], out[name + "_x"] = x_raw, np.array(ref.x)
        out[name + "_c_n"], out[name + "_wigner"] = np.array(ref.c_n), wig
        for b in range(ref.L + 1):
            out[name + "_s%d" % b] = np.asarray(ref.s[b], dtype=float)
        print("BMPS %s: oracle == reference; bonds %s, <N> = %s" % (name, [int(v) for v in ref.x], np.round(occ, 5)))
    np.savez(os.path.join(GOLD, "bmps.npz"), cases=json.dumps({k: {**v, "pack": {**v["pack"]}} for k, v in BMPS_CASES.items()}),
             **out)


if __name__ == "__main__":
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cases = {"canon": case_canon_testcanon, "two_body": case_two_body, "update_double": case_update_double,
             "aklt": case_aklt, "tfim": case_tfim_quench, "chi128": case_chi128_bond, "idmrg": case_idmrg,
             "bmps": case_bmps}
    for name in (sys.argv[1:] or list(cases)):          # `python oracle/validate_against_reference.py bmps` runs one
        cases[name]()
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/validate_against_reference.py", "reference": "/root/reference (peijunz-archive/DMRG, unmodified)",
                   "numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0]}, f, indent=1)
    print("all cases agree; goldens written to", GOLD)
