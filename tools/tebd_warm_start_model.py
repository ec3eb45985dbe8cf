"""Would warm-started Jacobi sweeps pay in a real TEBD run?  (design aid, not product)

A random canonical chain (L sites, d = 2, saturated at bond dimension chi) evolves under second-order Trotter steps of
the TFIM gate, with EVERY two-site split done by a numpy model of the GPU's sweeps (8-row blocks, modulus ordering,
de Rijk's rule: tools/jacobi_order_model.py), started

  * cold: from the pivoted-QR factor of m (what the library does), or
  * warm: from X0 = R V_prev with Q R = m V_prev^H, V_prev = ALL right singular vectors this bond's previous update
    ended with (rows of the final work matrix, normalised) - an exactly unitary left transform of m, no inverse of
    small singular values involved.  Because all splits of the warm chain are warm, the bond bases evolve
    continuously (rotations start from the identity), which is what the idea relies on.

Both chains follow the same trajectory up to gauge; sweeps and rotations of the saturated (2 chi x 2 chi) bonds are
summed per Trotter step.

    python tools/tebd_warm_start_model.py [chi] [L] [steps] [dt]      # defaults 32 12 6 0.05; a few minutes
"""
import os
import sys

import numpy as np
import scipy.linalg as la

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jacobi_order_model import precond, sweep_mod          # noqa: E402
from oracle.mps_oracle import OracleMPS                     # noqa: E402


def jacobi_rows(X0):
    X, rots = X0.copy(), []
    nblk = X.shape[0] // 8
    for _ in range(30):
        rots.append(sweep_mod(X, nblk, True))
        if rots[-1] == 0:
            break
    return X, len(rots), sum(rots)


class ModelMPS(OracleMPS):
    def __init__(self, *a, warm=False, **kw):
        super().__init__(*a, **kw)
        self.warm, self.vprev, self.log = warm, {}, []

    def update_double(self, U, i, unitary=True):
        Bi, Bj = self.M[i], self.M[i + 1]
        xl, d, _ = Bi.shape
        xr = Bj.shape[2]
        theta0 = np.tensordot(Bi, Bj, axes=([2], [0]))
        thetaB = np.tensordot(theta0, U, axes=([1, 2], [2, 3])).transpose(0, 2, 3, 1)
        m = (self.s[i][:, None, None, None] * thetaB).reshape(xl * d, d * xr)
        big = m.shape[0] == m.shape[1] and m.shape[0] % 16 == 0 and m.shape[0] >= 32
        if not big:
            return super().update_double(U, i, unitary)
        vp = self.vprev.get(i) if self.warm else None
        if vp is not None and vp.shape[0] == m.shape[1]:
            _, R = np.linalg.qr(m @ vp.conj().T)
            X0 = R @ vp
        else:
            X0 = precond(m)
        X, nsw, nrot = jacobi_rows(X0)
        self.log.append((nsw, nrot))
        nrm = np.linalg.norm(X, axis=1)
        order = np.argsort(-nrm)
        nrm, X = nrm[order], X[order]
        self.vprev[i] = X / nrm[:, None] if nrm[-1] > 1e-13 * nrm[0] else None
        s = nrm / nrm[0]
        k = min(int(np.sum(s > self.err)), self.trun)
        s = s[:k] / np.linalg.norm(s[:k])
        vh = X[:k] / nrm[:k, None]
        self.x[i + 1] = k
        self.s[i + 1] = s
        self.M[i] = (thetaB.reshape(xl * d, d * xr) @ vh.conj().T).reshape(xl, d, k)
        self.M[i + 1] = vh.reshape(k, d, xr)


if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    dt = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
    rs = np.random.RandomState(7)
    x = [min(2 ** i, 2 ** (L - i), chi) for i in range(L + 1)]
    Ms = [rs.randn(x[i], 2, x[i + 1]) + 1j * rs.randn(x[i], 2, x[i + 1]) for i in range(L)]
    Z, X_, I2 = np.diag([1.0, -1.0]), np.array([[0.0, 1.0], [1.0, 0.0]]), np.eye(2)
    H = -np.kron(Z, Z) - 0.5 * (np.kron(X_, I2) + np.kron(I2, X_))
    chains = {}
    for warm in (False, True):
        c = ModelMPS([m.copy() for m in Ms], trun=chi, err=1e-14, warm=warm)
        c.canon()
        chains[warm] = c
    for step in range(steps):
        line = []
        for warm in (False, True):
            c = chains[warm]
            c.log = []
            c.tebd(H, dt, 1)
            n = len(c.log)
            line.append(f"{'warm' if warm else 'cold'}: {n} splits, {np.mean([a for a, _ in c.log]):.2f} sweeps, "
                        f"{np.mean([b for _, b in c.log]) / 1e3:.2f} k rotations per split")
        ds = max(np.max(abs(a - b)) for a, b in zip(chains[False].s[1:L], chains[True].s[1:L]))
        print(f"step {step}: " + " | ".join(line) + f" | max Schmidt difference {ds:.1e}", flush=True)
