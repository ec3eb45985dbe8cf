"""Model of the planned kernel: block row-cyclic (wavefront-equivalent) one-sided Jacobi with
per-sweep presort by row norm, inner parallel-ordered two-sided Jacobi on the Gram block."""
import numpy as np, sys
sys.path.insert(0, 'tools')
from jacobi_model import rr_pairs

def inner(G, max_sweeps, tol):
    nb = G.shape[0]; G = G.copy(); Q = np.eye(nb, dtype=complex); nrot = 0; nsw = 0
    for sw in range(max_sweeps):
        r = 0
        for st in range(nb - 1):
            J = np.eye(nb, dtype=complex)
            for p, q in rr_pairs(nb, st):
                a = G[p, p].real; b = G[q, q].real; c = G[p, q]; ac = abs(c)
                if not (ac > tol * np.sqrt(abs(a * b))) or ac == 0:
                    continue
                r += 1
                tau = (b - a) / (2 * ac); t = (1. if tau >= 0 else -1.) / (abs(tau) + np.sqrt(1 + tau * tau))
                cs = 1 / np.sqrt(1 + t * t); sn = t * cs; ph = c / ac
                J[p, p] = cs; J[p, q] = sn * ph; J[q, p] = -sn * np.conj(ph); J[q, q] = cs
            if r:
                G = J.conj().T @ G @ J; Q = Q @ J
        nsw += 1
        nrot += r
        if r == 0:
            break
    return Q, nrot, nsw

def block_jacobi(X, b, inner_sweeps, tol=1e-15, max_sweeps=60, order="cyc", presort=True, verbose=False):
    X = X.copy(); n = X.shape[0]; nblk = n // b; hist = []; inner_total = 0; pairs_rot = 0
    for sweep in range(max_sweeps):
        if presort:
            o = np.argsort(-(np.abs(X) ** 2).sum(1), kind='stable'); X = X[o]
        if order == "rr":
            seq = [pr for rnd in range(nblk - 1) for pr in rr_pairs(nblk, rnd)]
        else:
            seq = [(I, J) for I in range(nblk) for J in range(I + 1, nblk)]
        tot = 0
        for I, J in seq:
            idx = np.r_[I * b:(I + 1) * b, J * b:(J + 1) * b]; P = X[idx]; G = P @ P.conj().T
            Q, nr, nsw = inner(G, inner_sweeps, tol); tot += nr; inner_total += nsw
            if nr:
                X[idx] = Q.conj().T @ P; pairs_rot += 1
        hist.append(tot)
        if verbose: print("sweep", sweep, "rot", tot)
        if tot == 0: break
    s = np.sort(np.linalg.norm(X, axis=1))[::-1]
    return s, hist, inner_total, pairs_rot

if __name__ == "__main__":
    n = int(sys.argv[1]); b = int(sys.argv[2])
    rs = np.random.RandomState(0)
    u, _ = np.linalg.qr(rs.randn(n, n) + 1j * rs.randn(n, n)); v, _ = np.linalg.qr(rs.randn(n, n) + 1j * rs.randn(n, n))
    for decay in [0., 10., 40.]:
        if decay == 0.: m = rs.randn(n, n) + 1j * rs.randn(n, n)
        else:
            sv = np.exp(-np.arange(n) * decay / n); m = (u * sv) @ v
        sref = np.linalg.svd(m, compute_uv=False)
        for isw in [1, 2, 3, 30]:
            s, h, it, pr = block_jacobi(m, b, isw)
            npairs = (n // b) * (n // b - 1) // 2
            print(f"decay {decay} inner<={isw}: outer {len(h)} pairs_rotated {pr} (={pr/npairs:.1f} sweeps-equiv) inner_sweeps_total {it} (avg {it/max(pr,1):.1f}/pair) err {np.max(np.abs(s-sref))/sref[0]:.1e}")
