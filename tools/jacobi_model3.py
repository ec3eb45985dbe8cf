"""Exact emulation of the planned kernel ordering: blocks of b rows; per sweep: intra-block tournament,
then round-robin over block pairs, within a block pair b steps of b disjoint cross pairs."""
import numpy as np, sys, time
def rr_pairs(n, step):
    m = n - 1
    pairs = [(m, step % m)]
    for k in range(1, n // 2):
        a = (step + k) % m; b = (step - k) % m
        pairs.append((min(a, b), max(a, b)))
    return pairs
def rotate(X, p, q, Ld, tol, negl, stats):
    xp = X[p]; xq = X[q]
    a = np.vdot(xp[:Ld], xp[:Ld]).real; b = np.vdot(xq[:Ld], xq[:Ld]).real
    c = np.dot(xp[:Ld], xq[:Ld].conj())
    ac = abs(c)
    if not (ac > tol * np.sqrt(a * b)) or min(a,b) <= negl:
        # still swap so that larger norm first? no
        return
    stats[0] += 1
    tau = (b - a) / (2 * ac)
    t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1 + tau * tau))
    cs = 1 / np.sqrt(1 + t * t); sn = t * cs; ph = c / ac
    yp = cs * xp - sn * ph * xq
    yq = sn * np.conj(ph) * xp + cs * xq
    X[p] = yp; X[q] = yq
def svd_model(m, b=16, tol=None, max_sweeps=30, aug=None, verbose=False):
    n, Ld = m.shape
    X = m.copy() if aug is None else np.concatenate([m, aug], axis=1)
    npad = -(-n // (2 * b)) * 2 * b
    X = np.concatenate([X, np.zeros((npad - n, X.shape[1]), complex)], axis=0)
    if tol is None: tol = np.sqrt(Ld) * 2.2e-16
    nblk = npad // b
    gmax = max((abs(X[:, :Ld]) ** 2).sum(1))
    negl = gmax * 1e-34
    for sweep in range(max_sweeps):
        st = [0]
        for I in range(nblk):
            for s in range(b - 1):
                for (p, q) in rr_pairs(b, s):
                    rotate(X, I * b + p, I * b + q, Ld, tol, negl, st)
        for rnd in range(nblk - 1):
            for (I, J) in rr_pairs(nblk, rnd):
                for s in range(b):
                    for w in range(b):
                        rotate(X, I * b + w, J * b + (w + s) % b, Ld, tol, negl, st)
        if verbose: print("sweep", sweep, "rot", st[0])
        if st[0] == 0: break
    s = np.sqrt((abs(X[:, :Ld]) ** 2).sum(1))
    o = np.argsort(-s, kind='stable')
    return s[o], X[o], sweep + 1
if __name__ == "__main__":
    n = int(sys.argv[1]); b = int(sys.argv[2]); kind = sys.argv[3]
    rs = np.random.RandomState(0)
    if kind == "rand": m = rs.randn(n, n) + 1j * rs.randn(n, n)
    elif kind == "lowrank":
        m = (rs.randn(n, n//2) + 1j * rs.randn(n, n//2)) @ (rs.randn(n//2, n) + 1j * rs.randn(n//2, n))
    else:
        u, _ = np.linalg.qr(rs.randn(n, n) + 1j * rs.randn(n, n)); v, _ = np.linalg.qr(rs.randn(n, n) + 1j * rs.randn(n, n))
        m = (u * np.exp(-np.arange(n) * 40.0 / n)) @ v
    t = time.time()
    s, X, nsw = svd_model(m, b, verbose=True, aug=np.eye(n, dtype=complex))
    sref = np.linalg.svd(m, compute_uv=False)
    print("sweeps", nsw, "time", time.time() - t)
    print("abs err/s0", np.max(np.abs(s[:n] - sref)) / sref[0])
    k = (sref > 1e-12 * sref[0]).sum()
    print("k", k, "rel err top-k", np.max(np.abs(s[:k] - sref[:k]) / sref[:k]))
    Vh = X[:k, :n] / s[:k, None]; U = X[:k, n:].conj().T
    print("recon", np.abs((U * s[:k]) @ Vh - m).max() / sref[0], "orthU", np.abs(U.conj().T @ U - np.eye(k)).max(), "orthV", np.abs(Vh @ Vh.conj().T - np.eye(k)).max())
