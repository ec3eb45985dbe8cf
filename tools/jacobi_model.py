"""Numpy model of the block row-Jacobi SVD the CUDA kernels implement (design aid, not product).

Rows of X (n x L) are orthogonalised by left unitary block rotations:
for each block pair (I,J): G = P P^H (P = the 2b rows), Q from `inner` sweeps of
parallel-ordered two-sided Jacobi on G, P <- Q^H P.
"""
import numpy as np, sys, time

def rr_pairs(n, step):
    """round-robin (circle method): n even, step in [0,n-1) -> n/2 disjoint pairs"""
    m = n - 1
    pairs = [(m, step % m)]
    for k in range(1, n // 2):
        a = (step + k) % m
        b = (step - k) % m
        pairs.append((min(a, b), max(a, b)))
    out = []
    for p, q in pairs:
        out.append((min(p, q), max(p, q)))
    return out

def inner_jacobi(G, sweeps, tol):
    nb = G.shape[0]
    G = G.copy()
    Q = np.eye(nb, dtype=complex)
    nrot = 0
    for sw in range(sweeps):
        rot_this = 0
        for st in range(nb - 1):
            J = np.eye(nb, dtype=complex)
            for p, q in rr_pairs(nb, st):
                a = G[p, p].real; b = G[q, q].real; c = G[p, q]
                ac = abs(c)
                if ac <= tol * np.sqrt(a * b) or ac == 0.0:
                    continue
                rot_this += 1
                tau = (b - a) / (2 * ac)
                t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1 + tau * tau))
                cs = 1 / np.sqrt(1 + t * t); sn = t * cs
                ph = c / ac
                J[p, p] = cs; J[p, q] = sn * ph
                J[q, p] = -sn * np.conj(ph); J[q, q] = cs
            G = J.conj().T @ G @ J
            Q = Q @ J
        nrot += rot_this
        if rot_this == 0:
            break
    return Q, nrot

def block_jacobi(X, b=16, inner=1, tol=1e-15, max_sweeps=30, verbose=False):
    X = X.copy()
    n, L = X.shape
    nb = 2 * b
    assert n % nb == 0
    nblk = n // b
    for sweep in range(max_sweeps):
        nrot_total = 0; maxoff = 0
        for rnd in range(nblk - 1):
            for (I, J) in rr_pairs(nblk, rnd):
                idx = np.r_[I * b:(I + 1) * b, J * b:(J + 1) * b]
                P = X[idx]
                G = P @ P.conj().T
                d = np.sqrt(np.abs(np.diag(G).real))
                with np.errstate(invalid='ignore', divide='ignore'):
                    off = np.abs(G) / np.outer(d, d)
                off[~np.isfinite(off)] = 0
                np.fill_diagonal(off, 0)
                mo = off.max()
                maxoff = max(maxoff, mo)
                if mo <= tol:
                    continue
                Q, nrot = inner_jacobi(G, inner, tol)
                nrot_total += nrot
                X[idx] = Q.conj().T @ P
        if verbose:
            print(f"sweep {sweep}: maxoff={maxoff:.3e} rotations={nrot_total}")
        if nrot_total == 0:
            break
    s = np.linalg.norm(X, axis=1)
    order = np.argsort(-s)
    return s[order], X[order], sweep + 1

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    inner = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    rs = np.random.RandomState(0)
    kind = sys.argv[4] if len(sys.argv) > 4 else "rand"
    if kind == "rand":
        m = rs.randn(n, n) + 1j * rs.randn(n, n)
    else:  # decaying spectrum like an MPS bond
        u, _ = np.linalg.qr(rs.randn(n, n) + 1j * rs.randn(n, n))
        v, _ = np.linalg.qr(rs.randn(n, n) + 1j * rs.randn(n, n))
        sv = np.exp(-np.arange(n) * 40.0 / n)
        m = (u * sv) @ v
    t = time.time()
    s, Xs, nsw = block_jacobi(m, b=b, inner=inner, verbose=True)
    sref = np.linalg.svd(m, compute_uv=False)
    print("sweeps", nsw, "time", time.time() - t)
    print("max abs err s / s0:", np.max(np.abs(s - sref)) / sref[0])
    print("max rel err s:", np.max(np.abs(s - sref) / sref))
    Vh = Xs / s[:, None]
    k = n // 2
    print("orth err V (top half):", np.abs(Vh[:k] @ Vh[:k].conj().T - np.eye(k)).max())
