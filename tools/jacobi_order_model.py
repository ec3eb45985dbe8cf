"""Numpy model of the block pair ORDERING of the one-sided Jacobi sweeps (design aid, not product).

Compares, on the benchmark matrices (chi x d = n rows, QR-preconditioned exactly as the kernels do), the number of
sweeps and rotations of
  * the round-robin ordering of 8-row blocks (round 0: complete 16-row tournaments; rounds 1..: cross pairs), and
  * the modulus ordering (step k: block pairs with I + J = k mod nblk; blocks with 2 I = k meet themselves),
    with and without de Rijk's rule (after a rotation the row of the lower block keeps the larger norm).

    python tools/jacobi_order_model.py [chi]        # chi = 128 is the benchmark (takes about a minute)

Measured here (chi = 128, n = 256): round-robin 12 sweeps / 272.8 k rotations (the CUDA kernels: 12 / 271 k);
modulus 10 / 238.5 k; modulus + de Rijk 10 / 218.0 k.
"""
import sys

import numpy as np
import scipy.linalg as la

TOL = 16 * 2.220446049250313e-16


def rr_pairs(n, step):
    m = n - 1
    return [(m, step % m)] + [((step + k) % m, (step - k) % m) for k in range(1, n // 2)]


def bench_matrix(chi, d, seed):
    rs = np.random.RandomState(seed)

    def iso():
        q, _ = np.linalg.qr(rs.randn(d * chi, chi) + 1j * rs.randn(d * chi, chi))
        return np.ascontiguousarray(q.conj().T).reshape(chi, d, chi)
    Bi, Bj = iso(), iso()
    s = np.linalg.svd(rs.randn(chi, chi) + 1j * rs.randn(chi, chi), compute_uv=False)
    A = rs.randn(d * d, d * d) + 1j * rs.randn(d * d, d * d)
    U = la.expm(-0.05j * (A + A.conj().T)).reshape([d] * 4)
    th = np.einsum('lcr,rjk,abcj->labk', Bi, Bj, U)
    return ((s / np.linalg.norm(s))[:, None, None, None] * th).reshape(chi * d, d * chi)


def precond(m):
    order = np.argsort(-np.linalg.norm(m, axis=0))
    _, r = np.linalg.qr(m[:, order])
    X = np.zeros_like(m)
    X[:, order] = r
    return X


def rms_cos(X):
    G = X @ X.conj().T
    d = np.sqrt(np.real(np.diag(G)))
    C = abs(G) / np.outer(d, d)
    np.fill_diagonal(C, 0)
    return np.sqrt((C ** 2).sum() / (C.shape[0] * (C.shape[0] - 1)))


def rot_pair(X, p, q, sort):
    xp, xq = X[p], X[q]
    a, b, c = np.vdot(xp, xp).real, np.vdot(xq, xq).real, np.vdot(xq, xp)
    ac = abs(c)
    if ac <= TOL * np.sqrt(a * b):
        return 0
    tau = (b - a) / (2 * ac)
    t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1 + tau * tau))
    cs = 1 / np.sqrt(1 + t * t)
    sn, ph = t * cs, c / ac
    np_, nq_ = cs * xp - sn * ph * xq, sn * np.conj(ph) * xp + cs * xq
    if sort and (a - t * ac) < (b + t * ac):
        np_, nq_ = nq_, np_
    X[p], X[q] = np_, nq_
    return 1


def cross_task(X, I, J, sort, B=8):          # the kernel's order: warp w holds rows 2w, 2w+1 of I, step s meets row (2w+s)%8
    n = 0
    for s in range(B):
        for w in range(B // 2):
            q = (2 * w + s) % B
            n += rot_pair(X, I * B + 2 * w, J * B + q, sort)
            n += rot_pair(X, I * B + 2 * w + 1, J * B + q, sort)
    return n


def self_task(X, I, B=8):
    return sum(rot_pair(X, I * B + p, I * B + q, False) for p in range(B - 1) for q in range(p + 1, B))


def full16_task(X, I, J, B=8):
    idx = [I * B + r for r in range(B)] + [J * B + r for r in range(B)]
    return sum(rot_pair(X, idx[min(p, q)], idx[max(p, q)], False) for st in range(2 * B - 1) for p, q in rr_pairs(2 * B, st))


def sweep_rr(X, nblk, sort):
    n = 0
    for rnd in range(nblk - 1):
        for I, J in rr_pairs(nblk, rnd):
            I, J = min(I, J), max(I, J)
            n += full16_task(X, I, J) if rnd == 0 else cross_task(X, I, J, False)
    return n


def sweep_mod(X, nblk, sort):
    n = 0
    for k in range(nblk):
        for I in range(nblk):
            J = (k - I) % nblk
            if I < J:
                n += cross_task(X, I, J, sort)
            elif I == J:
                n += self_task(X, I)
    return n


if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    m = bench_matrix(chi, 2, 1000)
    X0 = precond(m)
    sref = np.linalg.svd(m, compute_uv=False)
    nblk = X0.shape[0] // 8
    for name, fn, sort in (("round-robin (round-1 kernels)", sweep_rr, False), ("modulus", sweep_mod, False),
                           ("modulus + de Rijk (sort on rotation, cross tasks)", sweep_mod, True)):
        X, hist, rots = X0.copy(), [], []
        for _ in range(20):
            hist.append(rms_cos(X))
            rots.append(fn(X, nblk, sort))
            if rots[-1] == 0:
                break
        s = np.sort(np.linalg.norm(X, axis=1))[::-1]
        print(f"{name}: {len(rots)} sweeps, {sum(rots)} rotations, max |s - s_lapack| / s_0 = {np.max(abs(s - sref)) / sref[0]:.1e}")
        print("   rms |cos| at sweep start:", " ".join(f"{h:.1e}" for h in hist))
        print("   rotations per sweep     :", rots)
