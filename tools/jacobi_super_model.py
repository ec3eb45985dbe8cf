"""Numpy model of the SUPER-BLOCK modulus ordering of jacobi_quad_kernel against the 8-row-block modulus ordering
(design aid, not product).

    python tools/jacobi_super_model.py [chi] [seed]

chi = 128: 10 sweeps / 221.7 k rotations (8-row blocks) against 10 sweeps / 219.9 k (super blocks of 16 rows).
"""
import sys

import numpy as np

from jacobi_order_model import bench_matrix, cross_task, precond, rms_cos, self_task, sweep_mod

def mod_pairs(m, k):
    out = []
    for I in range(m):
        J = (k - I) % m
        if I < J: out.append((I, J))
        elif I == J: out.append((I, I))
    return out

def sweep_super(X, nblk, sort):
    n = 0
    ns = nblk // 2
    for K in range(ns):
        for SI, SJ in mod_pairs(ns, K):
            if SI < SJ:
                for a in (0, 1):
                    for b in (0, 1):
                        n += cross_task(X, 2 * SI + a, 2 * SJ + b, sort)
            else:
                n += self_task(X, 2 * SI); n += self_task(X, 2 * SI + 1)
                n += cross_task(X, 2 * SI, 2 * SI + 1, sort)
    return n

if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    m = bench_matrix(chi, 2, seed)
    X0 = precond(m)
    sref = np.linalg.svd(m, compute_uv=False)
    nblk = X0.shape[0] // 8
    for name, fn, sort in (("modulus + de Rijk", sweep_mod, True), ("super-block modulus + de Rijk", sweep_super, True)):
        X, hist, rots = X0.copy(), [], []
        for _ in range(20):
            hist.append(rms_cos(X))
            rots.append(fn(X, nblk, sort))
            if rots[-1] == 0:
                break
        s = np.sort(np.linalg.norm(X, axis=1))[::-1]
        print(f"{name}: {len(rots)} sweeps, {sum(rots)} rotations, max |s - s_lapack| / s_0 = {np.max(abs(s - sref)) / sref[0]:.1e}", flush=True)
        print("   rotations per sweep     :", rots, flush=True)
