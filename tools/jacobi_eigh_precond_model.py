"""How many sweeps are left if the rows start from the eigenvectors of the Gram matrix?  (design aid, not product)

X0 = Q^H R with R the pivoted-QR factor of the benchmark matrix and Q the eigenvectors of R R^H from LAPACK's zheevd
(an FP64 Hermitian eigensolver standing in for a batched GPU one: tridiagonalisation + QL / divide and conquer, GEMM-rich
back-transformation).  Squaring the matrix costs the small singular values their relative accuracy, so X0 is not yet the
answer - but it is close, and the one-sided sweeps (kernel schedule of tools/jacobi_order_model.py) only have to clean up.

    python tools/jacobi_eigh_precond_model.py [chi]

chi = 128 (n = 256): cold start 10 sweeps / 221.7 k rotations; from the Gram eigenvectors 3 sweeps / 10.7 k rotations
(10 723 + 21 + 0: the last two sweeps only verify), same singular values, the smallest 16 to 7e-14 relative.
"""
import sys

import numpy as np

from jacobi_order_model import bench_matrix, precond, rms_cos, sweep_mod

if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    m = bench_matrix(chi, 2, 1000)
    R = precond(m)
    nblk = R.shape[0] // 8
    sref = np.linalg.svd(m, compute_uv=False)
    for name in ("cold (pivoted QR)", "Gram eigenvectors"):
        X = R.copy()
        if name.startswith("Gram"):
            w, Q = np.linalg.eigh(R @ R.conj().T)
            X = Q[:, ::-1].conj().T @ R
        rots, start = [], rms_cos(X)
        for _ in range(20):
            rots.append(sweep_mod(X, nblk, True))
            if rots[-1] == 0:
                break
        s = np.sort(np.linalg.norm(X, axis=1))[::-1]
        print(f"{name}: start rms |cos| = {start:.1e}, {len(rots)} sweeps, {sum(rots)} rotations {rots}, "
              f"max |s - s_lapack| / s_0 = {np.max(abs(s - sref)) / sref[0]:.1e}, "
              f"max relative error of the smallest 16: {np.max(abs(s[-16:] - sref[-16:]) / sref[-16:]):.1e}", flush=True)
