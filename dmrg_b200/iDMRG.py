"""Infinite-size DMRG growth step with the interface of the reference's `DMRG/iDMRG.py`, executed
matrix-free on the GPU.

The reference builds the dense effective Hamiltonian (`Hamiltonian`, :38-40 — (chi^2 d^2)^2 entries,
68.7 GB at chi = 128) and gives it to ARPACK (:57).  Here `next_LR` never forms it: the Lanczos solver
in libmps_b200 applies L.W.W.R to the two-site tensor as GEMM -> fused W.W kernel -> GEMM, the SVD
split runs in the Jacobi kernel and the environment growth in the fused env kernels.

Index conventions (unchanged): L[A,a,i] (bra bond, ket bond, mpo), R[D,d,k], M[i,j,B,b]
(mpo-row, mpo-col, bra, ket).  Differences from the reference, both deliberate:
  * the environment growth conjugates the bra leg (mathematically correct); pass `swap_conj=True`
    to get the reference's as-written contraction — identical whenever the Schmidt vectors are real;
  * the Lanczos start vector is seeded (`seed`), the reference's ARPACK start is random.
"""
import numpy as np

from . import _lib
from .MPO import MPO  # noqa: F401  (re-exported like the reference does)
from .MPS import _device_svd

KRYLOV = 24          # Lanczos basis size between restarts
MAX_RESTARTS = 60
TOL = 1e-13          # residual target relative to max(1, |E|)
last_info = {}       # diagnostics of the most recent next_LR: n_matvec, residual, S (singular values)


def halfshape(A):
    """Shape (chi_l, d, d, chi_r) of the two-site tensor H_eff acts on (`DMRG/iDMRG.py:13-14`)."""
    return A.shape[:4]


def matrixify(A):
    """Reshape a (chi,d,d,chi,chi,d,d,chi) tensor into a square matrix (`DMRG/iDMRG.py:43-46`)."""
    n = np.rint(np.sqrt(A.size)).astype(int)
    return A.reshape([n, n])


def halve(psi, sh, trunc=10):
    """SVD-split psi[(chi_l d), (d chi_r)] and keep `trunc` states (`DMRG/iDMRG.py:17-35`):
    returns U[chi_l, d, k], the FULL unnormalised S, V[k, d, chi_r] with k = min(trunc, S.size).
    Accepts a numpy array (returns numpy) or a device tensor (returns device tensors)."""
    on_host = isinstance(psi, np.ndarray)
    lsize, rsize = sh[0] * sh[1], sh[2] * sh[3]
    dev = _lib.to_device(psi.reshape(lsize, rsize)) if on_host else psi.reshape(lsize, rsize)
    u, s_all, vh, k = _device_svd(dev, lsize, rsize, trunc, 0.0, want_u=True, normalise=0)
    U = u.reshape(sh[0], sh[1], k)
    V = vh.reshape(k, sh[2], sh[3])
    if on_host:
        return _lib.to_host(U), _lib.to_host(s_all), _lib.to_host(V)
    return U, s_all, V


def _halve_split(theta, sh, trunc):
    """U and V of the two-site tensor from two independent row-orthogonalisations (of psi and of psi^H) instead of
    one with carried identity columns: rows stay half as long, which keeps chi <= 128 on the register-resident
    Jacobi kernel.  U and V then carry unrelated phases per singular vector - harmless for `next_LR`, whose two
    environment updates use them separately (DMRG/iDMRG.py:59-60)."""
    lsize, rsize = sh[0] * sh[1], sh[2] * sh[3]
    A = theta.reshape(lsize, rsize)
    _, s_all, vh, k = _device_svd(A, lsize, rsize, trunc, 0.0, want_u=False, normalise=0)
    _, _, uh, k2 = _device_svd(_lib.conj_transpose(A), rsize, lsize, trunc, 0.0, want_u=False, normalise=0)
    assert k == k2
    U = _lib.conj_transpose(uh)
    return U.reshape(sh[0], sh[1], k), s_all, vh.reshape(k, sh[2], sh[3])


_halve_public = halve


def _heff_apply(L, M, R, theta):
    lib = _lib.load()
    xl, xr, w, d = L.shape[0], R.shape[0], M.shape[0], M.shape[2]
    out = _lib.empty(tuple(theta.shape))
    ws, wsn = _lib.workspace(lib.mpsb_heff_workspace(xl, xr, d, w))
    _lib.check(lib.mpsb_heff_matvec(xl, xr, d, w, _lib.ptr(L), _lib.ptr(M), _lib.ptr(R), _lib.ptr(theta), _lib.ptr(out),
                                    ws, wsn, _lib.stream_ptr()), "mpsb_heff_matvec")
    return out


def heff_matvec(L, M, R, theta):
    """out = H_eff theta without forming H_eff (new; host arrays in and out)."""
    Ld, Md, Rd, td = (_lib.to_device(a) for a in (L, M, R, theta))
    return _lib.to_host(_heff_apply(Ld, Md, Rd, td))


def Hamiltonian(L, M, R):
    """Dense H_eff[A,B,C,D,a,b,c,d] (`DMRG/iDMRG.py:38-40`).  Kept for interface parity only — built column
    by column from the matrix-free product, so use it for small chi; `next_LR` does not call it."""
    L, M, R = np.asarray(L), np.asarray(M), np.asarray(R)
    sh = (L.shape[0], M.shape[2], M.shape[2], R.shape[0])
    n = int(np.prod(sh))
    Ld, Md, Rd = (_lib.to_device(a) for a in (L, M, R))
    H = np.empty((n, n), dtype=complex)
    basis = np.zeros(n, dtype=complex)
    for col in range(n):
        basis[col] = 1
        H[:, col] = _lib.to_host(_heff_apply(Ld, Md, Rd, _lib.to_device(basis.reshape(sh)))).ravel()
        basis[col] = 0
    return H.reshape(sh + sh)


def ground_state(Ld, Md, Rd, seed=0, theta0=None):
    """Lowest eigenpair of H_eff on device tensors: (E, theta) with theta normalised."""
    lib = _lib.load()
    xl, xr, w, d = Ld.shape[0], Rd.shape[0], Md.shape[0], Md.shape[2]
    if theta0 is None:      # seeded start vector, generated on the device (the reference's ARPACK start is random too)
        t = _lib.torch()
        gen = t.Generator(device=_lib.device())
        gen.manual_seed(int(seed))
        theta0 = t.randn((xl, d, d, xr, 2), dtype=t.float64, device=_lib.device(), generator=gen)
        theta0 = t.view_as_complex(theta0)
    theta = theta0.clone()
    import ctypes
    E, resid, nmv = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
    ws, wsn = _lib.workspace(lib.mpsb_lanczos_workspace(xl, xr, d, w, KRYLOV))
    _lib.check(lib.mpsb_lanczos_ground(xl, xr, d, w, _lib.ptr(Ld), _lib.ptr(Md), _lib.ptr(Rd), _lib.ptr(theta), KRYLOV,
                                       MAX_RESTARTS, TOL, ctypes.byref(E), ctypes.byref(resid), ctypes.byref(nmv), ws,
                                       wsn, _lib.stream_ptr()), "mpsb_lanczos_ground")
    last_info.update(n_matvec=nmv.value, residual=resid.value)
    return E.value, theta


def _env(left, Ed, Td, Md, swap_conj):
    lib = _lib.load()
    w, d = Md.shape[0], Md.shape[2]
    chi = Ed.shape[0]
    chin = Td.shape[2] if left else Td.shape[0]
    out = _lib.empty((chin, chin, w))
    ws, wsn = _lib.workspace(lib.mpsb_env_workspace(chi, chin, d, w))
    fn = lib.mpsb_env_left if left else lib.mpsb_env_right
    _lib.check(fn(chi, chin, d, w, _lib.ptr(Ed), _lib.ptr(Td), _lib.ptr(Md), int(swap_conj), _lib.ptr(out), ws, wsn,
                  _lib.stream_ptr()), "mpsb_env")
    return out


def next_LR_device(Ld, Rd, Md, trunc=None, swap_conj=False, seed=0):
    """`next_LR` on device-resident tensors (no host round trip): returns (L', R', E, S_all)."""
    E, theta = ground_state(Ld, Md, Rd, seed=seed)
    sh = tuple(theta.shape)
    if trunc is None and halve is not _halve_public:       # somebody rebound the module-level halve (reference idiom)
        U, S, V = halve(theta, sh)
    elif sh[0] * sh[1] <= 256:                             # carried-identity SVD is faster while rows stay short
        U, S, V = halve(theta, sh, 10 if trunc is None else trunc)
    else:
        U, S, V = _halve_split(theta, sh, 10 if trunc is None else trunc)
    Ln = _env(True, Ld, U.contiguous(), Md, swap_conj)
    Rn = _env(False, Rd, V.contiguous(), Md, swap_conj)
    last_info["S"] = _lib.to_host(S)
    return Ln, Rn, E, S


def next_LR(L, R, M, trunc=None, swap_conj=False, seed=0):
    """One iDMRG growth step (`DMRG/iDMRG.py:49-61`): returns (L', R', E) where E is the total energy
    of the enlarged open chain.  numpy in, numpy out."""
    Ld, Rd, Md = (_lib.to_device(np.asarray(a)) for a in (L, R, M))
    Ln, Rn, E, _ = next_LR_device(Ld, Rd, Md, trunc=trunc, swap_conj=swap_conj, seed=seed)
    return _lib.to_host(Ln), _lib.to_host(Rn), E


if __name__ == "__main__":
    # Transverse-field Ising chain H = -sum X - sum Z Z, as in the reference's driver (`DMRG/iDMRG.py:64-76`)
    M = (- MPO('sx') - MPO('sz') ** 2).tomatrix()
    L = np.array([[[0, 0, 1]]])
    R = np.array([[[1, 0, 0]]])
    for i in range(40):
        L, R, val = next_LR(L, R, M)
        print('Energy {} is'.format(2 * i + 2), val / (i * 2 + 2))
