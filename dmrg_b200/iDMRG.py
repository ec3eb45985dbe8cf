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

REAL_PATH = True     # real L, W, R (MPO.tomatrix(dtype='double'), DMRG/MPO.py:68-76) run in real arithmetic (DGEMM on DMMA)
# large splits: U = theta V^H S^-1 (+ two Newton-Schulz steps) while S_k / S_0 stays above this.  The columns start
# orthonormal to e = eps S_0 / S_k; the symmetric re-orthonormalisation spreads that defect over ALL columns, the
# dominant ones included, so e <= 2e-11 is what keeps them at the 1e-10 level of the parity target
SPLIT_CUTOFF = 1e-5
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


def _is_real_t(t):
    return not t.is_complex()


def _all_real(*arrays):
    """True when every host array is real-valued (real dtype, or complex dtype with an exactly zero imaginary part)."""
    if not REAL_PATH:
        return False
    for a in arrays:
        a = np.asarray(a)
        if np.iscomplexobj(a) and np.any(a.imag != 0):
            return False
    return True


def _to_device_auto(real, *arrays):
    if real:
        return tuple(_lib.to_device(np.asarray(a).real, np.float64) for a in arrays)
    return tuple(_lib.to_device(np.asarray(a)) for a in arrays)


def _as_complex(x):
    t = _lib.torch()
    return t.complex(x, t.zeros_like(x)) if not x.is_complex() else x


def halve(psi, sh, trunc=10):
    """SVD-split psi[(chi_l d), (d chi_r)] and keep `trunc` states (`DMRG/iDMRG.py:17-35`):
    returns U[chi_l, d, k], the FULL unnormalised S, V[k, d, chi_r] with k = min(trunc, S.size).
    Accepts a numpy array (returns numpy) or a device tensor (returns device tensors)."""
    on_host = isinstance(psi, np.ndarray)
    lsize, rsize = sh[0] * sh[1], sh[2] * sh[3]
    if on_host:
        dev, = _to_device_auto(np.isrealobj(psi) and REAL_PATH, psi.reshape(lsize, rsize))
    else:
        dev = psi.reshape(lsize, rsize)      # a real (float64) tensor takes the real SVD (mpsb_svd_trunc_real)
    u, s_all, vh, k = _device_svd(dev, lsize, rsize, trunc, 0.0, want_u=True, normalise=0)
    U = u.reshape(sh[0], sh[1], k)
    V = vh.reshape(k, sh[2], sh[3])
    if on_host:
        return _lib.to_host(U), _lib.to_host(s_all), _lib.to_host(V)
    return U, s_all, V


def _halve_split(theta, sh, trunc):
    """Split of a two-site tensor whose rows are too long to carry an identity through the Jacobi sweeps: V and S
    from a row-orthogonalisation of psi (rows stay short: register-resident kernel up to chi = 128, half the tile
    traffic beyond), then U = psi V^H S^-1 re-orthonormalised by two Newton-Schulz steps (`mpsb_left_vectors`) - the
    Schmidt partners of V, also inside a degenerate multiplet that `trunc` cuts.  When the kept spectrum reaches below
    SPLIT_CUTOFF * S_0 that product is not orthonormal to working precision any more and U comes from a second,
    independent row-orthogonalisation of psi^H (orthonormal always; its phases and its choice inside a cut multiplet
    are then unrelated to V's - harmless for `next_LR`, whose two environment updates use U and V separately,
    DMRG/iDMRG.py:59-60)."""
    lsize, rsize = sh[0] * sh[1], sh[2] * sh[3]
    A = theta.reshape(lsize, rsize)
    _, s_all, vh, k = _device_svd(A, lsize, rsize, trunc, 0.0, want_u=False, normalise=0)
    U = _left_vectors(A.reshape(1, lsize, rsize), vh.reshape(1, k, rsize), s_all.reshape(1, -1), k)
    if U is not None:                                       # Schmidt partners of V from three GEMMs
        return U.reshape(sh[0], sh[1], k), s_all, vh.reshape(k, sh[2], sh[3])
    _, _, uh, k2 = _device_svd(_lib.conj_transpose(A), rsize, lsize, trunc, 0.0, want_u=False, normalise=0)
    assert k == k2
    U = _lib.conj_transpose(uh)
    return U.reshape(sh[0], sh[1], k), s_all, vh.reshape(k, sh[2], sh[3])


def _left_vectors(A, vh, s_all, k):
    """U [n, rows, k] = A Vh^H S^-1, re-orthonormalised by two Newton-Schulz steps (`mpsb_left_vectors`), or None when
    some chain's kept spectrum reaches below SPLIT_CUTOFF * S_0 (the columns of the small values would then not be
    orthonormal to working precision and the caller takes them from a second SVD)."""
    s_host = _lib.to_host(s_all)[:, :k]
    if not np.all(s_host[:, k - 1] > SPLIT_CUTOFF * s_host[:, 0]):
        return None
    lib = _lib.load()
    n, rows, cols = (int(v) for v in A.shape)
    real = _is_real_t(A)
    U = _lib.empty((n, rows, k), "f64" if real else "c128")
    A = A.contiguous()
    ws, wsn = _lib.workspace(lib.mpsb_left_vectors_workspace(n, rows, cols, k))
    _lib.check(lib.mpsb_left_vectors(n, rows, cols, k, _lib.ptr(A), cols, rows * cols, _lib.ptr(vh.contiguous()),
                                     _lib.ptr(s_all), int(s_all.shape[1]), _lib.ptr(U), int(real), ws, wsn,
                                     _lib.stream_ptr()), "mpsb_left_vectors")
    return U


_halve_public = halve


def _heff_apply(L, M, R, theta):
    lib = _lib.load()
    xl, xr, w, d = L.shape[0], R.shape[0], M.shape[0], M.shape[2]
    ws, wsn = _lib.workspace(lib.mpsb_heff_workspace(xl, xr, d, w))
    if _is_real_t(L):
        out = _lib.empty(tuple(theta.shape), "f64")
        _lib.check(lib.mpsb_heff_matvec_batched_real(1, xl, xr, d, w, _lib.ptr(L), 0, _lib.ptr(M), 0, _lib.ptr(R), 0,
                                                     _lib.ptr(theta), 0, _lib.ptr(out), 0, ws, wsn, _lib.stream_ptr()),
                   "mpsb_heff_matvec_batched_real")
        return out
    out = _lib.empty(tuple(theta.shape))
    _lib.check(lib.mpsb_heff_matvec(xl, xr, d, w, _lib.ptr(L), _lib.ptr(M), _lib.ptr(R), _lib.ptr(theta), _lib.ptr(out),
                                    ws, wsn, _lib.stream_ptr()), "mpsb_heff_matvec")
    return out


def heff_matvec(L, M, R, theta):
    """out = H_eff theta without forming H_eff (new; host arrays in and out; real arithmetic when all four are real)."""
    Ld, Md, Rd, td = _to_device_auto(_all_real(L, M, R, theta), L, M, R, theta)
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
    real = _is_real_t(Ld)
    if theta0 is None:      # seeded start vector, generated on the device (the reference's ARPACK start is random too)
        t = _lib.torch()
        gen = t.Generator(device=_lib.device())
        gen.manual_seed(int(seed))
        theta0 = t.randn((xl, d, d, xr, 2), dtype=t.float64, device=_lib.device(), generator=gen)
        theta0 = theta0[..., 0].contiguous() if real else t.view_as_complex(theta0)
    theta = theta0.clone()
    import ctypes
    E, resid, nmv = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
    ws, wsn = _lib.workspace(lib.mpsb_lanczos_workspace(xl, xr, d, w, KRYLOV))
    if real:
        _lib.check(lib.mpsb_lanczos_ground_batched_real(1, xl, xr, d, w, _lib.ptr(Ld), 0, _lib.ptr(Md), 0, _lib.ptr(Rd), 0,
                                                        _lib.ptr(theta), xl * d * d * xr, KRYLOV, MAX_RESTARTS, TOL,
                                                        ctypes.byref(E), ctypes.byref(resid), ctypes.byref(nmv), ws, wsn,
                                                        _lib.stream_ptr()), "mpsb_lanczos_ground_batched_real")
    else:
        _lib.check(lib.mpsb_lanczos_ground(xl, xr, d, w, _lib.ptr(Ld), _lib.ptr(Md), _lib.ptr(Rd), _lib.ptr(theta), KRYLOV,
                                           MAX_RESTARTS, TOL, ctypes.byref(E), ctypes.byref(resid), ctypes.byref(nmv), ws,
                                           wsn, _lib.stream_ptr()), "mpsb_lanczos_ground")
    last_info.update(n_matvec=nmv.value, residual=resid.value)
    _warn_if_residual_large(E.value, resid.value)
    return E.value, theta


def _env(left, Ed, Td, Md, swap_conj):
    lib = _lib.load()
    w, d = Md.shape[0], Md.shape[2]
    chi = Ed.shape[0]
    chin = Td.shape[2] if left else Td.shape[0]
    ws, wsn = _lib.workspace(lib.mpsb_env_workspace(chi, chin, d, w))
    if _is_real_t(Ed):
        out = _lib.empty((chin, chin, w), "f64")
        fn = lib.mpsb_env_left_batched_real if left else lib.mpsb_env_right_batched_real
        _lib.check(fn(1, chi, chin, d, w, _lib.ptr(Ed), 0, _lib.ptr(Td), 0, _lib.ptr(Md), 0, _lib.ptr(out), 0, ws, wsn,
                      _lib.stream_ptr()), "mpsb_env_real")
        return out
    out = _lib.empty((chin, chin, w))
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
    elif sh[0] * sh[1] <= (512 if _is_real_t(theta) else 256):   # carried-identity SVD is faster while rows stay short
        U, S, V = halve(theta, sh, 10 if trunc is None else trunc)
    else:
        U, S, V = _halve_split(theta, sh, 10 if trunc is None else trunc)
    Ln = _env(True, Ld, U.contiguous(), Md, swap_conj)
    Rn = _env(False, Rd, V.contiguous(), Md, swap_conj)
    last_info["S"] = _lib.to_host(S)
    return Ln, Rn, E, S


def _overlap(a, b):
    """<a|b> of two device tensors of equal shape: one batched 1x1 GEMM per row, summed on the host."""
    rows = int(a.shape[0]) * int(a.shape[1])
    cols = a.numel() // rows
    real = _is_real_t(a)
    part = _lib.empty((rows,), "f64" if real else "c128")
    (_lib.dgemm if real else _lib.zgemm)(a.reshape(rows, cols), b.reshape(rows, cols), part, 1, 1, cols, cols, 1, 1,
                                         opA=_lib.OP_J, opB=_lib.OP_T, batch1=rows, sA=(cols, 0), sB=(cols, 0), sC=(1, 0))
    return complex(_lib.to_host(part).sum())


class Chain:
    """iDMRG run that keeps its matrix product state (new).

    The reference's `next_LR` (`DMRG/iDMRG.py:49-61`) returns the grown environments and forgets U, S, V.  `Chain`
    keeps them: `A` (left-orthonormal, = U), `B` (right-orthonormal, = V) and the normalised Schmidt values `S` of
    the centre bond, plus `S_prev` of the step before.  That buys two things the reference cannot do:

    * observables of the infinite chain - `to_mps()` hands the two centre sites to `dmrg_b200.MPS` (measure, corr);
    * wavefunction prediction - every step after the second starts Lanczos from (S B) S_prev^-1 (A S) instead of
      a random vector (`mpsb_predict_theta`), which cuts the matvec count several-fold near convergence;
      `fidelity[-1]` = |<prediction|ground state>| is the usual iDMRG convergence measure.

    Energies are the same as `next_LR`'s (the start vector does not change the eigenpair Lanczos converges to)."""

    def __init__(self, M, trunc=10, L=None, R=None, swap_conj=False, seed=0, predict=True, cutoff=1e-12):
        M = np.asarray(M)
        w = M.shape[0]
        if L is None:                                      # open boundaries of a lower-triangular MPO (:69-70)
            L = np.zeros((1, 1, w), dtype=complex)
            L[0, 0, w - 1] = 1
        if R is None:
            R = np.zeros((1, 1, w), dtype=complex)
            R[0, 0, 0] = 1
        self.real = _all_real(M, L, R)                     # real MPO and boundaries: the whole run stays real
        self.M, self.L, self.R = _to_device_auto(self.real, M, L, R)
        self.trunc, self.swap_conj, self.seed, self.predict, self.cutoff = trunc, swap_conj, seed, predict, cutoff
        self.A = self.B = self.S = self.S_prev = None
        self.energies, self.n_matvec, self.fidelity = [], [], []

    @property
    def n_sites(self):
        return 2 * len(self.energies)

    def _prediction(self):
        lib = _lib.load()
        chip, d, chi = (int(v) for v in self.A.shape)
        if len(self.S_prev) != chip or len(self.S) != chi:
            return None
        # Schmidt values that are rounding noise (rank-deficient states: AKLT keeps 4 of 64) are cut on BOTH bonds:
        # the guess then has exact zeros on the dead states instead of products of noise that underflow step by step.
        inv = np.where(self.S_prev > self.cutoff * self.S_prev[0], 1.0 / np.maximum(self.S_prev, 1e-300), 0.0)
        cur = np.where(self.S > self.cutoff * self.S[0], self.S, 0.0)
        S, Sinv = _lib.to_device(cur, np.float64), _lib.to_device(inv, np.float64)
        ws, wsn = _lib.workspace(lib.mpsb_predict_workspace(chi, chip, d))
        if self.real:
            theta = _lib.empty((chi, d, d, chi), "f64")
            _lib.check(lib.mpsb_predict_theta_batched_real(1, chi, chip, d, _lib.ptr(self.A), _lib.ptr(self.B), _lib.ptr(S),
                                                           _lib.ptr(Sinv), _lib.ptr(theta), ws, wsn, _lib.stream_ptr()),
                       "mpsb_predict_theta_batched_real")
            return theta
        theta = _lib.empty((chi, d, d, chi))
        _lib.check(lib.mpsb_predict_theta(chi, chip, d, _lib.ptr(self.A), _lib.ptr(self.B), _lib.ptr(S), _lib.ptr(Sinv),
                                          _lib.ptr(theta), ws, wsn, _lib.stream_ptr()), "mpsb_predict_theta")
        return theta

    def step(self):
        """One growth step (two sites); returns the total energy of the enlarged chain like `next_LR`."""
        guess = self._prediction() if (self.predict and self.S_prev is not None) else None
        E, theta = ground_state(self.L, self.M, self.R, seed=self.seed, theta0=guess)
        self.n_matvec.append(last_info["n_matvec"])
        if guess is not None:
            nrm = np.sqrt(abs(_overlap(guess, guess)))
            self.fidelity.append(abs(_overlap(guess, theta)) / nrm if nrm > 0 else 0.0)
        sh = tuple(theta.shape)
        U, S_all, V = _halve_public(theta, sh, self.trunc)     # one SVD with carried identity: U and V share a gauge
        self.L = _env(True, self.L, U.contiguous(), self.M, self.swap_conj)
        self.R = _env(False, self.R, V.contiguous(), self.M, self.swap_conj)
        s = _lib.to_host(S_all)[:U.shape[2]]
        self.S_prev, self.S = self.S, s / np.linalg.norm(s)
        self.A, self.B = U.contiguous(), V.contiguous()
        self.energies.append(E)
        last_info["S"] = _lib.to_host(S_all)
        return E

    def run(self, n):
        for _ in range(n):
            self.step()
        return self.energies[-1]

    def energy_per_site(self):
        """Bond energy (E_n - E_{n-1}) / 2: the estimate of e_inf that converges fastest with the chain length."""
        if len(self.energies) < 2:
            return self.energies[-1] / 2
        return (self.energies[-1] - self.energies[-2]) / 2

    def entropy(self):
        p = self.S ** 2
        p = p[p > 0]
        return float(-np.sum(p * np.log(p)))

    def to_mps(self):
        """The two centre sites as a canonical `MPS` (open legs = the orthonormal block bases): site 0 = A S, site 1 = B."""
        from .MPS import MPS
        a = _lib.to_host(self.A) * self.S[np.newaxis, np.newaxis, :]
        mps = MPS([a, _lib.to_host(self.B)], trun=max(int(self.trunc), len(self.S)), err=0.0)
        mps.canon()
        return mps


# ---- many independent chains in lock step (BASELINE config 4: coupling sweeps over h/J) -------------------------------
def _transpose_batched(a, conj=True):
    """[n, rows, cols] device tensor -> [n, cols, rows] (conjugated) through the library's tile-transpose kernel."""
    n, rows, cols = (int(v) for v in a.shape)
    if _is_real_t(a):
        out = _lib.empty((n, cols, rows), "f64")
        _lib.check(_lib.load().mpsb_transpose_batched_real(n, rows, cols, _lib.ptr(a), cols, rows * cols, _lib.ptr(out), rows,
                                                           rows * cols, _lib.stream_ptr()), "mpsb_transpose_batched_real")
        return out
    out = _lib.empty((n, cols, rows))
    _lib.check(_lib.load().mpsb_transpose_batched(n, rows, cols, _lib.ptr(a), cols, rows * cols, _lib.ptr(out), rows,
                                                  rows * cols, int(conj), _lib.stream_ptr()), "mpsb_transpose_batched")
    return out


def _svd_batched(A, trunc, want_u):
    """`halve`'s SVD (no eps cut, raw singular values) of n matrices [n, rows, cols] in one call: (Uh | None, S_all, Vh)
    with Uh [n, k, rows] = U^H, Vh [n, k, cols], k = min(trunc, rows, cols) for every chain."""
    lib = _lib.load()
    n, rows, cols = (int(v) for v in A.shape)
    real = _is_real_t(A)
    dt = "f64" if real else "c128"
    k = max(1, min(rows, cols, int(trunc)))
    vh = _lib.empty((n, k, cols), dt)
    uh = _lib.empty((n, k, rows), dt) if want_u else None
    s = _lib.empty((n, k), "f64")
    sraw = _lib.empty((n, min(rows, cols)), "f64")
    rank = _lib.empty((n,), "i32")
    wsq, fn = (lib.mpsb_svd_trunc_real_workspace, lib.mpsb_svd_trunc_real) if real else (lib.mpsb_svd_trunc_workspace,
                                                                                         lib.mpsb_svd_trunc)
    ws, wsn = _lib.workspace(wsq(n, rows, cols, int(want_u)))
    _lib.check(fn(n, rows, cols, _lib.ptr(A), cols, rows * cols, int(trunc), 0.0, 0, k, _lib.ptr(vh),
                  _lib.ptr(uh), _lib.ptr(s), _lib.ptr(sraw), _lib.ptr(rank), ws, wsn, _lib.stream_ptr()),
               "mpsb_svd_trunc")
    _lib.warn_if_unconverged("halve (batched)")
    return uh, sraw, vh


def halve_batched(theta, trunc):
    """`halve` (DMRG/iDMRG.py:17-35) for n two-site tensors [n, xl, d, d, xr] at once: U [n, xl, d, k], S_all [n, .],
    V [n, k, d, xr] (device tensors; every chain keeps k = min(trunc, S.size) states, so the shapes stay common)."""
    n, xl, d, _, xr = (int(v) for v in theta.shape)
    A = theta.reshape(n, xl * d, d * xr)
    if xl * d <= (512 if _is_real_t(theta) else 256):       # (real rows pack two columns into one element)                                       # carried-identity SVD: U and V from one decomposition
        uh, s_all, vh = _svd_batched(A, trunc, True)
        U = _transpose_batched(uh)
    else:                                                   # two row-orthogonalisations (see _halve_split)
        _, s_all, vh = _svd_batched(A, trunc, False)
        U = _left_vectors(A, vh, s_all, int(vh.shape[1]))
        if U is None:
            _, _, uh = _svd_batched(_transpose_batched(A), trunc, False)
            U = _transpose_batched(uh)
    k = int(vh.shape[1])
    return U.reshape(n, xl, d, k), s_all, vh.reshape(n, k, d, xr)


def ground_state_batched(Ld, Md, Rd, seed=0, theta0=None):
    """Lowest eigenpairs of n effective Hamiltonians in lock step: Ld [n, xl, xl, w], Rd [n, xr, xr, w], Md
    [w, w, d, d] (one MPO for all chains) or [n, w, w, d, d] (one per chain).  Returns (E [n] numpy, theta
    [n, xl, d, d, xr] device, normalised per chain)."""
    import ctypes
    lib = _lib.load()
    n, xl, xr, w = int(Ld.shape[0]), int(Ld.shape[1]), int(Rd.shape[1]), int(Ld.shape[3])
    d = int(Md.shape[-1])
    sW = w * w * d * d if Md.dim() == 5 else 0
    nel = xl * d * d * xr
    real = _is_real_t(Ld)
    if theta0 is None:
        t = _lib.torch()
        gen = t.Generator(device=_lib.device())
        gen.manual_seed(int(seed))
        theta0 = t.randn((n, xl, d, d, xr, 2), dtype=t.float64, device=_lib.device(), generator=gen)
        theta0 = theta0[..., 0].contiguous() if real else t.view_as_complex(theta0)
    theta = theta0.clone()
    E = (ctypes.c_double * n)()
    resid = (ctypes.c_double * n)()
    nmv = ctypes.c_int(0)
    ws, wsn = _lib.workspace(lib.mpsb_lanczos_batched_workspace(n, xl, xr, d, w, KRYLOV))
    fn = lib.mpsb_lanczos_ground_batched_real if real else lib.mpsb_lanczos_ground_batched
    _lib.check(fn(n, xl, xr, d, w, _lib.ptr(Ld), xl * xl * w, _lib.ptr(Md), sW, _lib.ptr(Rd), xr * xr * w, _lib.ptr(theta),
                  nel, KRYLOV, MAX_RESTARTS, TOL, E, resid, ctypes.byref(nmv), ws, wsn, _lib.stream_ptr()),
               "mpsb_lanczos_ground_batched")
    E, resid = np.array(E[:]), np.array(resid[:])
    last_info.update(n_matvec=nmv.value, residual=float(resid.max()),
                     chain_matvecs=int(lib.mpsb_last_lanczos_chain_matvecs()))
    _warn_if_residual_large(E, resid)
    return E, theta


def _warn_if_residual_large(E, resid):
    """The reference's ARPACK raises ArpackNoConvergence; here an eigenpair whose true residual ||H v - E v|| is far from
    the target after MAX_RESTARTS cycles is reported instead of flowing silently into the split."""
    bad = np.atleast_1d(resid) > 1e3 * TOL * np.maximum(1.0, np.abs(np.atleast_1d(E)))
    if bad.any():
        import warnings
        warnings.warn(f"Lanczos ground state: residual {float(np.max(resid)):.2e} exceeds 1e3 x TOL for {int(bad.sum())} "
                      f"chain(s) after {MAX_RESTARTS} restarts")


def _env_batched(left, Ed, Td, Md, swap_conj):
    lib = _lib.load()
    n, chi, w = int(Ed.shape[0]), int(Ed.shape[1]), int(Ed.shape[3])
    d = int(Md.shape[-1])
    chin = int(Td.shape[3]) if left else int(Td.shape[1])
    sW = w * w * d * d if Md.dim() == 5 else 0
    ws, wsn = _lib.workspace(lib.mpsb_env_batched_workspace(n, chi, chin, d, w))
    if _is_real_t(Ed):
        out = _lib.empty((n, chin, chin, w), "f64")
        fn = lib.mpsb_env_left_batched_real if left else lib.mpsb_env_right_batched_real
        _lib.check(fn(n, chi, chin, d, w, _lib.ptr(Ed), chi * chi * w, _lib.ptr(Td), chi * d * chin, _lib.ptr(Md), sW,
                      _lib.ptr(out), chin * chin * w, ws, wsn, _lib.stream_ptr()), "mpsb_env_batched_real")
        return out
    out = _lib.empty((n, chin, chin, w))
    fn = lib.mpsb_env_left_batched if left else lib.mpsb_env_right_batched
    _lib.check(fn(n, chi, chin, d, w, _lib.ptr(Ed), chi * chi * w, _lib.ptr(Td), chi * d * chin, _lib.ptr(Md), sW,
                  int(swap_conj), _lib.ptr(out), chin * chin * w, ws, wsn, _lib.stream_ptr()), "mpsb_env_batched")
    return out


class ChainBatch:
    """n independent iDMRG runs in lock step on one GPU - the coupling sweep of BASELINE config 4 (e.g. TFIM with
    g / J on np.linspace(0.2, 2.0, n)).  The reference would run `next_LR` (DMRG/iDMRG.py:49-61) once per
    parameter in separate processes; here every Lanczos product, the split and the environment growth are single
    batched launches over all chains.  All chains share (w, d) and, because `halve` keeps min(trunc, S.size) states
    without an eps cut, the same bond dimension at every step.

        W = np.stack([(-g * MPO('sx') - MPO('sz') ** 2).tomatrix() for g in gs])     # [n, w, w, d, d]
        run = ChainBatch(W, trunc=128)
        for _ in range(steps):
            E = run.step()            # total energies of the enlarged chains, array [n]
    """

    def __init__(self, W, trunc=10, left_index=None, right_index=0, swap_conj=False, predict=True, cutoff=1e-12):
        W = np.asarray(W)
        assert W.ndim == 5, "W: one MPO per chain, [n, w, w, d, d]"
        self.n, self.w, self.d = int(W.shape[0]), int(W.shape[1]), int(W.shape[3])
        self.trunc, self.swap_conj = int(trunc), bool(swap_conj)
        self.predict, self.cutoff = bool(predict), float(cutoff)
        self.A = self.B = self.Sn = self.Sn_prev = None
        self.real = _all_real(W)                            # real MPOs (TFIM, XXZ, Heisenberg): real arithmetic throughout
        L = np.zeros((self.n, 1, 1, self.w), dtype=complex)
        R = np.zeros((self.n, 1, 1, self.w), dtype=complex)
        L[:, 0, 0, self.w - 1 if left_index is None else left_index] = 1     # boundary vectors of the lower-triangular MPO
        R[:, 0, 0, right_index] = 1
        self.W, self.L, self.R = _to_device_auto(self.real, W, L, R)
        self.energies, self.n_matvec, self.S = [], [], None
        self.chain_matvecs = []       # H_eff products per step summed over the chains that formed them (frozen chains drop out)
        self.sites = 0

    def _prediction(self):
        """Start vectors (S B) S_prev^-1 (A S) of all chains from the tensors of the last two splits (`Chain._prediction`
        batched; needs U and V of one SVD, i.e. chi d <= 256, and a bond dimension that has stopped growing)."""
        if not self.predict or self.Sn_prev is None or self.A is None:
            return None
        n, chip, d, chi = (int(v) for v in self.A.shape)
        if self.Sn_prev.shape[1] != chip or self.Sn.shape[1] != chi or chi * d > 256:
            return None
        lib = _lib.load()
        inv = np.where(self.Sn_prev > self.cutoff * self.Sn_prev[:, :1], 1.0 / np.maximum(self.Sn_prev, 1e-300), 0.0)
        cur = np.where(self.Sn > self.cutoff * self.Sn[:, :1], self.Sn, 0.0)
        theta = _lib.empty((n, chi, d, d, chi), "f64" if self.real else "c128")
        S, Sinv = _lib.to_device(cur, np.float64), _lib.to_device(inv, np.float64)
        ws, wsn = _lib.workspace(lib.mpsb_predict_batched_workspace(n, chi, chip, d))
        fn = lib.mpsb_predict_theta_batched_real if self.real else lib.mpsb_predict_theta_batched
        _lib.check(fn(n, chi, chip, d, _lib.ptr(self.A), _lib.ptr(self.B), _lib.ptr(S), _lib.ptr(Sinv), _lib.ptr(theta), ws,
                      wsn, _lib.stream_ptr()), "mpsb_predict_theta_batched")
        return theta

    def step(self, seed=0):
        """One growth step for every chain; returns the total energies [n]."""
        E, theta = ground_state_batched(self.L, self.W, self.R, seed=seed + self.sites, theta0=self._prediction())
        U, S, V = halve_batched(theta, self.trunc)
        self.L = _env_batched(True, self.L, U.contiguous(), self.W, self.swap_conj)
        self.R = _env_batched(False, self.R, V.contiguous(), self.W, self.swap_conj)
        self.S = _lib.to_host(S)
        k = int(U.shape[3])
        sk = self.S[:, :k]
        self.Sn_prev, self.Sn = self.Sn, sk / np.linalg.norm(sk, axis=1, keepdims=True)
        self.A, self.B = U.contiguous(), V.contiguous()
        self.sites += 2
        self.energies.append(E)
        self.n_matvec.append(last_info["n_matvec"])
        self.chain_matvecs.append(last_info["chain_matvecs"])
        return E

    def energy_per_site(self):
        """Energy per site from the last two growth steps (bond energy of the infinite chain), array [n]."""
        assert len(self.energies) >= 2
        return (self.energies[-1] - self.energies[-2]) / 2

    def entropies(self):
        """Entanglement entropy (nats) of the kept spectrum of the last split, array [n]."""
        s = self.S[:, :self.trunc]
        p = s ** 2 / np.sum(s ** 2, axis=1, keepdims=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            return -np.where(p > 0, p * np.log(p), 0.0).sum(axis=1)


def next_LR(L, R, M, trunc=None, swap_conj=False, seed=0):
    """One iDMRG growth step (`DMRG/iDMRG.py:49-61`): returns (L', R', E) where E is the total energy
    of the enlarged open chain.  numpy in, numpy out."""
    Ld, Rd, Md = _to_device_auto(_all_real(L, R, M), L, R, M)
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
