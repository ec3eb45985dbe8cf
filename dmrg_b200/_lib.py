"""ctypes binding of libmps_b200.so (include/mpsb.h) and the device-buffer helpers around it.

PyTorch is used for exactly three things: allocating device memory, host<->device copies and the
current CUDA stream.  All arithmetic happens in the library's own sm_100a kernels.  There is no CPU
fallback: if the shared library or a CUDA device is missing, the first call raises.
"""
import ctypes
import os
import threading as _threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmps_b200.so")

_c = ctypes
_vp, _i, _ll, _d, _sz = _c.c_void_p, _c.c_int, _c.c_longlong, _c.c_double, _c.c_size_t
_pd, _pi = _c.POINTER(_c.c_double), _c.POINTER(_c.c_int)

# name -> (restype, argtypes); mirrors include/mpsb.h one to one
SIGNATURES = {
    "mpsb_version": (_i, []),
    "mpsb_last_error": (_c.c_char_p, []),
    "mpsb_launch_count": (_c.c_ulonglong, []),
    "mpsb_last_svd_sweeps": (_i, []),
    "mpsb_debug_sweep_schedule": (_i, [_i, _i, _pi, _i]),
    "mpsb_profile_enable": (None, [_i]),
    "mpsb_profile_read": (_i, [_pd, _c.POINTER(_c.c_ulonglong), _c.POINTER(_c.c_ulonglong)]),
    "mpsb_zgemm": (_i, [_i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _ll, _i, _vp]),
    "mpsb_dgemm": (_i, [_i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _ll, _i, _vp]),
    "mpsb_svd_trunc_workspace": (_sz, [_i, _i, _i, _i]),
    "mpsb_svd_trunc": (_i, [_i, _i, _i, _vp, _i, _ll, _i, _d, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_svd_trunc_real_workspace": (_sz, [_i, _i, _i, _i]),
    "mpsb_svd_trunc_real": (_i, [_i, _i, _i, _vp, _i, _ll, _i, _d, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_tebd_two_site_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "mpsb_tebd_two_site": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _i, _d, _i, _vp, _ll,
                                _vp, _ll, _vp, _ll, _vp, _vp, _sz, _vp]),
    "mpsb_one_site": (_i, [_i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp]),
    "mpsb_mpo_site": (_i, [_i, _i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp]),
    "mpsb_predict_workspace": (_sz, [_i, _i, _i]),
    "mpsb_predict_theta": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_predict_batched_workspace": (_sz, [_i, _i, _i, _i]),
    "mpsb_predict_theta_batched": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_scale_rows": (_i, [_i, _i, _vp, _i, _vp, _i, _vp, _i, _vp]),
    "mpsb_transpose": (_i, [_i, _i, _vp, _i, _vp, _i, _i, _vp]),
    "mpsb_transpose_batched": (_i, [_i, _i, _i, _vp, _i, _ll, _vp, _i, _ll, _i, _vp]),
    "mpsb_heff_batched_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "mpsb_heff_matvec_batched": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _sz, _vp]),
    "mpsb_heff_matvec_batched_real": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _sz,
                                           _vp]),
    "mpsb_lanczos_ground_batched_real": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _d, _pd, _pd,
                                              _pi, _vp, _sz, _vp]),
    "mpsb_env_left_batched_real": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _sz, _vp]),
    "mpsb_env_right_batched_real": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _sz, _vp]),
    "mpsb_predict_theta_batched_real": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_transpose_batched_real": (_i, [_i, _i, _i, _vp, _i, _ll, _vp, _i, _ll, _vp]),
    "mpsb_last_lanczos_chain_matvecs": (_c.c_ulonglong, []),
    "mpsb_left_vectors_workspace": (_sz, [_i, _i, _i, _i]),
    "mpsb_left_vectors": (_i, [_i, _i, _i, _i, _vp, _i, _ll, _vp, _vp, _ll, _vp, _i, _vp, _sz, _vp]),
    "mpsb_lanczos_batched_workspace": (_sz, [_i, _i, _i, _i, _i, _i]),
    "mpsb_lanczos_ground_batched": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _d, _pd, _pd, _pi,
                                         _vp, _sz, _vp]),
    "mpsb_env_batched_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "mpsb_env_left_batched": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _i, _vp, _ll, _vp, _sz, _vp]),
    "mpsb_env_right_batched": (_i, [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _i, _vp, _ll, _vp, _sz, _vp]),
    "mpsb_gauge_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "mpsb_gauge_batched_workspace": (_sz, [_i, _i, _i, _i, _i, _i]),
    "mpsb_gauge_left_batched": (_i, [_i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _d, _i, _vp, _ll, _vp, _ll, _vp,
                                     _ll, _vp, _vp, _sz, _vp]),
    "mpsb_gauge_right_batched": (_i, [_i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _d, _i, _vp, _ll, _vp, _ll, _vp,
                                      _ll, _vp, _vp, _sz, _vp]),
    "mpsb_transfer_batched_workspace": (_sz, [_i, _i, _i, _i, _i, _i]),
    "mpsb_transfer_step_batched": (_i, [_i, _i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _sz,
                                        _vp]),
    "mpsb_gauge_left": (_i, [_i, _i, _i, _vp, _vp, _vp, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_gauge_right": (_i, [_i, _i, _i, _vp, _vp, _vp, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_transfer_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "mpsb_transfer_step": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_heff_workspace": (_sz, [_i, _i, _i, _i]),
    "mpsb_heff_matvec": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mpsb_lanczos_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "mpsb_lanczos_ground": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _d, _pd, _pd, _pi, _vp, _sz, _vp]),
    "mpsb_env_workspace": (_sz, [_i, _i, _i, _i]),
    "mpsb_env_left": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "mpsb_env_right": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
}


class MpsbError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpsbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). dmrg_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise MpsbError(f"{what} failed (code {rc}): {load().mpsb_last_error().decode()}")


def launch_count():
    return int(load().mpsb_launch_count())


# ---- device buffers (torch = allocator + copies + stream, nothing else) ----------------------------
_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        if not _t.cuda.is_available():
            raise MpsbError("dmrg_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
        _torch = _t
    return _torch


def device():
    t = torch()
    return t.device("cuda", t.cuda.current_device())


def bind_host_thread_to_gpu():
    """Pin the calling host thread to the CPUs next to the current GPU (NVML's ideal affinity), so that pinned host
    buffers allocated afterwards live on that NUMA node.  Multi-GPU runs that stream host-resident batches
    (`batched.HostBondUpdater`) otherwise push half of their PCIe traffic across the socket interconnect.  Returns
    the CPU set chosen, or None when NVML is unavailable (the call is then a no-op)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        t = torch()
        uuid = str(t.cuda.get_device_properties(t.cuda.current_device()).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def stream_ptr():
    return _c.c_void_p(torch().cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return _c.c_void_p(0)
    return _c.c_void_p(t.data_ptr())


def to_device(a, dtype=np.complex128):
    """Host ndarray -> contiguous device tensor of the given dtype."""
    t = torch()
    arr = np.ascontiguousarray(np.asarray(a), dtype=dtype)
    return t.from_numpy(arr).to(device())


def to_host(t):
    return t.detach().cpu().numpy()


def empty(shape, dtype="c128"):
    t = torch()
    dt = {"c128": t.complex128, "f64": t.float64, "i32": t.int32, "u8": t.uint8}[dtype]
    return t.empty(shape, dtype=dt, device=device())


def zeros(shape, dtype="c128"):
    t = torch()
    dt = {"c128": t.complex128, "f64": t.float64, "i32": t.int32, "u8": t.uint8}[dtype]
    return t.zeros(shape, dtype=dt, device=device())


class Workspace:
    """Grow-only scratch buffer handed to the library (it never allocates on its own)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes):
        nbytes = int(nbytes)
        if nbytes <= 0:
            raise MpsbError("workspace query failed: " + load().mpsb_last_error().decode())
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = empty((nbytes + (1 << 20),), "u8")
        return self.buf


# One scratch buffer per (host thread, device, stream): calls issued on different streams or devices - or from
# different threads - never share theta / work-matrix / Krylov scratch (include/mpsb.h: one host thread per
# (device, stream)).  Kernels of one stream run in order, so reusing that stream's buffer call after call is safe.
_ws_local = _threading.local()


def _workspace_for_current_stream():
    t = torch()
    key = (t.cuda.current_device(), int(t.cuda.current_stream().cuda_stream))
    table = getattr(_ws_local, "table", None)
    if table is None:
        table = _ws_local.table = {}
    ws = table.get(key)
    if ws is None:
        ws = table[key] = Workspace()
    return ws


def workspace(nbytes):
    buf = _workspace_for_current_stream().get(nbytes)
    return _c.c_void_p(buf.data_ptr()), _c.c_size_t(buf.numel())


# ---- thin typed wrappers -----------------------------------------------------------------------------
OP_N, OP_T, OP_C, OP_J = 0, 1, 2, 3


def zgemm(A, B, C, M, N, K, lda, ldb, ldc, opA=OP_N, opB=OP_N, batch1=1, batch2=1, sA=(0, 0), sB=(0, 0), sC=(0, 0),
          accumulate=False):
    check(load().mpsb_zgemm(opA, opB, M, N, K, ptr(A), lda, ptr(B), ldb, ptr(C), ldc, batch1, batch2, sA[0], sA[1],
                            sB[0], sB[1], sC[0], sC[1], int(accumulate), stream_ptr()), "mpsb_zgemm")


def dgemm(A, B, C, M, N, K, lda, ldb, ldc, opA=OP_N, opB=OP_N, batch1=1, batch2=1, sA=(0, 0), sB=(0, 0), sC=(0, 0),
          accumulate=False):
    check(load().mpsb_dgemm(opA, opB, M, N, K, ptr(A), lda, ptr(B), ldb, ptr(C), ldc, batch1, batch2, sA[0], sA[1],
                            sB[0], sB[1], sC[0], sC[1], int(accumulate), stream_ptr()), "mpsb_dgemm")


def matmul(A, B, opA=OP_N, opB=OP_N):
    """Plain 2-D product of device matrices through the DMMA kernel (shim convenience)."""
    ar, ac = A.shape
    br, bc = B.shape
    M, K = (ar, ac) if opA in (OP_N, OP_J) else (ac, ar)
    K2, N = (br, bc) if opB in (OP_N, OP_J) else (bc, br)
    if K != K2:
        raise MpsbError(f"matmul: inner dimensions differ ({K} vs {K2})")
    C = empty((M, N))
    if M and N:
        if K == 0:
            C.zero_()
        else:
            zgemm(A, B, C, M, N, K, ac, bc, N, opA, opB)
    return C


def conj_transpose(A, conj=True):
    """(rows x cols) device matrix -> its (conjugate) transpose, through the library's tile-transpose kernel."""
    rows, cols = A.shape
    if not A.is_complex():
        out = empty((cols, rows), "f64")
        if rows and cols:
            check(load().mpsb_transpose_batched_real(1, rows, cols, ptr(A), cols, 0, ptr(out), rows, 0, stream_ptr()),
                  "mpsb_transpose_batched_real")
        return out
    out = empty((cols, rows))
    if rows and cols:
        check(load().mpsb_transpose(rows, cols, ptr(A), cols, ptr(out), rows, int(conj), stream_ptr()), "mpsb_transpose")
    return out


def warn_if_unconverged(where):
    """The Jacobi iteration stops after MPSB_SVD_MAX_SWEEPS (40) sweeps; say so instead of staying silent."""
    if int(load().mpsb_last_svd_sweeps()) >= int(os.environ.get("MPSB_SVD_MAX_SWEEPS", "40")):
        import warnings
        warnings.warn(f"{where}: Jacobi SVD stopped at the sweep limit; results may be less accurate than 1e-10")
