import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dmrg_b200 import iDMRG, _lib
lib = _lib.load()
rs = np.random.RandomState(0)
for chi in (32, 128):
    w, d = 3, 2
    L = _lib.to_device(rs.randn(chi, chi, w) + 0j); R = _lib.to_device(rs.randn(chi, chi, w) + 0j)
    W = _lib.to_device(rs.randn(w, w, d, d) + 0j); th = _lib.to_device(rs.randn(chi, d, d, chi) + 0j)
    out = _lib.empty((chi, d, d, chi))
    ws, wsn = _lib.workspace(lib.mpsb_heff_workspace(chi, chi, d, w))
    def mv():
        _lib.check(lib.mpsb_heff_matvec(chi, chi, d, w, _lib.ptr(L), _lib.ptr(W), _lib.ptr(R), _lib.ptr(th), _lib.ptr(out), ws, wsn, _lib.stream_ptr()), "mv")
    for _ in range(10): mv()
    torch.cuda.synchronize(); l0 = _lib.launch_count(); t0 = time.perf_counter()
    for _ in range(200): mv()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    nl = _lib.launch_count() - l0
    print(f"chi={chi}: heff_matvec {1e6*(t2-t0)/200:.1f} us per call ({nl/200:.0f} launches), host-side enqueue {1e6*(t1-t0)/200:.1f} us")
    A = _lib.to_device(rs.randn(96, 32) + 0j); B = _lib.to_device(rs.randn(32, 128) + 0j); C = _lib.empty((96, 128))
    for _ in range(10): _lib.zgemm(A, B, C, 96, 128, 32, 32, 128, 128)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(500): _lib.zgemm(A, B, C, 96, 128, 32, 32, 128, 128)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"   small zgemm: {1e6*(t2-t0)/500:.1f} us per call, host enqueue {1e6*(t1-t0)/500:.1f} us")
    E, theta = iDMRG.ground_state(L, W, R, seed=1)   # non-hermitian junk env: just exercise; ignore result
