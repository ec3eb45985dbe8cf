"""DGEMM / ZGEMM throughput at the shapes of the iDMRG contractions (development aid).  Flops: 2 M N K (real), 8 M N K (complex)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmrg_b200 import _lib

def bench(real, M, N, K, opB, batch):
    dt = torch.float64 if real else torch.complex128
    A = torch.randn((batch, M, K), dtype=dt, device="cuda")
    B = torch.randn((batch, N, K) if opB else (batch, K, N), dtype=dt, device="cuda")
    C = torch.empty((batch, M, N), dtype=dt, device="cuda")
    f = _lib.dgemm if real else _lib.zgemm
    run = lambda: f(A, B, C, M, N, K, K, K if opB else N, N, 0, opB, batch1=batch, sA=(M * K, 0), sB=(N * K, 0), sC=(M * N, 0))
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = (2.0 if real else 8.0) * M * N * K * batch
    ref = torch.matmul(A[0], B[0].T if opB else B[0])
    err = float((C[0] - ref).abs().max())
    print(f"{'d' if real else 'z'}gemm M={M} N={N} K={K} opB={opB} batch={batch}: {ms:8.3f} ms  {fl / ms / 1e9:7.2f} TFLOP/s  err {err:.1e}")

for real in (True, False):
    bench(real, 5 * 1024, 4 * 1024, 1024, 0, 1)      # chi = 1024 Heisenberg: L.theta
    bench(real, 4 * 1024, 1024, 5 * 1024, 1, 1)      # T3.R
    bench(real, 3 * 128, 4 * 128, 128, 0, 256)       # 256 TFIM chains chi = 128
    bench(real, 4 * 128, 128, 3 * 128, 1, 256)
    bench(real, 3 * 128, 4 * 128, 128, 0, 1)         # one chain chi = 128
    bench(real, 4096, 4096, 4096, 0, 1)
