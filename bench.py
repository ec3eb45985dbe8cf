"""Benchmark of the hot path: batched TEBD two-site updates (fp64 complex, chi = 128, d = 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--chains C] [--impl reference]

One step = one call of mpsb_tebd_two_site over C independent chi = 128 bonds resident on the GPU (the
per-GPU shard of BASELINE config 4: 4096 chains / 8 GPUs = 512).  Weak scaling: every rank owns C chains, no
data-path collective.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHI, D, TRUN, ERR = 128, 2, 128, 1e-14
METRIC = "two-site updates/sec (fp64, chi=128, batched)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (profiling recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_inputs(n, seed0):
    """n synthetic saturated bonds (host, complex128).  Recipe shared with the CPU leg (oracle/cpu_bench.py) but
    re-stated here so the GPU arm does not import anything under oracle/."""
    import scipy.linalg as la
    Bi = np.empty((n, CHI, D, CHI), dtype=complex)
    Bj = np.empty((n, CHI, D, CHI), dtype=complex)
    Sl = np.empty((n, CHI))
    U = np.empty((n, D, D, D, D), dtype=complex)
    for c in range(n):
        rs = np.random.RandomState(seed0 + c)
        for B in (Bi, Bj):
            q, _ = np.linalg.qr(rs.randn(D * CHI, CHI) + 1j * rs.randn(D * CHI, CHI))
            B[c] = q.conj().T.reshape(CHI, D, CHI)
        s = np.linalg.svd(rs.randn(CHI, CHI) + 1j * rs.randn(CHI, CHI), compute_uv=False)
        Sl[c] = s / np.linalg.norm(s)
        A = rs.randn(D * D, D * D) + 1j * rs.randn(D * D, D * D)
        U[c] = la.expm(-0.05j * (A + A.conj().T)).reshape([D] * 4)
    return Bi, Bj, Sl, U


def make_config(chains, world):
    """The `config` object of the JSON line - identical for the GPU arm and the reference arm."""
    site = CHI * D * CHI
    return {"workload": f"config 4 shard: {chains} independent chains/GPU, one saturated chi={CHI} d={D} complex128 "
                        f"two-site update each per step (trun={TRUN}, err={ERR})",
            "chains_per_gpu": chains, "chi": CHI, "d": D,
            "inputs": "right-isometric B tensors, Gaussian Schmidt spectrum, per-chain GUE gate dt=0.05",
            "l2": f"inputs {2 * chains * site * 16 / 2**20:.0f} MiB per step > 126 MB L2 (no flush needed)",
            "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"}


def cpu_leg(per_worker, ref_per_worker=None):
    """CPU baseline on the host cores: the UNMODIFIED reference (oracle/_ref/DMRG, staged by build()) when it is
    there, else the oracle port; the port's rate is always reported as `port_value`."""
    from oracle import cpu_bench
    ups, cores, total, wall = cpu_bench.run(per_worker, CHI, D, TRUN, ERR)
    port = {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port",
            "sample": f"{total} update_double calls on synthetic chi={CHI} d={D} bonds ({per_worker} per worker, one "
                      f"worker per core, OMP_NUM_THREADS=1), oracle/mps_oracle.py (numpy tensordot + LAPACK zgesdd); "
                      f"wall {wall:.1f} s; rate = updates / busiest worker's loop time (pool start-up excluded)"}
    if not cpu_bench.reference_available():
        return port
    rp = ref_per_worker or max(2, per_worker)
    ups_r, cores, total_r, wall_r = cpu_bench.run(rp, CHI, D, TRUN, ERR, kind="reference")
    return {"value": ups_r, "unit": "updates/s", "cores": cores, "kind": "reference",
            "sample": f"{total_r} calls of the unmodified DMRG.MPS.MPS.update_double (DMRG/MPS.py:428-457, staged copy "
                      f"oracle/_ref/DMRG) on synthetic chi={CHI} d={D} bonds ({rp} per worker, one worker process per "
                      f"core, OMP_NUM_THREADS=1); wall {wall_r:.1f} s; rate = updates / busiest worker's loop time "
                      f"(pool start-up excluded)",
            "port_value": ups, "port_sample": port["sample"]}


def idmrg_leg(with_cpu):
    """Secondary metric of BASELINE.json: time per iDMRG sweep (= one next_LR growth step), TFIM g = J = 1
    (DMRG/iDMRG.py:64-76) at trunc = 32 (config 1) and 128, device-resident environments; CPU = the oracle's
    matrix-free restatement with ARPACK on one core (the reference's dense H_eff build takes 9-12 s at chi = 32)."""
    import torch
    from dmrg_b200 import iDMRG, _lib as L
    from dmrg_b200.MPO import MPO
    W = (-MPO('sx') - MPO('sz') ** 2).tomatrix()
    out = {"model": "TFIM g=J=1, d=2, w=3, real MPO -> real arithmetic (DGEMM on DMMA); *_complex legs force complex128",
           "unit": "ms per next_LR"}
    for chi in (32, 128):
        Ld, Rd, Wd = iDMRG._to_device_auto(True, np.array([[[0, 0, 1]]]), np.array([[[1, 0, 0]]]), W)
        times, nmv = [], []
        for i in range(16):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            Ld, Rd, E, _ = iDMRG.next_LR_device(Ld, Rd, Wd, trunc=chi)
            torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
            nmv.append(iDMRG.last_info["n_matvec"])
        out[f"chi{chi}"] = {"ms": 1e3 * statistics.mean(times[-6:]), "matvecs": statistics.mean(nmv[-6:]),
                            "energy_per_site_step16": E / 32}
    # stored-state run with wavefunction prediction (dmrg_b200.iDMRG.Chain): same energies, fewer H_eff products
    def chain_leg(Wm, chi, steps, tail):
        ch = iDMRG.Chain(Wm, trunc=chi)
        times = []
        for i in range(steps):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ch.step()
            torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
        w, d = Wm.shape[0], Wm.shape[2]
        nmv = statistics.mean(ch.n_matvec[-tail:])
        sec = statistics.mean(times[-tail:])
        c = 2.0 if ch.real else 8.0          # flops per multiply-add, real / complex
        flop = c * (nmv * (2 * w * d * d * chi ** 3 + 2 * w * w * d ** 3 * chi ** 2) + 4 * w * d * chi ** 3 + 2 * w * w * d * d * chi ** 2)
        return {"ms": 1e3 * sec, "matvecs": nmv, "bond": int(ch.A.shape[2]), "one_minus_fidelity": 1 - ch.fidelity[-1],
                "e_bond": ch.energy_per_site(), "contraction_tflops": flop / sec / 1e12,
                "arithmetic": "real (c = 2)" if ch.real else "complex (c = 8)"}
    for chi in (32, 128):
        out[f"chi{chi}_predict"] = chain_leg(W, chi, 16, 6)
    from dmrg_b200 import spin
    Wh = (2 * MPO('P', op=spin.P()) * MPO('N', op=spin.N()) + 2 * MPO('N', op=spin.N()) * MPO('P', op=spin.P())
          + MPO('sz') ** 2).tomatrix()
    ss = sum(np.kron(a, a) for a in spin.S(1)[1:])
    from dmrg_b200.MPO import two_site_mpo
    Wa = two_site_mpo((ss + ss @ ss / 3).real, 3).astype(complex)
    ak = chain_leg(Wa, 64, 10, 3)
    ak.update(model="AKLT spin-1 (config 2), d=3, w=%d" % Wa.shape[0], e_bond_error=abs(ak["e_bond"] + 2.0 / 3.0))
    out["aklt_chi64_predict"] = ak
    out["heisenberg_chi1024_predict"] = dict(chain_leg(Wh, 1024, 13, 2), model="Heisenberg (config 5), d=2, w=5, real MPO",
                                             note="contraction_tflops on SURVEY 8d's F_iDMRG-step with c = 2 (real arithmetic)")
    iDMRG.REAL_PATH = False                  # the same run forced through the complex128 kernels, for comparison
    try:
        out["chi128_predict_complex"] = chain_leg(W, 128, 16, 6)
        out["heisenberg_chi1024_predict_complex"] = chain_leg(Wh, 1024, 13, 2)
    finally:
        iDMRG.REAL_PATH = True
    if with_cpu:
        from threadpoolctl import threadpool_limits
        from oracle import idmrg_oracle as io_
        Lc, Rc = io_.boundary(3)
        t = []
        with threadpool_limits(limits=1):        # 64 x 64 matrices: BLAS threads only get in each other's way
            for i in range(8):                   # bond dimension 32 is reached at step 5
                t0 = time.perf_counter()
                Lc, Rc, Ec = io_.next_LR(Lc, Rc, W, trunc=32)
                t.append(time.perf_counter() - t0)
        out["cpu_chi32"] = {"ms": 1e3 * statistics.mean(t[-2:]), "kind": "port (matrix-free numpy + ARPACK, 1 process, 1 BLAS thread)",
                            "note": "the reference's own next_LR builds the dense 4096^2 H_eff: 9-12 s per step (SURVEY section 6)"}
    return out


def idmrg_batch_leg(n_chains=256, chi=128):
    """BASELINE config 4 on the iDMRG side: a TFIM coupling sweep (g / J on np.linspace(0.2, 2.0, n)) of
    n_chains independent chains grown in lock step by dmrg_b200.iDMRG.ChainBatch; ms per growth step at bond
    dimension chi and the contraction rate of the whole step on SURVEY 8(d)'s F_iDMRG-step formula (complex flops)."""
    import torch
    from dmrg_b200 import iDMRG, _lib as L_
    from dmrg_b200.MPO import MPO
    gs = np.linspace(0.2, 2.0, n_chains)
    Ws = np.stack([(-g * MPO('sx') - MPO('sz') ** 2).tomatrix() for g in gs])
    run = iDMRG.ChainBatch(Ws, trunc=chi)
    times = []
    while True:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        run.step()
        torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
        if int(run.L.shape[1]) >= chi and len(times) >= int(np.log2(chi)) + 3:
            break
    w, d = 3, 2
    # the batched H_eff product alone (DMRG/iDMRG.py:38-40 + the dense matvec inside eigsh, :57), CUDA events
    lib = L_.load()
    theta = torch.randn((n_chains, chi, d, d, chi), dtype=torch.float64 if run.real else torch.complex128, device=run.L.device)
    out = torch.empty_like(theta)
    nel = chi * d * d * chi
    ws, wsn = L_.workspace(lib.mpsb_heff_batched_workspace(n_chains, chi, chi, d, w))
    mv = lib.mpsb_heff_matvec_batched_real if run.real else lib.mpsb_heff_matvec_batched
    c = 2.0 if run.real else 8.0             # flops per multiply-add, real / complex

    def matvec():
        L_.check(mv(n_chains, chi, chi, d, w, L_.ptr(run.L), chi * chi * w, L_.ptr(run.W), w * w * d * d,
                    L_.ptr(run.R), chi * chi * w, L_.ptr(theta), nel, L_.ptr(out), nel, ws, wsn,
                    L_.stream_ptr()), "mpsb_heff_matvec_batched")
    for _ in range(3):
        matvec()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        matvec()
    e1.record(); torch.cuda.synchronize()
    mv_ms = e0.elapsed_time(e1) / 20
    mv_flop = n_chains * c * (2 * w * d * d * chi ** 3 + 2 * w * w * d ** 3 * chi ** 2)
    nmv = statistics.mean(run.n_matvec[-2:])
    cmv = statistics.mean(run.chain_matvecs[-2:])      # products actually formed (converged chains leave the batch)
    sec = statistics.mean(times[-2:])
    flop = c * (cmv * (2 * w * d * d * chi ** 3 + 2 * w * w * d ** 3 * chi ** 2) + n_chains * (4 * w * d * chi ** 3 + 2 * w * w * d * d * chi ** 2))
    e = run.energy_per_site()
    return {"workload": f"{n_chains} TFIM chains, g/J = 0.2 .. 2.0, chi = {chi}, d = 2, w = 3, "
                        f"{'real arithmetic (real MPO)' if run.real else 'complex128'}, lock-step growth",
            "flops_per_multiply_add": c,
            "ms_per_growth_step": 1e3 * sec, "chain_steps_per_s": n_chains / sec, "matvecs_per_step": nmv,
            "matvecs_per_chain_mean": cmv / n_chains,
            "contraction_tflops": flop / sec / 1e12,
            "matvec_only": {"ms": mv_ms, "tflops": mv_flop / (mv_ms * 1e-3) / 1e12,
                            "note": "mpsb_heff_matvec_batched(_real) alone (incl. its per-call MPO / environment preparation), "
                                    "SURVEY 8(d) flops c (2 w d^2 chi^3 + 2 w^2 d^3 chi^2) per chain"},
            "e_site_g0.2": float(e[0]), "e_site_g2.0": float(e[-1]),
            "e_site_g1_exact": -4.0 / np.pi}


def tebd_chain_leg():
    """BASELINE config 3 shape: ONE chain, L = 128, chi = 128, d = 2 - every half sweep is a batch of 63/64 saturated
    bonds through MPS.iTEBD_double (random canonical state, TFIM gate); ms per Trotter step = two half sweeps."""
    import torch
    from dmrg_b200.MPS import MPS
    from dmrg_b200.spin import sigma
    rs = np.random.RandomState(7)
    Lc = 128
    x = [min(2 ** i, 2 ** (Lc - i), CHI) for i in range(Lc + 1)]
    s = MPS([rs.randn(x[i], D, x[i + 1]) + 1j * rs.randn(x[i], D, x[i + 1]) for i in range(Lc)], trun=CHI, err=1e-14)
    s.canon()
    Z, X, I2 = sigma[3], sigma[1], np.eye(2)
    Hb = -np.kron(Z, Z) - 0.5 * (np.kron(X, I2) + np.kron(I2, X))
    s.iTEBD_double(Hb, 0.05, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    k = 4
    s.iTEBD_double(Hb, 0.05 * k, k)
    torch.cuda.synchronize(); el = time.perf_counter() - t0
    return {"workload": "one chain L=128 chi=128 d=2, iTEBD_double (2nd order), all bonds saturated",
            "ms_per_trotter_step": 1e3 * el / k, "bond_updates_per_s": (2 * k + 1) * (Lc - 1) / 2 / el,
            "max_bond": int(s.x.max())}


def finite_dmrg_leg():
    """Finite-system two-site DMRG (dmrg_b200.DMRG.FiniteChain), TFIM g = J = 1, L = 32, chi = 64: ms per bond
    optimisation in the first and in the third sweep (the latter restarts every solve from converged tensors)."""
    import torch
    from dmrg_b200 import DMRG
    from dmrg_b200.MPO import MPO
    W = (-MPO('sx') - MPO('sz') ** 2).tomatrix()
    ch = DMRG.FiniteChain(W, 32, trunc=64)
    t = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ch.sweep_right(); ch.sweep_left()
        torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    nb = 2 * 31
    return {"workload": "TFIM g=J=1, L=32, chi=64, complex128, full sweeps (right + left)", "ms_per_bond_sweep1": 1e3 * t[0] / nb,
            "ms_per_bond_sweep3": 1e3 * t[2] / nb, "matvecs_per_bond_sweep3": statistics.mean(ch.n_matvec[-nb:]),
            "energy": ch.energies[-1], "max_discarded_weight": max(ch.discarded[-nb:])}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (rank 0 only):
    the unmodified DMRG.MPS.MPS.update_double when oracle/_ref/DMRG was staged, else the oracle port.  Every step is
    a bounded sample of the workload; `value` is the mean rate over the timed steps."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import __graft_entry__ as g
    g.stage_reference()
    times, rates = [], []
    for _ in range(args.warmup):
        cpu_leg(2, 1)
    cb = None
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cb = cpu_leg(args.cpu_updates)
        times.append(time.perf_counter() - t0)
        rates.append(cb["value"])
    cb = dict(cb, value=statistics.mean(rates), per_step_values=rates)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": make_config(args.chains, args.gpus),
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "updates/s", "h2d_bytes_per_step": 0,
                                        "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chains", type=int, default=512, help="independent chi=128 bonds per GPU per step")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cpu-updates", type=int, default=24, help="oracle updates per CPU worker in the baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-idmrg", action="store_true", help="skip the secondary metric (time per iDMRG growth step)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    g.build()
    from dmrg_b200 import _lib as L
    lib = L.load()
    import ctypes
    numa_cpus = L.bind_host_thread_to_gpu() if world > 1 else None     # pinned buffers on the GPU's NUMA node

    n = args.chains
    Bi_h, Bj_h, Sl_h, U_h = make_inputs(n, 1000 + rank * n)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    Bi_p, Bj_p, Sl_p, U_p = pin(Bi_h), pin(Bj_h), pin(Sl_h), pin(U_h)
    dev = torch.device("cuda", local)
    Bi, Bj, Sl, U = (t.to(dev) for t in (Bi_p, Bj_p, Sl_p, U_p))
    Bio, Bjo = torch.empty_like(Bi), torch.empty_like(Bj)
    Sro = torch.empty((n, CHI), dtype=torch.float64, device=dev)
    rk = torch.empty((n,), dtype=torch.int32, device=dev)
    Bio_p, Bjo_p = torch.empty_like(Bi_p).pin_memory(), torch.empty_like(Bj_p).pin_memory()
    Sro_p, rk_p = torch.empty((n, CHI), dtype=torch.float64).pin_memory(), torch.empty((n,), dtype=torch.int32).pin_memory()
    ws, wsn = L.workspace(lib.mpsb_tebd_two_site_workspace(n, CHI, CHI, CHI, D))
    site = CHI * D * CHI

    def step():
        L.check(lib.mpsb_tebd_two_site(n, CHI, CHI, CHI, D, L.ptr(Bi), site, L.ptr(Bj), site, L.ptr(Sl), CHI, L.ptr(U),
                                       D ** 4, TRUN, ERR, CHI, L.ptr(Bio), site, L.ptr(Bjo), site, L.ptr(Sro), CHI,
                                       L.ptr(rk), ws, wsn, L.stream_ptr()), "mpsb_tebd_two_site")

    from dmrg_b200.batched import HostBondUpdater
    host_update = HostBondUpdater(n, CHI, CHI, CHI, D, TRUN, ERR)

    def e2e_run(k):
        """k batches streamed through the package's host-buffer entry point: every batch is copied host -> device,
        updated and copied device -> host inside the timed region; the copies of neighbouring batches overlap the
        kernels (double-buffered staging, see HostBondUpdater)."""
        host_update.stage(Bi_p, Bj_p, Sl_p, U_p)
        for i in range(k):
            if i + 1 < k:
                host_update.stage(Bi_p, Bj_p, Sl_p, U_p)
            host_update.run(Bio_p, Bjo_p, Sro_p, rk_p)
        host_update.finish()

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        fence()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step()
    # --- device-resident throughput, with phase timing on the library's stream --------------------------------
    lib.mpsb_profile_enable(int(os.environ.get('MPSB_PROFILE_LEVEL', '1')))
    phases = np.zeros(6)
    rot = sweeps = 0
    launches0 = L.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    fence()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
        ph = (ctypes.c_double * 6)()
        r_, s_ = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
        L.check(lib.mpsb_profile_read(ph, ctypes.byref(r_), ctypes.byref(s_)), "mpsb_profile_read")
        phases += np.array(list(ph)); rot += r_.value; sweeps += s_.value
    e1.record()
    fence()
    clocks = sampler.stop()
    launches = L.launch_count() - launches0
    lib.mpsb_profile_enable(0)
    ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total = float(ms_t.item())
    value = world * n * args.steps / (ms_total * 1e-3)
    ranks_ok = bool((rk == CHI).all().item())
    sv_sweeps = lib.mpsb_last_svd_sweeps()

    # --- end to end through host buffers ---------------------------------------------------------------------------
    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    e2e_value = world * n * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in (Bi_p, Bj_p, Sl_p, U_p))
    d2h = sum(t.numel() * t.element_size() for t in (Bio_p, Bjo_p, Sro_p, rk_p))

    # per-chain observables of the last step, computed through the batched transfer entry point and gathered over the
    # ranks with the path's one collective (outside the timed region): entanglement entropy of the new bond, kept rank,
    # <sigma_z> on the left site and the TFIM bond energy <-ZZ - (XI + IX)/2> of the updated pair
    # (DMRG/MPS.py:358-410 with E_0 = diag(Sl^2); the synthetic bonds have open ends, so these are the reference's
    # `corr` / `measure` sums for the two updated tensors taken as they are)
    from dmrg_b200.batched import gather_observables

    def transfer(E, A, Bt, op, dd):
        out = torch.empty((n, A.shape[-1], Bt.shape[-1]), dtype=torch.complex128, device=dev)
        wsb, wsbn = L.workspace(lib.mpsb_transfer_batched_workspace(n, CHI, CHI, dd, A.shape[-1], Bt.shape[-1]))
        opd = L.to_device(op)
        L.check(lib.mpsb_transfer_step_batched(n, CHI, CHI, dd, A.shape[-1], Bt.shape[-1], L.ptr(E), CHI * CHI, L.ptr(A),
                                               A[0].numel(), L.ptr(Bt), Bt[0].numel(), L.ptr(opd), 0, L.ptr(out),
                                               out[0].numel(), wsb, wsbn, L.stream_ptr()), "mpsb_transfer_step_batched")
        return out
    E0 = torch.zeros((n, CHI, CHI), dtype=torch.complex128, device=dev)
    E0.diagonal(dim1=1, dim2=2).copy_(L.to_device(L.to_host(Sl) ** 2))
    sz, sx, i2 = np.diag([1.0, -1.0]), np.array([[0.0, 1.0], [1.0, 0.0]]), np.eye(2)
    z_left = transfer(E0, Bio, Bio, sz, D).diagonal(dim1=1, dim2=2).contiguous()
    blk = torch.empty((n, CHI, D * D, CHI), dtype=torch.complex128, device=dev)
    L.zgemm(Bio, Bjo, blk, CHI * D, D * CHI, CHI, CHI, D * CHI, D * CHI, batch1=n, sA=(site, 0), sB=(site, 0),
            sC=(CHI * D * D * CHI, 0))
    hb = -np.kron(sz, sz) - 0.5 * (np.kron(sx, i2) + np.kron(i2, sx))
    e_bond = transfer(E0, blk, blk, hb.T, D * D).diagonal(dim1=1, dim2=2).contiguous()
    p2 = L.to_host(Sro) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        ent = -np.where(p2 > 0, p2 * np.log(p2), 0.0).sum(axis=1)
    local = np.stack([ent, L.to_host(rk).astype(float), L.to_host(z_left).sum(axis=1).real,
                      L.to_host(e_bond).sum(axis=1).real], axis=1)
    obs = gather_observables(torch.from_numpy(local).to(dev))
    obs_summary = {"chains": int(obs.shape[0]), "mean_entropy": float(obs[:, 0].mean().item()),
                   "mean_rank": float(obs[:, 1].mean().item()), "mean_sigma_z_left": float(obs[:, 2].mean().item()),
                   "mean_bond_energy": float(obs[:, 3].mean().item()),
                   "gathered": "entropy, rank, <sigma_z>, <h_bond> per chain, one all-gather"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # --- FP64 roofline denominator: cuBLAS DGEMM measured here (MEASURED_PEAKS.json has no FP64 entry) ---------
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); torch.matmul(a, b); t1.record(); torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    fp64_peak = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
    del a, b

    nrow, ncol = CHI * D, D * CHI
    npairs = nrow * (nrow - 1) // 2
    k_steps = args.steps
    # Algorithmic flops of the SVD phase.  Primary figure: SURVEY 8(d)'s count for one truncated SVD with vectors,
    # 4 * 21 n^3 real flops at n = d chi = 256 -> 1.409 GFLOP, over the time of the WHOLE SVD phase (QR preconditioning
    # + Jacobi sweeps) - it does not grow with the number of sweeps.  Second figure ("lean"): what one-sided Jacobi has
    # to do per visited pair with cached norms and scaled rotations (inner product 8 flops per column; 2 x 2 update of
    # two complex rows 16 flops per column), from the library's own sweep / rotation counters, over the Jacobi time.
    svd_flop = k_steps * n * 4.0 * 21.0 * float(nrow) ** 3
    svd_s = (phases[2] + phases[3]) * 1e-3
    jac_flop = sweeps * npairs * 8.0 * ncol + rot * 16.0 * ncol
    jac_flop_textbook = sweeps * npairs * 16.0 * ncol + rot * 24.0 * ncol
    jac_s = phases[3] * 1e-3
    gemm_flop = k_steps * n * 8.0 * (nrow * CHI * ncol + nrow * ncol * CHI)
    gemm_s = (phases[0] + phases[5]) * 1e-3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    traffic = None
    try:   # dram bytes per launch of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2d_dominant_kernel_traffic.json")))
        # captured at 256 problems per launch; a launch of this run covers n problems (same bytes per problem)
        traffic = int(tj["dram_bytes_per_launch"] * n / tj["problems_in_capture"])
    except (OSError, ValueError, KeyError):
        pass
    peak_note = ("cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; profiles/r2_fp64_peak.json "
                 "holds the DFMA / DMMA / mixed micro-benchmark of this pool: 35.5 / 37.1 / 33.8 TFLOP/s - DMMA and DFMA "
                 "share one execution unit on this part)")
    roofline = {"kernel": "jacobi_quad_kernel (all Jacobi sweeps of one step: modulus ordering on 16-row super blocks, four cross tasks "
                          "per CTA; self tasks on a side stream next to the launches; calls below 1776 CTAs per step use "
                          "jacobi_cross_kernel)",
                "bound": "fp64",
                "achieved": svd_flop / svd_s / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": svd_flop / svd_s / 1e12 / fp64_peak,
                "traffic": traffic,
                "traffic_note": "DRAM read + write bytes of one step launch (ncu --set full, profiles/r2d_dominant_kernel_traffic.json: "
                                "482 MB at 256 problems against 537 MB algorithmic), scaled to this run's problems per launch",
                "flop_model": "SURVEY 8(d): 4*21*n^3 = 1.409 GFLOP per truncated SVD (n = 256), over QR + Jacobi time",
                "lean": {"achieved": jac_flop / jac_s / 1e12, "frac": jac_flop / jac_s / 1e12 / fp64_peak,
                         "flop_model": "8*L per visited pair + 16*L per rotation (L = 256 columns), library counters, Jacobi time only",
                         "achieved_textbook": jac_flop_textbook / jac_s / 1e12},
                "peak_source": peak_note,
                "share_of_step": float((phases[2] + phases[3]) / phases.sum())}
    roof_gemm = {"kernel": "zgemm_kernel (theta formation + back-projection, DMMA m8n8k4)", "bound": "tensor",
                 "achieved": gemm_flop / gemm_s / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                 "frac": gemm_flop / gemm_s / 1e12 / fp64_peak, "traffic": None,
                 "share_of_step": float((phases[0] + phases[5]) / phases.sum())}
    hbm_bytes = k_steps * n * (2 * site * 16 + CHI * 8 + 2 * site * 16 + CHI * 8)
    line = {"metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": make_config(n, world), "host_thread_bound_to_gpu_numa_node": bool(numa_cpus),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "roofline": roofline, "roofline_contractions": roof_gemm,
            "phases_ms_per_step": {k: float(v / k_steps) for k, v in zip(
                ["theta_gemm", "gate", "qr", "jacobi", "finalise", "backproj_gemm"], phases)},
            "svd": {"sweeps_slowest_problem": int(sv_sweeps), "rotations_per_update": rot / (k_steps * n),
                    "sweeps_per_update": sweeps / (k_steps * n), "all_ranks_full": ranks_ok},
            "observables": obs_summary,
            "hbm_mandatory_gbs": hbm_bytes / (ms_total * 1e-3) / 1e9,
            "hbm_peak_gbs": peaks.get("hbm_gbs")}
    if world == 1 and not args.no_idmrg:
        line["idmrg"] = idmrg_leg(not args.no_cpu)
        line["idmrg_batch"] = idmrg_batch_leg()
        line["tebd_chain"] = tebd_chain_leg()
        line["finite_dmrg"] = finite_dmrg_leg()
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_leg(args.cpu_updates)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
