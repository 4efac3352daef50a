#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json:

    iLQG iters/sec (backward+forward), batch=65536 n=32 m=8 T=256, FP64, at 1/2/4/8 B200.

One "step" = one backward sweep (ddp_back_pass_f64) + one forward rollout with alpha = 1
(ddp_forward_pass_f64) + the batch statistics (ddp_batch_stats_f64, all-reduced over NCCL when
N > 1) over 65 536 trajectories PER GPU (weak scaling; the trajectories are independent, so the
batch shards with no data-path collective -- the only exchange is the 64-byte statistics vector).
`value` counts 65 536-trajectory iterations per second summed over all ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (default N=1)
    python bench.py --impl reference ...                         # CPU arm: restated reference (C++/OpenMP port)

Besides the headline line this run measures, in the same JSON line under "configs", every other configuration of
BASELINE.json that runs on the GPUs it was given: C3 (pendulum on a cart with control limits, 262 144 trajectories, N = 1 only),
C4 (one KL-constrained iteration on C2's system at 65 536 trajectories, N = 1 only) and C5's per-GPU share (262 144
trajectories per GPU through the chunked device iteration ddp_ilqg_iter_f64; at --gpus 8 that IS config 5).  Every
configuration carries CUDA-event timings, its roofline fractions and a SAMPLED ORACLE CHECK AT FULL SIZE: >= 32 (C4: 16) random
trajectories of the very batch that was timed are re-computed by the CPU oracle (oracle/ddp_oracle.py, outside every timed
region, as the checker only) and the largest element-wise relative error of K, k, Vx, Vxx1, dV, xnew, unew, cost is reported
together with the exact equality of `diverge` and of the clamped sets.

Inputs are synthetic (seeded): per-trajectory LTI dynamics A_b = exp(h(G-G')), B_b = h N(0,1),
Q = hI, R = 0.1hI, x pre-rolled from x0 = 1 + 0.1 N(0,1) with u = 0.1 N(0,1), cx = Qx, cu = Ru
(SURVEY.md section 8d, config C2).  The working set (~56 GB) is far larger than the 126 MB L2,
so no explicit L2 flush is needed between timed iterations.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_X, M_U, T_H = 32, 8, 256
BATCH = 65536
H_STEP = 0.01
# algorithmic work per trajectory-iteration, SURVEY.md section 8(d) / DESIGN.md
FLOPS_BACK_STEP = 213419.0            # per backward timestep at n=32, m=8 (reference formulation)
FLOPS_BACK_STEP_SYM = 336 * 512.0     # the same step with symmetric products counted once = what the tile kernel issues
FLOPS_FWD_STEP = 5264.0
BYTES_BACK = 92168.0 + 606228.0       # backward read + write per trajectory
BYTES_FWD = 632832.0 + 81928.0        # forward read + write per trajectory
FP64_TENSOR_PEAK_FALLBACK = 37.08     # profiles/microbench/ubench_r01_b200.txt; used only if the in-run self test fails
FP64_DFMA_PEAK_FALLBACK = 34.14
# config 3 (n=4, m=1, T=600, lims): SURVEY.md 8d
C3_T, C3_B = 600, 262144
C3_BYTES_BACK = 124824.0 + 43220.0
C3_BYTES_FWD = 48000.0 + 28808.0
# config 4: KL-augmented sweep, rollout, KL evaluation as a separate pass
C4_BYTES_BACK = 747536.0 + 868372.0
C4_BYTES_KL = 8.0 * (2 * 256 * 256 + 2 * 256 * 32 + 256 * 8 + 3 * 256 * 64 + 1)    # K_prev, K_new, xnew, xold, k_new, Sig_new, Sig_prev, Sigi_prev
C4_DMMA_BACK = 336 + 4 + 4 + 20       # tiles per step of bp_tile32x8_kernel<GPS>: + Sigma_i K_prev (8) and K_prev' S (20)
C4_DMMA_KL = 248
C5_B = 262144
METRIC = "iLQG iters/sec (backward+forward), batch=65536 n=32 m=8 T=256"
WORKLOAD = ("C2: batched LTI linear dynamics n=32 m=8 T=256, 65536 trajectories per GPU, lambda=1 regType=1 "
            "no lims, alpha=1 (BASELINE.json configs[1])")
UNIT = "iters/s (1 iter = backward+forward sweep over 65536 trajectories, FP64)"
PARITY_TOL = 1e-8                     # north star: K, k, Vx, Vxx within 1e-8 relative; integer outcomes exact


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU arm: the restated reference (oracle/cpu_ref.cpp) on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------

def cpu_sample_inputs(nsample, seed=1234):
    import scipy.linalg as sla
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((nsample, N_X, N_X))
    A = np.stack([sla.expm(H_STEP * (g - g.T)) for g in G])
    Bm = H_STEP * rng.standard_normal((nsample, N_X, M_U))
    x0 = 1.0 + 0.1 * rng.standard_normal((nsample, N_X))
    u = 0.1 * rng.standard_normal((nsample, T_H, M_U))
    x = np.zeros((nsample, T_H, N_X))
    x[:, 0] = x0
    for t in range(T_H - 1):
        x[:, t + 1] = np.einsum("bij,bj->bi", A, x[:, t]) + np.einsum("bia,ba->bi", Bm, u[:, t])
    Q = H_STEP * np.eye(N_X)
    R = 0.1 * H_STEP * np.eye(M_U)
    return A, Bm, Q, R, x, u, x @ Q.T, u @ R.T


def cpu_step_time(inputs, nthreads):
    """One backward + forward sweep of the C++/OpenMP restated reference over the sample."""
    from oracle import cpu_ref as CR
    A, Bm, Q, R, x, u, cx, cu = inputs
    fx = np.ascontiguousarray(np.swapaxes(A, -1, -2))
    fu = np.ascontiguousarray(np.swapaxes(Bm, -1, -2))
    t0 = time.perf_counter()
    dv, K, k, Vx, _, _, _, dV = CR.back_pass(cx, cu, Q.T, np.zeros((M_U, N_X)), R.T, fx, fu, 1.0, 1, None, u,
                                             want_Vxx=False, nthreads=nthreads)
    CR.forward_pass_linear(K, k, x[:, 0].copy(), x, u, 1.0, None, fx, fu, Q.T, R.T, nthreads=nthreads)
    return time.perf_counter() - t0


def run_cpu_baseline(nsample, steps, warmup, cpus=None):
    """The CPU arm.  The OpenMP thread count is set EXPLICITLY to the CPUs this process may run on (torchrun exports
    OMP_NUM_THREADS=1, which would otherwise make this a one-thread baseline)."""
    from oracle import cpu_ref as CR
    CR.load()
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
        except Exception:
            pass
    cores = CR.host_cores()
    inputs = cpu_sample_inputs(nsample)
    for _ in range(warmup):
        cpu_step_time(inputs, cores)
    ts = [cpu_step_time(inputs, cores) for _ in range(steps)]
    t = float(np.mean(ts))
    iters_per_s = 1.0 / (t * BATCH / nsample)       # trajectories are independent: linear extrapolation to 65 536
    return dict(value=iters_per_s, unit=UNIT, cores=cores, kind="port", extrapolated=True, sample_trajectories=nsample,
                sample_seconds_per_step=t, build="g++ -O3 -march=native -fopenmp (oracle/Makefile), rebuilt on this host",
                sample=f"{nsample} of 65536 trajectories per step (n=32,m=8,T=256), {t:.3f} s/step on {cores} OpenMP threads "
                       f"(set explicitly from the process's CPU affinity), scaled linearly to the full batch; C++/OpenMP restated "
                       f"reference (oracle/cpu_ref.cpp), not Julia"), t


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, args.steps)
    cb, t = run_cpu_baseline(args.cpu_sample, steps, min(args.warmup, 2))
    line = dict(impl="reference", metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps,
                warmup=min(args.warmup, 2), ms_per_step=1e3 / cb["value"], higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD,
                            impl_note="restated reference (C++/OpenMP port of back_pass + forward_pass), all host threads; "
                                      "each step is a bounded sample of the workload, the value is extrapolated linearly"),
                cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in out.strip().splitlines():
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        load = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return dict(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), samples=len(sm),
                    reasons=sorted(reasons))


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI device) before any pinned host
    buffer is allocated, so that the e2e path's host memory sits on the GPU's NUMA node.  Returns the CPU set or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        bus = bus[-12:] if len(bus) > 12 else bus            # 00000000:1b:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


# ---- sampled oracle checks (outside every timed region; the oracle is the checker, never the thing measured) --------------

def _relerr_elem(a, b, floor=1e-4):
    """Largest element-wise relative error |a-b| / max(|b|, floor*max|b|) (entries below `floor` of the tensor's largest are
    measured against that floor: FP64 cancellation makes a relative figure on them meaningless)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    if not (np.all(np.isfinite(a)) and np.all(np.isfinite(b))):
        return 0.0 if np.array_equal(a, b, equal_nan=True) else float("inf")
    scale = float(np.max(np.abs(b)))
    den = np.maximum(np.abs(b), floor * (scale if scale > 0 else 1.0))
    return float(np.max(np.abs(a - b) / den))


def _h(t):
    return t.detach().cpu().numpy()


def _math(t):                      # device column-major blocks -> math layout (swap the last two axes)
    return np.swapaxes(_h(t), -1, -2)


class ErrTable:
    def __init__(self):
        self.err, self.exact = {}, {}

    def rel(self, name, got, ref, floor=1e-4):
        self.err[name] = max(self.err.get(name, 0.0), _relerr_elem(got, ref, floor))

    def same(self, name, ok):
        self.exact[name] = bool(self.exact.get(name, True) and ok)

    def report(self, nsamp, tol=PARITY_TOL, loose=()):
        worst = max([v for k, v in self.err.items() if k not in loose] + [0.0])
        return dict(samples=nsamp, tolerance=tol, max_elementwise_rel_err=self.err, exact=self.exact,
                    within_tolerance=bool(worst <= tol and all(self.exact.values())), worst=worst,
                    error_measure="max over sampled trajectories and elements of |gpu - oracle| / max(|oracle|, 1e-4 max|oracle|)",
                    oracle="oracle/ddp_oracle.py (NumPy restatement of the reference), run on the host after the timed region")


def check_linear(idx, fx, fu, x, u, cx, cu, Q, R, lam, K, k, Vx, Vxx1, dV, diverge, xnew, unew, cost, alpha=1.0):
    """C2 / C5: oracle back_pass + forward_pass on the sampled trajectories `idx` of the timed batch."""
    from oracle import ddp_oracle as O
    import torch
    ii = torch.as_tensor(idx, device=x.device)
    g = lambda t: None if t is None else _h(t[ii])
    A_s, B_s = np.swapaxes(g(fx), -1, -2), np.swapaxes(g(fu), -1, -2)
    x_s, u_s, cx_s, cu_s = g(x), g(u), g(cx), g(cu)
    K_s = None if K is None else np.swapaxes(g(K), -1, -2)
    k_s, Vx_s, V1_s = g(k), g(Vx), (None if Vxx1 is None else np.swapaxes(g(Vxx1), -1, -2))
    dV_s, dv_s, xn_s, un_s, c_s, lam_s = g(dV), g(diverge), g(xnew), g(unew), g(cost), g(lam)
    Qh, Rh = _h(Q), _h(R)
    n, m = Qh.shape[0], Rh.shape[0]
    E = ErrTable()
    for j in range(len(idx)):
        cxj = cx_s[j] if cx_s is not None else x_s[j] @ Qh.T
        cuj = cu_s[j] if cu_s is not None else u_s[j] @ Rh.T
        d0, p0, Vx0, Vxx0, dV0 = O.back_pass(cxj, cuj, Qh, np.zeros((n, m)), Rh, A_s[j], B_s[j], float(lam_s[j]), 1, None, x_s[j], u_s[j])
        om = O.LinearModel(A_s[j], B_s[j], Qh, Rh)
        xn0, un0, c0 = O.forward_pass(p0, x_s[j, 0], u_s[j], x_s[j], alpha, om.f, om.costfun, None)
        E.same("diverge", int(dv_s[j]) == d0)
        if K_s is not None:
            E.rel("K", K_s[j], p0.K); E.rel("k", k_s[j], p0.k); E.rel("Vx", Vx_s[j], Vx0)
        if V1_s is not None:
            E.rel("Vxx1", V1_s[j], Vxx0[0])
        E.rel("dV", dV_s[j], dV0); E.rel("xnew", xn_s[j], xn0); E.rel("unew", un_s[j], un0); E.rel("cost", c_s[j], c0)
    return E.report(len(idx))


def _nvml_free_gb(dev):
    import torch
    free, total = torch.cuda.mem_get_info(dev)
    return free / 2**30, total / 2**30


def main_gpu(args):
    import torch
    import torch.distributed as dist
    import ddp_b200 as ddp
    from ddp_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this arm has no CPU fallback (use --impl reference for the CPU arm)")
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    n, m, T = N_X, M_U, T_H
    f64 = torch.float64
    want = set(s for s in args.configs.split(",") if s)
    empty = lambda *s: torch.empty(*s, dtype=f64, device=dev)
    tn = lambda t_, sb, st: L.Tensor(t_.data_ptr(), sb, st)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    stream_ptr = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=f64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gen_linear(Bg, seed, chunk=32768):
        """Per-trajectory LTI systems of SURVEY 8d (A = exp(h(G-G')), B = h N(0,1)), x0, u -- generated on the device in chunks."""
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        fx, fu = empty(Bg, n, n), empty(Bg, m, n)
        x0, u = empty(Bg, n), empty(Bg, T, m)
        for b0 in range(0, Bg, chunk):
            nb = min(chunk, Bg - b0)
            G = torch.randn(nb, n, n, dtype=f64, device=dev, generator=gen)
            A = torch.linalg.matrix_exp(H_STEP * (G - G.transpose(1, 2)))
            fx[b0:b0 + nb] = A.transpose(1, 2)               # column-major per trajectory == reference layout (n,n,B)
            fu[b0:b0 + nb] = (H_STEP * torch.randn(nb, n, m, dtype=f64, device=dev, generator=gen)).transpose(1, 2)
            x0[b0:b0 + nb] = 1.0 + 0.1 * torch.randn(nb, n, dtype=f64, device=dev, generator=gen)
            u[b0:b0 + nb] = 0.1 * torch.randn(nb, T, m, dtype=f64, device=dev, generator=gen)
            del G, A
        return fx, fu, x0, u

    # ---- the FP64 denominators, measured on this device in this run (MEASURED_PEAKS.json has no FP64 figure)
    eng0 = ddp.Engine(n, m, T, 8, device=local_rank)
    eng0.set_stream(stream_ptr)
    try:
        dmma_peak, dmma_ms = eng0.selftest_peak("dmma", 3)
        dfma_peak, dfma_ms = eng0.selftest_peak("dfma", 3)
        try:
            mixed_peak, _ = eng0.selftest_peak("mixed", 3)
        except Exception:
            mixed_peak = None
        peak_how = (f"ddp_selftest_peak_f64 in this run, before the timed region: mma.sync.m8n8k4.f64 {dmma_peak:.2f} TFLOP/s ({dmma_ms:.2f} ms/launch), "
                    f"fma.rn.f64 {dfma_peak:.2f} TFLOP/s ({dfma_ms:.2f} ms/launch); burst figures of a kernel timed alone")
    except Exception as exc:
        dmma_peak, dfma_peak, mixed_peak = FP64_TENSOR_PEAK_FALLBACK, FP64_DFMA_PEAK_FALLBACK, None
        peak_how = f"fallback constants of profiles/microbench/ubench_r01_b200.txt (self test failed: {exc})"
    eng0.close()

    # ---- synthetic inputs, generated on the device (seeded per rank)
    fx, fu, x0, u = gen_linear(B, 1000 + rank)
    Q = (H_STEP * torch.eye(n, dtype=f64, device=dev)).contiguous()
    R = (0.1 * H_STEP * torch.eye(m, dtype=f64, device=dev)).contiguous()
    cxu = torch.zeros(m, n, dtype=f64, device=dev)
    lam = torch.ones(B, dtype=f64, device=dev)
    x, cost0 = empty(B, T, n), empty(B)
    K, k, Vx, dV = empty(B, T, n, m), empty(B, T, m), empty(B, T, n), empty(B, 2)
    Vxx1 = empty(B, n, n)
    xnew, unew, cost = empty(B, T, n), empty(B, T, m), empty(B)
    cx, cu = empty(B, T, n), empty(B, T, m)
    diverge = torch.empty(B, dtype=torch.int32, device=dev)
    stats = torch.zeros(8, dtype=f64, device=dev)

    eng = ddp.Engine(n, m, T, B, device=local_rank)
    eng.set_stream(stream_ptr)

    model = L.Model()
    model.kind = 1
    model.A, model.Bm = tn(fx, n * n, 0), tn(fu, n * m, 0)
    model.Q, model.R = tn(Q, 0, 0), tn(R, 0, 0)
    model.flags = 1                                   # Q = h*I is diagonal (checked on the host: isdiag(Q))
    # pre-roll x with the library's own forward kernel (empty policy), which also yields cx = Qx, cu = Ru
    fa0 = L.ForwardPassArgs()
    fa0.x0, fa0.u = tn(x0, n, 0), tn(u, T * m, m)
    fa0.alpha_scalar, fa0.u_scale = 1.0, 1.0
    unew0 = empty(B, T, m)
    fa0.xnew, fa0.unew, fa0.cost, fa0.cx, fa0.cu = x.data_ptr(), unew0.data_ptr(), cost0.data_ptr(), cx.data_ptr(), cu.data_ptr()
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa0)))
    torch.cuda.synchronize()
    del unew0

    ba = L.BackPassArgs()
    ba.cx, ba.cu = tn(cx, T * n, n), tn(cu, T * m, m)
    ba.cxx, ba.cxu, ba.cuu = tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
    ba.fx, ba.fu = tn(fx, n * n, 0), tn(fu, n * m, 0)
    ba.lam, ba.reg_type = lam.data_ptr(), 1
    ba.diverge, ba.K, ba.k, ba.Vx, ba.dV = diverge.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr()
    ba.Vxx1 = Vxx1.data_ptr()
    fa = L.ForwardPassArgs()
    fa.K, fa.k = K.data_ptr(), k.data_ptr()
    fa.x0, fa.x, fa.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
    fa.alpha_scalar, fa.u_scale = 1.0, 1.0
    fa.xnew, fa.unew, fa.cost = xnew.data_ptr(), unew.data_ptr(), cost.data_ptr()

    back_ms, fwd_ms = [], []

    def step(timed):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
        e1.record()
        eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa)))
        e2.record()
        eng._ck(eng.lib.ddp_batch_stats_f64(eng.h, cost0.data_ptr(), cost.data_ptr(), dV.data_ptr(), None, 1.0,
                                            diverge.data_ptr(), None, stats.data_ptr()))
        if world > 1:                                 # the line-search cost reduction: 64 bytes over NVLink
            if lib_comm:
                eng.allreduce_stats(stats.data_ptr())  # ncclAllReduce inside libddp (ddp_comm_allreduce_stats_f64), on the handle's stream
            else:
                dist.all_reduce(stats)
        if timed:
            back_ms.append((e0, e1)); fwd_ms.append((e1, e2))

    # the library's own communicator for the one collective of the path; torch.distributed carries the 128-byte id
    lib_comm = False
    if world > 1:
        try:
            idt = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            eng.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
            lib_comm = True
        except Exception as exc:                      # fall back to torch.distributed's all-reduce
            sys.stderr.write(f"[bench] ddp_comm_init unavailable ({exc}); using torch.distributed.all_reduce\n")
        flag = torch.tensor([1.0 if lib_comm else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        lib_comm = bool(flag.item() > 0.5)
    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    torch.cuda.synchronize()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        step(True)
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(t0.elapsed_time(t1))
    launches = eng.launch_count - launches0 + (args.steps if world > 1 else 0)
    ms_per_step = total_ms / args.steps
    bk = float(np.mean([a.elapsed_time(b) for a, b in back_ms]))
    fw = float(np.mean([a.elapsed_time(b) for a, b in fwd_ms]))
    stats_h = stats.cpu().numpy()
    n_div = int((diverge > 0).sum().item())

    # ---- size-independent properties of the full-size result (outside the timed region; never fatal)
    props = {}
    try:
        expected = -(dV[:, 0] + dV[:, 1])                         # iLQG.jl:271 at alpha = 1
        ratio = (cost0 - cost) / expected
        props["ratio_min"], props["ratio_max"] = float(ratio.min().item()), float(ratio.max().item())
        props["terminal_Vx_is_cx"] = bool(torch.equal(Vx[:, T - 1], cx[:, T - 1]))          # backward_pass.jl:21
        props["terminal_gains_zero"] = bool((K[:, T - 1] == 0).all().item() and (k[:, T - 1] == 0).all().item())   # quirk Q7
        props["xnew0_is_x0"] = bool(torch.equal(xnew[:, 0], x0))                            # forward_pass.jl:13
        props["Vxx1_exactly_symmetric"] = bool(torch.equal(Vxx1, Vxx1.transpose(1, 2)))     # backward_pass.jl:71-72
    except Exception as exc:
        props["error"] = str(exc)
    # ---- sampled oracle check at full size: 32 random trajectories of the timed batch, element-wise
    oracle_check = None
    if rank == 0 and args.oracle_samples > 0:
        try:
            idx = np.sort(np.random.default_rng(7).choice(B, size=min(args.oracle_samples, B), replace=False))
            tchk = time.perf_counter()
            oracle_check = check_linear(idx, fx, fu, x, u, cx, cu, Q, R, lam, K, k, Vx, Vxx1, dV, diverge, xnew, unew, cost)
            oracle_check["seconds"] = time.perf_counter() - tchk
            oracle_check["trajectories"] = idx.tolist()
        except Exception as exc:
            oracle_check = dict(error=str(exc))

    cost_dev_h, div_dev_h = _h(cost), _h(diverge)
    results = dict()
    # what the other configurations need of this scope (an explicit dict: locals() would pin every tensor of this frame)
    common = dict(ddp=ddp, L=L, dev=dev, local_rank=local_rank, tn=tn, ev=ev, args=args, rank=rank, world=world, n=n, m=m, T=T, Q=Q, R=R,
                  gen_linear=gen_linear, barrier=barrier, max_over_ranks=max_over_ranks, lib_comm=lib_comm, dmma_peak=dmma_peak)

    # =============================================================================================================
    # C4: one KL-constrained iteration (back_pass_gps + forward + KL evaluation) on C2's system, B = 65 536, N = 1 only
    # =============================================================================================================
    if "c4" in want and world == 1:
        try:
            results["c4"] = run_c4(dict(L=L, eng=eng, dev=dev, tn=tn, empty=empty, ev=ev, n=n, m=m, T=T, B=B, fx=fx, fu=fu, x=x, u=u, cx=cx, cu=cu,
                                        Q=Q, R=R, cxu=cxu, K=K, ba=ba, model=model, xnew=xnew, unew=unew, cost=cost, args=args, rank=rank,
                                        dmma_peak=dmma_peak, ddp=ddp, local_rank=local_rank))
        except Exception as exc:
            results["c4"] = dict(error=f"{type(exc).__name__}: {exc}")
        torch.cuda.empty_cache()

    # ---- end-to-end: same step through ddp_ilqg_iter_host_f64 on pinned host buffers
    e2e = None
    host_in = {name: _h(src) for name, src in (("fx", fx), ("fu", fu), ("x", x), ("u", u), ("lam", lam))}
    Qh, Rh = _h(Q), _h(R)
    K_probe = {b: _h(K[b]) for b in (0, B // 2, B - 1)}
    cost0_h = _h(cost0)
    # =============================================================================================================
    # C2 with control limits (row a4 at the headline shape: the box-QP branch of the tile kernel) and the line search over
    # 10 step sizes (row f3: multi-alpha tensor-tile rollout vs serial rollouts), B = 65 536, N = 1 only
    # =============================================================================================================
    if world == 1:
        ctx2 = dict(L=L, eng=eng, dev=dev, tn=tn, empty=empty, ev=ev, n=n, m=m, T=T, B=B, fx=fx, fu=fu, x=x, u=u, cx=cx, cu=cu,
                    Q=Q, R=R, cxu=cxu, lam=lam, K=K, k=k, model=model, xnew=xnew, unew=unew, cost=cost, args=args, rank=rank, dmma_peak=dmma_peak,
                    bk_plain=bk)
        if "linesearch" in want:
            try:
                results["line_search"] = run_line_search(ctx2)
            except Exception as exc:
                results["line_search"] = dict(error=f"{type(exc).__name__}: {exc}")
        if "c2lims" in want:
            try:
                results["c2_lims"] = run_c2_lims(ctx2)
            except Exception as exc:
                results["c2_lims"] = dict(error=f"{type(exc).__name__}: {exc}")
        del ctx2
        torch.cuda.empty_cache()

    # free the device-resident working set: the e2e path holds its own (full-batch mirrors + the resident policy)
    del K, k, Vx, Vxx1, xnew, unew, cx, cu, x, u, fx, fu, x0, dV, cost, cost0, diverge, ba, fa, fa0, model
    eng.close()
    torch.cuda.empty_cache()
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 1 << 62
    per_traj_host = 8 * (n * n + n * m + 4 * T * n + 4 * T * m + 4) + 4
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    Be = B
    while Be > 1024 and Be * per_traj_host * local_world > 0.5 * avail:
        Be //= 2
    try:
        eng_e = ddp.Engine(n, m, T, Be, device=local_rank)
        eng_e.set_stream(stream_ptr)
        it = ddp.HostIteration(eng_e, np.zeros((n, n)), np.zeros((m, m)), reg_type=1, alpha=1.0, chunk=args.chunk, device_derivs=True,
                               keep_policy=True)
        it.Q[:] = Qh; it.R[:] = Rh
        it.args.q_diagonal = 1
        for name, src in host_in.items():
            it.bufs[name][...] = src[:Be]
        del host_in
        e_steps = max(1, min(args.steps, args.e2e_steps))
        it.run()                                       # warm-up (allocates the pipeline)
        it.run()
        barrier()
        w0 = time.perf_counter()
        for _ in range(e_steps):
            h2d, d2h = it.run()                        # synchronous: returns when the results are in host memory
        w1 = time.perf_counter()
        e_ms = max_over_ranks((w1 - w0) * 1e3 / e_steps)
        ok = bool(np.array_equal(it.bufs["diverge"], div_dev_h[:Be])) and \
            bool(np.allclose(it.bufs["cost"], cost_dev_h[:Be], rtol=1e-12, atol=0))
        # the policy of EVERY trajectory is still on the device after the call: first, middle and last compared bit for bit with
        # what ddp_back_pass_f64 produced in the device-resident run
        pol_ok = True
        for b, Kref in K_probe.items():
            if b < Be:
                Kb, _, _ = it.policy(b, b + 1)
                pol_ok = pol_ok and bool(np.array_equal(np.swapaxes(Kb[0], -1, -2), Kref))
        e2e = dict(value=world * (Be / BATCH) * 1e3 / e_ms, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                   ms_per_step=e_ms, steps=e_steps, batch_per_gpu=Be, chunk=(args.chunk or "ramped: 1,2,4,...,4,2,1 rounds of sm_count*8 trajectories"),
                   api="ddp_ilqg_iter_host_f64: x,u,fx,fu,lambda from pinned host memory -> df (cx=Qx, cu=Ru) + backward + forward on the device -> "
                       "xnew,unew,cost,dV,diverge in host memory; the policy K,k,Vx of the WHOLE batch stays resident on the device "
                       "(keep_policy = 1, ddp_iter_host_policy) -- 38.8 GB that never cross PCIe",
                   h2d_gbs=h2d / e_ms * 1e-6, d2h_gbs=d2h / e_ms * 1e-6,
                   matches_device_path=ok, policy_resident_bitwise_equal=pol_ok,
                   host_cpus_bound_to_gpu_numa_node=(len(numa_cpus) if numa_cpus else None))
        # variant: the inputs stay on the device between iterations (what a real iLQG loop does): only xnew, unew, cost cross PCIe
        try:
            it.run(commit_accepted=True, cost_prev=cost0_h[:Be])
            barrier()
            w0 = time.perf_counter()
            for _ in range(e_steps):
                h2d_r, d2h_r = it.run(inputs_resident=True)
            w1 = time.perf_counter()
            r_ms = max_over_ranks((w1 - w0) * 1e3 / e_steps)
            e2e["resident_inputs"] = dict(value=world * (Be / BATCH) * 1e3 / r_ms, ms_per_step=r_ms, h2d_bytes_per_step=h2d_r, d2h_bytes_per_step=d2h_r,
                                          d2h_gbs=d2h_r / r_ms * 1e-6,
                                          note="inputs_resident = 1: fx,fu,x,u,lambda are the device copies left by the previous call (x,u committed "
                                               "from the accepted rollout on the device); not the headline e2e, which re-uploads every input each step")
        except Exception as exc:
            e2e["resident_inputs"] = dict(error=str(exc))
        it.close()
        eng_e.close()
    except Exception as exc:                           # report, never fake
        e2e = dict(value=None, unit=UNIT, error=str(exc))
    torch.cuda.empty_cache()

    # =============================================================================================================
    # whole solve end to end: iLQG(f,costfun,df,x0,u0) moves x0,u0 up once and x,u,K,k down once (iLQG.jl:143-341)
    # =============================================================================================================
    if "c1" in want and world == 1:
        try:
            results["c1"] = run_c1(common)
        except Exception as exc:
            results["c1"] = dict(error=f"{type(exc).__name__}: {exc}")
        torch.cuda.empty_cache()

    if "solve" in want and world == 1:
        try:
            results["solve_e2e"] = run_solve_e2e(common)
        except Exception as exc:
            results["solve_e2e"] = dict(error=f"{type(exc).__name__}: {exc}")
        torch.cuda.empty_cache()

    # =============================================================================================================
    # C3: pendulum on a cart, control limits (boxQP branch), n=4 m=1 T=600, B = 262 144, N = 1 only
    # =============================================================================================================
    if "c3" in want and world == 1:
        try:
            results["c3"] = run_c3(common)
        except Exception as exc:
            results["c3"] = dict(error=f"{type(exc).__name__}: {exc}")
        torch.cuda.empty_cache()

    # =============================================================================================================
    # C5 per-GPU share: 262 144 trajectories per GPU through the chunked device iteration
    # =============================================================================================================
    if "c5" in want:
        try:
            results["c5"] = run_c5(common)
        except Exception as exc:
            results["c5"] = dict(error=f"{type(exc).__name__}: {exc}")
        torch.cuda.empty_cache()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        flops_back = FLOPS_BACK_STEP_SYM * (T - 1) * B
        ach_tf = flops_back / (bk * 1e-3) * 1e-12
        ref_tf = FLOPS_BACK_STEP * (T - 1) * B / (bk * 1e-3) * 1e-12
        roofline = dict(
            bound="tensor", kernel="bp_tile32x8_kernel (backward sweep, FP64 mma.sync m8n8k4 tiles)",
            achieved=ach_tf, peak=dmma_peak, unit="TFLOP/s", frac=ach_tf / dmma_peak,
            peak_source="FP64 tensor (DMMA) peak " + peak_how + " (MEASURED_PEAKS.json has no FP64 figure); "
                        "algorithmic flops = 172032/step (the symmetric halves of F'VF and of the Vxx update counted once: 336 m8n8k4 "
                        "tiles) x 255 steps x 65536 trajectories; SURVEY.md 8d's reference formulation (213419/step, full products) is "
                        "reported as reference_formulation_tflops and would read 1.0+ of the peak",
            reference_formulation_tflops=ref_tf, reference_formulation_frac=ref_tf / dmma_peak,
            fp64_dfma_peak=dfma_peak, fp64_dfma_frac=ach_tf / dfma_peak,
            fp64_mixed_tflops=mixed_peak,
            fp64_pipe_note="fp64_mixed_tflops: 8 DMMA + 32 DFMA per warp-iteration interleaved, all flops counted -- it stays below the DMMA peak, "
                           "i.e. tensor tiles and scalar FP64 instructions share ONE datapath on this GPU: the sweep's ~205 scalar FP64 instructions "
                           "per step (Gauss-Jordan, k, Vx, dV) take ~7 % of the pipe next to the 336 tiles (ncu: DMMA sub-pipe + FP64 pipe = 89 %)",
            kernel_ms=bk, share_of_step=bk / ms_per_step,
            hbm=dict(achieved=BYTES_BACK * B / (bk * 1e-3) * 1e-9, peak=hbm_peak, unit="GB/s",
                     frac=BYTES_BACK * B / (bk * 1e-3) * 1e-9 / hbm_peak, peak_source=f"MEASURED_PEAKS.json ({peak_kind})"),
            forward=dict(kernel="fwd_lin32x8_kernel", kernel_ms=fw, bound="hbm", achieved=BYTES_FWD * B / (fw * 1e-3) * 1e-9,
                         peak=hbm_peak, unit="GB/s", frac=BYTES_FWD * B / (fw * 1e-3) * 1e-9 / hbm_peak),
            step_hbm_frac=(BYTES_BACK + BYTES_FWD) * B / (ms_per_step * 1e-3) * 1e-9 / hbm_peak,
            traffic=None)
        for tag in ("r02", "r01"):
            prof = os.path.join(ROOT, "profiles", f"traffic_{tag}.json")
            if os.path.exists(prof):
                try:
                    roofline["traffic"] = json.load(open(prof)).get("bp_tile32x8_kernel_dram_bytes_per_launch")
                    roofline["traffic_source"] = f"profiles/traffic_{tag}.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at B = 65536"
                    break
                except Exception:
                    pass
        try:                                               # the one FP64 datapath's busy share under ncu (DMMA sub-pipe + scalar FP64 pipe)
            import re as _re
            txt = open(os.path.join(ROOT, "profiles", "ncu_bp_tile_r02.txt")).read()
            pd = float(_re.search(r"pipe_tensor_subpipe_dmma\S*\s+([0-9.]+)", txt).group(1))
            pf = float(_re.search(r"sm__inst_executed_pipe_fp64\S*\s+([0-9.]+)", txt).group(1))
            roofline["fp64_datapath_busy_under_ncu"] = dict(dmma_pct=pd, scalar_fp64_pct=pf, total_pct=pd + pf,
                                                            source="profiles/ncu_bp_tile_r02.txt (one ncu --set full capture of this kernel at B = 65536)")
        except Exception:
            pass
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu_baseline, _ = run_cpu_baseline(args.cpu_sample, 2, 1, cpus=all_cpus)      # every host core again
            except Exception as exc:
                cpu_baseline = dict(value=None, error=str(exc))
        value = world * (B / BATCH) * 1e3 / ms_per_step
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=WORKLOAD,
                                batch_per_gpu=B, l2="inputs (~56 GB working set) are larger than the 126 MB L2: no flush needed",
                                parallelism=(f"batch sharded over {world} GPU(s); one 64-byte NCCL all-reduce per step "
                                             f"({'inside libddp: ddp_comm_allreduce_stats_f64' if lib_comm else 'torch.distributed'})") if world > 1
                                else "single GPU", kernel_variant="tile32x8"),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu_baseline,
                    check=dict(diverged=n_div, properties=props, oracle=oracle_check, mean_cost_new=float(stats_h[0] / max(stats_h[5], 1)),
                               accepted_frac=float(stats_h[3] / max(stats_h[5], 1))),
                    configs=results)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
# the other BASELINE configurations (each: timed with CUDA events, roofline fractions, sampled oracle check)
# ------------------------------------------------------------------------------------------------

def _timed(fn, steps, warmup, sync, record_parts=None):
    """fn(parts|None) runs one step; returns (ms_per_step, [mean ms of each part])."""
    import torch
    for _ in range(warmup):
        fn(None)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    parts = []
    e0.record()
    for _ in range(steps):
        p = []
        fn(p)
        parts.append(p)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1) / steps
    means = []
    if parts and parts[0]:
        for j in range(len(parts[0]) - 1):
            means.append(float(np.mean([p[j].elapsed_time(p[j + 1]) for p in parts])))
    return ms, means


def run_c4(g):
    """BASELINE configs[3]: iLQGkl KL-constrained, linear n=32 m=8 T=256 batch=65536: one back_pass_gps + forward(alpha=1) + KL
    evaluation; traj_prev = gains of one plain back pass, k_prev = 0, Sigma_i_prev = Quu, Sigma_prev = Quu^-1, eta = 1, R1 = 1e-4 I."""
    import torch
    from oracle import ddp_oracle as O
    L, eng, dev, tn, empty, ev = g["L"], g["eng"], g["dev"], g["tn"], g["empty"], g["ev"]
    n, m, T, B = g["n"], g["m"], g["T"], g["B"]
    fx, fu, x, u, cx, cu, Q, R, cxu = g["fx"], g["fu"], g["x"], g["u"], g["cx"], g["cu"], g["Q"], g["R"], g["cxu"]
    args, rank = g["args"], g["rank"]
    steps, warm = max(2, min(args.steps, 5)), 3
    # traj_prev: one plain back pass WITH the Quu history (untimed); K of the headline run is that pass's gains
    K_prev = g["K"]
    Sigi_prev = empty(B, T, m, m)
    ba = g["ba"]
    ba.Quu = Sigi_prev.data_ptr()
    eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
    ba.Quu = None
    torch.cuda.synchronize()
    Sig_prev = empty(B, T, m, m)
    for b0 in range(0, B, 4096):
        inv = torch.linalg.inv(Sigi_prev[b0:b0 + 4096])
        Sig_prev[b0:b0 + 4096] = 0.5 * (inv + inv.transpose(-1, -2))
        del inv
    eta = torch.ones(B, dtype=torch.float64, device=dev)
    R1 = (1e-4 * torch.eye(n, dtype=torch.float64, device=dev)).contiguous()
    Kn, kn, Vxn, dVn = empty(B, T, n, m), empty(B, T, m), empty(B, T, n), empty(B, 2)
    Quu, Quui = empty(B, T, m, m), empty(B, T, m, m)
    xnew, unew, cost, klm = g["xnew"], g["unew"], g["cost"], empty(B)
    dvn = torch.empty(B, dtype=torch.int32, device=dev)
    a = L.BackPassArgs()
    a.cx, a.cu = tn(cx, T * n, n), tn(cu, T * m, m)
    a.cxx, a.cxu, a.cuu = tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
    a.fx, a.fu = tn(fx, n * n, 0), tn(fu, n * m, 0)
    a.diverge, a.K, a.k, a.Vx, a.dV, a.Quu = dvn.data_ptr(), Kn.data_ptr(), kn.data_ptr(), Vxn.data_ptr(), dVn.data_ptr(), Quu.data_ptr()
    gp = L.GpsArgs()
    gp.K_prev, gp.Sigi_prev = tn(K_prev, T * n * m, n * m), tn(Sigi_prev, T * m * m, m * m)
    gp.eta, gp.Quui = eta.data_ptr(), Quui.data_ptr()
    model = g["model"]
    fa = L.ForwardPassArgs()
    fa.K, fa.k = Kn.data_ptr(), kn.data_ptr()
    fa.x0, fa.x, fa.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
    fa.alpha_scalar, fa.u_scale = 1.0, 1.0
    fa.xnew, fa.unew, fa.cost = xnew.data_ptr(), unew.data_ptr(), cost.data_ptr()
    ka = L.KlArgs()
    ka.fx, ka.R1 = tn(fx, n * n, 0), tn(R1, 0, 0)
    ka.xnew, ka.xold, ka.K_new, ka.k_new, ka.Sig_new = xnew.data_ptr(), x.data_ptr(), Kn.data_ptr(), kn.data_ptr(), Quui.data_ptr()
    ka.K_prev, ka.Sig_prev, ka.Sigi_prev = tn(K_prev, T * n * m, n * m), tn(Sig_prev, T * m * m, m * m), tn(Sigi_prev, T * m * m, m * m)
    ka.kl_mean = klm.data_ptr()

    def one(parts):
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_back_pass_gps_f64(eng.h, C.byref(a), C.byref(gp)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_kl_div_f64(eng.h, C.byref(ka)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()

    ms, (bk, fw, kl) = _timed(one, steps, warm, torch.cuda.synchronize)
    peaks, _ = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    dm = g["dmma_peak"]
    res = dict(workload="C4: one iLQGkl iteration (back_pass_gps + forward_pass(alpha=1) + forward_covariance/kl_div_wiki), linear n=32 m=8 T=256, "
                        "65536 trajectories, eta=1, kl_step=1, R1=1e-4 I (BASELINE.json configs[3])",
               batch=B, steps=steps, warmup=warm, ms_per_iter=ms, iters_per_s=1e3 / ms,
               back_pass_gps_ms=bk, forward_ms=fw, kl_ms=kl,
               fp64_frac=dict(back_pass_gps=C4_DMMA_BACK * 512.0 * (T - 1) * B / (bk * 1e-3) * 1e-12 / dm,
                              kl=C4_DMMA_KL * 512.0 * T * B / (kl * 1e-3) * 1e-12 / dm, peak_tflops=dm),
               hbm_frac=dict(back_pass_gps=C4_BYTES_BACK * B / (bk * 1e-3) * 1e-9 / hbm, forward=BYTES_FWD * B / (fw * 1e-3) * 1e-9 / hbm,
                             kl=C4_BYTES_KL * B / (kl * 1e-3) * 1e-9 / hbm,
                             iteration=(C4_BYTES_BACK + BYTES_FWD + C4_BYTES_KL) * B / (ms * 1e-3) * 1e-9 / hbm),
               binding="FP64 datapath (DMMA tiles)", diverged=int((dvn > 0).sum().item()),
               mean_kl=float(klm.mean().item()))
    if rank == 0 and args.oracle_samples > 0:
        ns = max(4, args.oracle_samples // 2)
        idx = np.sort(np.random.default_rng(8).choice(B, size=ns, replace=False))
        ii = torch.as_tensor(idx, device=dev)
        gg = lambda t: _h(t[ii])
        A_s, B_s = np.swapaxes(gg(fx), -1, -2), np.swapaxes(gg(fu), -1, -2)
        x_s, u_s, cx_s, cu_s = gg(x), gg(u), gg(cx), gg(cu)
        Kp_s, Si_s, Sp_s = np.swapaxes(gg(K_prev), -1, -2), np.swapaxes(gg(Sigi_prev), -1, -2), np.swapaxes(gg(Sig_prev), -1, -2)
        Kn_s, kn_s, Vx_s, dV_s = np.swapaxes(gg(Kn), -1, -2), gg(kn), gg(Vxn), gg(dVn)
        Qu_s, Qi_s = np.swapaxes(gg(Quu), -1, -2), np.swapaxes(gg(Quui), -1, -2)
        xn_s, un_s, c_s, kl_s, dv_s = gg(xnew), gg(unew), gg(cost), gg(klm), gg(dvn)
        Qh, Rh, R1h = _h(Q), _h(R), _h(R1)
        rep = lambda t: np.tile(t, (T, 1, 1))
        E = ErrTable()
        t0 = time.perf_counter()
        for j in range(ns):
            prev = O.GaussianPolicy(T, n, m, Kp_s[j], np.zeros((T, m)), Sp_s[j], Si_s[j])
            d0, p0, Vx0, Vxx0, dV0 = O.back_pass_gps(cx_s[j], cu_s[j], rep(Qh), rep(np.zeros((n, m))), rep(Rh), rep(A_s[j]), rep(B_s[j]), None,
                                                     x_s[j], u_s[j], (O.grad_kl(prev), np.array([1e-8, 1.0, 1e16])))
            om = O.LinearModel(A_s[j], B_s[j], Qh, Rh)
            xn0, un0, c0 = O.forward_pass(p0, x_s[j, 0], u_s[j], x_s[j], 1.0, om.f, om.costfun, None)
            sig = O.forward_covariance(A_s[j], R1h, p0)
            kl0 = float(np.mean(O.kl_div_wiki(xn0, x_s[j], sig, p0, prev)))
            E.same("diverge", int(dv_s[j]) == d0)
            E.rel("K", Kn_s[j], p0.K); E.rel("k", kn_s[j], p0.k); E.rel("Vx", Vx_s[j], Vx0); E.rel("dV", dV_s[j], dV0)
            E.rel("Sigma_i (Quu)", Qu_s[j], p0.Sigmai); E.rel("Sigma (Quu^-1)", Qi_s[j], p0.Sigma)
            E.rel("xnew", xn_s[j], xn0); E.rel("unew", un_s[j], un0); E.rel("cost", c_s[j], c0); E.rel("kl_mean", kl_s[j], kl0)
        res["oracle"] = E.report(ns)
        res["oracle"]["seconds"] = time.perf_counter() - t0
    klm_recompute = klm.clone()
    # the same iteration with forward_covariance's state block kept from an earlier eta iteration (it depends on fx and R1 only):
    # ddp_kl_args.Sx_tri, mode 1 = store while computing, mode 2 = read back -- what ddp_ilqgkl_solve_f64 does from iteration 2 on.
    # Next to this benchmark's two 32 GiB gain histories the 66 GiB cache of all 65536 trajectories does not fit in 178 GiB:
    # Sx_count = as many leading trajectories as the free memory holds, the rest is propagated as before (one launch each).
    cached = None
    try:
        per_gb = T * 528 * 8 / 2**30
        free_gb = torch.cuda.mem_get_info()[0] / 2**30 + (torch.cuda.memory_reserved() - torch.cuda.memory_allocated()) / 2**30
        nc = int(min(B, (free_gb - 3.0) / per_gb))
        nc -= nc % 1184                                    # whole rounds of the resident warp set of a B200
        if nc < B // 8:
            raise MemoryError(f"{per_gb * B:.1f} GiB for the full (528,T,B) cache, {free_gb:.1f} GiB free")
        torch.cuda.empty_cache()
        Sx = empty(nc, T, 528)
        ka.Sx_tri, ka.Sx_mode, ka.Sx_count = Sx.data_ptr(), 1, (nc if nc < B else 0)
        e0, e1 = ev(), ev()
        e0.record()
        eng._ck(eng.lib.ddp_kl_div_f64(eng.h, C.byref(ka)))
        e1.record(); torch.cuda.synchronize()
        store_ms = e0.elapsed_time(e1)
        same_store = bool(torch.equal(klm, klm_recompute))
        ka.Sx_mode = 2
        ms_c, (bk_c, fw_c, kl_c) = _timed(one, steps, warm, torch.cuda.synchronize)
        cached = dict(batch=B, cached_trajectories=nc, cache_gib=nc * per_gb, ms_per_iter=ms_c, iters_per_s=1e3 / ms_c,
                      back_pass_gps_ms=bk_c, forward_ms=fw_c, kl_ms=kl_c, kl_store_pass_ms=store_ms,
                      iteration_speedup=ms / ms_c, kl_speedup=kl / kl_c,
                      kl_mean_bitwise_equal_to_recompute=bool(torch.equal(klm, klm_recompute)) and same_store,
                      note="iterations 2.. of one iLQGkl solve at the full C4 batch: Sigma_x(t) of the first cached_trajectories read from the "
                           "cache written by iteration 1 (kl_store_pass_ms), the others propagated as in the uncached run; kl_mean compared "
                           "bitwise with the recomputing run")
        ka.Sx_tri, ka.Sx_mode, ka.Sx_count = None, 0, 0
        del Sx
    except Exception as exc:
        cached = dict(error=f"{type(exc).__name__}: {exc}")
        ka.Sx_tri, ka.Sx_mode, ka.Sx_count = None, 0, 0
    res["covariance_cached"] = cached
    return res


def run_line_search(g):
    """Row f3: the cost of 10 step sizes (iLQG.jl:145's alpha list) on the headline policy -- ddp_forward_costs_multi_f64 (one launch,
    gains read once, FP64 tensor tiles) vs 10 serial rollouts; costs compared with each other and, on samples, with the oracle."""
    import torch
    from oracle import ddp_oracle as O
    L, eng, dev, tn, empty = g["L"], g["eng"], g["dev"], g["tn"], g["empty"]
    n, m, T, B = g["n"], g["m"], g["T"], g["B"]
    fx, fu, x, u, Q, R, K, k = g["fx"], g["fu"], g["x"], g["u"], g["Q"], g["R"], g["K"], g["k"]
    args, rank = g["args"], g["rank"]
    alphas = np.ascontiguousarray(10.0 ** np.linspace(0, -3, 10))
    NA = len(alphas)
    costs, cs = empty(NA, B), empty(NA, B)
    fa = L.ForwardPassArgs()
    fa.K, fa.k = K.data_ptr(), k.data_ptr()
    fa.x0, fa.x, fa.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
    fa.alpha_scalar, fa.u_scale = 1.0, 1.0
    xs_, us_ = empty(B, T, n), empty(B, T, m)
    fa.xnew, fa.unew = xs_.data_ptr(), us_.data_ptr()
    model = g["model"]
    l0 = eng.launch_count

    def multi(parts):
        eng._ck(eng.lib.ddp_forward_costs_multi_f64(eng.h, C.byref(model), C.byref(fa), NA, alphas.ctypes.data_as(C.POINTER(C.c_double)), costs.data_ptr()))

    def serial(parts):
        for j, al in enumerate(alphas):
            fa.alpha_scalar = float(al)
            fa.cost = cs[j].data_ptr()
            eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa)))
        fa.alpha_scalar = 1.0

    steps, warm = max(2, min(args.steps, 5)), 3
    ms_multi, _ = _timed(multi, steps, warm, torch.cuda.synchronize)
    launches_multi = (eng.launch_count - l0) / (steps + warm)
    ms_serial, _ = _timed(serial, steps, warm, torch.cuda.synchronize)
    rel = ((costs - cs).abs() / cs.abs().clamp_min(1e-300)).max().item()
    res = dict(workload="line search: cost of 10 step sizes 10^linspace(0,-3,10) on the headline policy, n=32 m=8 T=256, 65536 trajectories",
               batch=B, n_alpha=NA, steps=steps, warmup=warm, multi_alpha_ms=ms_multi, serial_rollouts_ms=ms_serial, speedup=ms_serial / ms_multi,
               launches_per_search=launches_multi, max_rel_diff_multi_vs_serial=rel,
               fp64_frac=100.0 * 512.0 * (T - 1) * B / (ms_multi * 1e-3) * 1e-12 / g["dmma_peak"],
               note="fp64_frac: 100 m8n8k4 tiles per step and trajectory (16 step sizes fit the same tiles) over the measured DMMA peak")
    if rank == 0 and args.oracle_samples > 0:
        ns = max(4, args.oracle_samples // 4)
        idx = np.sort(np.random.default_rng(9).choice(B, size=ns, replace=False))
        ii = torch.as_tensor(idx, device=dev)
        gg = lambda t: _h(t[ii])
        A_s, B_s = np.swapaxes(gg(fx), -1, -2), np.swapaxes(gg(fu), -1, -2)
        K_s, k_s, x_s, u_s = np.swapaxes(gg(K), -1, -2), gg(k), gg(x), gg(u)
        c_s = _h(costs[:, ii])
        Qh, Rh = _h(Q), _h(R)
        E = ErrTable()
        t0 = time.perf_counter()
        for j in range(ns):
            pol = O.GaussianPolicy(T, n, m, K_s[j], k_s[j], None, None)
            om = O.LinearModel(A_s[j], B_s[j], Qh, Rh)
            for a_i, al in enumerate(alphas):
                _, _, c0 = O.forward_pass(pol, x_s[j, 0], u_s[j], x_s[j], float(al), om.f, om.costfun, None)
                E.rel("cost", np.array([c_s[a_i, j]]), np.array([float(np.sum(c0))]))
        res["oracle"] = E.report(ns)
        res["oracle"]["seconds"] = time.perf_counter() - t0
    return res


def run_c2_lims(g):
    """Row a4 at the headline shape: back_pass with control limits (boxQP branch, backward_pass.jl:43-62) + the clamped rollout
    (forward_pass.jl:22-28) on C2's systems; limits +-0.15 around zero with u ~ 0.1 N(0,1)."""
    import torch
    from oracle import ddp_oracle as O
    L, eng, dev, tn, empty, ev = g["L"], g["eng"], g["dev"], g["tn"], g["empty"], g["ev"]
    n, m, T, B = g["n"], g["m"], g["T"], g["B"]
    fx, fu, x, u, cx, cu, Q, R, cxu, lam = g["fx"], g["fu"], g["x"], g["u"], g["cx"], g["cu"], g["Q"], g["R"], g["cxu"], g["lam"]
    args, rank = g["args"], g["rank"]
    lims = torch.tensor([-0.15] * m + [0.15] * m, dtype=torch.float64, device=dev)      # (m,2) column-major: lower column, upper column
    K2, k2, Vx2, dV2 = empty(B, T, n, m), empty(B, T, m), empty(B, T, n), empty(B, 2)
    dv2 = torch.empty(B, dtype=torch.int32, device=dev)
    xn2, un2, c2 = empty(B, T, n), empty(B, T, m), empty(B)
    ba = L.BackPassArgs()
    ba.cx, ba.cu = tn(cx, T * n, n), tn(cu, T * m, m)
    ba.cxx, ba.cxu, ba.cuu = tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
    ba.fx, ba.fu = tn(fx, n * n, 0), tn(fu, n * m, 0)
    ba.lam, ba.reg_type = lam.data_ptr(), 1
    ba.lims, ba.u = lims.data_ptr(), tn(u, T * m, m)
    ba.diverge, ba.K, ba.k, ba.Vx, ba.dV = dv2.data_ptr(), K2.data_ptr(), k2.data_ptr(), Vx2.data_ptr(), dV2.data_ptr()
    fa = L.ForwardPassArgs()
    fa.K, fa.k = K2.data_ptr(), k2.data_ptr()
    fa.x0, fa.x, fa.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
    fa.alpha_scalar, fa.u_scale = 1.0, 1.0
    fa.lims = lims.data_ptr()
    fa.xnew, fa.unew, fa.cost = xn2.data_ptr(), un2.data_ptr(), c2.data_ptr()
    model = g["model"]

    def one(parts):
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()

    steps, warm = max(2, min(args.steps, 3)), 3
    ms, (bk, fw) = _timed(one, steps, warm, torch.cuda.synchronize)
    rows_clamped = float((K2[:, :T - 1].abs().sum(dim=2) == 0).double().mean().item())     # K2 is (B,T,n,m): row i of K_t = K2[b,t,:,i]
    res = dict(workload="C2 with control limits: back_pass (box-QP branch on bp_tile32x8_kernel<LIMS>, warp-cooperative projected Newton) + clamped "
                        "forward_pass, linear n=32 m=8 T=256, 65536 trajectories, lims = +-0.15",
               batch=B, steps=steps, warmup=warm, ms_per_iter=ms, iters_per_s=1e3 / ms, back_pass_ms=bk, forward_ms=fw,
               clamped_control_frac=rows_clamped, diverged=int((dv2 > 0).sum().item()),
               vs_no_limits=dict(back_pass_ms_no_limits=g.get("bk_plain"), note="the same sweep without limits is the headline's backward kernel"))
    if rank == 0 and args.oracle_samples > 0:
        ns = max(4, args.oracle_samples // 4)
        idx = np.sort(np.random.default_rng(10).choice(B, size=ns, replace=False))
        ii = torch.as_tensor(idx, device=dev)
        gg = lambda t: _h(t[ii])
        A_s, B_s = np.swapaxes(gg(fx), -1, -2), np.swapaxes(gg(fu), -1, -2)
        x_s, u_s, cx_s, cu_s, lam_s = gg(x), gg(u), gg(cx), gg(cu), gg(lam)
        K_s, k_s, Vx_s, dV_s, dv_s = np.swapaxes(gg(K2), -1, -2), gg(k2), gg(Vx2), gg(dV2), gg(dv2)
        xn_s, un_s, c_s = gg(xn2), gg(un2), gg(c2)
        Qh, Rh = _h(Q), _h(R)
        lims_h = np.stack([np.full(m, -0.15), np.full(m, 0.15)], axis=1)
        E = ErrTable()
        t0 = time.perf_counter()
        for j in range(ns):
            d0, p0, Vx0, Vxx0, dV0 = O.back_pass(cx_s[j], cu_s[j], Qh, np.zeros((n, m)), Rh, A_s[j], B_s[j], float(lam_s[j]), 1, lims_h, x_s[j], u_s[j])
            om = O.LinearModel(A_s[j], B_s[j], Qh, Rh)
            xn0, un0, c0 = O.forward_pass(p0, x_s[j, 0], u_s[j], x_s[j], 1.0, om.f, om.costfun, lims_h)
            E.same("diverge", int(dv_s[j]) == d0)
            E.same("clamped rows of K (free sets of every step)", bool(np.array_equal(np.abs(K_s[j]).sum(axis=2) == 0, np.abs(p0.K).sum(axis=2) == 0)))
            E.same("clamped controls of the rollout", bool(np.array_equal(np.abs(un_s[j]) == 0.15, np.abs(un0) == 0.15)))
            E.rel("K", K_s[j], p0.K); E.rel("k", k_s[j], p0.k); E.rel("Vx", Vx_s[j], Vx0); E.rel("dV", dV_s[j], dV0)
            E.rel("xnew", xn_s[j], xn0); E.rel("unew", un_s[j], un0); E.rel("cost", c_s[j], c0)
        res["oracle"] = E.report(ns)
        res["oracle"]["seconds"] = time.perf_counter() - t0
    return res


def run_c1(g):
    """BASELINE configs[0]'s problem as a batch: demo_linear (demo_linear.jl:5-60) n=10 m=2 T=1000, 16384 trajectories -- one
    back_pass + forward_pass on the coverage kernels (bp_generic_kernel in its warp-per-trajectory form, fwd_generic_kernel)."""
    import torch
    from oracle import ddp_oracle as O
    ddp, L, dev, tn, ev = g["ddp"], g["L"], g["dev"], g["tn"], g["ev"]
    args, rank = g["args"], g["rank"]
    n, m, T, B, h = 10, 2, 1000, 16384, 0.01
    f64 = torch.float64
    empty = lambda *s_: torch.empty(*s_, dtype=f64, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(31)
    G = torch.randn(B, n, n, dtype=f64, device=dev, generator=gen)
    fx = torch.linalg.matrix_exp(h * (G - G.transpose(1, 2))).transpose(1, 2).contiguous()
    fu = (h * torch.randn(B, n, m, dtype=f64, device=dev, generator=gen)).transpose(1, 2).contiguous()
    x0 = torch.ones(B, n, dtype=f64, device=dev)
    u = 0.1 * torch.randn(B, T, m, dtype=f64, device=dev, generator=gen)
    Q = (h * torch.eye(n, dtype=f64, device=dev)).contiguous(); R = (0.1 * h * torch.eye(m, dtype=f64, device=dev)).contiguous()
    cxu = torch.zeros(m, n, dtype=f64, device=dev); lam = torch.ones(B, dtype=f64, device=dev)
    eng = ddp.Engine(n, m, T, B, device=g["local_rank"])
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    model = L.Model(); model.kind = 1
    model.A, model.Bm, model.Q, model.R, model.flags = tn(fx, n * n, 0), tn(fu, n * m, 0), tn(Q, 0, 0), tn(R, 0, 0), 1
    x, c0, un, cx, cu = empty(B, T, n), empty(B), empty(B, T, m), empty(B, T, n), empty(B, T, m)
    fa = L.ForwardPassArgs(); fa.x0, fa.u = tn(x0, n, 0), tn(u, T * m, m); fa.alpha_scalar = fa.u_scale = 1.0
    fa.xnew, fa.unew, fa.cost, fa.cx, fa.cu = x.data_ptr(), un.data_ptr(), c0.data_ptr(), cx.data_ptr(), cu.data_ptr()
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa)))
    K, k, Vx, dV = empty(B, T, n, m), empty(B, T, m), empty(B, T, n), empty(B, 2)
    dv = torch.empty(B, dtype=torch.int32, device=dev)
    ba = L.BackPassArgs()
    ba.cx, ba.cu, ba.cxx, ba.cxu, ba.cuu = tn(cx, T * n, n), tn(cu, T * m, m), tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
    ba.fx, ba.fu, ba.lam, ba.reg_type = tn(fx, n * n, 0), tn(fu, n * m, 0), lam.data_ptr(), 1
    ba.diverge, ba.K, ba.k, ba.Vx, ba.dV = dv.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr()
    xn, unw, cst = empty(B, T, n), empty(B, T, m), empty(B)
    fp = L.ForwardPassArgs()
    fp.K, fp.k = K.data_ptr(), k.data_ptr()
    fp.x0, fp.x, fp.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
    fp.alpha_scalar, fp.u_scale = 1.0, 1.0
    fp.xnew, fp.unew, fp.cost = xn.data_ptr(), unw.data_ptr(), cst.data_ptr()

    def one(parts):
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fp)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()

    steps, warm = max(2, min(args.steps, 3)), 3
    ms, (bk, fw) = _timed(one, steps, warm, torch.cuda.synchronize)
    res = dict(workload="C1 batched: demo_linear n=10 m=2 T=1000 (BASELINE.json configs[0]'s problem; the reference runs one trajectory), 16384 "
                        "trajectories, one back_pass + forward_pass(alpha=1) on the coverage kernels (bp_generic_kernel<warp per trajectory>, fwd_generic_kernel)",
               batch=B, steps=steps, warmup=warm, ms_per_iter=ms, iters_per_s=1e3 / ms, back_pass_ms=bk, forward_ms=fw,
               trajectory_steps_per_s=B * T / (ms * 1e-3), diverged=int((dv > 0).sum().item()))
    if rank == 0 and args.oracle_samples > 0:
        ns = max(2, args.oracle_samples // 8)
        idx = np.sort(np.random.default_rng(12).choice(B, size=ns, replace=False))
        rep = check_linear(idx, fx, fu, x, u, cx, cu, Q, R, lam, K, k, Vx, None, dV, dv, xn, unw, cst)
        res["oracle"] = rep
    eng.close()
    return res


def run_solve_e2e(g):
    """What iLQG(f,costfun,df,x0,u0) actually moves (iLQG.jl:143-341): x0,u0 up once, then the whole outer loop on the device
    (ddp_ilqg_solve_f64), then x,u,K,k,cost down once.  Reported per accepted-or-rejected outer iteration."""
    import torch
    ddp, L, dev, local_rank = g["ddp"], g["L"], g["dev"], g["local_rank"]
    n, m, T = g["n"], g["m"], g["T"]
    Bs = 4096
    rng = np.random.default_rng(11)
    import scipy.linalg as sla
    # one random system per 64 trajectories keeps the host-side generation short; the device work is independent of that
    nsys = Bs // 64
    A = np.stack([sla.expm(H_STEP * (gm - gm.T)) for gm in rng.standard_normal((nsys, n, n))])
    A = np.repeat(A, 64, axis=0)
    Bm = H_STEP * rng.standard_normal((Bs, n, m))
    x0 = 1.0 + 0.1 * rng.standard_normal((Bs, n))
    u0 = 0.1 * rng.standard_normal((Bs, T, m))
    Q, R = H_STEP * np.eye(n), 0.1 * H_STEP * np.eye(m)
    eng = ddp.Engine(n, m, T, Bs, device=local_rank)
    model = ddp.LinearModel(A[:, None], Bm[:, None], Q, R)
    warm = ddp.LinearModel(A[:256, None], Bm[:256, None], Q, R)
    ddp.iLQG(warm.f, warm.costfun, warm.df, x0[:256], u0[:256], max_iter=3)               # warm the allocator / module
    t0 = time.perf_counter()
    xs, us, pol, Vx, Vxx, cost, tr = ddp.iLQG(model.f, model.costfun, model.df, x0, u0, engine=eng)
    dt = time.perf_counter() - t0
    eng.close()
    h2d = 8 * (A.size + Bm.size + x0.size + u0.size)
    d2h = 8 * (xs.size + us.size + pol.K.size + pol.k.size + Vx.size + Vxx.size)
    n_outer = int(tr["n_outer"])
    return dict(workload="whole iLQG solve through the host mirror of iLQG(f,costfun,df,x0,u0): 4096 LQ problems n=32 m=8 T=256 to tol_grad",
                seconds=dt, outer_iterations=n_outer, status_counts={str(s): int((tr["status"] == s).sum()) for s in np.unique(tr["status"])},
                h2d_bytes=int(h2d), d2h_bytes=int(d2h), iters_per_s_65536_equiv=(n_outer * Bs / BATCH) / dt,
                note="wall clock of the Python call: host packing (transposes), one upload, the device-resident loop (backward + line search per "
                     "outer iteration), one download of x,u,K,k,Vx; iters_per_s_65536_equiv = outer iterations x batch / 65536 / seconds")


def run_c3(g):
    """BASELINE configs[2]: pendcart n=4 m=1 T=600 with control lims (boxQP path) batch=262144: x0_b = [pi-0.6+0.2 U(-1,1),0,0,0],
    u = 0, lims = +-5, regType = 2, lambda = 1; fx,fu (B,T,.) from the device df (ZoH), boxQP branch active."""
    import torch
    from oracle import ddp_oracle as O
    ddp, L, dev, local_rank, tn, ev = g["ddp"], g["L"], g["dev"], g["local_rank"], g["tn"], g["ev"]
    args, rank = g["args"], g["rank"]
    f64 = torch.float64
    n, m, T, B = 4, 1, C3_T, (C3_B if args.batch == BATCH else max(1024, args.batch * 4))
    empty = lambda *s: torch.empty(*s, dtype=f64, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    x0 = torch.zeros(B, n, dtype=f64, device=dev)
    x0[:, 0] = math.pi - 0.6 + 0.2 * (2.0 * torch.rand(B, dtype=f64, device=dev, generator=gen) - 1.0)
    u = torch.zeros(B, T, m, dtype=f64, device=dev)
    Q = torch.diag(torch.tensor([10.0, 1.0, 2.0, 1.0], dtype=f64, device=dev)).contiguous()
    R = torch.ones(1, 1, dtype=f64, device=dev)
    goal = torch.tensor([math.pi, 0.0, 0.0, 0.0], dtype=f64, device=dev)
    lims = torch.tensor([-5.0, 5.0], dtype=f64, device=dev)          # [lower(m); upper(m)]
    cxu = torch.zeros(m, n, dtype=f64, device=dev)
    lam = torch.ones(B, dtype=f64, device=dev)
    eng = ddp.Engine(n, m, T, B, device=local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    model = L.Model()
    model.kind = 2
    model.Q, model.R, model.goal = tn(Q, 0, 0), tn(R, 0, 0), goal.data_ptr()
    for i, v in enumerate((9.82, 0.35, 0.01, 0.99)):
        model.p[i] = v
    model.terminal_cost, model.flags = 1, 1
    x, cost0 = empty(B, T, n), empty(B)
    un0 = empty(B, T, m)
    fa0 = L.ForwardPassArgs()
    fa0.x0, fa0.u = tn(x0, n, 0), tn(u, T * m, m)
    fa0.alpha_scalar, fa0.u_scale = 1.0, 1.0
    fa0.xnew, fa0.unew, fa0.cost = x.data_ptr(), un0.data_ptr(), cost0.data_ptr()
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa0)))
    del un0
    fx, fu, cx, cu = empty(B, T, n, n), empty(B, T, m, n), empty(B, T, n), empty(B, T, m)
    d0, d1 = ev(), ev()
    d0.record()
    eng._ck(eng.lib.ddp_model_derivs_f64(eng.h, C.byref(model), x.data_ptr(), u.data_ptr(), fx.data_ptr(), fu.data_ptr(), cx.data_ptr(), cu.data_ptr()))
    d1.record()
    torch.cuda.synchronize()
    df_ms = d0.elapsed_time(d1)
    K, k, Vx, dV, Vxx1 = empty(B, T, n, m), empty(B, T, m), empty(B, T, n), empty(B, 2), empty(B, n, n)
    xnew, unew, cost = empty(B, T, n), empty(B, T, m), empty(B)
    dv = torch.empty(B, dtype=torch.int32, device=dev)
    stats = torch.zeros(8, dtype=f64, device=dev)
    ba = L.BackPassArgs()
    ba.cx, ba.cu = tn(cx, T * n, n), tn(cu, T * m, m)
    ba.cxx, ba.cxu, ba.cuu = tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
    ba.fx, ba.fu = tn(fx, T * n * n, n * n), tn(fu, T * n * m, n * m)
    ba.lam, ba.reg_type, ba.lims, ba.u = lam.data_ptr(), 2, lims.data_ptr(), tn(u, T * m, m)
    ba.diverge, ba.K, ba.k, ba.Vx, ba.dV, ba.Vxx1 = dv.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr(), Vxx1.data_ptr()
    fa = L.ForwardPassArgs()
    fa.K, fa.k = K.data_ptr(), k.data_ptr()
    fa.x0, fa.x, fa.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
    fa.alpha_scalar, fa.u_scale, fa.lims = 1.0, 1.0, lims.data_ptr()
    fa.xnew, fa.unew, fa.cost = xnew.data_ptr(), unew.data_ptr(), cost.data_ptr()

    def one(parts):
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa)))
        if parts is not None:
            parts.append(ev()); parts[-1].record()
        eng._ck(eng.lib.ddp_batch_stats_f64(eng.h, cost0.data_ptr(), cost.data_ptr(), dV.data_ptr(), None, 1.0, dv.data_ptr(), None, stats.data_ptr()))

    steps, warm = max(3, min(args.steps, 10)), 3
    ms, (bk, fw) = _timed(one, steps, warm, torch.cuda.synchronize)
    peaks, _ = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    clamped_frac = float(((unew == 5.0) | (unew == -5.0)).double().mean().item())
    kzero_frac = float((K[:, :T - 1].abs().sum(dim=(2, 3)) == 0).double().mean().item())
    res = dict(workload="C3: pendulum on a cart n=4 m=1 T=600, control limits +-5 (boxQP branch), regType=2, lambda=1, 262144 trajectories "
                        "(BASELINE.json configs[2]); fx,fu (B,T,.) ZoH Jacobians from the device df",
               batch=B, steps=steps, warmup=warm, ms_per_iter=ms, iters_per_s=1e3 / ms, backward_ms=bk, forward_ms=fw, df_ms_untimed_part=df_ms,
               kernels=dict(backward="bp_small_kernel<4,1>", forward="fwd_pend_staged_kernel"),
               hbm_frac=dict(backward=C3_BYTES_BACK * B / (bk * 1e-3) * 1e-9 / hbm, forward=C3_BYTES_FWD * B / (fw * 1e-3) * 1e-9 / hbm,
                             iteration=(C3_BYTES_BACK + C3_BYTES_FWD) * B / (ms * 1e-3) * 1e-9 / hbm, peak_gbs=hbm,
                             algorithmic_bytes_per_trajectory=dict(backward=C3_BYTES_BACK, forward=C3_BYTES_FWD)),
               binding="HBM", target_hbm_frac=0.70, diverged=int((dv > 0).sum().item()),
               clamped_control_frac_in_rollout=clamped_frac, clamped_step_frac_in_backward=kzero_frac)
    if rank == 0 and args.oracle_samples > 0:
        ns = args.oracle_samples
        idx = np.sort(np.random.default_rng(9).choice(B, size=ns, replace=False))
        ii = torch.as_tensor(idx, device=dev)
        gg = lambda t: _h(t[ii])
        fx_s, fu_s = np.swapaxes(gg(fx), -1, -2), np.swapaxes(gg(fu), -1, -2)
        x_s, u_s, cx_s, cu_s = gg(x), gg(u), gg(cx), gg(cu)
        K_s, k_s, Vx_s, dV_s, V1_s = np.swapaxes(gg(K), -1, -2), gg(k), gg(Vx), gg(dV), np.swapaxes(gg(Vxx1), -1, -2)
        xn_s, un_s, c_s, dv_s = gg(xnew), gg(unew), gg(cost), gg(dv)
        om = O.PendcartModel()
        limh = np.array([[-5.0, 5.0]])
        E = ErrTable()
        t0 = time.perf_counter()
        for j in range(ns):
            if j < 4:        # the device's ZoH Jacobians against the oracle's expm (0.8 s per trajectory: four of the samples)
                fx0, fu0, _, _, _, cx0, cu0, *_ = om.df(x_s[j].copy(), u_s[j].copy())
                E.rel("fx (df)", fx_s[j], fx0, floor=1e-12); E.rel("fu (df)", fu_s[j], fu0, floor=1e-12)
                E.rel("cx (df)", cx_s[j], cx0); E.rel("cu (df)", cu_s[j], cu0)
            d0_, p0, Vx0, Vxx0, dV0 = O.back_pass(cx_s[j], cu_s[j], om.Q, np.zeros((4, 1)), np.array([[om.R]]), fx_s[j], fu_s[j], 1.0, 2, limh,
                                                  x_s[j], u_s[j])
            xn0, un0, c0 = O.forward_pass(p0, x_s[j, 0], u_s[j], x_s[j], 1.0, om.f, om.costfun, limh)
            E.same("diverge", int(dv_s[j]) == d0_)
            E.same("clamped set of the backward pass (K_t == 0 pattern)", np.array_equal(K_s[j] == 0, p0.K == 0))
            E.same("clamped set of the rollout (u == +-5 pattern)", np.array_equal(np.abs(un_s[j]) == 5.0, np.abs(un0) == 5.0))
            E.rel("K", K_s[j], p0.K); E.rel("k", k_s[j], p0.k); E.rel("Vx", Vx_s[j], Vx0); E.rel("Vxx1", V1_s[j], Vxx0[0]); E.rel("dV", dV_s[j], dV0)
            E.rel("xnew", xn_s[j], xn0); E.rel("unew", un_s[j], un0); E.rel("cost", c_s[j], float(np.sum(c0)))
        res["oracle"] = E.report(ns)
        res["oracle"]["seconds"] = time.perf_counter() - t0
    eng.close()
    return res


def run_c5(g):
    """BASELINE configs[4]: n=32 m=8 T=256 batch=2097152 sharded 8xB200 = 262144 trajectories per GPU, processed by
    ddp_ilqg_iter_f64 in resident chunks (policy scratch of one chunk), one NCCL all-reduce of the statistics per iteration."""
    import torch
    import torch.distributed as dist
    ddp, L, dev, local_rank, tn, ev = g["ddp"], g["L"], g["dev"], g["local_rank"], g["tn"], g["ev"]
    args, rank, world = g["args"], g["rank"], g["world"]
    n, m, T = g["n"], g["m"], g["T"]
    f64 = torch.float64
    B5 = C5_B if args.batch == BATCH else args.batch * 4
    empty = lambda *s: torch.empty(*s, dtype=f64, device=dev)
    fx, fu, x0, u = g["gen_linear"](B5, 5000 + rank)
    Q, R = g["Q"], g["R"]
    lam = torch.ones(B5, dtype=f64, device=dev)
    eng = ddp.Engine(n, m, T, B5, device=local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    model = L.Model()
    model.kind = 1
    model.A, model.Bm = tn(fx, n * n, 0), tn(fu, n * m, 0)
    model.Q, model.R, model.flags = tn(Q, 0, 0), tn(R, 0, 0), 1
    x, cost0 = empty(B5, T, n), empty(B5)
    xnew, unew, cost, dV = empty(B5, T, n), empty(B5, T, m), empty(B5), empty(B5, 2)
    dv = torch.empty(B5, dtype=torch.int32, device=dev)
    stats = torch.zeros(8, dtype=f64, device=dev)
    fa0 = L.ForwardPassArgs()
    fa0.x0, fa0.u = tn(x0, n, 0), tn(u, T * m, m)
    fa0.alpha_scalar, fa0.u_scale = 1.0, 1.0
    fa0.xnew, fa0.unew, fa0.cost = x.data_ptr(), unew.data_ptr(), cost0.data_ptr()
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa0)))
    torch.cuda.synchronize()
    ia = L.IterArgs()
    ia.x, ia.u, ia.lam = x.data_ptr(), u.data_ptr(), lam.data_ptr()
    ia.alpha_scalar, ia.reg_type = 1.0, 1
    ia.xnew, ia.unew, ia.cost, ia.dV, ia.diverge = xnew.data_ptr(), unew.data_ptr(), cost.data_ptr(), dV.data_ptr(), dv.data_ptr()
    ia.chunk = args.c5_chunk
    lib_comm = g["lib_comm"]
    if world > 1 and lib_comm:
        try:
            idt = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            eng.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
        except Exception:
            lib_comm = False
        flag = torch.tensor([1.0 if lib_comm else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        lib_comm = bool(flag.item() > 0.5)

    def one(parts):
        eng._ck(eng.lib.ddp_ilqg_iter_f64(eng.h, C.byref(model), C.byref(ia)))
        eng._ck(eng.lib.ddp_batch_stats_f64(eng.h, cost0.data_ptr(), cost.data_ptr(), dV.data_ptr(), None, 1.0, dv.data_ptr(), None, stats.data_ptr()))
        if world > 1:
            if lib_comm:
                eng.allreduce_stats(stats.data_ptr())
            else:
                dist.all_reduce(stats)

    steps, warm = max(2, min(args.steps, 4)), 3
    l0 = None
    for _ in range(warm):
        one(None)
    g["barrier"]()
    l0 = eng.launch_count
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(steps):
        one(None)
    e1.record()
    g["barrier"]()
    ms = g["max_over_ranks"](e0.elapsed_time(e1) / steps)
    launches_per_iter = (eng.launch_count - l0) / steps
    peaks, _ = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    st = stats.cpu().numpy()
    res = dict(workload=f"C5: n=32 m=8 T=256, {B5} trajectories per GPU x {world} GPU(s) = {B5 * world} (BASELINE.json configs[4] is 2097152 on 8 GPUs), "
                        "chunked device iteration ddp_ilqg_iter_f64 (derivative step + backward + forward per chunk, policy in a chunk-sized scratch), "
                        "one 64-byte all-reduce per iteration",
               is_config_5=bool(world == 8 and B5 == C5_B), batch_per_gpu=B5, n_gpus=world, chunks_per_iteration=int(ia.n_chunks),
               chunk_trajectories=(int(args.c5_chunk) if args.c5_chunk else "library default: sm_count x 8 x 55 = 65120 on a B200"),
               launches_per_iteration_per_gpu=launches_per_iter,
               steps=steps, warmup=warm, ms_per_iter=ms, iters_per_s_of_the_whole_batch=1e3 / ms,
               c2_equivalent_iters_per_s=world * (B5 / BATCH) * 1e3 / ms,
               fp64_frac=FLOPS_BACK_STEP_SYM * (T - 1) * B5 / (ms * 1e-3) * 1e-12 / g["dmma_peak"],
               hbm_frac=(BYTES_BACK + BYTES_FWD + 8.0 * 2 * (T * n + T * m)) * B5 / (ms * 1e-3) * 1e-9 / hbm,
               all_reduce=("ddp_comm_allreduce_stats_f64 (ncclAllReduce inside libddp)" if lib_comm else "torch.distributed") if world > 1 else None,
               diverged_all_ranks=int(st[4]), trajectories_counted_all_ranks=int(st[5]),
               note="fp64_frac counts the backward sweep's executed tiles over the whole iteration time (derivative step and rollout included); "
                    "hbm_frac adds the derivative step's read of x,u and write of cx,cu")
    if rank == 0 and args.oracle_samples > 0:
        idx = np.sort(np.random.default_rng(10).choice(B5, size=args.oracle_samples, replace=False))
        t0 = time.perf_counter()
        res["oracle"] = check_linear(idx, fx, fu, x, u, None, None, Q, R, lam, None, None, None, None, dV, dv, xnew, unew, cost)
        res["oracle"]["seconds"] = time.perf_counter() - t0
        res["oracle"]["note"] = "the policy lives in the chunk scratch only: K,k,Vx are checked through dV, the rollout and its cost"
    eng.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="trajectories per GPU (the metric is quoted at 65536)")
    ap.add_argument("--chunk", type=int, default=0, help="e2e pipeline chunk (trajectories); 0 = library default")
    ap.add_argument("--c5-chunk", type=int, default=0, help="chunk of the C5 device iteration; 0 = library default")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=1024, help="trajectories in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="c1,c3,c4,c5,solve,c2lims,linesearch", help="other BASELINE configurations to measure in the same run (comma list; '' = none)")
    ap.add_argument("--oracle-samples", type=int, default=32, help="trajectories of each timed batch re-computed by the CPU oracle (0 = off)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
