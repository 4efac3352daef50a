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

Inputs are synthetic (seeded): per-trajectory LTI dynamics A_b = exp(h(G-G')), B_b = h N(0,1),
Q = hI, R = 0.1hI, x pre-rolled from x0 = 1 + 0.1 N(0,1) with u = 0.1 N(0,1), cx = Qx, cu = Ru
(SURVEY.md section 8d, config C2).  The working set (~56 GB) is far larger than the 126 MB L2,
so no explicit L2 flush is needed between timed iterations.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
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
FP64_TENSOR_PEAK_TFLOPS = 37.08       # measured: profiles/microbench/ubench_r01_b200.txt (DMMA m8n8k4)
FP64_DFMA_PEAK_TFLOPS = 34.14         # measured, same file (DFMA pipe, sustained)
METRIC = "iLQG iters/sec (backward+forward), batch=65536 n=32 m=8 T=256"
WORKLOAD = ("C2: batched LTI linear dynamics n=32 m=8 T=256, 65536 trajectories per GPU, lambda=1 regType=1 "
            "no lims, alpha=1 (BASELINE.json configs[1])")
UNIT = "iters/s (1 iter = backward+forward sweep over 65536 trajectories, FP64)"


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


def cpu_step_time(inputs, nthreads=0):
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


def run_cpu_baseline(nsample, steps, warmup):
    from oracle import cpu_ref as CR
    CR.load()
    cores = CR.num_threads()
    inputs = cpu_sample_inputs(nsample)
    for _ in range(warmup):
        cpu_step_time(inputs)
    ts = [cpu_step_time(inputs) for _ in range(steps)]
    t = float(np.mean(ts))
    iters_per_s = 1.0 / (t * BATCH / nsample)       # trajectories are independent: linear extrapolation to 65 536
    return dict(value=iters_per_s, unit=UNIT, cores=cores, kind="port",
                sample=f"{nsample} of 65536 trajectories per step (n=32,m=8,T=256), {t:.3f} s/step on {cores} threads, "
                       f"scaled linearly to the full batch; C++/OpenMP restated reference (oracle/cpu_ref.cpp), not Julia"), t


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
                            impl_note="restated reference (C++/OpenMP port of back_pass + forward_pass), all host threads"),
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

    # ---- synthetic inputs, generated on the device (seeded per rank)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 + rank)
    G = torch.randn(B, n, n, dtype=f64, device=dev, generator=gen)
    A = torch.linalg.matrix_exp(H_STEP * (G - G.transpose(1, 2)))
    del G
    Bm = H_STEP * torch.randn(B, n, m, dtype=f64, device=dev, generator=gen)
    fx = A.transpose(1, 2).contiguous()              # column-major per trajectory == reference layout (n,n,B)
    fu = Bm.transpose(1, 2).contiguous()
    x0 = 1.0 + 0.1 * torch.randn(B, n, dtype=f64, device=dev, generator=gen)
    u = 0.1 * torch.randn(B, T, m, dtype=f64, device=dev, generator=gen)
    Q = (H_STEP * torch.eye(n, dtype=f64, device=dev)).contiguous()
    R = (0.1 * H_STEP * torch.eye(m, dtype=f64, device=dev)).contiguous()
    cxu = torch.zeros(m, n, dtype=f64, device=dev)
    lam = torch.ones(B, dtype=f64, device=dev)
    empty = lambda *s: torch.empty(*s, dtype=f64, device=dev)
    x, cost0 = empty(B, T, n), empty(B)
    K, k, Vx, dV = empty(B, T, n, m), empty(B, T, m), empty(B, T, n), empty(B, 2)
    xnew, unew, cost = empty(B, T, n), empty(B, T, m), empty(B)
    cx, cu = empty(B, T, n), empty(B, T, m)
    diverge = torch.empty(B, dtype=torch.int32, device=dev)
    stats = torch.zeros(8, dtype=f64, device=dev)
    del A, Bm

    eng = ddp.Engine(n, m, T, B, device=local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    tn = lambda t_, sb, st: L.Tensor(t_.data_ptr(), sb, st)

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
    fa = L.ForwardPassArgs()
    fa.K, fa.k = K.data_ptr(), k.data_ptr()
    fa.x0, fa.x, fa.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
    fa.alpha_scalar, fa.u_scale = 1.0, 1.0
    fa.xnew, fa.unew, fa.cost = xnew.data_ptr(), unew.data_ptr(), cost.data_ptr()

    ev = lambda: torch.cuda.Event(enable_timing=True)
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
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
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
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t0.elapsed_time(t1)
    launches = eng.launch_count - launches0 + (args.steps if world > 1 else 0)
    tms = torch.tensor([total_ms], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms = float(tms.item())
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
        # batch independence: the first S trajectories alone (another grid, another warp <-> trajectory map) give the same bits
        S = 1000
        eng_s = ddp.Engine(n, m, T, S, device=local_rank)
        eng_s.set_stream(torch.cuda.current_stream().cuda_stream)
        K_s, k_s, Vx_s, dV_s = empty(S, T, n, m), empty(S, T, m), empty(S, T, n), empty(S, 2)
        xn_s, un_s, c_s = empty(S, T, n), empty(S, T, m), empty(S)
        dv_s = torch.empty(S, dtype=torch.int32, device=dev)
        bs = L.BackPassArgs()
        for name in ("cx", "cu", "cxx", "cxu", "cuu", "fx", "fu", "lam", "reg_type"):
            setattr(bs, name, getattr(ba, name))
        bs.diverge, bs.K, bs.k, bs.Vx, bs.dV = dv_s.data_ptr(), K_s.data_ptr(), k_s.data_ptr(), Vx_s.data_ptr(), dV_s.data_ptr()
        fs = L.ForwardPassArgs()
        fs.K, fs.k = K_s.data_ptr(), k_s.data_ptr()
        fs.x0, fs.x, fs.u = fa.x0, fa.x, fa.u
        fs.alpha_scalar, fs.u_scale = 1.0, 1.0
        fs.xnew, fs.unew, fs.cost = xn_s.data_ptr(), un_s.data_ptr(), c_s.data_ptr()
        eng_s._ck(eng_s.lib.ddp_back_pass_f64(eng_s.h, C.byref(bs)))
        eng_s._ck(eng_s.lib.ddp_forward_pass_f64(eng_s.h, C.byref(model), C.byref(fs)))
        torch.cuda.synchronize()
        props["slice_bitwise_equal"] = bool(torch.equal(K_s, K[:S]) and torch.equal(k_s, k[:S]) and torch.equal(Vx_s, Vx[:S]) and
                                            torch.equal(xn_s, xnew[:S]) and torch.equal(c_s, cost[:S]) and torch.equal(dV_s, dV[:S]))
        eng_s.close()
        del K_s, k_s, Vx_s, dV_s, xn_s, un_s, c_s, dv_s
    except Exception as exc:
        props["error"] = str(exc)

    # ---- end-to-end: same step through ddp_ilqg_iter_host_f64 on pinned host buffers
    e2e = None
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
        # free the device-resident working set that the e2e path does not use
        eng_e = ddp.Engine(n, m, T, Be, device=local_rank)
        eng_e.set_stream(torch.cuda.current_stream().cuda_stream)
        it = ddp.HostIteration(eng_e, np.zeros((n, n)), np.zeros((m, m)), reg_type=1, alpha=1.0, chunk=args.chunk, device_derivs=True)
        it.Q[:] = Q.cpu().numpy(); it.R[:] = R.cpu().numpy()
        it.args.q_diagonal = 1
        for name, src in (("fx", fx), ("fu", fu), ("x", x), ("u", u), ("lam", lam)):
            it.bufs[name][...] = src[:Be].cpu().numpy()
        e_steps = max(1, min(args.steps, args.e2e_steps))
        it.run()                                       # warm-up (allocates the chunk pipeline)
        it.run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        w0 = time.perf_counter()
        for _ in range(e_steps):
            h2d, d2h = it.run()                        # synchronous: returns when the results are in host memory
        w1 = time.perf_counter()
        e_ms = (w1 - w0) * 1e3 / e_steps
        tme = torch.tensor([e_ms], dtype=f64, device=dev)
        if world > 1:
            dist.all_reduce(tme, op=dist.ReduceOp.MAX)
        e_ms = float(tme.item())
        ok = bool(np.array_equal(it.bufs["diverge"], diverge[:Be].cpu().numpy())) and \
            bool(np.allclose(it.bufs["cost"], cost[:Be].cpu().numpy(), rtol=1e-12, atol=0))
        e2e = dict(value=world * (Be / BATCH) * 1e3 / e_ms, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                   ms_per_step=e_ms, steps=e_steps, batch_per_gpu=Be, chunk=(args.chunk or "ramped: 1,2,4,...,4,2,1 rounds of sm_count*8 trajectories"),
                   api="ddp_ilqg_iter_host_f64: x,u,fx,fu,lambda from pinned host memory -> df (cx=Qx, cu=Ru) + backward + forward on the device -> xnew,unew,cost,dV,diverge in host memory; policy K stays on the device",
                   matches_device_path=ok, host_cpus_bound_to_gpu_numa_node=(len(numa_cpus) if numa_cpus else None))
        it.close()
        eng_e.close()
    except Exception as exc:                           # report, never fake
        e2e = dict(value=None, unit=UNIT, error=str(exc))

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        flops_back = FLOPS_BACK_STEP_SYM * (T - 1) * B
        ach_tf = flops_back / (bk * 1e-3) * 1e-12
        ref_tf = FLOPS_BACK_STEP * (T - 1) * B / (bk * 1e-3) * 1e-12
        roofline = dict(
            bound="tensor", kernel="bp_tile32x8_kernel (backward sweep, FP64 mma.sync m8n8k4 tiles)",
            achieved=ach_tf, peak=FP64_TENSOR_PEAK_TFLOPS, unit="TFLOP/s", frac=ach_tf / FP64_TENSOR_PEAK_TFLOPS,
            peak_source="FP64 tensor (DMMA) peak measured on this pool's B200 by profiles/microbench (MEASURED_PEAKS.json has no FP64 figure); "
                        "algorithmic flops = 172032/step (the symmetric halves of F'VF and of the Vxx update counted once: 336 m8n8k4 "
                        "tiles) x 255 steps x 65536 trajectories; SURVEY.md 8d's reference formulation (213419/step, full products) is "
                        "reported as reference_formulation_tflops and would read 1.0+ of the peak",
            reference_formulation_tflops=ref_tf, reference_formulation_frac=ref_tf / FP64_TENSOR_PEAK_TFLOPS,
            fp64_dfma_frac=ach_tf / FP64_DFMA_PEAK_TFLOPS,
            kernel_ms=bk, share_of_step=bk / ms_per_step,
            hbm=dict(achieved=BYTES_BACK * B / (bk * 1e-3) * 1e-9, peak=hbm_peak, unit="GB/s",
                     frac=BYTES_BACK * B / (bk * 1e-3) * 1e-9 / hbm_peak, peak_source=f"MEASURED_PEAKS.json ({peak_kind})"),
            forward=dict(kernel="fwd_lin32x8_kernel", kernel_ms=fw, bound="hbm", achieved=BYTES_FWD * B / (fw * 1e-3) * 1e-9,
                         peak=hbm_peak, unit="GB/s", frac=BYTES_FWD * B / (fw * 1e-3) * 1e-9 / hbm_peak),
            step_hbm_frac=(BYTES_BACK + BYTES_FWD) * B / (ms_per_step * 1e-3) * 1e-9 / hbm_peak,
            traffic=None)
        prof = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(prof):
            try:
                roofline["traffic"] = json.load(open(prof)).get("bp_tile32x8_kernel_dram_bytes_per_launch")
            except Exception:
                pass
        cpu_baseline = None
        try:
            os.sched_setaffinity(0, all_cpus)              # the CPU baseline uses every host core again
        except Exception:
            pass
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu_baseline, _ = run_cpu_baseline(args.cpu_sample, 2, 1)
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
                                else "single GPU", kernel_variant=eng.kernel_variant),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu_baseline,
                    check=dict(diverged=n_div, properties=props, mean_cost_new=float(stats_h[0] / max(stats_h[5], 1)),
                               accepted_frac=float(stats_h[3] / max(stats_h[5], 1))))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="trajectories per GPU (the metric is quoted at 65536)")
    ap.add_argument("--chunk", type=int, default=0, help="e2e pipeline chunk (trajectories); 0 = library default")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=1024, help="trajectories in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
