"""Device timings of the other BASELINE.json configs (parity-test cases, not bench lines):
   C3 pendcart n=4 m=1 T=600 with lims (boxQP path), C4 KL-augmented sweep on C2's system.
   python scripts/bench_configs.py c3 [B]   |   python scripts/bench_configs.py c4 [B]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ddp_b200 as ddp
from ddp_b200 import _lib as L

dev = torch.device("cuda:0")
f64 = torch.float64
tn = lambda t, sb, st: L.Tensor(t.data_ptr(), sb, st)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = ev(), ev(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def c3(B):
    n, m, T = 4, 1, 600
    g = torch.Generator(device=dev); g.manual_seed(1)
    x0 = torch.zeros(B, n, dtype=f64, device=dev)
    x0[:, 0] = np.pi - 0.6 + 0.2 * (2 * torch.rand(B, dtype=f64, device=dev, generator=g) - 1)
    u = torch.zeros(B, T, m, dtype=f64, device=dev)
    Q = torch.diag(torch.tensor([10.0, 1, 2, 1], dtype=f64, device=dev)).contiguous(); R = torch.ones(1, 1, dtype=f64, device=dev)
    goal = torch.tensor([np.pi, 0, 0, 0], dtype=f64, device=dev)
    lims = torch.tensor([-5.0, 5.0], dtype=f64, device=dev)
    E = lambda *s: torch.empty(*s, dtype=f64, device=dev)
    x, un, cost, cx, cu, fx, fu = E(B, T, n), E(B, T, m), E(B), E(B, T, n), E(B, T, m), E(B, T, n, n), E(B, T, n, m)
    K, k, Vx, dV, xnew, unew, cnew = E(B, T, n, m), E(B, T, m), E(B, T, n), E(B, 2), E(B, T, n), E(B, T, m), E(B)
    dv = torch.empty(B, dtype=torch.int32, device=dev); lam = torch.ones(B, dtype=f64, device=dev); cxu = torch.zeros(m, n, dtype=f64, device=dev)
    eng = ddp.Engine(n, m, T, B); eng.set_stream(torch.cuda.current_stream().cuda_stream)
    M = L.Model(); M.kind = 2; M.Q, M.R = tn(Q, 0, 0), tn(R, 0, 0); M.goal = goal.data_ptr(); M.terminal_cost = 1
    for i, v in enumerate((9.82, 0.35, 0.01, 0.99)): M.p[i] = v
    fa = L.ForwardPassArgs(); fa.x0, fa.u = tn(x0, n, 0), tn(u, T * m, m); fa.alpha_scalar = fa.u_scale = 1.0; fa.lims = lims.data_ptr()
    fa.xnew, fa.unew, fa.cost = x.data_ptr(), un.data_ptr(), cost.data_ptr()
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(M), C.byref(fa)))
    t_df = timeit(lambda: eng._ck(eng.lib.ddp_model_derivs_f64(eng.h, C.byref(M), x.data_ptr(), un.data_ptr(), fx.data_ptr(), fu.data_ptr(), cx.data_ptr(), cu.data_ptr())))
    ba = L.BackPassArgs(); ba.cx, ba.cu = tn(cx, T * n, n), tn(cu, T * m, m); ba.cxx, ba.cxu, ba.cuu = tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
    ba.fx, ba.fu = tn(fx, T * n * n, n * n), tn(fu, T * n * m, n * m); ba.lam, ba.reg_type = lam.data_ptr(), 2; ba.lims = lims.data_ptr(); ba.u = tn(un, T * m, m)
    ba.diverge, ba.K, ba.k, ba.Vx, ba.dV = dv.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr()
    t_b = timeit(lambda: eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba))))
    f2 = L.ForwardPassArgs(); f2.K, f2.k = K.data_ptr(), k.data_ptr(); f2.x0, f2.x, f2.u = tn(x0, n, 0), tn(x, T * n, n), tn(un, T * m, m)
    f2.alpha_scalar = f2.u_scale = 1.0; f2.lims = lims.data_ptr(); f2.xnew, f2.unew, f2.cost = xnew.data_ptr(), unew.data_ptr(), cnew.data_ptr()
    t_f = timeit(lambda: eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(M), C.byref(f2))))
    nclamp = int(((K[:, :-1].abs().sum(dim=(2, 3)) == 0)).sum().item())
    bytes_b, bytes_f = (124824 + 43220) * B, (48000 + 28808) * B
    print(json.dumps(dict(config="C3 pendcart n=4 m=1 T=600 lims=+-5 regType=2 (boxQP branch)", batch=B, variant=eng.kernel_variant,
                          back_ms=t_b, fwd_ms=t_f, df_ms=t_df, iters_per_s=1e3 / (t_b + t_f), diverged=int((dv > 0).sum().item()),
                          clamped_steps=nclamp, back_hbm_GBs=bytes_b / t_b * 1e-6, fwd_hbm_GBs=bytes_f / t_f * 1e-6,
                          hbm_frac_of_6552=(bytes_b + bytes_f) / (t_b + t_f) * 1e-6 / 6552.6)))


def c4(B):
    n, m, T, h = 32, 8, 256, 0.01
    g = torch.Generator(device=dev); g.manual_seed(0)
    G = torch.randn(B, n, n, dtype=f64, device=dev, generator=g)
    A = torch.linalg.matrix_exp(h * (G - G.transpose(1, 2))); Bm = h * torch.randn(B, n, m, dtype=f64, device=dev, generator=g)
    fx, fu = A.transpose(1, 2).contiguous(), Bm.transpose(1, 2).contiguous()
    Q = (h * torch.eye(n, dtype=f64, device=dev)).contiguous(); R = (0.1 * h * torch.eye(m, dtype=f64, device=dev)).contiguous()
    E = lambda *s: torch.empty(*s, dtype=f64, device=dev)
    cx, cu = 0.01 * torch.randn(B, T, n, dtype=f64, device=dev, generator=g), 0.001 * torch.randn(B, T, m, dtype=f64, device=dev, generator=g)
    K, k, Vx, dV, Quu, Quui = E(B, T, n, m), E(B, T, m), E(B, T, n), E(B, 2), E(B, T, m, m), E(B, T, m, m)
    dv = torch.empty(B, dtype=torch.int32, device=dev); lam = torch.ones(B, dtype=f64, device=dev); cxu = torch.zeros(m, n, dtype=f64, device=dev)
    eng = ddp.Engine(n, m, T, B); eng.set_stream(torch.cuda.current_stream().cuda_stream)
    ba = L.BackPassArgs(); ba.cx, ba.cu = tn(cx, T * n, n), tn(cu, T * m, m); ba.cxx, ba.cxu, ba.cuu = tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
    ba.fx, ba.fu = tn(fx, n * n, 0), tn(fu, n * m, 0); ba.lam, ba.reg_type = lam.data_ptr(), 1
    ba.diverge, ba.K, ba.k, ba.Vx, ba.dV, ba.Quu = dv.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr(), Quu.data_ptr()
    eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba))); torch.cuda.synchronize()
    Kp, Sip = K.clone(), Quu.clone()                       # traj_prev: gains of one plain back pass, Sigma_i = Quu
    eta = torch.ones(B, dtype=f64, device=dev)
    gp = L.GpsArgs(); gp.K_prev, gp.Sigi_prev = tn(Kp, T * n * m, n * m), tn(Sip, T * m * m, m * m); gp.eta = eta.data_ptr(); gp.Quui = Quui.data_ptr()
    t_g = timeit(lambda: eng._ck(eng.lib.ddp_back_pass_gps_f64(eng.h, C.byref(ba), C.byref(gp))), reps=2)
    # forward rollout (alpha = 1) with the new policy and the KL evaluation (forward_covariance + kl_div_wiki), iLQGkl.jl:134-143
    x0 = 1.0 + 0.1 * torch.randn(B, n, dtype=f64, device=dev, generator=g)
    u = 0.1 * torch.randn(B, T, m, dtype=f64, device=dev, generator=g)
    x, xnew, unew, cnew = E(B, T, n), E(B, T, n), E(B, T, m), E(B)
    M = L.Model(); M.kind = 1; M.A, M.Bm = tn(fx, n * n, 0), tn(fu, n * m, 0); M.Q, M.R = tn(Q, 0, 0), tn(R, 0, 0); M.flags = 1
    f0 = L.ForwardPassArgs(); f0.x0, f0.u = tn(x0, n, 0), tn(u, T * m, m); f0.alpha_scalar = f0.u_scale = 1.0
    f0.xnew, f0.unew, f0.cost = x.data_ptr(), unew.data_ptr(), cnew.data_ptr()
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(M), C.byref(f0)))
    f2 = L.ForwardPassArgs(); f2.K, f2.k = K.data_ptr(), k.data_ptr(); f2.x0, f2.x, f2.u = tn(x0, n, 0), tn(x, T * n, n), tn(u, T * m, m)
    f2.alpha_scalar = f2.u_scale = 1.0; f2.xnew, f2.unew, f2.cost = xnew.data_ptr(), unew.data_ptr(), cnew.data_ptr()
    t_f = timeit(lambda: eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(M), C.byref(f2))))
    Sp = torch.linalg.inv(Sip).contiguous()
    R1 = (1e-4 * torch.eye(n, dtype=f64, device=dev)).contiguous()
    klm = E(B)
    ka = L.KlArgs(); ka.fx, ka.R1 = tn(fx, n * n, 0), tn(R1, 0, 0); ka.xnew, ka.xold = xnew.data_ptr(), x.data_ptr()
    ka.K_new, ka.k_new, ka.Sig_new = K.data_ptr(), k.data_ptr(), Quui.data_ptr()
    ka.K_prev, ka.Sig_prev, ka.Sigi_prev = tn(Kp, T * n * m, n * m), tn(Sp, T * m * m, m * m), tn(Sip, T * m * m, m * m)
    ka.kl_mean = klm.data_ptr()
    t_k = timeit(lambda: eng._ck(eng.lib.ddp_kl_div_f64(eng.h, C.byref(ka))), reps=2)
    # multi-alpha line-search costs: 10 step sizes in one pass over K
    al = (C.c_double * 10)(*[10.0 ** (-0.3 * i) for i in range(10)])
    mc = E(10, B)
    t_m = timeit(lambda: eng._ck(eng.lib.ddp_forward_costs_multi_f64(eng.h, C.byref(M), C.byref(f2), 10, al, mc.data_ptr())), reps=2)
    sc = 65536 / B
    print(json.dumps(dict(config="C4 KL-augmented sweep + forward + KL evaluation on C2's system (eta=1)", batch=B, gps_back_ms=t_g, fwd_ms=t_f,
                          kl_div_ms=t_k, multi_alpha10_ms=t_m, iter_ms_scaled_to_65536=(t_g + t_f + t_k) * sc,
                          gps_back_ms_scaled=t_g * sc, kl_div_ms_scaled=t_k * sc, multi_alpha10_ms_scaled=t_m * sc,
                          kl_mean=float(klm.mean().item()), diverged=int((dv > 0).sum().item()), variant=eng.kernel_variant)))


def solve(which, B):
    """Whole device-resident iLQG solves (ddp_ilqg_solve_f64): wall time, outer iterations, final status histogram."""
    import time
    if which == "c2":
        n, m, T, h = 32, 8, 256, 0.01
        g = torch.Generator(device=dev); g.manual_seed(0)
        G = torch.randn(B, n, n, dtype=f64, device=dev, generator=g)
        A = torch.linalg.matrix_exp(h * (G - G.transpose(1, 2))); Bm = h * torch.randn(B, n, m, dtype=f64, device=dev, generator=g)
        fx, fu = A.transpose(1, 2).contiguous(), Bm.transpose(1, 2).contiguous()
        Q = (h * torch.eye(n, dtype=f64, device=dev)).contiguous(); R = (0.1 * h * torch.eye(m, dtype=f64, device=dev)).contiguous()
        M = L.Model(); M.kind = 1; M.A, M.Bm = tn(fx, n * n, 0), tn(fu, n * m, 0); M.Q, M.R = tn(Q, 0, 0), tn(R, 0, 0); M.flags = 1
        x0 = 1.0 + 0.1 * torch.randn(B, n, dtype=f64, device=dev, generator=g)
        u0 = 0.1 * torch.randn(B, T, m, dtype=f64, device=dev, generator=g)
        alpha = [10.0 ** (-3.0 * i / 10) for i in range(11)]
        o = L.IlqgOpts(); o.tol_fun, o.tol_grad, o.max_iter = 1e-7, 1e-4, 500; o.reg_type = 1
        o.lam, o.dlam, o.lam_factor, o.lam_max, o.lam_min = 1.0, 1.0, 1.6, 1e10, 1e-6
        keep = (fx, fu, Q, R)
    else:
        n, m, T = 4, 1, 600
        g = torch.Generator(device=dev); g.manual_seed(1)
        x0 = torch.zeros(B, n, dtype=f64, device=dev)
        x0[:, 0] = np.pi - 0.6 + 0.2 * (2 * torch.rand(B, dtype=f64, device=dev, generator=g) - 1)
        u0 = torch.zeros(B, T, m, dtype=f64, device=dev)
        Q = torch.diag(torch.tensor([10.0, 1, 2, 1], dtype=f64, device=dev)).contiguous(); R = torch.ones(1, 1, dtype=f64, device=dev)
        goal = torch.tensor([np.pi, 0, 0, 0], dtype=f64, device=dev)
        lims = torch.tensor([-5.0, 5.0], dtype=f64, device=dev)
        M = L.Model(); M.kind = 2; M.Q, M.R = tn(Q, 0, 0), tn(R, 0, 0); M.goal = goal.data_ptr(); M.terminal_cost = 1
        for i, v in enumerate((9.82, 0.35, 0.01, 0.99)): M.p[i] = v
        alpha = [10.0 ** (0.2 - 3.2 * i / 5) for i in range(6)]                    # system_pendcart.jl:197-206
        o = L.IlqgOpts(); o.tol_fun, o.tol_grad, o.max_iter = 1e-8, 1e-8, int(os.environ.get("C3_MAX_ITER", "30")); o.reg_type = 2
        o.lam, o.dlam, o.lam_factor, o.lam_max, o.lam_min = 1.0, 1.0, 1.6, 1e15, 1e-6
        o.lims = lims.data_ptr()
        keep = (Q, R, goal, lims)
    o.n_alpha = len(alpha)
    for i, v in enumerate(alpha): o.alpha[i] = v
    E = lambda *s_: torch.zeros(*s_, dtype=f64, device=dev)
    x, u, K, k, Vx, Vxx1 = E(B, T, n), E(B, T, m), E(B, T, n, m), E(B, T, m), E(B, T, n), E(B, n, n)
    st = torch.zeros(B, C.sizeof(L.IlqgState), dtype=torch.uint8, device=dev)
    eng = ddp.Engine(n, m, T, B); eng.set_stream(torch.cuda.current_stream().cuda_stream)
    nouter = C.c_int32(0)
    walls = []
    for rep in range(2):                                  # the second call re-uses the handle's workspace arena
        torch.cuda.synchronize(); l0 = eng.launch_count; t0 = time.perf_counter()
        eng._ck(eng.lib.ddp_ilqg_solve_f64(eng.h, C.byref(M), C.byref(o), x0.data_ptr(), u0.data_ptr(), x.data_ptr(), u.data_ptr(), K.data_ptr(),
                                           k.data_ptr(), Vx.data_ptr(), Vxx1.data_ptr(), st.data_ptr(), C.byref(nouter)))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        walls.append(dt)
    sa = np.frombuffer(st.cpu().numpy().tobytes(), dtype=np.dtype([("lam", "f8"), ("dlam", "f8"), ("cost", "f8"), ("g_norm", "f8"), ("last_dcost", "f8"),
                                                                   ("last_alpha", "f8"), ("iter", "i4"), ("accepted_iter", "i4"), ("status", "i4"), ("pad", "i4")]))
    hist = {int(s_): int((sa["status"] == s_).sum()) for s_ in np.unique(sa["status"])}
    print(json.dumps(dict(config=f"whole iLQG solve ({which}), device resident", batch=B, wall_s=dt, wall_s_first_call=walls[0], n_outer=int(nouter.value), launches=eng.launch_count - l0,
                          status_hist=hist, iter_mean=float(sa["iter"].mean()), iter_max=int(sa["iter"].max()), cost_mean=float(sa["cost"].mean()),
                          trajectory_iterations_per_s=float(sa["iter"].sum() / dt))))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    if which == "c3":
        c3(int(sys.argv[2]) if len(sys.argv) > 2 else 262144)
    elif which == "solve":
        solve(sys.argv[2], int(sys.argv[3]))
    else:
        c4(int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
