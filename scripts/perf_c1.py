#!/usr/bin/env python
"""BASELINE configs[0]'s shape as a batch: demo_linear n=10 m=2 T=1000 (demo_linear.jl:5-60) -- one back_pass + forward_pass.
usage: python scripts/perf_c1.py [B]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ddp_b200 as ddp
from ddp_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n, m, T, h = 10, 2, 1000, 0.01
dev = torch.device("cuda", 0)
f64 = torch.float64
gen = torch.Generator(device=dev); gen.manual_seed(3)
G = torch.randn(B, n, n, dtype=f64, device=dev, generator=gen)
fx = torch.linalg.matrix_exp(h * (G - G.transpose(1, 2))).transpose(1, 2).contiguous()
fu = (h * torch.randn(B, n, m, dtype=f64, device=dev, generator=gen)).transpose(1, 2).contiguous()
x0 = torch.ones(B, n, dtype=f64, device=dev)
u = 0.1 * torch.randn(B, T, m, dtype=f64, device=dev, generator=gen)
Q = (h * torch.eye(n, dtype=f64, device=dev)).contiguous(); R = (0.1 * h * torch.eye(m, dtype=f64, device=dev)).contiguous()
cxu = torch.zeros(m, n, dtype=f64, device=dev); lam = torch.ones(B, dtype=f64, device=dev)
e = lambda *s: torch.empty(*s, dtype=f64, device=dev)
tn = lambda t_, sb, st: L.Tensor(t_.data_ptr(), sb, st)
eng = ddp.Engine(n, m, T, B)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
model = L.Model(); model.kind = 1
model.A, model.Bm, model.Q, model.R, model.flags = tn(fx, n * n, 0), tn(fu, n * m, 0), tn(Q, 0, 0), tn(R, 0, 0), 1
x, c0, un, cx, cu = e(B, T, n), e(B), e(B, T, m), e(B, T, n), e(B, T, m)
fa = L.ForwardPassArgs(); fa.x0, fa.u = tn(x0, n, 0), tn(u, T * m, m); fa.alpha_scalar = fa.u_scale = 1.0
fa.xnew, fa.unew, fa.cost, fa.cx, fa.cu = x.data_ptr(), un.data_ptr(), c0.data_ptr(), cx.data_ptr(), cu.data_ptr()
eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa)))
K, k, Vx, dV = e(B, T, n, m), e(B, T, m), e(B, T, n), e(B, 2)
dv = torch.empty(B, dtype=torch.int32, device=dev)
ba = L.BackPassArgs()
ba.cx, ba.cu, ba.cxx, ba.cxu, ba.cuu = tn(cx, T * n, n), tn(cu, T * m, m), tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
ba.fx, ba.fu, ba.lam, ba.reg_type = tn(fx, n * n, 0), tn(fu, n * m, 0), lam.data_ptr(), 1
ba.diverge, ba.K, ba.k, ba.Vx, ba.dV = dv.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr()
xn, unw, cst = e(B, T, n), e(B, T, m), e(B)
fp = L.ForwardPassArgs()
fp.K, fp.k = K.data_ptr(), k.data_ptr()
fp.x0, fp.x, fp.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
fp.alpha_scalar, fp.u_scale = 1.0, 1.0
fp.xnew, fp.unew, fp.cost = xn.data_ptr(), unw.data_ptr(), cst.data_ptr()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / reps


bk = timed(lambda: eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba))))
fw = timed(lambda: eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fp))))
byt_b = 8.0 * (T * (n + m) + T * (n * m + m + n) + n * n + n * m) * B
byt_f = 8.0 * (T * (n * m + m + 2 * n + m) + T * (n + m)) * B
print(json.dumps(dict(B=B, n=n, m=m, T=T, back_ms=bk, fwd_ms=fw, back_hbm_gbs=byt_b / bk * 1e-6, fwd_hbm_gbs=byt_f / fw * 1e-6,
                      diverged=int((dv > 0).sum().item()), iters_per_s=1e3 / (bk + fw))))
