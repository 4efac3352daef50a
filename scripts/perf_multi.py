#!/usr/bin/env python
"""Line search at the headline shape: the cost of NA step sizes on one policy -- NA serial rollouts (fwd_lin32x8_kernel) vs the
FMA multi-alpha kernel vs the FP64 tensor-tile multi-alpha kernel.  usage: perf_multi.py [B] [NA]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ddp_b200 as ddp
from ddp_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
NA = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n, m, T, h = 32, 8, 256, 0.01
dev = torch.device("cuda", 0)
f64 = torch.float64
gen = torch.Generator(device=dev); gen.manual_seed(3)
fx = torch.empty(B, n, n, dtype=f64, device=dev)
for b0 in range(0, B, 16384):
    G = torch.randn(min(16384, B - b0), n, n, dtype=f64, device=dev, generator=gen)
    fx[b0:b0 + G.shape[0]] = torch.linalg.matrix_exp(h * (G - G.transpose(1, 2))).transpose(1, 2)
    del G
fu = (h * torch.randn(B, n, m, dtype=f64, device=dev, generator=gen)).transpose(1, 2).contiguous()
x0 = 1.0 + 0.1 * torch.randn(B, n, dtype=f64, device=dev, generator=gen)
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
eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
xn, unw, cst = e(B, T, n), e(B, T, m), e(B)
fp = L.ForwardPassArgs()
fp.K, fp.k = K.data_ptr(), k.data_ptr()
fp.x0, fp.x, fp.u = tn(x, T * n, 0), tn(x, T * n, n), tn(u, T * m, m)
fp.alpha_scalar, fp.u_scale = 1.0, 1.0
fp.xnew, fp.unew, fp.cost = xn.data_ptr(), unw.data_ptr(), cst.data_ptr()
alphas = np.ascontiguousarray(10.0 ** np.linspace(0, -3, NA))
costs = e(NA, B)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / reps


def serial():
    for a in alphas:
        fp.alpha_scalar = float(a)
        eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fp)))


def multi():
    eng._ck(eng.lib.ddp_forward_costs_multi_f64(eng.h, C.byref(model), C.byref(fp), NA, alphas.ctypes.data_as(L.c_double_p), costs.data_ptr()))


out = dict(B=B, n_alpha=NA)
out["serial_ms"] = timed(serial)
ref = []
for a in alphas:
    fp.alpha_scalar = float(a)
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fp)))
    ref.append(cst.clone())
ref = torch.stack(ref)
os.environ.pop("DDP_MULTI_NO_TILE", None)
out["multi_tile_ms"] = timed(multi)
out["multi_tile_max_rel_err_vs_serial"] = float(((costs - ref).abs() / ref.abs()).max().item())
os.environ["DDP_MULTI_NO_TILE"] = "1"
out["multi_fma_ms"] = timed(multi)
out["multi_fma_bitwise_equal_serial"] = bool(torch.equal(costs, ref))
out["speedup_tile_vs_serial"] = out["serial_ms"] / out["multi_tile_ms"]
out["dmma_frac_of_peak_37tf"] = (100 if NA > 8 else 52) * 512.0 * T * B / (out["multi_tile_ms"] * 1e-3) * 1e-12 / 37.09
print(json.dumps(out))
