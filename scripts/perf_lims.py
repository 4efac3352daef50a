#!/usr/bin/env python
"""C2 with control limits: the n=32, m=8 backward sweep's box-QP branch on the tile kernel vs the generic kernel
(VERDICT r01 item 4).  usage: python scripts/perf_lims.py [B] [lims_scale]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ddp_b200 as ddp
from ddp_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
sc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
n, m, T, h = 32, 8, 256, 0.01
dev = torch.device("cuda", 0)
f64 = torch.float64
gen = torch.Generator(device=dev); gen.manual_seed(3)
G = torch.randn(B, n, n, dtype=f64, device=dev, generator=gen)
fx = torch.linalg.matrix_exp(h * (G - G.transpose(1, 2))).transpose(1, 2).contiguous()
fu = (h * torch.randn(B, n, m, dtype=f64, device=dev, generator=gen)).transpose(1, 2).contiguous()
x0 = 1.0 + 0.1 * torch.randn(B, n, dtype=f64, device=dev, generator=gen)
u = 0.1 * torch.randn(B, T, m, dtype=f64, device=dev, generator=gen)
Q = (h * torch.eye(n, dtype=f64, device=dev)).contiguous(); R = (0.1 * h * torch.eye(m, dtype=f64, device=dev)).contiguous()
cxu = torch.zeros(m, n, dtype=f64, device=dev); lam = torch.full((B,), 1e-2, dtype=f64, device=dev)
lims = torch.cat([-sc * (0.2 + 0.3 * torch.rand(m, dtype=f64, device=dev, generator=gen)), sc * (0.2 + 0.3 * torch.rand(m, dtype=f64, device=dev, generator=gen))]).contiguous()
e = lambda *s: torch.empty(*s, dtype=f64, device=dev)
tn = lambda t_, sb, st: L.Tensor(t_.data_ptr(), sb, st)
out = {}
for name, generic in (("tile", False), ("generic", True)):
    eng = ddp.Engine(n, m, T, B, force_generic=generic)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    model = L.Model(); model.kind = 1
    model.A, model.Bm, model.Q, model.R, model.flags = tn(fx, n * n, 0), tn(fu, n * m, 0), tn(Q, 0, 0), tn(R, 0, 0), 1
    x, c0, un, cx, cu = e(B, T, n), e(B), e(B, T, m), e(B, T, n), e(B, T, m)
    fa = L.ForwardPassArgs(); fa.x0, fa.u = tn(x0, n, 0), tn(u, T * m, m); fa.alpha_scalar = fa.u_scale = 1.0
    fa.xnew, fa.unew, fa.cost, fa.cx, fa.cu = x.data_ptr(), un.data_ptr(), c0.data_ptr(), cx.data_ptr(), cu.data_ptr()
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(model), C.byref(fa)))
    K, k, Vx, dV = e(B, T, n, m), e(B, T, m), e(B, T, n), e(B, 2)
    dv = torch.empty(B, dtype=torch.int32, device=dev)
    res = {}
    for use_lims in (False, True):
        ba = L.BackPassArgs()
        ba.cx, ba.cu, ba.cxx, ba.cxu, ba.cuu = tn(cx, T * n, n), tn(cu, T * m, m), tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
        ba.fx, ba.fu, ba.lam, ba.reg_type = tn(fx, n * n, 0), tn(fu, n * m, 0), lam.data_ptr(), 1
        if use_lims:
            ba.lims, ba.u = lims.data_ptr(), tn(u, T * m, m)
        ba.diverge, ba.K, ba.k, ba.Vx, ba.dV = dv.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr()
        for _ in range(2):
            eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 3
        rows = (K[:, :T - 1].abs().sum(dim=2) == 0).double().mean().item() if use_lims else 0.0
        res["lims" if use_lims else "no_lims"] = dict(ms=ms, ms_per_65536=ms * 65536 / B, diverged=int((dv > 0).sum().item()), clamped_row_frac=rows)
        if use_lims:
            out.setdefault("_k", {})[name] = k.clone()
    out[name] = res
    eng.close()
out["k_bitwise_equal_tile_vs_generic"] = bool(torch.equal(out["_k"]["tile"], out["_k"]["generic"]))
del out["_k"]
out["B"] = B
print(json.dumps(out))
