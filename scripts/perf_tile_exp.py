#!/usr/bin/env python
"""Backward tile kernel at the headline shape (n=32, m=8, T=256) for the experimental schedules selected by
DDP_TILE_EXP.  usage: python scripts/perf_tile_exp.py [B] [exp,exp,...]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ddp_b200 as ddp
from ddp_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n, m, T, h = 32, 8, 256, 0.01
dev = torch.device("cuda", 0)
f64 = torch.float64
gen = torch.Generator(device=dev); gen.manual_seed(3)
G = torch.randn(B, n, n, dtype=f64, device=dev, generator=gen)
fx = torch.linalg.matrix_exp(h * (G - G.transpose(1, 2))).transpose(1, 2).contiguous()
del G
fu = (h * torch.randn(B, n, m, dtype=f64, device=dev, generator=gen)).transpose(1, 2).contiguous()
cx = 0.01 * torch.randn(B, T, n, dtype=f64, device=dev, generator=gen)
cu = 0.01 * torch.randn(B, T, m, dtype=f64, device=dev, generator=gen)
Q = (h * torch.eye(n, dtype=f64, device=dev)).contiguous(); R = (0.1 * h * torch.eye(m, dtype=f64, device=dev)).contiguous()
cxu = torch.zeros(m, n, dtype=f64, device=dev); lam = torch.full((B,), 1e-2, dtype=f64, device=dev)
e = lambda *s: torch.empty(*s, dtype=f64, device=dev)
tn = lambda t_, sb, st: L.Tensor(t_.data_ptr(), sb, st)
eng = ddp.Engine(n, m, T, B)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
K, k, Vx, dV = e(B, T, n, m), e(B, T, m), e(B, T, n), e(B, 2)
dv = torch.empty(B, dtype=torch.int32, device=dev)
ba = L.BackPassArgs()
ba.cx, ba.cu, ba.cxx, ba.cxu, ba.cuu = tn(cx, T * n, n), tn(cu, T * m, m), tn(Q, 0, 0), tn(cxu, 0, 0), tn(R, 0, 0)
ba.fx, ba.fu, ba.lam, ba.reg_type = tn(fx, n * n, 0), tn(fu, n * m, 0), lam.data_ptr(), 1
ba.diverge, ba.K, ba.k, ba.Vx, ba.dV = dv.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr()


def run(reps=4):
    for _ in range(2):
        eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(ba)))
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


out = []
os.environ.pop("DDP_TILE_STAGGER", None)
os.environ.pop("DDP_TILE_EXP", None)
base = run()
Kref, kref, Vxref = K.clone(), k.clone(), Vx.clone()
out.append(dict(exp=0, ms=base))
for ex in [int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4"])]:
    os.environ["DDP_TILE_EXP"] = str(ex)
    ms = run()
    err = max(((K - Kref).abs().max() / Kref.abs().max()).item(), ((Vx - Vxref).abs().max() / Vxref.abs().max()).item())
    out.append(dict(exp=ex, ms=ms, max_rel_diff_vs_base=err, diverged=int((dv > 0).sum().item())))
    print(out[-1], file=sys.stderr)
os.environ.pop("DDP_TILE_EXP", None)
out.append(dict(exp=0, ms=run()))
print(json.dumps(dict(B=B, runs=out)))
