"""Quick device-timing probe for the backward/forward kernels (development aid, not the bench)."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddp_b200 as ddp
from ddp_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
LTV = len(sys.argv) > 2 and sys.argv[2] == "ltv"          # time-varying dynamics: fx, fu materialised as [B,T,.,.]
n, m, T = 32, 8, 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
h = 0.01
G = torch.randn(B, n, n, dtype=torch.float64, device=dev)
A = torch.linalg.matrix_exp(h * (G - G.transpose(1, 2)))
Bm = h * torch.randn(B, n, m, dtype=torch.float64, device=dev)
fx = A.transpose(1, 2).contiguous()      # column-major per trajectory
fu = Bm.transpose(1, 2).contiguous()
cx = 0.01 * torch.randn(B, T, n, dtype=torch.float64, device=dev)
cu = 0.001 * torch.randn(B, T, m, dtype=torch.float64, device=dev)
Q = (h * torch.eye(n, dtype=torch.float64, device=dev)).contiguous()
R = (0.1 * h * torch.eye(m, dtype=torch.float64, device=dev)).contiguous()
cxu = torch.zeros(m, n, dtype=torch.float64, device=dev)
lam = torch.ones(B, dtype=torch.float64, device=dev)
K = torch.empty(B, T, n, m, dtype=torch.float64, device=dev)
k = torch.empty(B, T, m, dtype=torch.float64, device=dev)
Vx = torch.empty(B, T, n, dtype=torch.float64, device=dev)
dV = torch.empty(B, 2, dtype=torch.float64, device=dev)
dv = torch.empty(B, dtype=torch.int32, device=dev)
eng = ddp.Engine(n, m, T, B)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
a = L.BackPassArgs()
t = lambda x, sb, st: L.Tensor(x.data_ptr(), sb, st)
a.cx, a.cu = t(cx, T * n, n), t(cu, T * m, m)
a.cxx, a.cxu, a.cuu = t(Q, 0, 0), t(cxu, 0, 0), t(R, 0, 0)
if LTV:
    fx = fx[:, None].expand(B, T, n, n).contiguous(); fu = fu[:, None].expand(B, T, m, n).contiguous()
    a.fx, a.fu = t(fx, T * n * n, n * n), t(fu, T * n * m, n * m)
else:
    a.fx, a.fu = t(fx, n * n, 0), t(fu, n * m, 0)
a.lam, a.reg_type = lam.data_ptr(), 1
a.diverge, a.K, a.k, a.Vx, a.dV = dv.data_ptr(), K.data_ptr(), k.data_ptr(), Vx.data_ptr(), dV.data_ptr()
print("variant", eng.kernel_variant)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(a)))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flops = 213419.0 * (T - 1) * B
    print(f"back_pass {'LTV ' if LTV else ''}B={B}: {ms:.2f} ms  -> {flops / ms * 1e-9:.2f} TFLOP/s algorithmic; full-batch(65536) est {ms * 65536 / B:.1f} ms; diverged {int((dv>0).sum())}")
