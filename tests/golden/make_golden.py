"""Generates tests/golden/*.npz from the CPU oracle (oracle/ddp_oracle.py) on seeded inputs.

The reference ships no golden vectors and cannot be executed here (Julia is absent), so these
fixtures freeze the ORACLE's outputs -- the restatement that the analytic known-answer tests pin
(tests/test_oracle_kat.py).  They guard both the oracle and the CUDA path against regressions:
tests/test_golden.py checks the oracle against them on CPU and the GPU path against them with
-m gpu.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import make_batch_lq, make_lq  # noqa: E402
from oracle import ddp_oracle as O  # noqa: E402


def back_pass_case(name, n, m, N, regType, lims, lam, seed):
    A, Bm, Q, R, x, u = make_batch_lq(seed, 2, n, m, N)
    cxu = 0.01 * np.random.default_rng(seed + 1).standard_normal((n, m))
    lim = None if lims is None else np.tile(np.array([[-lims, lims]]), (m, 1))
    out = dict(A=A, Bm=Bm, Q=Q, R=R, x=x, u=u, cxu=cxu, lam=lam, regType=regType, lims=np.zeros((0, 2)) if lim is None else lim)
    Ks, ks, Vxs, Vxxs, dVs, dvs = [], [], [], [], [], []
    for b in range(2):
        d, p, Vx, Vxx, dV = O.back_pass(x[b] @ Q.T, u[b] @ R.T, Q, cxu, R, A[b], Bm[b], lam, regType, lim, x[b], u[b])
        Ks.append(p.K); ks.append(p.k); Vxs.append(Vx); Vxxs.append(Vxx); dVs.append(dV); dvs.append(d)
    out.update(K=np.array(Ks), k=np.array(ks), Vx=np.array(Vxs), Vxx=np.array(Vxxs), dV=np.array(dVs), diverge=np.array(dvs))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def boxqp_case():
    rng = np.random.default_rng(77)
    m, B = 5, 24
    Hs, xs, rs, frees, nfs = [], [], [], [], []
    g = 3 * rng.standard_normal((B, m)); lo = -rng.random((B, m)); up = rng.random((B, m)); x0 = rng.standard_normal((B, m))
    for b in range(B):
        G = rng.standard_normal((m, m)); H = G @ G.T + 0.05 * np.eye(m); H = (H + H.T) / 2
        x, r, Hf, free, nf = O.boxQP(H, g[b], lo[b], up[b], x0[b])
        Hs.append(H); xs.append(x); rs.append(r); frees.append(free); nfs.append(nf)
    np.savez_compressed(os.path.join(HERE, "boxqp_m5.npz"), H=np.array(Hs), g=g, lower=lo, upper=up, x0=x0, x=np.array(xs),
                        result=np.array(rs), free=np.array(frees), nfactor=np.array(nfs))


def ilqg_case():
    rng = np.random.default_rng(5)
    n, m, N = 8, 2, 60
    A, Bm, Q, R = make_lq(rng, n, m)
    u0 = 0.1 * rng.standard_normal((N, m))
    om = O.LinearModel(A, Bm, Q, R)
    x, u, pol, Vx, Vxx, cost, tr = O.iLQG(om.f, om.costfun, om.df, np.ones(n), u0)
    np.savez_compressed(os.path.join(HERE, "ilqg_lq_n8.npz"), A=A, Bm=Bm, Q=Q, R=R, u0=u0, x=x, u=u, K=pol.K, k=pol.k, cost=np.sum(cost),
                        status=tr["status"], iters=tr["iters"], lam_final=tr["lam_final"],
                        cost_trace=np.array([c for _, c in tr["cost"]]), alpha_trace=np.array([a for _, a in tr["alpha"]]))


if __name__ == "__main__":
    back_pass_case("back_pass_n10_m2_chol", 10, 2, 40, 1, None, 1.0, 101)
    back_pass_case("back_pass_n32_m8_reg2", 32, 8, 16, 2, None, 0.5, 102)
    back_pass_case("back_pass_n4_m1_lims", 4, 1, 50, 2, 0.05, 1e-3, 103)
    back_pass_case("back_pass_n6_m3_lims", 6, 3, 30, 1, 0.05, 1e-3, 104)
    boxqp_case()
    ilqg_case()
    print(sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
