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


def ilqgkl_case():
    """KL-constrained path at the headline shape (n=32, m=8): one KL-augmented sweep, the KL evaluation and a whole iLQGkl solve."""
    from helpers import rollout
    n, m, N = 32, 8, 16
    rng = np.random.default_rng(9)
    A, Bm, Q, R = make_lq(rng, n, m, h=0.1)
    u = 0.1 * rng.standard_normal((N, m))
    x = rollout(A, Bm, np.ones(n), u)
    d, p, _, _, _ = O.back_pass(x @ Q.T, u @ R.T, Q, np.zeros((n, m)), R, A, Bm, 1.0, 1, None, x, u)
    Sigi = p.Sigmai.copy()
    Sig = np.array([np.linalg.inv(s) for s in Sigi])
    R1 = 1e-3 * np.eye(n)
    om = O.LinearModel(A, Bm, Q, R)
    cost0 = om.costfun(x, u)
    rep = lambda a: np.tile(a, (N, 1, 1))
    prev0 = O.GaussianPolicy(N, n, m, p.K.copy(), np.zeros((N, m)), Sig.copy(), Sigi.copy())
    eta = np.array([1e-8, 1.3, 1e16])
    dg, pg, Vxg, Vxxg, dVg = O.back_pass_gps(x @ Q.T, u @ R.T, rep(Q), rep(np.zeros((n, m))), rep(R), rep(A), rep(Bm), None, x, u, (O.grad_kl(prev0), eta))
    xn, un, cn = O.forward_pass(pg, x[0], u, x, 1, om.f, om.costfun, None)
    kl = O.kl_div_wiki(xn, x, O.forward_covariance(A, R1, pg), pg, prev0)
    prev = O.GaussianPolicy(N, n, m, p.K.copy(), u.copy(), Sig.copy(), Sigi.copy())
    r = O.iLQGkl(om.f, om.costfun, lambda xx, uu: om.df(xx, uu, time_varying=True), x, prev, A, R1, kl_step=1.5, cost=cost0)
    tr = r[6]
    np.savez_compressed(os.path.join(HERE, "ilqgkl_n32_m8.npz"), A=A, Bm=Bm, Q=Q, R=R, x=x, u=u, K_prev=p.K, Sig_prev=Sig, Sigi_prev=Sigi, R1=R1,
                        cost0=np.sum(cost0), eta=eta, gps_K=pg.K, gps_k=pg.k, gps_Sigma=pg.Sigma, gps_Sigmai=pg.Sigmai, gps_Vx=Vxg, gps_dV=dVg,
                        gps_diverge=dg, fwd_x=xn, fwd_u=un, kl_t=kl, kl_step=1.5, sol_x=r[0], sol_u=r[1], sol_K=r[2].K, sol_cost=np.sum(r[5]),
                        sol_iters=tr["iters"], sol_satisfied=tr["satisfied"], sol_etabracket=tr["etabracket"],
                        sol_divergence=tr["divergence"][-1][1])


if __name__ == "__main__":
    back_pass_case("back_pass_n10_m2_chol", 10, 2, 40, 1, None, 1.0, 101)
    back_pass_case("back_pass_n32_m8_reg2", 32, 8, 16, 2, None, 0.5, 102)
    back_pass_case("back_pass_n4_m1_lims", 4, 1, 50, 2, 0.05, 1e-3, 103)
    back_pass_case("back_pass_n6_m3_lims", 6, 3, 30, 1, 0.05, 1e-3, 104)
    boxqp_case()
    ilqg_case()
    ilqgkl_case()
    print(sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
