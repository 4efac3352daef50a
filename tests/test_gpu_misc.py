"""GPU parity for ddp_boxqp_f64 (bit-exact), ddp_forward_pass_f64, ddp_back_pass_gps_f64 and
ddp_kl_div_f64 against the CPU oracle."""
import numpy as np
import pytest

from helpers import make_batch_lq, make_lq, relerr, rollout
from oracle import ddp_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-8


@pytest.mark.parametrize("m", [1, 2, 3, 5, 8, 16])
def test_boxqp_bit_exact(ddp, m):
    """x, result, free set, nfactor and the subspace factor are bit-identical to the oracle
    (fixed summation order, no FMA) on exactly symmetric H, as demoQP builds (boxQP.jl:193)."""
    rng = np.random.default_rng(100 + m)
    B = 64
    Hs, gs, los, ups, x0s = [], [], [], [], []
    for b in range(B):
        G = rng.standard_normal((m, m))
        H = G @ G.T + 0.05 * np.eye(m)
        H = (H + H.T) / 2
        Hs.append(H); gs.append(3 * rng.standard_normal(m))
        w = 0.1 + rng.random(m)
        los.append(-w); ups.append(w * rng.random(m) + 0.05); x0s.append(rng.standard_normal(m))
    Hs, gs, los, ups, x0s = map(np.array, (Hs, gs, los, ups, x0s))
    x, res, Hf, free, nf = ddp.boxQP(Hs, gs, los, ups, x0s)
    codes = set()
    for b in range(B):
        x0_, r0, Hf0, free0, nf0 = O.boxQP(Hs[b], gs[b], los[b], ups[b], x0s[b])
        assert res[b] == r0 and nf[b] == nf0
        assert np.array_equal(free[b], free0)
        assert np.array_equal(x[b], x0_), (b, x[b] - x0_)
        k = Hf0.shape[0] if r0 != 6 or nf0 > 0 else 0
        assert np.array_equal(Hf[b][:k, :k], Hf0[:k, :k])
        codes.add(int(r0))
    assert codes <= {4, 5, 6} and len(codes) >= 1


def test_boxqp_not_pd_and_unbatched(ddp):
    H = np.array([[1.0, 2.0], [2.0, 1.0]])      # indefinite
    with pytest.raises(ddp.PosDefException):
        ddp.boxQP(H, np.ones(2), -np.ones(2), np.ones(2), np.zeros(2))
    with pytest.raises(O.PosDefException):
        O.boxQP(H, np.ones(2), -np.ones(2), np.ones(2), np.zeros(2))
    H = np.array([[2.0, 0.5], [0.5, 1.0]])
    x, r, Hf, free, nf = ddp.boxQP(H, np.array([1.0, -4.0]), -np.ones(2), np.ones(2), np.zeros(2))
    x0, r0, Hf0, free0, nf0 = O.boxQP(H, np.array([1.0, -4.0]), -np.ones(2), np.ones(2), np.zeros(2))
    assert r == r0 and nf == nf0 and np.array_equal(x, x0) and np.array_equal(free, free0) and np.array_equal(Hf, Hf0)


@pytest.mark.parametrize("n,m,N", [(10, 2, 50), (32, 8, 30), (5, 3, 20)])
@pytest.mark.parametrize("lims", [None, 0.3])
def test_forward_linear(ddp, n, m, N, lims):
    B = 3
    A, Bm, Q, R, x, u = make_batch_lq(7, B, n, m, N)
    cx, cu = x @ Q.T, u @ R.T
    lim = None if lims is None else np.tile(np.array([[-lims, lims]]), (m, 1))
    alphas = np.array([1.0, 0.5, 0.125])
    model = ddp.LinearModel(A[:, None], Bm[:, None], Q, R)
    pols = []
    for b in range(B):
        d0, p0, _, _, _ = O.back_pass(cx[b], cu[b], Q, np.zeros((n, m)), R, A[b], Bm[b], 1.0, 1, None, x[b], u[b])
        pols.append(p0)
    pol = ddp.GaussianPolicy(N, n, m, np.array([p.K for p in pols]), np.array([p.k for p in pols]))
    xn, un, cn, (cxn, cun) = ddp.forward_pass(pol, x[:, 0], u, x, alphas, model.f, model.costfun, lim, want_derivs=True,
                                              force_generic=True)
    xn2, un2, ct = ddp.forward_pass(pol, x[:, 0], u, x, alphas, model.f, model.costfun, lim, per_step_cost=True, force_generic=True)
    for b in range(B):
        om = O.LinearModel(A[b], Bm[b], Q, R, per_step_cost=True)
        x0_, u0_, c0_ = O.forward_pass(pols[b], x[b, 0], u[b], x[b], alphas[b], om.f, om.costfun, lim)
        assert relerr(xn[b], x0_) < TOL and relerr(un[b], u0_) < TOL
        assert abs(cn[b] - np.sum(c0_)) < TOL * abs(np.sum(c0_))
        assert relerr(ct[b], c0_) < TOL
        assert relerr(cxn[b], x0_ @ Q.T) < TOL and relerr(cun[b], u0_ @ R.T) < TOL
        if lim is not None:
            assert np.all(un[b] <= lims) and np.all(un[b] >= -lims)
            assert np.array_equal(un[b] == lims, u0_ == lims) and np.array_equal(un[b] == -lims, u0_ == -lims)


def test_forward_empty_policy_and_pendcart(ddp):
    """initial rollout (iLQG.jl:185): empty policy, αi*u; pendcart Euler dynamics + T+1 cost entries."""
    N = 80
    rng = np.random.default_rng(3)
    B = 4
    u = 2.0 * rng.standard_normal((B, N, 1))
    x0 = np.stack([np.array([np.pi - 0.6 + 0.2 * rng.uniform(-1, 1), 0, 0, 0]) for _ in range(B)])
    pm = ddp.PendcartModel()
    lims = np.array([[-5.0, 5.0]])
    xn, un, ct = ddp.forward_pass(ddp.GaussianPolicy.empty(), x0, u, None, 1.0, pm.f, pm.costfun, lims, u_scale=0.5,
                                  per_step_cost=True, force_generic=True)
    om = O.PendcartModel()
    for b in range(B):
        x0_, u0_, c0_ = O.forward_pass(O.GaussianPolicy.empty(), x0[b], 0.5 * u[b], None, 1, om.f, om.costfun, lims)
        assert relerr(xn[b], x0_) < TOL and relerr(un[b], u0_) < TOL
        assert ct.shape[1] == N + 1 and relerr(ct[b], c0_) < TOL


def test_forward_rejects_host_callbacks(ddp):
    with pytest.raises(TypeError):
        ddp.forward_pass(ddp.GaussianPolicy.empty(), np.zeros(4), np.zeros((5, 1)), None, 1.0, lambda x, u, i: x,
                         lambda x, u: 0.0, None)


def _prev_policy(n, m, N, seed):
    A, Bm, Q, R, x, u = make_batch_lq(seed, 1, n, m, N)
    A, Bm, x, u = A[0], Bm[0], x[0], u[0]
    cx, cu = x @ Q.T, u @ R.T
    d, p, _, _, _ = O.back_pass(cx, cu, Q, np.zeros((n, m)), R, A, Bm, 1.0, 1, None, x, u)
    Sigi = p.Sigmai.copy()
    prev = O.GaussianPolicy(N, n, m, p.K.copy(), 0.02 * np.random.default_rng(seed).standard_normal((N, m)),
                            np.array([np.linalg.inv(s) for s in Sigi]), Sigi)
    return A, Bm, Q, R, x, u, cx, cu, prev


@pytest.mark.parametrize("n,m,N,lims", [(10, 2, 30, None), (6, 2, 25, 0.2), (32, 8, 12, None), (4, 1, 30, 0.1)])
def test_back_pass_gps(ddp, n, m, N, lims):
    A, Bm, Q, R, x, u, cx, cu, prev = _prev_policy(n, m, N, 11)
    lim = None if lims is None else np.tile(np.array([[-lims, lims]]), (m, 1))
    eta = np.array([1e-8, 0.7, 1e16])
    rep = lambda a: np.tile(a, (N, 1, 1))
    terms = (O.grad_kl(prev), eta)
    d0, p0, Vx0, Vxx0, dV0 = O.back_pass_gps(cx, cu, rep(Q), rep(np.zeros((n, m))), rep(R), rep(A), rep(Bm), lim, x, u, terms)
    gp = ddp.GaussianPolicy(N, n, m, prev.K, prev.k, prev.Sigma, prev.Sigmai)
    d1, p1, Vx1, Vxx1, dV1 = ddp.back_pass_gps(cx, cu, rep(Q), rep(np.zeros((n, m))), rep(R), rep(A), rep(Bm), lim, x, u,
                                               (gp, eta), force_generic=True)
    assert d0 == d1 == 0
    for a, b in ((p1.K, p0.K), (p1.k, p0.k), (Vx1, Vx0), (Vxx1, Vxx0), (dV1, dV0), (p1.Sigmai, p0.Sigmai), (p1.Sigma, p0.Sigma)):
        assert relerr(a, b) < TOL


@pytest.mark.parametrize("n,m,N,dense_r1", [(8, 2, 30, False), (32, 8, 20, False), (32, 8, 9, True)])
def test_kl_div(ddp, n, m, N, dense_r1):
    """forward_covariance + kl_div_wiki (forward_pass.jl:37-56, klutils.jl:70-100): generic kernel and, for n=32, m=8,
    the FP64 tensor-tile kernel (kl_tile.cu), per-step divergence vs the oracle."""
    A, Bm, Q, R, x, u, cx, cu, prev = _prev_policy(n, m, N, 21)
    rep = lambda a: np.tile(a, (N, 1, 1))
    d0, pnew, _, _, _ = O.back_pass_gps(cx, cu, rep(Q), rep(np.zeros((n, m))), rep(R), rep(A), rep(Bm), None, x, u,
                                        (O.grad_kl(prev), np.array([1e-8, 2.0, 1e16])))
    om = O.LinearModel(A, Bm, Q, R)
    xnew, unew, _ = O.forward_pass(pnew, x[0], u, x, 1, om.f, om.costfun, None)
    R1 = 1e-4 * np.eye(n)
    if dense_r1:
        Wn = np.random.default_rng(3).standard_normal((n, n))
        R1 = R1 + 1e-4 * (Wn @ Wn.T) / n
    sig = O.forward_covariance(A, R1, pnew)
    kl0 = O.kl_div_wiki(xnew, x, sig, pnew, prev)
    gpn = ddp.GaussianPolicy(N, n, m, pnew.K, pnew.k, pnew.Sigma, pnew.Sigmai)
    gpp = ddp.GaussianPolicy(N, n, m, prev.K, prev.k, prev.Sigma, prev.Sigmai)
    klt, klm = ddp.kl_div_wiki(xnew, x, A, R1, gpn, gpp)
    assert np.all(kl0 > 0)
    assert relerr(klt, kl0) < TOL and abs(klm - kl0.mean()) < TOL * kl0.mean()


# ---- specialised forward kernels (dispatch by shape; compare with the oracle and the generic kernel)

@pytest.mark.parametrize("lims,goal", [(None, False), (0.3, False), (None, True)])
def test_forward_fast_lin32x8(ddp, lims, goal):
    B, n, m, N = 7, 32, 8, 41
    A, Bm, Q, R, x, u = make_batch_lq(17, B, n, m, N)
    rng = np.random.default_rng(5)
    Qf = Q + 0.001 * (lambda M: M @ M.T)(rng.standard_normal((n, n)))      # dense Q
    Rf = R + 0.0005 * (lambda M: M @ M.T)(rng.standard_normal((m, m)))
    cx, cu = x @ Qf.T, u @ Rf.T
    lim = None if lims is None else np.tile(np.array([[-lims, lims]]), (m, 1))
    alphas = np.linspace(1.0, 0.1, B)
    model = ddp.LinearModel(A[:, None], Bm[:, None], Qf, Rf)
    if goal:
        model.goal = 0.1 * rng.standard_normal(n)
    pols = [O.back_pass(cx[b], cu[b], Qf, np.zeros((n, m)), Rf, A[b], Bm[b], 1.0, 1, None, x[b], u[b])[1] for b in range(B)]
    pol = ddp.GaussianPolicy(N, n, m, np.array([p.K for p in pols]), np.array([p.k for p in pols]))
    xn, un, ct, (cxn, cun) = ddp.forward_pass(pol, x[:, 0], u, x, alphas, model.f, model.costfun, lim, per_step_cost=True, want_derivs=True)
    xg, ug, cg = ddp.forward_pass(pol, x[:, 0], u, x, alphas, model.f, model.costfun, lim, force_generic=True)
    gl = model.goal if goal else np.zeros(n)
    for b in range(B):
        om = O.LinearModel(A[b], Bm[b], Qf, Rf, per_step_cost=True)
        costf = (lambda xx, uu: 0.5 * np.sum((xx - gl) * ((xx - gl) @ Qf.T), axis=1) + 0.5 * np.sum(uu * (uu @ Rf.T), axis=1))
        x0_, u0_, c0_ = O.forward_pass(pols[b], x[b, 0], u[b], x[b], alphas[b], om.f, costf, lim)
        assert relerr(xn[b], x0_) < TOL and relerr(un[b], u0_) < TOL and relerr(ct[b], c0_) < TOL
        assert relerr(cxn[b], (x0_ - gl) @ Qf.T) < TOL and relerr(cun[b], u0_ @ Rf.T) < TOL
        assert relerr(xg[b], x0_) < TOL and abs(cg[b] - c0_.sum()) < TOL * abs(c0_.sum())


@pytest.mark.parametrize("N", [120, 30, 31, 2])
def test_forward_fast_pendcart(ddp, N):
    """N = 120: whole 4-step blocks of the staged kernel; 30: ragged last block; 31: odd horizon (rows are not 16-byte
    aligned => the unstaged kernel); 2: shorter than one block.  B = 37 leaves a ragged warp."""
    B = 37
    rng = np.random.default_rng(8)
    x0 = np.stack([np.array([np.pi - 0.6 + 0.2 * rng.uniform(-1, 1), 0, 0, 0]) for _ in range(B)])
    u = rng.standard_normal((B, N, 1))
    pm, om = ddp.PendcartModel(), O.PendcartModel()
    lims = np.array([[-5.0, 5.0]])
    # a rollout to linearise around, then one oracle back pass per trajectory for a policy
    xr, ur, _ = ddp.forward_pass(ddp.GaussianPolicy.empty(), x0, u, None, 1.0, pm.f, pm.costfun, lims)
    pols = []
    for b in range(B):
        xo, uo, _ = O.forward_pass(O.GaussianPolicy.empty(), x0[b], u[b], None, 1, om.f, om.costfun, lims)
        assert relerr(xr[b], xo) < TOL
        fx, fu, _, _, _, cx, cu, cxx, cxu, cuu = om.df(xo, uo)
        d, p, _, _, _ = O.back_pass(cx, cu, cxx, cxu, cuu, fx, fu, 1.0, 2, lims, xo, uo)
        assert d == 0
        pols.append((p, xo, uo))
    pol = ddp.GaussianPolicy(N, 4, 1, np.array([p[0].K for p in pols]), np.array([p[0].k for p in pols]))
    xs, us = np.array([p[1] for p in pols]), np.array([p[2] for p in pols])
    xn, un, ct, (cxn, cun) = ddp.forward_pass(pol, x0, us, xs, 0.7, pm.f, pm.costfun, lims, per_step_cost=True, want_derivs=True)
    xg, ug, cg = ddp.forward_pass(pol, x0, us, xs, 0.7, pm.f, pm.costfun, lims, per_step_cost=True, force_generic=True)
    for b in range(B):
        x0_, u0_, c0_ = O.forward_pass(pols[b][0], x0[b], us[b], xs[b], 0.7, om.f, om.costfun, lims)
        assert relerr(xn[b], x0_) < TOL and relerr(un[b], u0_) < TOL and relerr(ct[b], c0_) < TOL
        assert relerr(xg[b], x0_) < TOL and relerr(cg[b], c0_) < TOL
        dcx = om.df(x0_, u0_)
        assert relerr(cxn[b], dcx[5]) < TOL and relerr(cun[b], dcx[6]) < TOL       # fused df outputs cx = Q(x - goal), cu = R u


@pytest.mark.parametrize("tv", [False, True])
@pytest.mark.parametrize("with_kprev", [True, False])
def test_back_pass_gps_tile32x8(ddp, tv, with_kprev):
    """KL-augmented sweep on the n=32, m=8 DMMA kernel (LTI and LTV), vs the oracle and the generic kernel."""
    n, m, N = 32, 8, 14
    A, Bm, Q, R, x, u, cx, cu, prev = _prev_policy(n, m, N, 13)
    if not with_kprev:
        prev.k = np.zeros_like(prev.k)                       # what iLQGkl.jl:52 does
    eta = np.array([1e-8, 1.7, 1e16])
    rep = (lambda a: np.tile(a, (N, 1, 1))) if tv else (lambda a: a)
    repo = lambda a: np.tile(a, (N, 1, 1))
    d0, p0, Vx0, Vxx0, dV0 = O.back_pass_gps(cx, cu, repo(Q), repo(np.zeros((n, m))), repo(R), repo(A), repo(Bm), None, x, u,
                                             (O.grad_kl(prev), eta))
    gp = ddp.GaussianPolicy(N, n, m, prev.K, prev.k, prev.Sigma, prev.Sigmai)
    for generic in (False, True):
        d1, p1, Vx1, Vxx1, dV1 = ddp.back_pass_gps(cx, cu, rep(Q), rep(np.zeros((n, m))), rep(R), rep(A), rep(Bm), None, x, u,
                                                   (gp, eta), force_generic=generic)
        assert d0 == d1 == 0
        for a, b in ((p1.K, p0.K), (p1.k, p0.k), (Vx1, Vx0), (Vxx1, Vxx0), (dV1, dV0), (p1.Sigmai, p0.Sigmai), (p1.Sigma, p0.Sigma)):
            assert relerr(a, b) < TOL


def test_forward_pendcart_staged_equals_unstaged(ddp, monkeypatch):
    """fwd_pend_staged_kernel (time-blocked cp.async ring) and fwd_pend_kernel (diagnostic switch DDP_PEND_NOSTAGE) run
    the same per-trajectory arithmetic: identical bits, ragged batch and ragged last block."""
    N, B = 38, 75
    rng = np.random.default_rng(12)
    x0 = np.stack([np.array([np.pi - 0.6 + 0.2 * rng.uniform(-1, 1), 0, 0, 0]) for _ in range(B)])
    u = rng.standard_normal((B, N, 1))
    pm = ddp.PendcartModel()
    lims = np.array([[-5.0, 5.0]])
    xr, ur, _ = ddp.forward_pass(ddp.GaussianPolicy.empty(), x0, u, None, 1.0, pm.f, pm.costfun, lims)
    pol = ddp.GaussianPolicy(N, 4, 1, 0.3 * rng.standard_normal((B, N, 1, 4)), 0.2 * rng.standard_normal((B, N, 1)))
    outs = []
    for flag in (None, "1"):
        if flag is None:
            monkeypatch.delenv("DDP_PEND_NOSTAGE", raising=False)
        else:
            monkeypatch.setenv("DDP_PEND_NOSTAGE", flag)
        outs.append(ddp.forward_pass(pol, x0, ur, xr, 0.5, pm.f, pm.costfun, lims, per_step_cost=True, want_derivs=True))
    monkeypatch.delenv("DDP_PEND_NOSTAGE", raising=False)
    a, b = outs
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[3][0], b[3][0]) and np.array_equal(a[3][1], b[3][1])


def test_library_communicator_single_rank(ddp):
    """ddp_comm_unique_id / ddp_comm_init / ddp_comm_allreduce_stats_f64 (NCCL resolved at run time) on a world of one rank:
    the all-reduce is the identity and runs on the handle's stream."""
    eng = ddp.Engine(4, 1, 8, 3)
    uid = eng.comm_unique_id()
    assert len(uid) == 128 and any(uid)
    eng.comm_init(1, 0, uid)
    st = eng.upload(np.arange(8, dtype=np.float64) + 0.5)
    n0 = eng.launch_count
    eng.allreduce_stats(st.ptr)
    eng.synchronize()
    assert np.array_equal(st.numpy(), np.arange(8) + 0.5) and eng.launch_count == n0 + 1
    eng.close()
