"""GPU parity for the device-resident iLQG driver (ddp_ilqg_solve_f64) and the host-buffer
iteration (ddp_ilqg_iter_host_f64) against the CPU oracle's iLQG state machine.
Integer outcomes (status, iter, accepted_iter) exact; x, u, cost within the stated tolerance."""
import numpy as np
import pytest

from helpers import make_batch_lq, make_lq, relerr
from oracle import ddp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,N,generic", [(10, 2, 120, True), (32, 8, 60, False), (32, 8, 40, True)])
def test_ilqg_lq_batch(ddp, n, m, N, generic):
    B = 5
    rng = np.random.default_rng(42)
    As, Bs, x0s, u0s = [], [], [], []
    for b in range(B):
        A, Bm, Q, R = make_lq(rng, n, m)
        As.append(A); Bs.append(Bm); x0s.append(np.ones(n) * (1 + 0.1 * b)); u0s.append(0.1 * rng.standard_normal((N, m)))
    As, Bs, x0s, u0s = map(np.array, (As, Bs, x0s, u0s))
    model = ddp.LinearModel(As[:, None], Bs[:, None], Q, R)
    x, u, pol, Vx, Vxx, cost, tr = ddp.iLQG(model.f, model.costfun, model.df, x0s, u0s, force_generic=generic)
    for b in range(B):
        om = O.LinearModel(As[b], Bs[b], Q, R)
        x0_, u0_, p0, Vx0, Vxx0, c0, t0 = O.iLQG(om.f, om.costfun, om.df, x0s[b], u0s[b])
        assert tr["status"][b] == t0["status"] and tr["iter"][b] == t0["iters"]
        assert relerr(x[b], x0_) < 1e-7 and relerr(u[b], u0_) < 1e-7
        assert abs(cost[b] - np.sum(c0)) < 1e-8 * abs(np.sum(c0))
        assert abs(tr["lam"][b] - t0["lam_final"]) <= 1e-12 * t0["lam_final"]
        assert relerr(pol.K[b], p0.K) < 1e-7 and relerr(pol.k[b], p0.k) < 1e-6 and relerr(Vx[b], Vx0) < 1e-6
        assert relerr(Vxx[b], Vxx0[0]) < 1e-7


@pytest.mark.parametrize("lims", [None, 0.35])
def test_ilqg_multi_alpha_linesearch(ddp, monkeypatch, lims):
    """Multi-alpha line search (all remaining step sizes in one pass over K after a rejected alpha[0]) finds exactly the
    step the serial backtracking of iLQG.jl:267-281 finds: integer outcomes and trajectories identical to the serial
    device path, and equal to the oracle's.  alpha[0] = 10^0.6 overshoots a quadratic, so it is always rejected."""
    n, m, N, B = 32, 8, 30, 6
    rng = np.random.default_rng(5)
    As, Bs, x0s, u0s = [], [], [], []
    for b in range(B):
        A, Bm, Q, R = make_lq(rng, n, m, h=0.05)
        As.append(A); Bs.append(Bm); x0s.append(np.ones(n) * (1 + 0.2 * b)); u0s.append(0.1 * rng.standard_normal((N, m)))
    As, Bs, x0s, u0s = map(np.array, (As, Bs, x0s, u0s))
    alpha = 10.0 ** np.linspace(0.6, -3, 8)
    lim = None if lims is None else np.tile(np.array([[-lims, lims]]), (m, 1))
    kw = dict(alpha=alpha, max_iter=8, lims=lim)
    model = ddp.LinearModel(As[:, None], Bs[:, None], Q, R)
    monkeypatch.delenv("DDP_NO_MULTI_ALPHA", raising=False)
    rm = ddp.iLQG(model.f, model.costfun, model.df, x0s, u0s, **kw)
    monkeypatch.setenv("DDP_NO_MULTI_ALPHA", "1")
    rs = ddp.iLQG(model.f, model.costfun, model.df, x0s, u0s, **kw)
    monkeypatch.delenv("DDP_NO_MULTI_ALPHA", raising=False)
    tm, ts = rm[6], rs[6]
    for key in ("status", "iter", "accepted_iter", "last_alpha", "lam"):
        assert np.array_equal(tm[key], ts[key]), key
    assert np.array_equal(rm[0], rs[0]) and np.array_equal(rm[1], rs[1]) and np.array_equal(rm[5], rs[5])
    assert np.all(tm["last_alpha"] < alpha[0])                   # the multi-alpha branch ran for every trajectory
    if lims is None:
        for b in range(B):
            om = O.LinearModel(As[b], Bs[b], Q, R)
            x0_, u0_, p0, Vx0, Vxx0, c0, t0 = O.iLQG(om.f, om.costfun, om.df, x0s[b], u0s[b], alpha=alpha, max_iter=8)
            assert tm["status"][b] == t0["status"] and tm["iter"][b] == t0["iters"]
            assert relerr(rm[0][b], x0_) < 1e-7 and relerr(rm[1][b], u0_) < 1e-7


@pytest.mark.parametrize("n,m,lims,dense_q", [(32, 8, None, False), (32, 8, 0.2, True), (6, 2, None, False)])
def test_forward_costs_multi(ddp, n, m, lims, dense_q, monkeypatch):
    """ddp_forward_costs_multi_f64: the cost of every step size from one pass over K equals forward_pass's -- bit for bit on the FMA
    kernels (same per-lane arithmetic), to rounding (1e-13 relative) on the FP64 tensor-tile kernel that takes the headline shape
    with a diagonal Q (its sums run in tile order)."""
    monkeypatch.setenv("DDP_MULTI_NO_TILE", "1")
    B, N = 7, 33 if n == 6 else 26
    A, Bm, Q, R, x, u = make_batch_lq(3, B, n, m, N, h=0.05)
    if dense_q:
        rng = np.random.default_rng(1)
        W = rng.standard_normal((n, n)); Q = 0.01 * (W @ W.T) / n + Q
    pol = None
    Ks, ks = [], []
    for b in range(B):
        d, p, _, _, _ = O.back_pass(x[b] @ Q.T, u[b] @ R.T, Q, np.zeros((n, m)), R, A[b], Bm[b], 1.0, 1, None, x[b], u[b])
        Ks.append(p.K); ks.append(p.k)
    pol = ddp.GaussianPolicy(N, n, m, np.array(Ks), np.array(ks))
    model = ddp.LinearModel(A[:, None], Bm[:, None], Q, R)
    lim = None if lims is None else np.tile(np.array([[-lims, lims]]), (m, 1))
    alphas = 10.0 ** np.linspace(0.3, -3, 12)                    # 12 > 10: two launches of the multi kernel
    costs = ddp.forward_costs(pol, x[:, 0], u, x, alphas, model.f, model.costfun, lim)
    assert costs.shape == (len(alphas), B)
    monkeypatch.delenv("DDP_MULTI_NO_TILE", raising=False)
    costs_tile = ddp.forward_costs(pol, x[:, 0], u, x, alphas, model.f, model.costfun, lim)     # tile kernel where eligible, else the same FMA kernel
    for i, a in enumerate(alphas):
        _, _, c = ddp.forward_pass(pol, x[:, 0], u, x, float(a), model.f, model.costfun, lim)
        assert np.array_equal(costs[i], c), (i, np.max(np.abs(costs[i] - c)))
        assert np.max(np.abs(costs_tile[i] - c) / np.abs(c)) < 1e-13, (i, np.max(np.abs(costs_tile[i] - c) / np.abs(c)))


@pytest.mark.parametrize("lims", [None, 0.15])
def test_forward_costs_multi_tile_vs_oracle(ddp, lims):
    """The tensor-tile multi-alpha rollout (fwd_lin32x8_multi_tile_kernel: 16 step sizes as the columns of one state matrix) against
    the oracle's forward_pass, step size by step size, incl. clamped controls; 3, 8, 11 and 16 step sizes (one and two alpha tiles)."""
    B, n, m, N = 5, 32, 8, 40
    A, Bm, Q, R, x, u = make_batch_lq(9, B, n, m, N, h=0.05)
    Ks, ks = [], []
    for b in range(B):
        d, p, _, _, _ = O.back_pass(x[b] @ Q.T, u[b] @ R.T, Q, np.zeros((n, m)), R, A[b], Bm[b], 1.0, 1, None, x[b], u[b])
        Ks.append(p.K); ks.append(p.k)
    pol = ddp.GaussianPolicy(N, n, m, np.array(Ks), np.array(ks))
    model = ddp.LinearModel(A[:, None], Bm[:, None], Q, R)
    lim = None if lims is None else np.tile(np.array([[-lims, lims]]), (m, 1))
    for na in (3, 8, 11, 16):
        alphas = 10.0 ** np.linspace(0.3, -3, na)
        costs = ddp.forward_costs(pol, x[:, 0], u, x, alphas, model.f, model.costfun, lim)
        for b in range(B):
            om = O.LinearModel(A[b], Bm[b], Q, R)
            pb = O.GaussianPolicy(N, n, m, pol.K[b], pol.k[b], None, None)
            for i, a in enumerate(alphas):
                _, _, c0 = O.forward_pass(pb, x[b, 0], u[b], x[b], float(a), om.f, om.costfun, lim)
                assert abs(costs[i, b] - c0) <= 1e-12 * abs(c0), (na, b, i, costs[i, b], c0)


def test_forward_costs_multi_pendcart(ddp):
    """Pendulum-on-a-cart multi-alpha rollout (fwd_pend_multi_kernel): costs bit-identical to one rollout per step size."""
    N, B = 46, 41
    rng = np.random.default_rng(2)
    x0 = np.stack([np.array([np.pi - 0.6 + 0.2 * rng.uniform(-1, 1), 0, 0, 0]) for _ in range(B)])
    pm = ddp.PendcartModel()
    lims = np.array([[-5.0, 5.0]])
    xr, ur, _ = ddp.forward_pass(ddp.GaussianPolicy.empty(), x0, rng.standard_normal((B, N, 1)), None, 1.0, pm.f, pm.costfun, lims)
    pol = ddp.GaussianPolicy(N, 4, 1, 0.3 * rng.standard_normal((B, N, 1, 4)), 0.2 * rng.standard_normal((B, N, 1)))
    alphas = 10.0 ** np.linspace(0.2, -3, 11)                   # 11 > 8: two launches
    costs = ddp.forward_costs(pol, x0, ur, xr, alphas, pm.f, pm.costfun, lims)
    for i, a in enumerate(alphas):
        _, _, c = ddp.forward_pass(pol, x0, ur, xr, float(a), pm.f, pm.costfun, lims)
        assert np.array_equal(costs[i], c), (i, np.max(np.abs(costs[i] - c)))


def test_ilqg_pendcart_multi_alpha_equals_serial(ddp, monkeypatch):
    """demo_pendcart's settings: the multi-alpha line search and the serial one take identical decisions and trajectories."""
    N, B = 100, 9
    rng = np.random.default_rng(6)
    x0 = np.stack([np.array([np.pi - 0.6 + 0.3 * rng.uniform(-1, 1), 0.1 * rng.standard_normal(), 0, 0]) for _ in range(B)])
    u0 = np.zeros((B, N, 1))
    kw = dict(lims=np.array([[-5.0, 5.0]]), regType=2, alpha=10.0 ** np.linspace(0.2, -3, 6), lammax=1e15, tol_fun=1e-8, tol_grad=1e-8, max_iter=8)
    pm = ddp.PendcartModel()
    monkeypatch.delenv("DDP_NO_MULTI_ALPHA", raising=False)
    rm = ddp.iLQG(pm.f, pm.costfun, pm.df, x0, u0, **kw)
    monkeypatch.setenv("DDP_NO_MULTI_ALPHA", "1")
    rs = ddp.iLQG(pm.f, pm.costfun, pm.df, x0, u0, **kw)
    monkeypatch.delenv("DDP_NO_MULTI_ALPHA", raising=False)
    for key in ("status", "iter", "accepted_iter", "last_alpha", "lam"):
        assert np.array_equal(rm[6][key], rs[6][key]), key
    assert np.array_equal(rm[0], rs[0]) and np.array_equal(rm[1], rs[1]) and np.array_equal(rm[5], rs[5])
    assert np.any(rm[6]["last_alpha"] < kw["alpha"][0])          # some line searches went past alpha[0]


def test_ilqg_thresholds_of_test_readme(ddp):
    """test/test_readme.jl:82-84 on fresh instances of its problem distribution (n=10, m=2, T=1000)."""
    rng = np.random.default_rng(0)
    B, n, m, N = 10, 10, 2, 1000
    As, Bs, u0s = [], [], []
    for b in range(B):
        A, Bm, Q, R = make_lq(rng, n, m)
        As.append(A); Bs.append(Bm); u0s.append(0.1 * rng.standard_normal((N, m)))
    model = ddp.LinearModel(np.array(As)[:, None], np.array(Bs)[:, None], Q, R)
    x, u, pol, Vx, Vxx, cost, tr = ddp.iLQG(model.f, model.costfun, model.df, np.ones((B, n)), np.array(u0s))
    assert np.all(tr["status"] == 0)
    assert cost.max() < 25 and cost.mean() < 10 and cost.min() < 5


def test_ilqg_pendcart_lims(ddp):
    """demo_pendcart's settings (system_pendcart.jl:197-206) on a short horizon, a few iterations."""
    N, B = 120, 3
    x0 = np.array([[np.pi - 0.6, 0, 0, 0], [np.pi - 0.5, 0, 0, 0], [np.pi - 0.7, 0.1, 0, 0]])
    u0 = np.zeros((B, N, 1))
    lims = np.array([[-5.0, 5.0]])
    kw = dict(lims=lims, regType=2, alpha=10.0 ** np.linspace(0.2, -3, 6), lammax=1e15, tol_fun=1e-8, tol_grad=1e-8, max_iter=6)
    pm = ddp.PendcartModel()
    x, u, pol, Vx, Vxx, cost, tr = ddp.iLQG(pm.f, pm.costfun, pm.df, x0, u0, **kw)
    for b in range(B):
        om = O.PendcartModel()
        x0_, u0_, p0, Vx0, Vxx0, c0, t0 = O.iLQG(om.f, om.costfun, om.df, x0[b], u0[b], **kw)
        assert tr["status"][b] == t0["status"] and tr["iter"][b] == t0["iters"]
        assert relerr(x[b], x0_) < 1e-6 and relerr(u[b], u0_) < 1e-6
        assert abs(cost[b] - np.sum(c0)) < 1e-7 * abs(np.sum(c0))


def test_ilqg_unbatched_and_errors(ddp):
    rng = np.random.default_rng(1)
    A, Bm, Q, R = make_lq(rng, 6, 2)
    model = ddp.LinearModel(A, Bm, Q, R)
    u0 = 0.1 * rng.standard_normal((50, 2))
    x, u, pol, Vx, Vxx, cost, tr = ddp.iLQG(model.f, model.costfun, model.df, np.ones(6), u0)
    om = O.LinearModel(A, Bm, Q, R)
    x0_, u0_, p0, Vx0, Vxx0, c0, t0 = O.iLQG(om.f, om.costfun, om.df, np.ones(6), u0)
    assert tr["status"] == t0["status"] and tr["iter"] == t0["iters"] and relerr(x, x0_) < 1e-7
    # an unstable system with huge controls diverges for every alpha -> reference returns nothing
    Au = 40.0 * np.eye(6)
    mu = ddp.LinearModel(Au, Bm, Q, R)
    assert ddp.iLQG(mu.f, mu.costfun, mu.df, 1e6 * np.ones(6), 1e6 * np.ones((50, 2))) is None
    with pytest.raises(TypeError):
        ddp.iLQG(lambda x, u, i: x, lambda x, u: 0.0, lambda x, u: None, np.ones(6), u0)


def test_iter_host_matches_separate_calls(ddp):
    B, n, m, N = 50, 32, 8, 32
    A, Bm, Q, R, x, u = make_batch_lq(9, B, n, m, N)
    cx, cu = x @ Q.T, u @ R.T
    eng = ddp.Engine(n, m, N, B)
    it = ddp.HostIteration(eng, Q, R, reg_type=1, alpha=0.5, chunk=16)          # 4 ragged chunks
    it.bufs["fx"][:] = np.swapaxes(A, -1, -2); it.bufs["fu"][:] = np.swapaxes(Bm, -1, -2)
    it.bufs["cx"][:] = cx; it.bufs["cu"][:] = cu; it.bufs["x"][:] = x; it.bufs["u"][:] = u
    it.bufs["lam"][:] = 1.0 + 0.01 * np.arange(B)
    h2d, d2h = it.run()
    assert h2d >= B * 8 * (n * n + n * m + 2 * N * n + 2 * N * m + 1) and d2h >= B * 8 * (N * n + N * m + 3)
    dv, pol, Vx, Vxx, dV = ddp.back_pass(cx, cu, Q, np.zeros((n, m)), R, A[:, None], Bm[:, None], it.bufs["lam"].copy(), 1, None, x, u)
    model = ddp.LinearModel(A[:, None], Bm[:, None], Q, R)
    xn, un, cn = ddp.forward_pass(pol, x[:, 0], u, x, 0.5, model.f, model.costfun, None)
    assert np.array_equal(it.bufs["diverge"], dv)
    assert np.array_equal(it.bufs["xnew"], xn) and np.array_equal(it.bufs["unew"], un)
    assert np.array_equal(it.bufs["cost"], cn) and np.array_equal(it.bufs["dV"], dV)
    it.close()


@pytest.mark.parametrize("kl_step", [1.0, 20.0])
def test_ilqgkl_matches_oracle(ddp, kl_step):
    """iLQGkl single-KL-constraint branch (iLQGkl.jl:93-183): same eta sequence, iteration count and result."""
    from helpers import rollout
    n, m, N = 6, 2, 40
    rng = np.random.default_rng(31)
    A, Bm, Q, R = make_lq(rng, n, m, h=0.1)
    u = 0.1 * rng.standard_normal((N, m))
    x = rollout(A, Bm, np.ones(n), u)
    d, p, _, _, _ = O.back_pass(x @ Q.T, u @ R.T, Q, np.zeros((n, m)), R, A, Bm, 1.0, 1, None, x, u)
    Sigi = p.Sigmai.copy()
    Sig = np.array([np.linalg.inv(s) for s in Sigi])
    R1 = 1e-3 * np.eye(n)
    om = O.LinearModel(A, Bm, Q, R)
    cost0 = om.costfun(x, u)
    prev_o = O.GaussianPolicy(N, n, m, p.K.copy(), u.copy(), Sig.copy(), Sigi.copy())
    ro = O.iLQGkl(om.f, om.costfun, lambda xx, uu: om.df(xx, uu, time_varying=True), x, prev_o, A, R1, kl_step=kl_step, cost=cost0)
    model = ddp.LinearModel(A, Bm, Q, R)
    prev_d = ddp.GaussianPolicy(N, n, m, p.K.copy(), u.copy(), Sig.copy(), Sigi.copy())
    rd = ddp.iLQGkl(model.f, model.costfun, model.df, x, prev_d, A, R1, kl_step=kl_step, cost=cost0)
    to, td = ro[6], rd[6]
    assert td["iters"] == to["iters"] and td["satisfied"] == to["satisfied"]
    assert np.allclose([e for _, e in td["eta"]], [e for _, e in to["eta"]], rtol=1e-9)
    assert np.allclose([e for _, e in td["divergence"]], [e for _, e in to["divergence"]], rtol=1e-7)
    assert relerr(rd[0], ro[0]) < 1e-7 and relerr(rd[1], ro[1]) < 1e-7 and abs(rd[5] - np.sum(ro[5])) < 1e-8 * abs(np.sum(ro[5]))
    assert relerr(rd[2].K, ro[2].K) < 1e-7 and relerr(rd[2].Sigma, ro[2].Sigma) < 1e-7


@pytest.mark.parametrize("n,m", [(6, 2), (32, 8)])
def test_ilqgkl_device_batch_matches_oracle(ddp, n, m):
    """ddp_ilqgkl_solve_f64: the whole iLQGkl loop (iLQGkl.jl:93-183, klutils.jl:110-130) device resident for a batch of
    different problems with different kl_step-to-divergence regimes; every trajectory ends with the oracle's iteration
    count, status and eta bracket (integers exact, eta <= 1e-9 rel) and its trajectories/policy within 1e-7."""
    from helpers import rollout
    B, N = 5, 24
    rng = np.random.default_rng(77)
    xs, us, Ks, Sigs, Sigis, costs, As, Bs = [], [], [], [], [], [], [], []
    Q = R = None
    for b in range(B):
        A, Bm, Q, R = make_lq(rng, n, m, h=0.1)
        u = (0.05 + 0.1 * b) * rng.standard_normal((N, m))
        x = rollout(A, Bm, np.ones(n), u)
        d, p, _, _, _ = O.back_pass(x @ Q.T, u @ R.T, Q, np.zeros((n, m)), R, A, Bm, 1.0, 1, None, x, u)
        assert d == 0
        Sigi = p.Sigmai.copy()
        xs.append(x); us.append(u); Ks.append(p.K.copy()); Sigis.append(Sigi); Sigs.append(np.array([np.linalg.inv(s_) for s_ in Sigi]))
        As.append(A); Bs.append(Bm)
        costs.append(O.LinearModel(A, Bm, Q, R).costfun(x, u))
    R1 = 1e-3 * np.eye(n)
    kl_step = 2.0
    ref = []
    for b in range(B):
        om = O.LinearModel(As[b], Bs[b], Q, R)
        prev = O.GaussianPolicy(N, n, m, Ks[b].copy(), us[b].copy(), Sigs[b].copy(), Sigis[b].copy())
        ref.append(O.iLQGkl(om.f, om.costfun, lambda xx, uu, om=om: om.df(xx, uu, time_varying=True), xs[b], prev, As[b], R1,
                            kl_step=kl_step, cost=costs[b]))
    A4, B4 = np.stack(As)[:, None], np.stack(Bs)[:, None]
    model = ddp.LinearModel(A4, B4, Q, R)
    prev_d = ddp.GaussianPolicy(N, n, m, np.stack(Ks), np.stack(us), np.stack(Sigs), np.stack(Sigis))
    rd = ddp.iLQGkl_device(model.f, model.costfun, model.df, np.stack(xs), prev_d, A4, R1, kl_step=kl_step,
                           cost=np.array([np.sum(c) for c in costs]))
    tr = rd[6]
    iters = [r[6]["iters"] for r in ref]
    assert len(set(iters)) > 1 or B == 1, "the batch should exercise different iteration counts"
    for b in range(B):
        to = ref[b][6]
        assert tr["iter"][b] == to["iters"] and bool(tr["satisfied"][b]) == bool(to["satisfied"])
        assert np.allclose([tr["eta_min"][b], tr["eta"][b], tr["eta_max"][b]], to["etabracket"], rtol=1e-9)
        assert abs(tr["divergence"][b] - to["divergence"][-1][1]) <= 1e-7 * max(1.0, abs(to["divergence"][-1][1]))
        assert relerr(rd[0][b], ref[b][0]) < 1e-7 and relerr(rd[1][b], ref[b][1]) < 1e-7
        assert abs(rd[5][b] - np.sum(ref[b][5])) < 1e-8 * abs(np.sum(ref[b][5]))
        assert relerr(rd[2].K[b], ref[b][2].K) < 1e-7 and relerr(rd[2].Sigma[b], ref[b][2].Sigma) < 1e-7
        assert relerr(rd[2].k[b], ref[b][1]) < 1e-7                      # quirk Q11: k = u on return


def test_iter_host_device_derivs(ddp):
    """cx = cu = NULL: the derivative step runs on the device; results equal the host-derivative path."""
    B, n, m, N = 21, 32, 8, 24
    A, Bm, Q, R, x, u = make_batch_lq(19, B, n, m, N)
    eng = ddp.Engine(n, m, N, B)
    outs = []
    for dd in (False, True):
        it = ddp.HostIteration(eng, Q, R, reg_type=1, alpha=1.0, chunk=8, device_derivs=dd)
        it.bufs["fx"][:] = np.swapaxes(A, -1, -2); it.bufs["fu"][:] = np.swapaxes(Bm, -1, -2)
        it.bufs["cx"][:] = x @ Q.T; it.bufs["cu"][:] = u @ R.T; it.bufs["x"][:] = x; it.bufs["u"][:] = u; it.bufs["lam"][:] = 1.0
        h2d, d2h = it.run()
        outs.append((h2d, it.bufs["xnew"].copy(), it.bufs["unew"].copy(), it.bufs["cost"].copy(), it.bufs["diverge"].copy()))
        it.close()
    assert outs[1][0] < outs[0][0]
    assert relerr(outs[1][1], outs[0][1]) < 1e-12 and relerr(outs[1][2], outs[0][2]) < 1e-12 and relerr(outs[1][3], outs[0][3]) < 1e-12
    assert np.array_equal(outs[1][4], outs[0][4])
