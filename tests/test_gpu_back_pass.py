"""GPU parity: ddp_back_pass_f64 (through the C ABI) vs the CPU oracle on the same seeded inputs.
Tolerance (BASELINE.json north_star): K, k, Vx, Vxx within 1e-8 relative (max-norm per tensor),
`diverge` and boxQP-derived integer outcomes exact."""
import numpy as np
import pytest

from helpers import make_batch_lq, relerr
from oracle import ddp_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _check(ddp, B, n, m, N, regType, lims=None, lam=1.0, ltv=False, tv_cost=False, force_generic=False, seed=0):
    A, Bm, Q, R, x, u = make_batch_lq(seed, B, n, m, N)
    cx = x @ Q.T
    cu = u @ R.T
    cxu = 0.01 * np.random.default_rng(seed + 1).standard_normal((n, m))
    if ltv:
        rng = np.random.default_rng(seed + 2)
        fx = A[:, None] + 1e-3 * rng.standard_normal((B, N, n, n))
        fu = Bm[:, None] + 1e-3 * rng.standard_normal((B, N, n, m))
    else:
        fx, fu = A[:, None], Bm[:, None]
    if tv_cost:
        cxx = np.tile(Q, (N, 1, 1)) * (1 + 0.1 * np.arange(N))[:, None, None]
        cuu = np.tile(R, (N, 1, 1)) * (1 + 0.05 * np.arange(N))[:, None, None]
        cxu_ = np.tile(cxu, (N, 1, 1))
    else:
        cxx, cuu, cxu_ = Q, R, cxu
    lam_b = lam * (1 + 0.1 * np.arange(B))
    dv, pol, Vx, Vxx, dV = ddp.back_pass(cx, cu, cxx, cxu_, cuu, fx, fu, lam_b, regType, lims, x, u, force_generic=force_generic)
    for b in range(B):
        fxb = fx[b] if ltv else fx[b, 0]
        fub = fu[b] if ltv else fu[b, 0]
        d0, p0, Vx0, Vxx0, dV0 = O.back_pass(cx[b], cu[b], cxx, cxu_, cuu, fxb, fub, lam_b[b], regType, lims, x[b], u[b])
        assert dv[b] == d0
        assert relerr(pol.K[b], p0.K) < TOL
        assert relerr(pol.k[b], p0.k) < TOL
        assert relerr(Vx[b], Vx0) < TOL
        assert relerr(Vxx[b], Vxx0) < TOL
        assert relerr(dV[b], dV0) < TOL
        assert relerr(pol.Sigmai[b][d0:], p0.Sigmai[d0:]) < TOL    # Quu, written slices only


@pytest.mark.parametrize("n,m,N", [(10, 2, 40), (4, 1, 30), (7, 3, 25), (32, 8, 20), (64, 16, 6)])
@pytest.mark.parametrize("regType", [1, 2])
def test_generic_cholesky_lti(ddp, n, m, N, regType):
    _check(ddp, 3, n, m, N, regType, force_generic=True)


def test_generic_ltv_tvcost(ddp):
    _check(ddp, 3, 10, 2, 30, 1, ltv=True, tv_cost=True, force_generic=True)
    _check(ddp, 2, 6, 2, 30, 2, ltv=True, tv_cost=False, force_generic=True)


@pytest.mark.parametrize("n,m", [(4, 1), (10, 2), (12, 4)])
def test_generic_boxqp_branch(ddp, n, m):
    lims = np.tile(np.array([[-0.05, 0.05]]), (m, 1))
    _check(ddp, 4, n, m, 30, 1, lims=lims, lam=1e-3, force_generic=True, seed=3)
    _check(ddp, 4, n, m, 30, 2, lims=lims, lam=1e-3, force_generic=True, seed=4)


def test_generic_diverge_non_pd(ddp):
    """forced non-PD cuu => Cholesky fails at the first processed step: diverge == N-1 (1-based)."""
    B, n, m, N = 2, 6, 2, 12
    A, Bm, Q, R, x, u = make_batch_lq(5, B, n, m, N)
    cx, cu = x @ Q.T, u @ R.T
    Rneg = -np.eye(m)
    dv, pol, Vx, Vxx, dV = ddp.back_pass(cx, cu, Q, np.zeros((n, m)), Rneg, A[:, None], Bm[:, None], 0.0, 1, None, x, u, force_generic=True)
    d0, p0, Vx0, Vxx0, dV0 = O.back_pass(cx[0], cu[0], Q, np.zeros((n, m)), Rneg, A[0], Bm[0], 0.0, 1, None, x[0], u[0])
    assert d0 == N - 1 and np.all(dv == N - 1)
    assert np.all(pol.K == 0) and np.all(pol.k == 0)
    assert np.array_equal(Vx[0], Vx0) and np.array_equal(Vxx[0], Vxx0)


# ---- specialised n=32, m=8 DMMA kernel (ddp_kernel_variant == "tile32x8") -----------------------

@pytest.mark.parametrize("regType", [1, 2])
@pytest.mark.parametrize("ltv", [False, True])
def test_tile32x8_vs_oracle(ddp, regType, ltv):
    _check(ddp, 5, 32, 8, 24, regType, ltv=ltv, force_generic=False, seed=20)


def test_tile32x8_tvcost_and_many(ddp):
    _check(ddp, 3, 32, 8, 16, 1, ltv=True, tv_cost=True, force_generic=False, seed=21)
    _check(ddp, 37, 32, 8, 9, 1, force_generic=False, seed=22)        # ragged vs the 4-warp CTA


def test_tile32x8_diverge(ddp):
    B, n, m, N = 3, 32, 8, 10
    A, Bm, Q, R, x, u = make_batch_lq(5, B, n, m, N)
    cx, cu = x @ Q.T, u @ R.T
    # cuu = -0.5 I: lambda = 1 keeps QuuF positive definite, lambda = 0.1 does not (trajectory 1 only)
    Rneg = -0.5 * np.eye(m)
    lam = np.array([1.0, 0.1, 1.0])
    dv, pol, Vx, Vxx, dV = ddp.back_pass(cx, cu, Q, np.zeros((n, m)), Rneg, A[:, None], Bm[:, None], lam, 1, None, x, u)
    for b in range(B):
        d0, p0, Vx0, Vxx0, dV0 = O.back_pass(cx[b], cu[b], Q, np.zeros((n, m)), Rneg, A[b], Bm[b], lam[b], 1, None, x[b], u[b])
        assert dv[b] == d0
        assert relerr(pol.K[b], p0.K) < TOL and relerr(Vx[b], Vx0) < TOL and relerr(Vxx[b], Vxx0) < TOL
    assert dv[1] == N - 1 and dv[0] == 0 and dv[2] == 0
    assert np.all(pol.K[1] == 0) and np.all(Vxx[1][: N - 1] == 0)


def test_tile32x8_matches_generic_kernel(ddp):
    A, Bm, Q, R, x, u = make_batch_lq(30, 6, 32, 8, 40)
    cx, cu = x @ Q.T, u @ R.T
    args = (cx, cu, Q, np.zeros((32, 8)), R, A[:, None], Bm[:, None], 0.5, 1, None, x, u)
    r1 = ddp.back_pass(*args)
    r2 = ddp.back_pass(*args, force_generic=True)
    assert np.array_equal(r1[0], r2[0])
    for a, b in ((r1[1].K, r2[1].K), (r1[1].k, r2[1].k), (r1[2], r2[2]), (r1[3], r2[3]), (r1[4], r2[4])):
        assert relerr(a, b) < 1e-10
    assert np.array_equal(r1[3], np.swapaxes(r1[3], -1, -2))          # Vxx exactly symmetric


# ---- specialised small-system kernel (one thread per trajectory; config 3 is n=4, m=1 with lims) ----

@pytest.mark.parametrize("n,m", [(4, 1), (2, 1), (3, 1), (4, 2)])
@pytest.mark.parametrize("regType", [1, 2])
def test_small_kernel_cholesky_and_qp(ddp, n, m, regType):
    _check(ddp, 70, n, m, 33, regType, ltv=True, seed=40)
    lims = np.tile(np.array([[-0.05, 0.05]]), (m, 1))
    _check(ddp, 70, n, m, 33, regType, lims=lims, lam=1e-3, ltv=True, seed=41)
    _check(ddp, 5, n, m, 20, regType, lims=lims, lam=1e-3, ltv=False, tv_cost=True, seed=42)


def test_small_kernel_pendcart_lims_matches_oracle_and_generic(ddp):
    """config 3 at test size: pendcart ZoH Jacobians, lims = +-5, regType 2 (system_pendcart.jl:197-206)."""
    N, B = 150, 6
    rng = np.random.default_rng(1)
    om = O.PendcartModel()
    lims = np.array([[-5.0, 5.0]])
    xs, us, fxs, fus, cxs, cus = [], [], [], [], [], []
    for b in range(B):
        x0 = np.array([np.pi - 0.6 + 0.2 * rng.uniform(-1, 1), 0, 0, 0])
        u = 6.0 * rng.standard_normal((N, 1))                       # large controls so that many steps clamp
        x, un, _ = O.forward_pass(O.GaussianPolicy.empty(), x0, u, None, 1, om.f, om.costfun, lims)
        fx, fu, _, _, _, cx, cu, cxx, cxu, cuu = om.df(x, un)
        xs.append(x); us.append(un); fxs.append(fx); fus.append(fu); cxs.append(cx); cus.append(cu)
    xs, us, fxs, fus, cxs, cus = map(np.array, (xs, us, fxs, fus, cxs, cus))
    args = (cxs, cus, cxx, cxu, cuu, fxs, fus, 1.0, 2, lims, xs, us)
    r1 = ddp.back_pass(*args)
    r2 = ddp.back_pass(*args, force_generic=True)
    nclamped = 0
    for b in range(B):
        d0, p0, Vx0, Vxx0, dV0 = O.back_pass(cxs[b], cus[b], cxx, cxu, cuu, fxs[b], fus[b], 1.0, 2, lims, xs[b], us[b])
        for r in (r1, r2):
            assert r[0][b] == d0 == 0
            assert relerr(r[1].K[b], p0.K) < TOL and relerr(r[1].k[b], p0.k) < TOL and relerr(r[2][b], Vx0) < TOL and relerr(r[3][b], Vxx0) < TOL
            assert np.array_equal(r[1].K[b] == 0, p0.K == 0)        # clamped steps have K = 0: the active sets agree exactly
        nclamped += int(np.sum(np.all(p0.K[:-1] == 0, axis=(1, 2))))
    assert nclamped > 0


def test_small_kernel_staged_equals_direct_loads(ddp, monkeypatch):
    """The cp.async-staged fx ring of bp_small_kernel changes where operands come from, not the arithmetic: the staged and
    the direct-load build (diagnostic switch DDP_SMALL_NOSTAGE) agree bit for bit, also on a ragged batch (B % 32 != 0)."""
    rng = np.random.default_rng(3)
    B, n, m, N = 45, 4, 1, 37
    fx = np.eye(n) + 0.05 * rng.standard_normal((B, N, n, n)); fu = 0.1 * rng.standard_normal((B, N, n, m))
    x = rng.standard_normal((B, N, n)); u = 0.3 * rng.standard_normal((B, N, m))
    Q, R = np.diag([10.0, 1, 2, 1]), np.eye(m)
    cx, cu = x @ Q.T, u @ R.T
    lims = np.array([[-0.4, 0.4]])
    args = (cx, cu, Q, np.zeros((n, m)), R, fx, fu, 0.7, 2, lims, x, u)
    monkeypatch.delenv("DDP_SMALL_NOSTAGE", raising=False)
    r1 = ddp.back_pass(*args)
    monkeypatch.setenv("DDP_SMALL_NOSTAGE", "1")
    r2 = ddp.back_pass(*args)
    monkeypatch.delenv("DDP_SMALL_NOSTAGE", raising=False)
    assert np.array_equal(r1[0], r2[0])
    for a, b in ((r1[1].K, r2[1].K), (r1[1].k, r2[1].k), (r1[2], r2[2]), (r1[3], r2[3]), (r1[4], r2[4])):
        assert np.array_equal(a, b)
    assert np.any(np.all(r1[1].K[:, :-1] == 0, axis=(2, 3)))        # some steps are clamped
