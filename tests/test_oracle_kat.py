"""CPU tests pinning the oracle (oracle/ddp_oracle.py) with the analytic known answers the reference's
maths implies and with the reference's only assertions (test/test_readme.jl:82-84).  The reference
ships no golden vectors and Julia is not available, so these are what pins the oracle
(SURVEY.md section 8c)."""
import numpy as np
import pytest
import scipy.linalg as sla
from scipy.optimize import lsq_linear

from helpers import make_lq, rollout
from oracle import ddp_oracle as O


def test_lq_backpass_is_riccati_and_dare():
    """lambda = 0 backward pass == discrete Riccati recursion; long horizon -> DARE solution."""
    rng = np.random.default_rng(0)
    n, m, N = 6, 2, 400
    A, B, Q, R = make_lq(rng, n, m, h=0.1)
    x = rng.standard_normal((N, n)); u = rng.standard_normal((N, m))
    d, pol, Vx, Vxx, dV = O.back_pass(x @ Q, u @ R, Q, np.zeros((n, m)), R, A, B, 0.0, 1, None, x, u)
    assert d == 0
    P = Q.copy()
    for i in range(N - 2, -1, -1):
        Kr = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
        assert np.allclose(pol.K[i], -Kr, rtol=1e-9, atol=1e-12)
        P = Q + A.T @ P @ A - A.T @ P @ B @ Kr
        assert np.allclose(Vxx[i], P, rtol=1e-9, atol=1e-12)
    Pinf = sla.solve_discrete_are(A, B, Q, R)
    assert np.allclose(Vxx[0], Pinf, rtol=1e-6)
    assert np.array_equal(Vxx[0], Vxx[0].T)                  # backward_pass.jl:71-72: exactly symmetric


def test_time_varying_dispatch_variants_agree():
    """the three live back_pass methods (backward_pass.jl:162/179/217) compute the same thing."""
    rng = np.random.default_rng(1)
    n, m, N = 5, 2, 30
    A, B, Q, R = make_lq(rng, n, m)
    x = rng.standard_normal((N, n)); u = rng.standard_normal((N, m))
    cxu = 0.01 * rng.standard_normal((n, m))
    rep = lambda a: np.tile(a, (N, 1, 1))
    r1 = O.back_pass(x @ Q, u @ R, Q, cxu, R, A, B, 0.3, 2, None, x, u)
    r2 = O.back_pass(x @ Q, u @ R, Q, cxu, R, rep(A), rep(B), 0.3, 2, None, x, u)
    r3 = O.back_pass(x @ Q, u @ R, rep(Q), rep(cxu), rep(R), rep(A), rep(B), 0.3, 2, None, x, u)
    for r in (r2, r3):
        assert np.allclose(r[1].K, r1[1].K, rtol=1e-12, atol=1e-14) and np.allclose(r[3], r1[3], rtol=1e-12, atol=1e-14)


def test_forced_non_pd_diverges_at_first_step():
    rng = np.random.default_rng(2)
    n, m, N = 4, 2, 10
    A, B, Q, R = make_lq(rng, n, m)
    x = rng.standard_normal((N, n)); u = rng.standard_normal((N, m))
    d, pol, Vx, Vxx, dV = O.back_pass(x @ Q, u @ R, Q, np.zeros((n, m)), -np.eye(m), A, B, 0.0, 1, None, x, u)
    assert d == N - 1                                        # 1-based index of the first processed step
    assert np.all(pol.K == 0) and np.all(pol.k == 0) and np.all(Vxx[:-1] == 0)     # quirk Q10
    assert np.array_equal(Vx[-1], (x @ Q)[-1])


@pytest.mark.parametrize("m", [1, 2, 5, 9])
def test_boxqp_kkt_and_bounded_lsq(m):
    rng = np.random.default_rng(10 + m)
    for trial in range(20):
        G = rng.standard_normal((m, m)); H = G @ G.T + 0.1 * np.eye(m); H = (H + H.T) / 2
        g = 3 * rng.standard_normal(m)
        lo, up = -rng.random(m), rng.random(m)
        x, res, Hf, free, nf = O.boxQP(H, g, lo, up, np.zeros(m))
        assert res in (4, 5, 6)
        grad = g + H @ x
        for i in range(m):                                   # KKT
            if x[i] == lo[i]:
                assert grad[i] > -1e-7
            elif x[i] == up[i]:
                assert grad[i] < 1e-7
            else:
                assert abs(grad[i]) < 1e-6
        Lc = np.linalg.cholesky(H)                           # min ½|L'x + L^-1 g|² s.t. bounds
        ref = lsq_linear(Lc.T, -np.linalg.solve(Lc, g), bounds=(lo, up), tol=1e-14).x
        assert np.allclose(x, ref, atol=1e-6)
        if free.any():
            assert np.allclose(Hf.T @ Hf, H[np.ix_(free, free)], rtol=1e-10)


def test_boxqp_m1_closed_form_and_result_codes():
    x, res, Hf, free, nf = O.boxQP(np.array([[2.0]]), np.array([-1.0]), np.array([-5.0]), np.array([5.0]), np.array([0.0]))
    assert abs(x[0] - 0.5) < 1e-15 and res in (4, 5) and free[0]
    x, res, Hf, free, nf = O.boxQP(np.array([[2.0]]), np.array([-100.0]), np.array([-5.0]), np.array([5.0]), np.array([0.0]))
    assert x[0] == 5.0 and res == 6 and not free[0]          # all dimensions clamped
    with pytest.raises(O.PosDefException):
        O.boxQP(np.array([[-1.0]]), np.array([1.0]), np.array([-1.0]), np.array([1.0]), np.array([0.0]))
    # quirk Q3: the reference's cholesky(H[free,free]) refuses a matrix that is not exactly symmetric
    H = np.array([[2.0, 0.3], [0.3 + 1e-15, 1.0]])
    with pytest.raises(O.PosDefException):
        O.boxQP(H, np.ones(2), -np.ones(2), np.ones(2), np.zeros(2), hermitian_check=True)
    assert O.boxQP(H, np.ones(2), -np.ones(2), np.ones(2), np.zeros(2), hermitian_check=False)[1] >= 1


def test_lambda_schedule_quirk_q1():
    lam, dlam = 1.0, 1.0
    seq = []
    for _ in range(4):
        lam, dlam = O._lam_increase(lam, dlam, 1.6, 1e-6)
        seq.append(lam)
    assert np.allclose(seq, [1.0, 1.6, 1.6 * 2.56, 1.6 * 2.56 * 4.096])      # λ uses the OLD dλ (iLQG.jl:246)


def test_ilqg_on_lq_reaches_the_optimum_and_thresholds():
    """one LQ instance of test_readme.jl's distribution: converges to the LQR optimum, cost within
    the reference's thresholds (max < 25 covers every run, min < 5 is met by this seed)."""
    rng = np.random.default_rng(0)
    n, m, N = 10, 2, 300
    A, B, Q, R = make_lq(rng, n, m)
    om = O.LinearModel(A, B, Q, R)
    u0 = 0.1 * rng.standard_normal((N, m))
    x, u, pol, Vx, Vxx, cost, tr = O.iLQG(om.f, om.costfun, om.df, np.ones(n), u0)
    assert tr["status"] == 0 and tr["lam_final"] < 1e-5
    assert np.sum(cost) < 25
    # optimal cost of the finite-horizon LQ problem from the Riccati value function: ½ x0' P0 x0
    d, p, Vx0, Vxx0, dV = O.back_pass(0 * x, 0 * u, Q, np.zeros((n, m)), R, A, B, 0.0, 1, None, 0 * x, 0 * u)
    # (+ the last control's cost: u[:,N] is never optimised, quirk Q7, and keeps its initial value)
    opt = 0.5 * np.ones(n) @ Vxx0[0] @ np.ones(n) + 0.5 * u0[-1] @ R @ u0[-1]
    assert np.array_equal(u[-1], u0[-1])
    assert abs(np.sum(cost) - opt) < 1e-6 * opt
    # first accepted step takes alpha = 1 with ratio ~ 1 (LQ problem, quadratic model exact up to lambda)
    assert tr["alpha"][0][1] == 1.0


def test_ilqg_initial_divergence_returns_none_and_zero_iterations_raises():
    n, m, N = 3, 1, 40
    om = O.LinearModel(50.0 * np.eye(n), np.ones((n, m)), np.eye(n), np.eye(m))
    assert O.iLQG(om.f, om.costfun, om.df, 1e6 * np.ones(n), np.ones((N, m))) is None      # iLQG.jl:205-210
    rng = np.random.default_rng(3)
    A, B, Q, R = make_lq(rng, n, m)
    om = O.LinearModel(A, B, Q, R)
    with pytest.raises(RuntimeError):                                                       # iLQG.jl:335 (quirk Q5)
        O.iLQG(om.f, om.costfun, om.df, np.zeros(n), np.zeros((N, m)), lam=1e-9, dlam=1e-9)


def test_gps_reduces_to_plain_backpass_and_kl_zero_for_same_policy():
    rng = np.random.default_rng(4)
    n, m, N = 5, 2, 25
    A, B, Q, R = make_lq(rng, n, m)
    u = 0.1 * rng.standard_normal((N, m)); x = rollout(A, B, np.ones(n), u)
    cx, cu = x @ Q, u @ R
    rep = lambda a: np.tile(a, (N, 1, 1))
    # zero KL terms and eta = 1  =>  identical to back_pass with lambda = 0
    zero_terms = (np.zeros((N, n)), np.zeros((N, m)), np.zeros((N, n, n)), np.zeros((N, m, n)), np.zeros((N, m, m)))
    d1, p1, Vx1, Vxx1, dV1 = O.back_pass_gps(cx, cu, rep(Q), rep(np.zeros((n, m))), rep(R), rep(A), rep(B), None, x, u,
                                             (zero_terms, np.array([1e-8, 1.0, 1e16])))
    d0, p0, Vx0, Vxx0, dV0 = O.back_pass(cx, cu, Q, np.zeros((n, m)), R, A, B, 0.0, 1, None, x, u)
    assert np.allclose(p1.K, p0.K, rtol=1e-10, atol=1e-13) and np.allclose(Vxx1, Vxx0, rtol=1e-10, atol=1e-13)
    assert np.allclose(p1.Sigma[:-1] @ p1.Sigmai[:-1], np.eye(m), atol=1e-9)
    # KL(p || p) == 0 and kl > 0 for a perturbed policy
    pol = O.GaussianPolicy(N, n, m, p1.K, p1.k, p1.Sigma, p1.Sigmai)
    sig = O.forward_covariance(A, 1e-4 * np.eye(n), pol)
    assert np.allclose(O.kl_div_wiki(x, x, sig, pol, pol), 0, atol=1e-12)
    pol2 = O.GaussianPolicy(N, n, m, 1.1 * p1.K, p1.k + 0.01, p1.Sigma, p1.Sigmai)
    assert np.all(O.kl_div_wiki(x, x, sig, pol2, pol) > 0)


def test_calc_eta_bracket_update():
    eb = np.array([1e-8, 1.0, 1e16])
    pol = O.GaussianPolicy.identity(3, 1, 1)
    sig = np.tile(np.eye(2), (3, 1, 1))
    x = np.zeros((3, 1))
    # identical policies: divergence 0 < kl_step -> eta too big -> upper bracket shrinks (klutils.jl:119-122)
    eb2, sat, div = O.calc_eta(x, x, sig, eb.copy(), pol, pol, 1.0)
    assert not sat and div == 0 and eb2[2] == 1.0 and eb2[1] == max(np.sqrt(1e-8 * 1.0), 0.1)
    assert O.calc_eta(x, x, sig, eb.copy(), pol, pol, 0.0)[1] is True        # kl_step <= 0: satisfied immediately


def test_pendcart_model_matches_closed_form_jacobian():
    om = O.PendcartModel()
    x = np.array([[np.pi - 0.3, 0.2, 0.0, -0.1], [np.pi, 0, 0, 0]]); u = np.array([[0.7], [0.0]])
    fx, fu, *_ , cx, cu, cxx, cxu, cuu = om.df(x, u)
    # ZoH discretisation ~ I + h*fxc to first order
    g, l, h, d = om.g, om.l, om.h, om.d
    fxc = np.array([[0, 1, 0, 0], [-g / l * np.cos(x[0, 0]) - u[0, 0] / l * np.sin(x[0, 0]), -d, 0, 0], [0, 0, 0, 1], [0, 0, 0, 0]])
    assert np.allclose(fx[0], np.eye(4) + h * fxc, atol=5e-3) and np.allclose(fx[0], sla.expm(h * fxc), atol=1e-12)
    assert cuu.shape == (1, 1) and cxu.shape == (4, 1) and np.allclose(cx[1], 0)
    c = om.costfun(x, u)
    assert c.shape == (3,) and abs(c[-1] - 0.5 * (x[-1] - om.goal) @ om.Q @ (x[-1] - om.goal)) < 1e-15


def test_kl_div_is_the_expected_gaussian_kl_and_covariance_is_lyapunov():
    """kl_div_wiki (klutils.jl:70-100) is E_x[ KL( N(Kn x + kn, Sn) || N(Kp x + kp, Sp) ) ] for x ~ N(mu, St): checked against
    the textbook Gaussian KL evaluated by Monte Carlo-free algebra written independently (means and covariances of the two
    control distributions), and forward_covariance (forward_pass.jl:47-54) converges to the discrete Lyapunov solution."""
    rng = np.random.default_rng(4)
    n, m, N = 5, 2, 6
    def spd(k):
        W = rng.standard_normal((k, k)); return W @ W.T / k + 0.5 * np.eye(k)
    Kp, Kn = rng.standard_normal((N, m, n)), rng.standard_normal((N, m, n))
    kp, kn = rng.standard_normal((N, m)), rng.standard_normal((N, m))
    Sp, Sn = np.array([spd(m) for _ in range(N)]), np.array([spd(m) for _ in range(N)])
    prev = O.GaussianPolicy(N, n, m, Kp, kp, Sp, np.array([np.linalg.inv(s) for s in Sp]))
    new = O.GaussianPolicy(N, n, m, Kn, kn, Sn, np.array([np.linalg.inv(s) for s in Sn]))
    xold, xnew = rng.standard_normal((N, n)), rng.standard_normal((N, n))
    sig = np.zeros((N, n + m, n + m))
    for t in range(N):
        sig[t, :n, :n] = spd(n)
    got = O.kl_div_wiki(xnew, xold, sig, new, prev)
    for t in range(N):
        mu, St = xnew[t] - xold[t], sig[t, :n, :n]
        Sip = np.linalg.inv(Sp[t])
        # KL(N(a, Sn) || N(b, Sp)) = 1/2 [tr(Sp^-1 Sn) + (b - a)' Sp^-1 (b - a) - m + ln det Sp - ln det Sn]; here b - a = dk + dK x is
        # itself Gaussian in x, so the quadratic term's expectation is its value at the mean plus tr(dK' Sp^-1 dK St)
        dK, dk = Kp[t] - Kn[t], kp[t] - kn[t]
        e = dk + dK @ mu
        want = 0.5 * (np.trace(Sip @ Sn[t]) + e @ Sip @ e + np.trace(dK.T @ Sip @ dK @ St) - m
                      + np.linalg.slogdet(Sp[t])[1] - np.linalg.slogdet(Sn[t])[1])
        assert abs(got[t] - max(0.0, want)) <= 1e-10 * max(1.0, abs(want))
    # covariance propagation: Sigma_{t+1} = fx Sigma_t fx' + R1 from Sigma_1 = R1 tends to the discrete Lyapunov solution
    fx = 0.8 * sla.expm(0.3 * (lambda G: G - G.T)(rng.standard_normal((n, n))))
    R1 = spd(n)
    T = 200
    pol = O.GaussianPolicy(T, n, m, np.zeros((T, m, n)), np.zeros((T, m)), np.tile(np.eye(m), (T, 1, 1)), np.tile(np.eye(m), (T, 1, 1)))
    sg = O.forward_covariance(fx, R1, pol)
    assert np.allclose(sg[0, :n, :n], R1)
    assert np.allclose(sg[-1, :n, :n], sla.solve_discrete_lyapunov(fx, R1), rtol=1e-9, atol=1e-12)
