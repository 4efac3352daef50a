"""Committed golden fixtures (tests/golden/*.npz, generated from the oracle by tests/golden/make_golden.py):
the oracle must still reproduce them bit for bit (CPU), the C++ baseline and the CUDA path must match
them within the parity tolerance (integer outcomes exact)."""
import glob
import os

import numpy as np
import pytest

from helpers import relerr
from oracle import cpu_ref as CR
from oracle import ddp_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BP_CASES = sorted(glob.glob(os.path.join(G, "back_pass_*.npz")))
TOL = 1e-8


def _lims(d):
    return None if d["lims"].size == 0 else d["lims"]


@pytest.mark.parametrize("path", BP_CASES, ids=[os.path.basename(p) for p in BP_CASES])
def test_oracle_and_cpp_reproduce_back_pass_fixture(path):
    d = np.load(path)
    lim = _lims(d)
    for b in range(2):
        dv, p, Vx, Vxx, dV = O.back_pass(d["x"][b] @ d["Q"].T, d["u"][b] @ d["R"].T, d["Q"], d["cxu"], d["R"], d["A"][b], d["Bm"][b],
                                         float(d["lam"]), int(d["regType"]), lim, d["x"][b], d["u"][b])
        assert dv == d["diverge"][b]
        assert np.array_equal(p.K, d["K"][b]) and np.array_equal(Vxx, d["Vxx"][b]) and np.array_equal(dV, d["dV"][b])
    cx, cu = d["x"] @ d["Q"].T, d["u"] @ d["R"].T
    dv, K, k, Vx, Vxx, _, _, dV = CR.back_pass(cx, cu, d["Q"].T, d["cxu"].T, d["R"].T, np.swapaxes(d["A"], -1, -2), np.swapaxes(d["Bm"], -1, -2),
                                               float(d["lam"]), int(d["regType"]), lim, d["u"])
    assert np.array_equal(dv, d["diverge"])
    assert relerr(np.swapaxes(K, -1, -2), d["K"]) < 1e-9 and relerr(np.swapaxes(Vxx, -1, -2), d["Vxx"]) < 1e-9 and relerr(dV, d["dV"]) < 1e-9


def test_oracle_reproduces_boxqp_and_ilqg_fixtures():
    d = np.load(os.path.join(G, "boxqp_m5.npz"))
    for b in range(d["H"].shape[0]):
        x, r, Hf, free, nf = O.boxQP(d["H"][b], d["g"][b], d["lower"][b], d["upper"][b], d["x0"][b])
        assert np.array_equal(x, d["x"][b]) and r == d["result"][b] and np.array_equal(free, d["free"][b]) and nf == d["nfactor"][b]
    x, res, Hf, free, nf = CR.boxqp(d["H"], d["g"], d["lower"], d["upper"], d["x0"])
    assert np.array_equal(x, d["x"]) and np.array_equal(res, d["result"]) and np.array_equal(free, d["free"]) and np.array_equal(nf, d["nfactor"])
    d = np.load(os.path.join(G, "ilqg_lq_n8.npz"))
    om = O.LinearModel(d["A"], d["Bm"], d["Q"], d["R"])
    x, u, pol, Vx, Vxx, cost, tr = O.iLQG(om.f, om.costfun, om.df, np.ones(8), d["u0"])
    assert tr["status"] == d["status"] and tr["iters"] == d["iters"] and np.array_equal(x, d["x"]) and np.array_equal(u, d["u"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", BP_CASES, ids=[os.path.basename(p) for p in BP_CASES])
def test_gpu_back_pass_matches_fixture(ddp, path):
    d = np.load(path)
    lim = _lims(d)
    for generic in (False, True):
        dv, pol, Vx, Vxx, dV = ddp.back_pass(d["x"] @ d["Q"].T, d["u"] @ d["R"].T, d["Q"], d["cxu"], d["R"], d["A"][:, None], d["Bm"][:, None],
                                             float(d["lam"]), int(d["regType"]), lim, d["x"], d["u"], force_generic=generic)
        assert np.array_equal(dv, d["diverge"])
        for a, b in ((pol.K, d["K"]), (pol.k, d["k"]), (Vx, d["Vx"]), (Vxx, d["Vxx"]), (dV, d["dV"])):
            assert relerr(a, b) < TOL
        if lim is not None:
            assert np.array_equal(pol.K == 0, d["K"] == 0)          # clamped sets agree exactly


@pytest.mark.gpu
def test_gpu_boxqp_and_ilqg_match_fixtures(ddp):
    d = np.load(os.path.join(G, "boxqp_m5.npz"))
    x, res, Hf, free, nf = ddp.boxQP(d["H"], d["g"], d["lower"], d["upper"], d["x0"])
    assert np.array_equal(x, d["x"]) and np.array_equal(res, d["result"]) and np.array_equal(free, d["free"]) and np.array_equal(nf, d["nfactor"])
    d = np.load(os.path.join(G, "ilqg_lq_n8.npz"))
    model = ddp.LinearModel(d["A"], d["Bm"], d["Q"], d["R"])
    x, u, pol, Vx, Vxx, cost, tr = ddp.iLQG(model.f, model.costfun, model.df, np.ones(8), d["u0"])
    assert tr["status"] == d["status"] and tr["iter"] == d["iters"]
    assert relerr(x, d["x"]) < 1e-7 and relerr(u, d["u"]) < 1e-7 and abs(cost - d["cost"]) < 1e-8 * abs(d["cost"])
    assert abs(tr["lam"] - d["lam_final"]) <= 1e-12 * d["lam_final"]


def _kl_fixture():
    d = np.load(os.path.join(G, "ilqgkl_n32_m8.npz"))
    N, n, m = d["x"].shape[0], d["x"].shape[1], d["u"].shape[1]
    return d, N, n, m


def test_oracle_reproduces_ilqgkl_fixture():
    """KL-constrained path at n=32, m=8: sweep, KL evaluation and whole iLQGkl solve, bit for bit on CPU."""
    d, N, n, m = _kl_fixture()
    om = O.LinearModel(d["A"], d["Bm"], d["Q"], d["R"])
    rep = lambda a: np.tile(a, (N, 1, 1))
    prev0 = O.GaussianPolicy(N, n, m, d["K_prev"].copy(), np.zeros((N, m)), d["Sig_prev"].copy(), d["Sigi_prev"].copy())
    dg, pg, Vxg, _, dVg = O.back_pass_gps(d["x"] @ d["Q"].T, d["u"] @ d["R"].T, rep(d["Q"]), rep(np.zeros((n, m))), rep(d["R"]), rep(d["A"]),
                                          rep(d["Bm"]), None, d["x"], d["u"], (O.grad_kl(prev0), d["eta"]))
    assert dg == d["gps_diverge"] and np.array_equal(pg.K, d["gps_K"]) and np.array_equal(pg.Sigma, d["gps_Sigma"]) and np.array_equal(dVg, d["gps_dV"])
    xn, un, _ = O.forward_pass(pg, d["x"][0], d["u"], d["x"], 1, om.f, om.costfun, None)
    kl = O.kl_div_wiki(xn, d["x"], O.forward_covariance(d["A"], d["R1"], pg), pg, prev0)
    assert np.array_equal(xn, d["fwd_x"]) and np.array_equal(kl, d["kl_t"])
    prev = O.GaussianPolicy(N, n, m, d["K_prev"].copy(), d["u"].copy(), d["Sig_prev"].copy(), d["Sigi_prev"].copy())
    r = O.iLQGkl(om.f, om.costfun, lambda xx, uu: om.df(xx, uu, time_varying=True), d["x"], prev, d["A"], d["R1"], kl_step=float(d["kl_step"]),
                 cost=float(d["cost0"]))
    assert r[6]["iters"] == d["sol_iters"] and bool(r[6]["satisfied"]) == bool(d["sol_satisfied"]) and np.array_equal(r[0], d["sol_x"])


@pytest.mark.gpu
def test_gpu_kl_path_matches_fixture(ddp):
    """back_pass_gps (tile kernel), forward rollout, KL tile kernel and ddp_ilqgkl_solve_f64 against the committed fixture."""
    d, N, n, m = _kl_fixture()
    rep = lambda a: np.tile(a, (N, 1, 1))
    gp = ddp.GaussianPolicy(N, n, m, d["K_prev"], np.zeros((N, m)), d["Sig_prev"], d["Sigi_prev"])
    dg, pg, Vxg, _, dVg = ddp.back_pass_gps(d["x"] @ d["Q"].T, d["u"] @ d["R"].T, d["Q"], np.zeros((n, m)), d["R"], d["A"], d["Bm"], None, d["x"], d["u"],
                                            (gp, d["eta"]))
    assert dg == d["gps_diverge"]
    for a, b in ((pg.K, d["gps_K"]), (pg.k, d["gps_k"]), (pg.Sigma, d["gps_Sigma"]), (pg.Sigmai, d["gps_Sigmai"]), (Vxg, d["gps_Vx"]), (dVg, d["gps_dV"])):
        assert relerr(a, b) < TOL
    model = ddp.LinearModel(d["A"], d["Bm"], d["Q"], d["R"])
    xn, un, cn = ddp.forward_pass(pg, d["x"][0], d["u"], d["x"], 1.0, model.f, model.costfun, None)
    assert relerr(xn, d["fwd_x"]) < TOL and relerr(un, d["fwd_u"]) < TOL
    klt, klm = ddp.kl_div_wiki(xn, d["x"], d["A"], d["R1"], pg, gp)
    assert relerr(klt, d["kl_t"]) < 1e-7 and abs(klm - d["kl_t"].mean()) < 1e-7 * d["kl_t"].mean()
    prev = ddp.GaussianPolicy(N, n, m, d["K_prev"], d["u"], d["Sig_prev"], d["Sigi_prev"])
    r = ddp.iLQGkl_device(model.f, model.costfun, model.df, d["x"], prev, d["A"], d["R1"], kl_step=float(d["kl_step"]), cost=float(d["cost0"]))
    tr = r[6]
    assert tr["iter"] == d["sol_iters"] and bool(tr["satisfied"]) == bool(d["sol_satisfied"])
    assert np.allclose([tr["eta_min"], tr["eta"], tr["eta_max"]], d["sol_etabracket"], rtol=1e-9)
    assert abs(tr["divergence"] - d["sol_divergence"]) < 1e-7 * max(1.0, abs(d["sol_divergence"]))
    assert relerr(r[0], d["sol_x"]) < 1e-7 and relerr(r[1], d["sol_u"]) < 1e-7 and relerr(r[2].K, d["sol_K"]) < 1e-7
    assert abs(r[5] - d["sol_cost"]) < 1e-8 * abs(d["sol_cost"])
