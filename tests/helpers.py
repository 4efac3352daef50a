"""Seeded problem generators shared by the oracle and GPU parity tests."""
import numpy as np
import scipy.linalg as sla


def make_lq(rng, n, m, h=0.01):
    """demo_linear.jl:9-21 distribution: A = exp(h(G-G')), B = h randn, Q = hI, R = 0.1hI."""
    G = rng.standard_normal((n, n))
    A = sla.expm(h * (G - G.T))
    B = h * rng.standard_normal((n, m))
    return A, B, h * np.eye(n), 0.1 * h * np.eye(m)


def rollout(A, B, x0, u):
    N = u.shape[0]
    x = np.zeros((N, A.shape[0]))
    x[0] = x0
    for t in range(N - 1):
        x[t + 1] = A @ x[t] + B @ u[t]
    return x


def make_batch_lq(seed, B, n, m, N, h=0.01):
    rng = np.random.default_rng(seed)
    As, Bs, xs, us = [], [], [], []
    Q = R = None
    for _ in range(B):
        A, Bm, Q, R = make_lq(rng, n, m, h)
        x0 = 1.0 + 0.1 * rng.standard_normal(n)
        u = 0.1 * rng.standard_normal((N, m))
        As.append(A); Bs.append(Bm); us.append(u); xs.append(rollout(A, Bm, x0, u))
    return np.array(As), np.array(Bs), Q, R, np.array(xs), np.array(us)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def relerr_elem(a, b, floor=1e-4):
    """Largest ELEMENT-WISE relative error |a-b| / max(|b|, floor * max|b|): every entry is measured against its own
    magnitude; entries smaller than ``floor`` times the tensor's largest (where FP64 cancellation makes a relative
    figure meaningless) are measured against that floor instead."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    scale = np.max(np.abs(b))
    den = np.maximum(np.abs(b), floor * (scale if scale > 0 else 1.0))
    return float(np.max(np.abs(a - b) / den))
