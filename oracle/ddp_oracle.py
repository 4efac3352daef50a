"""CPU oracle: a NumPy FP64 restatement of the iLQG/DDP hot path of
baggepinnen/DifferentialDynamicProgramming.jl (v0.5.0).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The shipped path
(``differentialdynamicprogramming.jl_b200``) never does; it fails loudly without its CUDA library.

Pinning status: the reference ships NO golden vectors, known-answer tests or fixtures for this
path (test/runtests.jl:7-12 are assertion-free smoke runs) and Julia is not installed in the build
container, so the reference itself cannot be executed here.  BIT-LEVEL PARITY WITH JULIA IS
THEREFORE UNPINNED (status: "parity unpinned").  What pins this oracle instead (tests/test_oracle_*.py):
  * the reference's only assertions, the statistical cost thresholds of test/test_readme.jl:82-84,
    re-run on fresh instances of the same problem distribution;
  * analytic known answers the reference's maths implies: LQ backward pass == discrete Riccati
    recursion (and its long-horizon limit == scipy's DARE), boxQP == KKT / bounded least squares,
    one-step convergence of iLQG on LQ problems, forced non-PD ``cuu`` => ``diverge == N-1``,
    kl_div_wiki == the expected Gaussian KL written independently, forward_covariance -> discrete Lyapunov;
  * committed fixtures of its own outputs (tests/golden/, generator beside them), reproduced bit for bit.

Conventions.  Time is 0-based here; the reference (Julia) is 1-based.  ``diverge`` is returned as the
reference's 1-based timestep index (0 == success) so values compare equal to the reference's.
Per-step matrices are stored time-first in *math layout*:
    fx[t][i, j] = d f_i / d x_j  (N, n, n)   fu (N, n, m)   cx (N, n)   cu (N, m)
    cxx (n, n) | (N, n, n)   cxu (n, m) | (N, n, m)   cuu (m, m) | (N, m, m)
    K (N, m, n)   k (N, m)   Vx (N, n)   Vxx (N, n, n)   Quu (N, m, m)
A 2-D ``fx``/``fu`` means time-invariant (the reference's LTI method, backward_pass.jl:217).

Deliberate, documented definitions (Julia delegates to OpenBLAS/LAPACK whose summation order is
not part of the reference's source):
  * ``boxQP`` arithmetic (objective, gradient, Cholesky, triangular solves) uses a fixed
    sequential summation order with separate multiply and add (no FMA), so that the C++ baseline
    and the CUDA kernel can reproduce every branch decision and every output bit of boxQP.
  * The large products of the Q-expansion use NumPy/BLAS ``@``; parity there is by tolerance
    (1e-8 relative, stated in the tests).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

# --------------------------------------------------------------------------------------------
# GaussianPolicy  (reference: src/iLQG.jl:39-53)
# --------------------------------------------------------------------------------------------


@dataclass
class GaussianPolicy:
    """Carrier between backward and forward pass.  iLQG.jl:39-47 (fields T,n,m,K,k,Sigma,Sigmai)."""

    T: int = 0
    n: int = 0
    m: int = 0
    K: Optional[np.ndarray] = None       # (T, m, n)
    k: Optional[np.ndarray] = None       # (T, m)
    Sigma: Optional[np.ndarray] = None   # (T, m, m)
    Sigmai: Optional[np.ndarray] = None  # (T, m, m)

    @staticmethod
    def empty() -> "GaussianPolicy":     # iLQG.jl:50
        return GaussianPolicy()

    @staticmethod
    def identity(T: int, n: int, m: int) -> "GaussianPolicy":  # iLQG.jl:51
        eye = np.tile(np.eye(m), (T, 1, 1))
        return GaussianPolicy(T, n, m, np.zeros((T, m, n)), np.zeros((T, m)), eye.copy(), eye.copy())

    def isempty(self) -> bool:           # iLQG.jl:52
        return self.T == 0 and self.n == 0 and self.m == 0

    def __len__(self) -> int:            # iLQG.jl:53
        return self.T


# --------------------------------------------------------------------------------------------
# Small dense kernels with a *defined* operation order (shared by boxQP everywhere)
# --------------------------------------------------------------------------------------------


def seq_chol_upper(A: np.ndarray):
    """Upper Cholesky factor R (R'R = A) reading only the upper triangle of ``A``.

    Stands for ``cholesky(Hermitian(A))`` / LAPACK ``dpotrf('U')`` (backward_pass.jl:35,
    boxQP.jl:111).  Column-by-column, sequential sums, no FMA.  Fails (returns ok=False) when a
    pivot is <= 0 or NaN, which is LAPACK's failure condition.
    """
    n = A.shape[0]
    R = np.zeros((n, n))
    for j in range(n):
        for i in range(j):
            s = float(A[i, j])
            for p in range(i):
                s = s - float(R[p, i]) * float(R[p, j])
            R[i, j] = s / float(R[i, i])
        d = float(A[j, j])
        for p in range(j):
            d = d - float(R[p, j]) * float(R[p, j])
        if not (d > 0.0):          # catches d <= 0 and NaN
            return R, False
        R[j, j] = math.sqrt(d)
    return R, True


def seq_solve_upper_t(R: np.ndarray, b: np.ndarray) -> np.ndarray:
    """y = R' \\ b  (forward substitution with the transpose of an upper factor)."""
    n = R.shape[0]
    y = np.zeros(n)
    for i in range(n):
        s = float(b[i])
        for p in range(i):
            s = s - float(R[p, i]) * float(y[p])
        y[i] = s / float(R[i, i])
    return y


def seq_solve_upper(R: np.ndarray, y: np.ndarray) -> np.ndarray:
    """x = R \\ y  (back substitution)."""
    n = R.shape[0]
    x = np.zeros(n)
    for i in range(n - 1, -1, -1):
        s = float(y[i])
        for p in range(i + 1, n):
            s = s - float(R[i, p]) * float(x[p])
        x[i] = s / float(R[i, i])
    return x


def _qp_value(H, g, x):
    """x'g + 0.5 x'H x evaluated as Julia parses it: dot(x,g) + ((0.5x')H)x  (boxQP.jl:63,139)."""
    n = len(x)
    s1 = 0.0
    for i in range(n):
        s1 = s1 + float(x[i]) * float(g[i])
    s2 = 0.0
    for j in range(n):
        t = 0.0
        for i in range(n):
            t = t + (0.5 * float(x[i])) * float(H[i, j])
        s2 = s2 + t * float(x[j])
    return s1 + s2


def _matvec_seq(H, x):
    n = len(x)
    out = np.zeros(n)
    for i in range(n):
        s = 0.0
        for j in range(n):
            s = s + float(H[i, j]) * float(x[j])
        out[i] = s
    return out


# --------------------------------------------------------------------------------------------
# boxQP  (reference: src/boxQP.jl:29-188)
# --------------------------------------------------------------------------------------------


class PosDefException(Exception):
    pass


def _clamp_jl(x, lo, hi):
    """Julia's ``clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))`` (Base), elementwise: NaN stays NaN and an
    inverted interval resolves hi-first -- what ``clamp.(…)`` does at boxQP.jl:58,139,144 and forward_pass.jl:23."""
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > hi, hi, np.where(x < lo, lo, x))


def boxQP(H, g, lower, upper, x0, *, maxIter=100, minGrad=1e-8, minRelImprove=1e-8, stepDec=0.6,
          minStep=1e-22, Armijo=0.1, hermitian_check=True):
    """Projected-Newton box-constrained QP, boxQP.jl:29-188.

    Returns ``(x, result, Hfree, free, nfactor)``; raises :class:`PosDefException` where the
    reference's ``cholesky`` throws (callers wrap in try, backward_pass.jl:48-52).

    ``hermitian_check=True`` reproduces Julia's ``cholesky(H[free,free])`` refusing a matrix that is
    not *exactly* symmetric (quirk Q3).  The batched engine factors the upper triangle instead
    (``hermitian_check=False``), a superset that never changes the result when H is symmetric.
    """
    H = np.asarray(H, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    lower = np.asarray(lower, dtype=np.float64).reshape(-1)
    upper = np.asarray(upper, dtype=np.float64).reshape(-1)
    n = H.shape[0]
    clamped = np.zeros(n, dtype=bool)              # :46
    free = np.ones(n, dtype=bool)                  # :47
    oldvalue = 0.0
    result = 0
    nfactor = 0
    Hfree = np.zeros((n, n))                       # :53

    x = _clamp_jl(np.asarray(x0, dtype=np.float64).reshape(-1), lower, upper)  # :58 clamp
    value = _qp_value(H, g, x)                     # :63

    it = 1
    while it <= maxIter:                           # :71
        if result != 0:                            # :73
            break
        if it > 1 and (oldvalue - value) < minRelImprove * abs(oldvalue):   # :78
            result = 4
            break
        oldvalue = value
        grad = g + _matvec_seq(H, x)               # :85
        old_clamped = clamped
        clamped = np.zeros(n, dtype=bool)
        for i in range(n):                         # :92-94 exact FP equality on the bounds
            clamped[i] = (x[i] == lower[i] and grad[i] > 0) or (x[i] == upper[i] and grad[i] < 0)
        free = ~clamped
        if clamped.all():                          # :98
            result = 6
            break
        factorize = True if it == 1 else bool(np.any(old_clamped != clamped))   # :104-108
        if factorize:
            Hff = H[np.ix_(free, free)]
            if hermitian_check and not np.array_equal(Hff, Hff.T):
                raise PosDefException("matrix is not Hermitian; Cholesky factorization failed")
            Hfree, ok = seq_chol_upper(Hff)        # :111
            if not ok:
                raise PosDefException("matrix is not positive definite; Cholesky factorization failed")
            nfactor += 1
        gf = grad[free]
        s = 0.0
        for v in gf:                               # norm(grad[free]) :120
            s = s + float(v) * float(v)
        gnorm = math.sqrt(s)
        if gnorm < minGrad:
            result = 5
            break
        grad_clamped = g + _matvec_seq(H, x * clamped)                          # :127
        search = np.zeros(n)
        search[free] = -seq_solve_upper(Hfree, seq_solve_upper_t(Hfree, grad_clamped[free])) - x[free]  # :129
        sdotg = 0.0
        for i in range(n):                         # :132
            sdotg = sdotg + float(search[i]) * float(grad[i])
        if sdotg >= 0:                             # :133 (should not happen) -> leaves result == 0
            break
        step = 1.0                                 # :138
        xc = _clamp_jl(x + step * search, lower, upper)
        vc = _qp_value(H, g, xc)
        while (vc - oldvalue) / (step * sdotg) < Armijo:   # :142
            step = step * stepDec
            xc = _clamp_jl(x + step * search, lower, upper)
            vc = _qp_value(H, g, xc)
            if step < minStep:
                result = 2
                break
        x = xc                                     # :161-163
        value = vc
        it += 1
    if it == maxIter:                              # :167 (quirk Q4)
        result = 1
    return x, result, Hfree, free, nfactor


# --------------------------------------------------------------------------------------------
# back_pass  (reference: src/backward_pass.jl:3-79, 162-252)
# --------------------------------------------------------------------------------------------


def _t(a, i, nd_static):
    """Slice a possibly time-invariant tensor: ``a`` has ``nd_static`` dims when time-invariant."""
    return a if a.ndim == nd_static else a[i]


def _lims_active(lims) -> bool:
    """``!(isempty(lims) || lims[1,1] > lims[1,2])``  backward_pass.jl:31."""
    if lims is None:
        return False
    lims = np.asarray(lims)
    if lims.size == 0:
        return False
    if lims.ndim == 3:                         # time-varying extension: the test reads the first step's block
        lims = lims[0]
    return not (lims[0, 0] > lims[0, 1])


def vectens(a, b):
    """The contraction the 15-argument ``back_pass`` methods call (backward_pass.jl:107,113,118) but the reference never
    defines (quirk Q12; only the broken ``choleskyvectens`` at backward_pass.jl:1 exists, whose shape says what was meant):
    ``vectens(a, b)[p, q] = sum_k a[k] * b[k, q, p]`` for ``b`` of shape (n, d2, d3) -> result (d3, d2)."""
    return np.einsum("k,kqp->pq", a, b)


def back_pass(cx, cu, cxx, cxu, cuu, fx, fu, lam, regType, lims, x, u, *, boxqp_hermitian_check=False,
              fxx=None, fxu=None, fuu=None):
    """One backward sweep; covers the three live dispatch variants backward_pass.jl:162-252 and, with ``fxx``/``fxu``/``fuu``
    (each ``(n,·,·)`` or ``(N,n,·,·)``, ``None`` = ``isempty``), the second-order terms of the 15-argument methods
    (backward_pass.jl:81-160) with ``vectens`` as defined above.  ``lims`` may be ``(m,2)`` or, as an extension,
    ``(N,m,2)`` time-varying.

    Returns ``(diverge, GaussianPolicy, Vx, Vxx, dV)`` as backward_pass.jl:251.  The policy's
    ``Sigmai`` holds the unregularised ``Quu`` (slice N-1 = cuu) and ``Sigma`` is ``None`` (the
    reference returns uninitialised memory there, quirk Q2).  On failure at step ``i`` the outputs
    for steps < i stay zero (quirk Q10).
    """
    cx = np.asarray(cx, dtype=np.float64)
    cu = np.asarray(cu, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    N, m = u.shape
    n = cx.shape[1]
    assert cx.shape == (N, n) and cu.shape == (N, m)
    k = np.zeros((N, m))
    K = np.zeros((N, m, n))
    Vx = np.zeros((N, n))
    Vxx = np.zeros((N, n, n))
    Quu = np.full((N, m, m), np.nan)            # `undef` in the reference
    dV = np.zeros(2)
    Vx[N - 1] = cx[N - 1]                        # :21 / :234
    Vxx[N - 1] = _t(cxx, N - 1, 2)
    Quu[N - 1] = _t(cuu, N - 1, 2)
    use_qp = _lims_active(lims)
    diverge = 0
    for i in range(N - 2, -1, -1):               # i = N-1:-1:1 (1-based)
        fxi, fui = _t(fx, i, 2), _t(fu, i, 2)
        cxxi, cxui, cuui = _t(cxx, i, 2), _t(cxu, i, 2), _t(cuu, i, 2)
        Vxn, Vxxn = Vx[i + 1], Vxx[i + 1]
        Qu = cu[i] + fui.T @ Vxn                                 # :240
        Qx = cx[i] + fxi.T @ Vxn                                 # :241
        Qux = cxui.T + (fui.T @ Vxxn) @ fxi                      # :242
        if fxu is not None:                                      # :106-109
            fxuVx = vectens(Vxn, _t(np.asarray(fxu), i, 3))
            Qux = Qux + fxuVx
        Quu[i] = cuui + (fui.T @ Vxxn) @ fui                     # :243
        if fuu is not None:                                      # :112-115
            fuuVx = vectens(Vxn, _t(np.asarray(fuu), i, 3))
            Quu[i] = Quu[i] + fuuVx
        Qxx = cxxi + (fxi.T @ Vxxn) @ fxi                        # :244
        if fxx is not None:                                      # :118
            Qxx = Qxx + vectens(Vxn, _t(np.asarray(fxx), i, 3))
        Vxx_reg = Vxxn + (lam * np.eye(n) if regType == 2 else 0.0)                 # :245
        Qux_reg = cxui.T + (fui.T @ Vxx_reg) @ fxi                                  # :246
        if fxu is not None:                                      # :121
            Qux_reg = Qux_reg + fxuVx
        QuuF = cuui + (fui.T @ Vxx_reg) @ fui + (lam * np.eye(m) if regType == 1 else 0.0)  # :247
        if fuu is not None:                                      # :123
            QuuF = QuuF + fuuVx
        # ---- @end_backward_pass  :28-79
        if not use_qp:
            R, ok = seq_chol_upper(QuuF)                         # :35 upper triangle only
            if not ok:
                diverge = i + 1
                return diverge, GaussianPolicy(N, n, m, K, k, None, Quu), Vx, Vxx, dV
            k_i = -seq_solve_upper(R, seq_solve_upper_t(R, Qu))                      # :41
            K_i = np.zeros((m, n))
            for j in range(n):                                                       # :42
                K_i[:, j] = -seq_solve_upper(R, seq_solve_upper_t(R, Qux_reg[:, j]))
        else:
            li = lims if np.ndim(lims) == 2 else lims[i]         # time-varying limits: block i
            lower = li[:, 0] - u[i]                              # :45
            upper = li[:, 1] - u[i]
            try:
                k_i, result, R, free, _ = boxQP(QuuF, Qu, lower, upper, k[min(i + 1, N - 2)],
                                                hermitian_check=boxqp_hermitian_check)   # :49
            except PosDefException:
                result = 0                                       # :50-52
            if result < 1:                                       # :53
                diverge = i + 1
                return diverge, GaussianPolicy(N, n, m, K, k, None, Quu), Vx, Vxx, dV
            K_i = np.zeros((m, n))                               # :57
            if free.any():                                       # :58-61
                idx = np.nonzero(free)[0]
                for j in range(n):
                    K_i[idx, j] = -seq_solve_upper(R, seq_solve_upper_t(R, Qux_reg[idx, j]))
        Quuki = Quu[i] @ k_i                                     # :64
        dV = dV + np.array([k_i @ Qu, 0.5 * (k_i @ Quuki)])      # :65,:68
        Vx[i] = ((Qx + K_i.T @ Quuki) + K_i.T @ Qu) + Qux.T @ k_i                    # :69
        Vt = ((Qxx + (K_i.T @ Quu[i]) @ K_i) + K_i.T @ Qux) + Qux.T @ K_i            # :70
        Vxx[i] = (Vt + Vt.T) / 2                                 # :71-72
        k[i] = k_i                                               # :75-76
        K[i] = K_i
    return diverge, GaussianPolicy(N, n, m, K, k, None, Quu), Vx, Vxx, dV


# --------------------------------------------------------------------------------------------
# KL cost terms and back_pass_gps  (reference: src/klutils.jl:8-23, src/backward_pass.jl:259-350)
# --------------------------------------------------------------------------------------------


def grad_kl(traj_prev: GaussianPolicy):
    """The reference's ``∇kl`` (klutils.jl:8-23): cx,cu,cxx,cxu,cuu of the KL term; cxu is (T,m,n)."""
    if traj_prev.isempty():
        return (0, 0, 0, 0, 0)
    T, n, m = traj_prev.T, traj_prev.n, traj_prev.m
    cx, cu = np.zeros((T, n)), np.zeros((T, m))
    cxx, cuu, cxu = np.zeros((T, n, n)), np.zeros((T, m, m)), np.zeros((T, m, n))
    for t in range(T):
        K, k, Si = traj_prev.K[t], traj_prev.k[t], traj_prev.Sigmai[t]
        cx[t] = K.T @ Si @ k
        cu[t] = -Si @ k
        cxx[t] = K.T @ Si @ K
        cuu[t] = Si
        cxu[t] = -Si @ K
    return cx, cu, cxx, cxu, cuu


def back_pass_gps(cx, cu, cxx, cxu, cuu, fx, fu, lims, x, u, kl_cost_terms):
    """KL-augmented backward sweep, backward_pass.jl:259-350.

    ``kl_cost_terms = ((cxkl,cukl,cxxkl,cxukl,cuukl), etabracket)`` with ``etabracket`` a length-3
    vector (min, eta, max).  Time-varying 3-D ``cxx,cxu,cuu,fx,fu`` as the signature (:259) demands;
    2-D inputs are broadcast for convenience.
    """
    (cxkl, cukl, cxxkl, cxukl, cuukl), etabracket = kl_cost_terms
    eta = float(np.asarray(etabracket).reshape(-1)[1])          # :263
    cx = np.asarray(cx, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    N, m = u.shape
    n = cx.shape[1]
    k = np.zeros((N, m))
    K = np.zeros((N, m, n))
    Vx = np.zeros((N, n))
    Vxx = np.zeros((N, n, n))
    Quu = np.full((N, m, m), np.nan)
    Quui = np.full((N, m, m), np.nan)
    dV = np.zeros(2)
    Vx[N - 1] = cx[N - 1]                                        # :280 (not eta-scaled, quirk Q8)
    Vxx[N - 1] = _t(cxx, N - 1, 2)
    Quu[N - 1] = _t(cuu, N - 1, 2) / eta + cuukl[N - 1]          # :282
    Quui[N - 1] = np.linalg.inv(Quu[N - 1])                      # :283
    use_qp = _lims_active(lims)
    diverge = 0
    for i in range(N - 2, -1, -1):
        fxi, fui = _t(fx, i, 2), _t(fu, i, 2)
        cxxi, cxui, cuui = _t(cxx, i, 2), _t(cxu, i, 2), _t(cuu, i, 2)
        Vxn, Vxxn = Vx[i + 1], Vxx[i + 1]
        Qu = cu[i] + fui.T @ Vxn                                 # :287
        Qx = cx[i] + fxi.T @ Vxn
        Qux = cxui.T + (fui.T @ Vxxn) @ fxi
        Quu_i = cuui + (fui.T @ Vxxn) @ fui
        Qxx = cxxi + (fxi.T @ Vxxn) @ fxi
        Qu = Qu / eta + cukl[i]                                  # :295-299
        Qux = Qux / eta + cxukl[i]
        Quu_i = Quu_i / eta + cuukl[i]
        Qx = Qx / eta + cxkl[i]
        Qxx = Qxx / eta + cxxkl[i]
        Quu_i = 0.5 * (Quu_i + Quu_i.T)                          # :301
        Quu[i] = Quu_i
        if not use_qp:
            R, ok = seq_chol_upper(Quu_i)                        # :307
            if not ok:
                diverge = i + 1
                return diverge, GaussianPolicy(N, n, m, K, k, Quui, Quu), Vx, Vxx, dV
            k_i = -seq_solve_upper(R, seq_solve_upper_t(R, Qu))
            K_i = np.zeros((m, n))
            for j in range(n):
                K_i[:, j] = -seq_solve_upper(R, seq_solve_upper_t(R, Qux[:, j]))
        else:
            lower = lims[:, 0] - u[i]
            upper = lims[:, 1] - u[i]
            try:
                k_i, result, R, free, _ = boxQP(Quu_i, Qu, lower, upper, k[min(i + 1, N - 2)])   # :322
            except PosDefException:
                result = 0
            if result < 1:
                diverge = i + 1
                return diverge, GaussianPolicy(N, n, m, K, k, Quui, Quu), Vx, Vxx, dV
            K_i = np.zeros((m, n))
            if free.any():
                idx = np.nonzero(free)[0]
                for j in range(n):
                    K_i[idx, j] = -seq_solve_upper(R, seq_solve_upper_t(R, Qux[idx, j]))
        dV = dV + np.array([k_i @ Qu, 0.5 * (k_i @ Quu_i @ k_i)])                    # :338
        Vx[i] = Qx + K_i.T @ Quu_i @ k_i + K_i.T @ Qu + Qux.T @ k_i                  # :339
        Vt = Qxx + K_i.T @ Quu_i @ K_i + K_i.T @ Qux + Qux.T @ K_i                   # :340
        Vxx[i] = 0.5 * (Vt + Vt.T)                                                   # :341
        k[i] = k_i
        K[i] = K_i
        Quui[i] = np.linalg.inv(Quu_i)                                               # :346
    return diverge, GaussianPolicy(N, n, m, K, k, Quui, Quu), Vx, Vxx, dV


# --------------------------------------------------------------------------------------------
# forward_pass / forward_covariance  (reference: src/forward_pass.jl:9-56)
# --------------------------------------------------------------------------------------------


def forward_pass(traj_new: GaussianPolicy, x0, u, x, alpha, f: Callable, costfun: Callable, lims,
                 diff: Callable = lambda a, b: a - b):
    """Closed-loop rollout, forward_pass.jl:9-33.  ``f(x,u,i)`` gets the 1-based step index."""
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
    u = np.asarray(u, dtype=np.float64)
    N, m = u.shape
    n = x0.shape[0]
    xnew = np.zeros((N, n))
    xnew[0] = x0
    unew = u.copy()
    has_lims = lims is not None and np.asarray(lims).size > 0     # :22 (no inverted-lims check, Q6)
    for i in range(N):
        if not traj_new.isempty():
            unew[i] = unew[i] + traj_new.k[i] * alpha                        # :18
            dx = diff(xnew[i], x[i])
            unew[i] = unew[i] + traj_new.K[i] @ dx                           # :20
        if has_lims:
            li = lims if np.ndim(lims) == 2 else lims[i]
            unew[i] = _clamp_jl(unew[i], li[:, 0], li[:, 1])                # :23 (NaN survives; f zeroes it)
        xnewi = f(xnew[i], unew[i], i + 1)                                   # :25 (called at i=N too)
        if i < N - 1:
            xnew[i + 1] = xnewi
    cnew = costfun(xnew, unew)                                               # :30
    return xnew, unew, cnew


def forward_covariance(fx, R1, traj: GaussianPolicy):
    """State/control covariance propagation, forward_pass.jl:37-56.

    The reference obtains ``fx`` and ``R1`` from the un-vendored LinearTimeVaryingModelsBase
    (``df(model,x,u)``, ``covariance(model,x,u)``); here they are inputs.
    Returns sigmanew (N, n+m, n+m); the last slice's control blocks stay NaN (`undef`).
    """
    N = traj.T
    n, m = traj.n, traj.m
    sig = np.full((N, n + m, n + m), np.nan)
    sig[0, :n, :n] = R1
    for i in range(N - 1):
        K, S = traj.K[i], traj.Sigma[i]
        fxi = _t(fx, i, 2)
        sig[i + 1, :n, :n] = fxi @ sig[i, :n, :n] @ fxi.T + R1
        sig[i, n:, :n] = K @ sig[i, :n, :n]
        sig[i, :n, n:] = sig[i, :n, :n] @ K.T
        sig[i, n:, n:] = K @ sig[i, :n, :n] @ K.T + S
    return sig


# --------------------------------------------------------------------------------------------
# KL divergence and eta update  (reference: src/klutils.jl:70-130, 155-156)
# --------------------------------------------------------------------------------------------


def _logdet(A):
    sign, ld = np.linalg.slogdet(A)
    if sign <= 0:
        raise ValueError("logdet of a matrix with non-positive determinant")
    return ld


def kl_div_wiki(xnew, xold, sigma_new, traj_new: GaussianPolicy, traj_prev: GaussianPolicy):
    """Per-timestep KL divergence, klutils.jl:70-100.  Returns (T,) clipped at 0."""
    mu = xnew - xold
    T, m, n = traj_new.T, traj_new.m, traj_new.n
    kld = np.zeros(T)
    for t in range(T):
        mut = mu[t]
        St = sigma_new[t, :n, :n]
        Kp, Kn = traj_prev.K[t], traj_new.K[t]
        kp, kn = traj_prev.k[t], traj_new.k[t]
        Sp, Sn = traj_prev.Sigma[t], traj_new.Sigma[t]
        Sip = traj_prev.Sigmai[t]
        kd = kp - kn
        Kd = Kp - Kn
        try:
            v = 0.5 * (np.trace(Sip @ Sn) + kd @ Sip @ kd - m + _logdet(Sp) - _logdet(Sn))
        except ValueError:                     # klutils.jl:92-96: `catch e ... return Inf` (a scalar, for the whole call)
            return np.full(T, np.inf)
        v += 0.5 * (mut @ Kd.T @ Sip @ Kd @ mut + np.trace(Kd.T @ Sip @ Kd @ St))
        v += kd @ Sip @ Kd @ mut
        kld[t] = v
    return np.maximum(0, kld)


def geom(etabracket):                          # klutils.jl:156
    return math.sqrt(etabracket[0] * etabracket[2])


def calc_eta(xnew, xold, sigmanew, etabracket, traj_new, traj_prev, kl_step: float):
    """Scalar-constraint eta bracket update, klutils.jl:110-130.  Mutates ``etabracket``."""
    if not (kl_step > 0):
        return etabracket, True, 0.0
    divergence = float(np.mean(kl_div_wiki(xnew, xold, sigmanew, traj_new, traj_prev)))
    violation = divergence - kl_step
    satisfied = abs(violation) < 0.1 * kl_step
    if not satisfied:
        if violation < 0:                      # eta too big
            etabracket[2] = etabracket[1]
            etabracket[1] = max(geom(etabracket), 0.1 * etabracket[2])
        else:                                  # eta too small
            etabracket[0] = etabracket[1]
            etabracket[1] = min(geom(etabracket), 10.0 * etabracket[0])
    return etabracket, satisfied, divergence


def entropy(traj: GaussianPolicy):             # klutils.jl:104
    return float(np.mean([_logdet(traj.Sigma[t]) / 2 for t in range(traj.T)]) + traj.m * math.log(2 * math.pi) / 2)


# --------------------------------------------------------------------------------------------
# Models (reference: src/demo_linear.jl:5-60, test/test_readme.jl:17-70, src/system_pendcart.jl)
# --------------------------------------------------------------------------------------------


@dataclass
class LinearModel:
    """x+ = A x + B u, cost 0.5 sum x.(Qx) + 0.5 sum u.(Ru); demo_linear.jl:35-50."""

    A: np.ndarray
    B: np.ndarray
    Q: np.ndarray
    R: np.ndarray
    per_step_cost: bool = False            # demo_linear_kl's costf returns a per-step vector (:98)

    def f(self, x, u, i):
        u[np.isnan(u)] = 0                 # demo_linear.jl:42 (in place)
        return self.A @ x + self.B @ u

    def costfun(self, x, u):
        cxs = 0.5 * np.sum(x * (x @ self.Q.T), axis=1)
        cus = 0.5 * np.sum(u * (u @ self.R.T), axis=1)
        return cxs + cus if self.per_step_cost else float(np.sum(cxs) + np.sum(cus))

    def df(self, x, u, time_varying=False):
        u[np.isnan(u)] = 0
        N = u.shape[0]
        n, m = self.B.shape
        cx = x @ self.Q.T
        cu = u @ self.R.T
        cxu = np.zeros((n, m))
        if time_varying:                   # demo_linear_kl: repeat(A,1,1,T)  (:82-86)
            rep = lambda a: np.tile(a, (N, 1, 1))
            return rep(self.A), rep(self.B), None, None, None, cx, cu, rep(self.Q), rep(cxu), rep(self.R)
        return self.A, self.B, None, None, None, cx, cu, self.Q, cxu, self.R


@dataclass
class PendcartModel:
    """Pendulum on a cart, Euler step + ZoH Jacobians; system_pendcart.jl:42-154."""

    g: float = 9.82
    l: float = 0.35
    h: float = 0.01
    d: float = 0.99
    Q: np.ndarray = field(default_factory=lambda: np.diag([10.0, 1.0, 2.0, 1.0]))
    R: float = 1.0
    goal: np.ndarray = field(default_factory=lambda: np.array([math.pi, 0.0, 0.0, 0.0]))

    def f(self, x, u, i):                  # dfsys, system_pendcart.jl:83-89
        u[np.isnan(u)] = 0
        g, l, h, d = self.g, self.l, self.h, self.d
        return np.array([
            x[0] + h * x[1],
            x[1] + h * (-g / l * math.sin(x[0]) + u[0] / l * math.cos(x[0]) - d * x[1]),
            x[2] + h * x[3],
            x[3] + h * u[0],
        ])

    def costfun(self, x, u):               # cost_quadratic(::Matrix), :97-106 -> T+1 entries
        dd = x - self.goal
        T = u.shape[0]
        c = np.zeros(T + 1)
        for t in range(T):
            c[t] = 0.5 * (dd[t] @ self.Q @ dd[t] + self.R * u[t, 0] * u[t, 0])
        c[T] = 0.5 * (dd[-1] @ self.Q @ dd[-1])
        return c

    def df(self, x, u):                    # lin_dyn_df, :137-154
        from scipy.linalg import expm
        u[np.isnan(u)] = 0
        I = u.shape[0]
        D = 4
        g, l, h, d = self.g, self.l, self.h, self.d
        cx = (x[:I] - self.goal) @ self.Q.T
        cu = self.R * u
        fxd = np.zeros((I, D, D))
        fud = np.zeros((I, D, 1))
        for ii in range(I):
            fxc = np.array([[0, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 1], [0, 0, 0, 0]], dtype=np.float64)
            fuc = np.array([0.0, 0.0, 0.0, 1.0])
            fxc[1, 0] = -g / l * math.cos(x[ii, 0]) - u[ii, 0] / l * math.sin(x[ii, 0])
            fxc[1, 1] = -d
            fuc[1] = math.cos(x[ii, 0]) / l
            M = np.zeros((D + 1, D + 1))
            M[:D, :D] = fxc * h
            M[:D, D] = fuc * h
            ABd = expm(M)                  # ZoH sampling :148
            fxd[ii] = ABd[:D, :D]
            fud[ii, :, 0] = ABd[:D, D]
        return fxd, fud, None, None, None, cx, cu, self.Q, np.zeros((D, 1)), np.array([[self.R]])


# --------------------------------------------------------------------------------------------
# iLQG outer loop  (reference: src/iLQG.jl:143-341)
# --------------------------------------------------------------------------------------------

DEFAULT_ALPHA = 10.0 ** np.linspace(0, -3, 11)     # iLQG.jl:145


def _lam_increase(lam, dlam, lamfactor, lammin):
    """``dλ,λ = max(dλ*f, f), max(λ*dλ, λmin)`` -- tuple assignment, λ uses the OLD dλ (quirk Q1)."""
    return max(lam * dlam, lammin), max(dlam * lamfactor, lamfactor)


def iLQG(f, costfun, df, x0, u0, *, lims=None, alpha=None, tol_fun=1e-7, tol_grad=1e-4, max_iter=500,
         lam=1.0, dlam=1.0, lamfactor=1.6, lammax=1e10, lammin=1e-6, regType=1, reduce_ratio_min=0.0,
         diff_fun=lambda a, b: a - b, cost=None, boxqp_hermitian_check=False):
    """Restatement of iLQG.jl:143-341 (printing/plotting/timing omitted).

    Returns ``(x, u, traj_new, Vx, Vxx, cost, trace)`` or ``None`` when the initial control
    sequence diverges (:205-210).  ``trace`` is a dict of (iteration, value) lists plus
    ``trace['status']`` (0 = tol_grad success, 1 = tol_fun success, 2 = lambda > lambdamax,
    3 = max_iter) and ``trace['iters']``.  Raises RuntimeError where the reference calls
    ``error`` (:199, :335).
    """
    alpha = DEFAULT_ALPHA if alpha is None else np.asarray(alpha, dtype=np.float64)
    x0 = np.asarray(x0, dtype=np.float64)
    u = np.array(u0, dtype=np.float64)
    N, m = u.shape
    traj_new = GaussianPolicy.empty()
    trace = {key: [] for key in ("lam", "dlam", "cost", "grad_norm", "alpha", "improvement", "reduce_ratio")}
    trace["lam"].append((0, lam))
    trace["dlam"].append((0, dlam))
    if x0.ndim == 1 or x0.shape[0] == 1 and N != 1:     # only the initial state given (:181)
        x0v = x0.reshape(-1)
        diverge = True
        for a_i in alpha:
            x, un, cost_ = forward_pass(traj_new, x0v, a_i * u, None, 1, f, costfun, lims, diff_fun)
            if np.all(np.abs(x) < 1e8):                 # :187
                u = un
                cost = cost_
                diverge = False
                break
    elif x0.shape[0] == N:                              # pre-rolled (:193)
        x = x0
        x0v = x0[0]
        diverge = False
        if cost is None:
            cost = costfun(x, u)
    else:
        raise RuntimeError("pre-rolled initial trajectory must be of correct length (size(x0,2) == N)")
    if diverge:
        return None
    trace["cost"].append((0, float(np.sum(cost))))
    flg_change = True
    dcost = 0.0
    status = 3
    it = accepted_iter = 1
    Vx = Vxx = dV = None
    g_norm = float("nan")
    while accepted_iter <= max_iter:                    # :222
        reduce_ratio = 0.0
        if flg_change:                                  # STEP 1
            fx, fu, fxx, fxu, fuu, cx, cu, cxx, cxu, cuu = df(x, u)
            flg_change = False
        back_pass_done = False                          # STEP 2
        while not back_pass_done:
            diverge, traj_new, Vx, Vxx, dV = back_pass(cx, cu, cxx, cxu, cuu, fx, fu, lam, regType, lims, x, u,
                                                       boxqp_hermitian_check=boxqp_hermitian_check)
            if diverge > 0:
                lam, dlam = _lam_increase(lam, dlam, lamfactor, lammin)            # :246
                if lam > lammax:
                    break
                continue
            back_pass_done = True
        k = traj_new.k
        g_norm = float(np.mean(np.max(np.abs(k) / (np.abs(u) + 1), axis=1)))       # :256
        trace["grad_norm"].append((it, g_norm))
        if g_norm < tol_grad and lam < 1e-5:                                        # :258
            status = 0
            break
        fwd_pass_done = False                           # STEP 3
        a_i = float("nan")
        if back_pass_done:
            for a_i in alpha:
                xnew, unew, costnew = forward_pass(traj_new, x0v, u, x, a_i, f, costfun, lims, diff_fun)
                dcost = float(np.sum(cost) - np.sum(costnew))
                expected = -a_i * (dV[0] + a_i * dV[1])
                reduce_ratio = dcost / expected if expected > 0 else float(np.sign(dcost))   # :271-276
                if reduce_ratio > reduce_ratio_min:
                    fwd_pass_done = True
                    break
        if fwd_pass_done:                               # STEP 4 accept
            dlam = min(dlam / lamfactor, 1 / lamfactor)                             # :299
            lam = max(lam * dlam, lammin)                                           # :300
            x, u, cost = xnew.copy(), unew.copy(), np.copy(costnew)
            traj_new.k = u.copy()                                                   # :303 (quirk Q11)
            flg_change = True
            if dcost < tol_fun:                                                     # :306 break before trace
                status = 1
                break
            accepted_iter += 1
        else:
            a_i = float("nan")
            lam, dlam = _lam_increase(lam, dlam, lamfactor, lammin)                 # :313
            if lam > lammax:
                status = 2
                break
        trace["lam"].append((it, lam))
        trace["dlam"].append((it, dlam))
        trace["alpha"].append((it, float(a_i)))
        trace["improvement"].append((it, dcost))
        trace["cost"].append((it, float(np.sum(cost))))
        trace["reduce_ratio"].append((it, reduce_ratio))
        it += 1
    if it == 1:                                          # :335 (quirk Q5)
        raise RuntimeError("Failure: no iterations completed, something is wrong.")
    trace["status"] = status
    trace["iters"] = it
    trace["lam_final"] = lam
    trace["dlam_final"] = dlam
    trace["g_norm"] = g_norm
    return x, u, traj_new, Vx, Vxx, cost, trace


# --------------------------------------------------------------------------------------------
# iLQGkl outer loop, single-KL-constraint branch  (reference: src/iLQGkl.jl:25-183, 238-252)
# --------------------------------------------------------------------------------------------


def iLQGkl(dynamics, costfun, derivs, x0, traj_prev: GaussianPolicy, fx_model, R1, *, kl_step=1.0, lims=None,
           max_iter=50, etabracket=(1e-8, 1.0, 1e16), del0=1e-4, cost=None, diff_fun=lambda a, b: a - b,
           max_eta_retries=200):
    """Restatement of iLQGkl.jl:25-183 + 238-252 (``constrain_per_step=false``).

    ``fx_model`` and ``R1`` replace the un-vendored ``model`` argument: they are what
    ``df(model,x,u)`` / ``covariance(model,x,u)`` return (forward_pass.jl:38,42).
    ``max_eta_retries`` bounds the reference's unbounded eta-retry loop (quirk Q9) for safety.
    """
    u = traj_prev.k.copy()                               # :47
    x0 = np.asarray(x0, dtype=np.float64)
    N, m = u.shape
    k_old = traj_prev.k.copy()
    traj_prev.k = traj_prev.k * 0                        # :52
    etabracket = np.array(etabracket, dtype=np.float64)  # :53 copy
    if x0.ndim != 2 or x0.shape[0] != N:
        raise RuntimeError("pre-rolled initial trajectory must be of correct length (size(x0,2) == N)")
    x = x0
    if cost is None:
        raise RuntimeError("Initial trajectory supplied, initial cost must also be supplied")
    trace = {key: [] for key in ("cost", "grad_norm", "improvement", "reduce_ratio", "divergence", "eta")}
    trace["cost"].append((0, float(np.sum(cost))))
    fx, fu, fxx, fxu, fuu, cx, cu, cxx, cxu, cuu = derivs(x, u)       # :88 once (quirk Q9)
    kl_cost_terms = (grad_kl(traj_prev), etabracket)                  # :92
    satisfied = False
    divergence = 0.0
    xnew = unew = costnew = traj_new = Vx = Vxx = None
    it = 0
    for it in range(1, max_iter + 1):                                 # :93
        diverge = 1
        retries = 0
        while diverge > 0:                                            # :97
            diverge, traj_new, Vx, Vxx, dV = back_pass_gps(cx, cu, cxx, cxu, cuu, fx, fu, lims, x, u, kl_cost_terms)
            if diverge > 0:
                etabracket[1] += del0                                 # :104
                del0 *= 2
                retries += 1
                if retries > max_eta_retries:
                    raise RuntimeError("eta retry loop did not terminate")
        g_norm = float(np.mean(np.max(np.abs(traj_new.k) / (np.abs(u) + 1), axis=1)))   # :127
        trace["grad_norm"].append((it, g_norm))
        xnew, unew, costnew = forward_pass(traj_new, x0[0], u, x, 1, dynamics, costfun, lims, diff_fun)   # :134
        sigmanew = forward_covariance(fx_model, R1, traj_new)         # :135
        traj_new.k = traj_new.k + traj_prev.k                         # :136
        dcost = float(np.sum(cost) - np.sum(costnew))
        expected = -(dV[0] + dV[1])                                   # :138
        reduce_ratio = dcost / expected
        etabracket, satisfied, divergence = calc_eta(xnew, x, sigmanew, etabracket, traj_new, traj_prev, kl_step)  # :143
        trace["improvement"].append((it, dcost))
        trace["cost"].append((it, float(np.sum(costnew))))
        trace["reduce_ratio"].append((it, reduce_ratio))
        trace["divergence"].append((it, float(np.mean(divergence))))
        trace["eta"].append((it, float(etabracket[1])))
        if satisfied:                                                 # :173
            break
        if etabracket[1] > 0.999 * etabracket[2]:                     # :178
            break
    x, u, cost = xnew, unew, costnew                                  # :240
    traj_new.k = u.copy()                                             # :241
    traj_prev.k = k_old                                               # :247
    trace["iters"] = it
    trace["satisfied"] = satisfied
    trace["etabracket"] = etabracket
    return x, u, traj_new, Vx, Vxx, cost, trace
