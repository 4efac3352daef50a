"""Host-side mirror of the reference's method surface, bound to libddp.so through ctypes.

The reference is Julia (``iLQG``, ``iLQGkl``, ``back_pass``, ``back_pass_gps``, ``forward_pass``,
``boxQP``, ``GaussianPolicy`` -- src/DifferentialDynamicProgramming.jl:6); Julia is not available in
the build container, so this Python layer carries the same names, argument order/meaning and
error behaviour and is what the parity tests drive.  ``julia/DifferentialDynamicProgramming.jl``
is the equivalent ``ccall`` shim.  Nothing here computes: every call lands in a CUDA kernel, and
without libddp.so / a GPU the calls raise.

Array conventions (identical to ``oracle/``): time-first *math layout*, optional leading batch axis.
    cx (N,n) | (B,N,n)          fx (n,n) | (N,n,n) | (B, 1|N, n,n)       K (N,m,n) | (B,N,m,n)
A matrix tensor with 2 dims is shared and time-invariant, with 3 dims time-varying and shared
across the batch, with 4 dims ``(B|1, N|1, r, c)``.  The device layout is the reference's
column-major ``(r,c,T,B)`` (== C order ``[B][T][c][r]``); packing transposes the last two axes.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib as L

# ---------------------------------------------------------------------------------------------
# engine / device memory
# ---------------------------------------------------------------------------------------------


class DevArray:
    """A device buffer owned by an Engine (ddp_malloc / ddp_free)."""

    def __init__(self, eng: "Engine", shape, dtype=np.float64):
        self.eng = eng
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        eng._ck(eng.lib.ddp_malloc(eng.h, C.byref(p), max(self.nbytes, 8)))
        self.ptr = p.value

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        if self.nbytes:
            self.eng._ck(self.eng.lib.ddp_download(self.eng.h, out.ctypes.data, self.ptr, self.nbytes))
        return out

    def set(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        if self.nbytes:
            self.eng._ck(self.eng.lib.ddp_upload(self.eng.h, self.ptr, a.ctypes.data, self.nbytes))
        return self

    def zero(self):
        self.eng._ck(self.eng.lib.ddp_memset(self.eng.h, self.ptr, 0, self.nbytes))
        return self

    def free(self):
        if self.ptr is not None and self.eng.h is not None:
            self.eng.lib.ddp_free(self.eng.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One libddp handle: fixed problem size (n, m, T, B) on one GPU, one stream."""

    def __init__(self, n: int, m: int, T: int, B: int = 1, device: int = 0, force_generic: bool = False):
        self.lib = L.load()
        self.h = None
        h = C.c_void_p()
        rc = self.lib.ddp_create(C.byref(h), device, n, m, T, B, 1 if force_generic else 0)
        if rc != 0:
            raise L.DDPError(rc, self.lib.ddp_last_error(None).decode())
        self.h = h
        self.n, self.m, self.T, self.B, self.device = n, m, T, B, device

    def _ck(self, rc: int):
        if rc != 0:
            raise L.DDPError(rc, self.lib.ddp_last_error(self.h).decode())

    def close(self):
        if self.h is not None:
            self.lib.ddp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._ck(self.lib.ddp_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.lib.ddp_synchronize(self.h))

    @property
    def kernel_variant(self) -> str:
        return self.lib.ddp_kernel_variant(self.h).decode()

    @property
    def launch_count(self) -> int:
        return int(self.lib.ddp_launch_count(self.h))

    # ---- multi-GPU: the statistics all-reduce inside the library (ddp_comm_*; NCCL resolved at run time)
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.lib.ddp_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        self._ck(self.lib.ddp_comm_init(self.h, nranks, rank, C.create_string_buffer(unique_id, 128)))

    def allreduce_stats(self, stats_ptr: int):
        self._ck(self.lib.ddp_comm_allreduce_stats_f64(self.h, C.c_void_p(stats_ptr)))

    def selftest_peak(self, kind: str = "dmma", reps: int = 3):
        """FP64 peak of this device in TFLOP/s measured by the library's own microkernels (ddp_selftest_peak_f64):
        ``"dfma"`` (FMA pipe), ``"dmma"`` (mma.sync.m8n8k4.f64, the instruction of the n=32, m=8 sweeps) or ``"mixed"`` (both interleaved,
        all flops counted).  Returns (TFLOP/s, ms)."""
        tf, ms = C.c_double(0.0), C.c_double(0.0)
        self._ck(self.lib.ddp_selftest_peak_f64(self.h, {"dfma": 0, "dmma": 1, "mixed": 2}[kind], reps, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def empty(self, shape, dtype=np.float64) -> DevArray:
        return DevArray(self, shape, dtype)

    def upload(self, a: np.ndarray, dtype=np.float64) -> DevArray:
        a = np.ascontiguousarray(a, dtype=dtype)
        return DevArray(self, a.shape, dtype).set(a)


def _tensor(ptr: Optional[int], sb: int = 0, st: int = 0) -> L.Tensor:
    return L.Tensor(ptr, sb, st)


def _pack_vec(eng: Engine, a, B: int, N: int, d: int, name: str):
    """(N,d) | (B,N,d) -> device (d,T,B) + strides."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2:
        a = a[None]
    if a.shape[1:] != (N, d) or a.shape[0] not in (1, B):
        raise ValueError(f"size({name}) should be ({d}, {N}) per trajectory, got {a.shape}")
    dev = eng.upload(a)
    return dev, _tensor(dev.ptr, N * d if a.shape[0] == B else 0, d)


def _pack_mat(eng: Engine, a, B: int, N: int, r: int, c: int, name: str):
    """(r,c) | (N,r,c) | (B|1,N|1,r,c) math layout -> device column-major + strides."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2:
        a = a[None, None]
    elif a.ndim == 3:
        a = a[None]
    if a.ndim != 4 or a.shape[2:] != (r, c) or a.shape[0] not in (1, B) or a.shape[1] not in (1, N):
        raise ValueError(f"size({name}) should be ({r}, {c}[, {N}]), got {a.shape}")
    dev = eng.upload(np.swapaxes(a, -1, -2))
    Nx = a.shape[1]
    st = r * c if (Nx == N and N > 1) else 0
    sb = Nx * r * c if (a.shape[0] == B and B > 1) else 0
    return dev, _tensor(dev.ptr, sb, st)


def _lims_dev(eng: Engine, lims, m: int):
    """(m,2) -> device [lower(m); upper(m)]; (N,m,2) time-varying -> device (N,2,m) blocks.  Returns (DevArray|None, stride_t)."""
    if lims is None or np.asarray(lims).size == 0:
        return None, 0
    lims = np.asarray(lims, dtype=np.float64)
    if lims.ndim == 3:
        return eng.upload(np.ascontiguousarray(np.swapaxes(lims.reshape(-1, m, 2), -1, -2))), 2 * m
    lims = lims.reshape(m, 2)
    return eng.upload(np.ascontiguousarray(lims.T)), 0


# ---------------------------------------------------------------------------------------------
# GaussianPolicy  (iLQG.jl:39-53)
# ---------------------------------------------------------------------------------------------


@dataclass
class GaussianPolicy:
    T: int = 0
    n: int = 0
    m: int = 0
    K: Optional[np.ndarray] = None
    k: Optional[np.ndarray] = None
    Sigma: Optional[np.ndarray] = None
    Sigmai: Optional[np.ndarray] = None

    @staticmethod
    def empty() -> "GaussianPolicy":
        return GaussianPolicy()

    @staticmethod
    def identity(T: int, n: int, m: int) -> "GaussianPolicy":
        eye = np.tile(np.eye(m), (T, 1, 1))
        return GaussianPolicy(T, n, m, np.zeros((T, m, n)), np.zeros((T, m)), eye.copy(), eye.copy())

    def isempty(self) -> bool:
        return self.T == 0 and self.n == 0 and self.m == 0

    def __len__(self) -> int:
        return self.T


# ---------------------------------------------------------------------------------------------
# back_pass / back_pass_gps
# ---------------------------------------------------------------------------------------------


def _back_common(cx, cu, cxx, cxu, cuu, fx, fu, lims, u, force_generic, engine):
    cx = np.asarray(cx, dtype=np.float64)
    cu = np.asarray(cu, dtype=np.float64)
    batched = cx.ndim == 3
    B = cx.shape[0] if batched else 1
    N, n = cx.shape[-2:]
    m = cu.shape[-1]
    if cu.shape[-2] != N:
        raise ValueError("size(cu) should be (m, N)")
    eng = engine or Engine(n, m, N, B, force_generic=force_generic)
    keep = []
    a = L.BackPassArgs()
    for name, arr, d in (("cx", cx, n), ("cu", cu, m)):
        dev, t = _pack_vec(eng, arr, B, N, d, name)
        keep.append(dev)
        setattr(a, name, t)
    for name, arr, r, c in (("cxx", cxx, n, n), ("cxu", cxu, n, m), ("cuu", cuu, m, m), ("fx", fx, n, n), ("fu", fu, n, m)):
        dev, t = _pack_mat(eng, arr, B, N, r, c, name)
        keep.append(dev)
        setattr(a, name, t)
    ld, lst = _lims_dev(eng, lims, m)
    if ld is not None:
        keep.append(ld)
        a.lims = ld.ptr
        a.lims_stride_t = lst
        dev, t = _pack_vec(eng, u, B, N, m, "u")
        keep.append(dev)
        a.u = t
    return eng, a, keep, batched, B, N, n, m


def _pack_tens3(eng: Engine, a, B: int, N: int, d1: int, d2: int, d3: int, name: str):
    """Second-order dynamics tensor, math layout (d1,d2,d3) | (N,d1,d2,d3) | (B|1,N|1,d1,d2,d3) with the reference's index
    order (size(fxu) == (n,n,m[,N]), iLQG.jl:80) -> device column-major (first index fastest) + strides."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 3:
        a = a[None, None]
    elif a.ndim == 4:
        a = a[None]
    if a.ndim != 5 or a.shape[2:] != (d1, d2, d3) or a.shape[0] not in (1, B) or a.shape[1] not in (1, N):
        raise ValueError(f"size({name}) should be ({d1}, {d2}, {d3}[, {N}]), got {a.shape}")
    dev = eng.upload(np.ascontiguousarray(np.transpose(a, (0, 1, 4, 3, 2))))
    Nx = a.shape[1]
    blk = d1 * d2 * d3
    return dev, _tensor(dev.ptr, Nx * blk if (a.shape[0] == B and B > 1) else 0, blk if (Nx == N and N > 1) else 0)


def _unpack_tri(tri: np.ndarray, d: int) -> np.ndarray:
    """(..., d(d+1)/2) packed upper triangle (column by column) -> (..., d, d) symmetric."""
    r, c = np.triu_indices(d)
    order = np.argsort(c * (c + 1) // 2 + r, kind="stable")
    r, c = r[order], c[order]
    full = np.zeros(tri.shape[:-1] + (d, d))
    full[..., r, c] = tri
    full[..., c, r] = tri
    return full


def _back_outputs(eng, a, B, N, n, m, want_Vxx, want_Quu):
    out = dict(diverge=eng.empty((B,), np.int32), K=eng.empty((B, N, n, m)), k=eng.empty((B, N, m)),
               Vx=eng.empty((B, N, n)), dV=eng.empty((B, 2)), Vxx1=eng.empty((B, n, n)))
    if want_Vxx == "upper":                      # packed upper-triangle history: half the bytes
        out["Vxx_tri"] = eng.empty((B, N, n * (n + 1) // 2))
    elif want_Vxx:
        out["Vxx"] = eng.empty((B, N, n, n))
    if want_Quu:
        out["Quu"] = eng.empty((B, N, m, m))
    for key, dev in out.items():
        setattr(a, key, dev.ptr)
    return out


def back_pass(cx, cu, cxx, cxu, cuu, fx, fu, *rest, want_Vxx=True, force_generic=False, engine: Engine = None):
    """``back_pass(cx,cu,cxx,cxu,cuu,fx,fu,λ,regType,lims,x,u)`` -- backward_pass.jl:162-252, or the 15-argument form
    ``back_pass(cx,cu,cxx,cxu,cuu,fx,fu,fxx,fxu,fuu,λ,regType,lims,x,u)`` (backward_pass.jl:81-160) whose second-order
    tensors ``fxx (n,n,n[,N])``, ``fxu (n,n,m[,N])``, ``fuu (n,m,m[,N])`` may each be ``None``/empty.

    Returns ``(diverge, GaussianPolicy, Vx, Vxx, dV)``; with a leading batch axis on the inputs the
    outputs carry it too (``diverge`` becomes an int array, the policy holds batched arrays).
    ``want_Vxx="upper"`` stores the history as packed upper triangles on the device (half the bytes) and expands it here;
    ``lims`` may be ``(m,2)`` or time-varying ``(N,m,2)``.
    """
    if len(rest) == 8:
        fxx, fxu, fuu, lam, regType, lims, x, u = rest
    elif len(rest) == 5:
        fxx = fxu = fuu = None
        lam, regType, lims, x, u = rest
    else:
        raise TypeError("back_pass takes 12 or 15 positional arguments")
    eng, a, keep, batched, B, N, n, m = _back_common(cx, cu, cxx, cxu, cuu, fx, fu, lims, u, force_generic, engine)
    for name, arr, dims in (("fxx", fxx, (n, n, n)), ("fxu", fxu, (n, n, m)), ("fuu", fuu, (n, m, m))):
        if arr is not None and np.asarray(arr).size > 0:
            dev, t = _pack_tens3(eng, arr, B, N, *dims, name)
            keep.append(dev)
            setattr(a, name, t)
    lam_dev = eng.upload(np.broadcast_to(np.asarray(lam, dtype=np.float64), (B,)))
    a.lam = lam_dev.ptr
    a.reg_type = int(regType)
    out = _back_outputs(eng, a, B, N, n, m, want_Vxx, True)
    eng._ck(eng.lib.ddp_back_pass_f64(eng.h, C.byref(a)))
    eng.synchronize()
    res = _collect_back(out, batched, B, N, n, m, None)
    del keep
    return res


def _collect_back(out, batched, B, N, n, m, Quui):
    diverge = out["diverge"].numpy()
    K = np.swapaxes(out["K"].numpy(), -1, -2)                # (B,N,n,m) col-major -> (B,N,m,n)
    k = out["k"].numpy()
    Vx = out["Vx"].numpy()
    if "Vxx_tri" in out:
        Vxx = _unpack_tri(out["Vxx_tri"].numpy(), n)
    else:
        Vxx = np.swapaxes(out["Vxx"].numpy(), -1, -2) if "Vxx" in out else np.swapaxes(out["Vxx1"].numpy(), -1, -2)
    Quu = np.swapaxes(out["Quu"].numpy(), -1, -2) if "Quu" in out else None
    dV = out["dV"].numpy()
    Sig = np.swapaxes(Quui.numpy(), -1, -2) if Quui is not None else None
    if not batched:
        pol = GaussianPolicy(N, n, m, K[0], k[0], None if Sig is None else Sig[0], None if Quu is None else Quu[0])
        return int(diverge[0]), pol, Vx[0], Vxx[0], dV[0]
    return diverge, GaussianPolicy(N, n, m, K, k, Sig, Quu), Vx, Vxx, dV


def back_pass_gps(cx, cu, cxx, cxu, cuu, fx, fu, lims, x, u, kl_cost_terms, *, force_generic=False, engine: Engine = None):
    """``back_pass_gps(cx,cu,cxx,cxu,cuu,fx,fu,lims,x,u,kl_cost_terms)`` -- backward_pass.jl:259-350.

    ``kl_cost_terms = (traj_prev, ηbracket)``: the KL terms of ``∇kl`` (klutils.jl:8-23) are formed
    on the device from the previous policy, so the policy itself is passed instead of the five
    pre-multiplied tensors.  ``ηbracket`` is ``(3,)`` or ``(B,3)``.
    """
    traj_prev, etabracket = kl_cost_terms
    eng, a, keep, batched, B, N, n, m = _back_common(cx, cu, cxx, cxu, cuu, fx, fu, lims, u, force_generic, engine)
    eta = np.asarray(etabracket, dtype=np.float64)
    eta = np.broadcast_to(eta.reshape(-1, 3)[:, 1], (B,)) if eta.size >= 3 else np.broadcast_to(eta, (B,))
    eta_dev = eng.upload(eta)
    g = L.GpsArgs()
    Kp, g.K_prev = _pack_mat(eng, traj_prev.K if batched else traj_prev.K[None], B, N, m, n, "K_prev")
    Sp, g.Sigi_prev = _pack_mat(eng, traj_prev.Sigmai if batched else traj_prev.Sigmai[None], B, N, m, m, "Sigi_prev")
    kp, g.k_prev = _pack_vec(eng, traj_prev.k, B, N, m, "k_prev")
    g.eta = eta_dev.ptr
    out = _back_outputs(eng, a, B, N, n, m, True, True)
    Quui = eng.empty((B, N, m, m))
    g.Quui = Quui.ptr
    eng._ck(eng.lib.ddp_back_pass_gps_f64(eng.h, C.byref(a), C.byref(g)))
    eng.synchronize()
    res = _collect_back(out, batched, B, N, n, m, Quui)
    del keep, Kp, Sp, kp
    return res


# ---------------------------------------------------------------------------------------------
# boxQP
# ---------------------------------------------------------------------------------------------


class PosDefException(Exception):
    pass


def boxQP(H, g, lower, upper, x0, *, maxIter=100, minGrad=1e-8, minRelImprove=1e-8, stepDec=0.6, minStep=1e-22,
          Armijo=0.1, engine: Engine = None):
    """``boxQP(H,g,lower,upper,x0)`` -- boxQP.jl:29-188.  Returns ``(x, result, Hfree, free, nfactor)``.

    Unbatched calls raise :class:`PosDefException` where the reference's ``cholesky`` throws;
    batched calls (H of shape (B,m,m)) report ``result == -1`` for those problems instead.
    """
    H = np.asarray(H, dtype=np.float64)
    batched = H.ndim == 3
    Hb = H if batched else H[None]
    B, m, _ = Hb.shape
    if m > 16:                                            # demoQP-sized problems: one CTA per QP (ddp_boxqp_large_f64)
        return _boxQP_large(Hb, g, lower, upper, x0, batched, L.BoxQPOpts(maxIter, minGrad, minRelImprove, stepDec, minStep, Armijo), engine)
    eng = engine or Engine(max(m, 1), m, 1, B)
    vec = lambda v: eng.upload(np.broadcast_to(np.asarray(v, dtype=np.float64).reshape(-1, m), (B, m)))
    dH = eng.upload(np.swapaxes(Hb, -1, -2))
    dg, dl, du, dx0 = vec(g), vec(lower), vec(upper), vec(x0)
    x, res, Hf = eng.empty((B, m)), eng.empty((B,), np.int32), eng.empty((B, m, m))
    fm, nf = eng.empty((B,), np.uint32), eng.empty((B,), np.int32)
    o = L.BoxQPOpts(maxIter, minGrad, minRelImprove, stepDec, minStep, Armijo)
    eng._ck(eng.lib.ddp_boxqp_f64(eng.h, B, dH.ptr, dg.ptr, dl.ptr, du.ptr, dx0.ptr, C.byref(o), x.ptr, res.ptr, Hf.ptr,
                                  fm.ptr, nf.ptr))
    eng.synchronize()
    xs, rs, Hfs, fms, nfs = x.numpy(), res.numpy(), np.swapaxes(Hf.numpy(), -1, -2), fm.numpy(), nf.numpy()
    free = ((fms[:, None] >> np.arange(m)[None, :]) & 1).astype(bool)
    if batched:
        return xs, rs, Hfs, free, nfs
    if rs[0] < 0:
        raise PosDefException("matrix is not positive definite; Cholesky factorization failed")
    nfree = int(np.count_nonzero(np.diag(Hfs[0])))
    return xs[0], int(rs[0]), Hfs[0][:nfree, :nfree], free[0], int(nfs[0])


def _boxQP_large(Hb, g, lower, upper, x0, batched, o, engine):
    B, n, _ = Hb.shape
    eng = engine or Engine(1, 1, 1, 1)
    vec = lambda v: eng.upload(np.broadcast_to(np.asarray(v, dtype=np.float64).reshape(-1, n), (B, n)))
    dH = eng.upload(np.swapaxes(Hb, -1, -2))
    dg, dl, du, dx0 = vec(g), vec(lower), vec(upper), vec(x0)
    x, res, Hf = eng.empty((B, n)), eng.empty((B,), np.int32), eng.empty((B, n, n))
    fr, nf = eng.empty((B, n), np.uint8), eng.empty((B,), np.int32)
    eng._ck(eng.lib.ddp_boxqp_large_f64(eng.h, n, B, dH.ptr, dg.ptr, dl.ptr, du.ptr, dx0.ptr, C.byref(o), x.ptr, res.ptr, Hf.ptr, fr.ptr, nf.ptr))
    eng.synchronize()
    xs, rs, Hfs, free, nfs = x.numpy(), res.numpy(), np.swapaxes(Hf.numpy(), -1, -2), fr.numpy().astype(bool), nf.numpy()
    if batched:
        return xs, rs, Hfs, free, nfs
    if rs[0] < 0:
        raise PosDefException("matrix is not positive definite; Cholesky factorization failed")
    nfree = int(np.count_nonzero(np.diag(Hfs[0])))
    return xs[0], int(rs[0]), Hfs[0][:nfree, :nfree], free[0], int(nfs[0])


def demoQP(n=500, seed=None, **kw):
    """``demoQP`` of boxQP.jl:190-199: a random n = 500 box QP (``H = A*A'``, bounds +-1) solved on the device."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal(n)
    A = rng.standard_normal((n, n))
    return boxQP(A @ A.T, g, -np.ones(n), np.ones(n), rng.standard_normal(n), **kw)


# ---------------------------------------------------------------------------------------------
# models: the reference's user callbacks (f, costfun, df) as device descriptors
# ---------------------------------------------------------------------------------------------


class _DeviceCallback:
    """Stands where the reference takes a Julia closure.  It cannot be called on the host."""

    def __init__(self, model, role):
        self.model, self.role = model, role

    def __call__(self, *a, **k):
        raise RuntimeError(f"{type(self.model).__name__}.{self.role} is a device model descriptor; it is "
                           "evaluated inside the CUDA kernels (there is no CPU fallback)")


class LinearModel:
    """x+ = A x + B u, cost ½Σx'Qx + ½Σu'Ru  (demo_linear.jl:35-50).  A: (n,n)|(N,n,n)|(B,1|N,n,n)."""

    kind = 1
    terminal_cost = 0

    def __init__(self, A, B, Q, R):
        self.A, self.B, self.Q, self.R = (np.asarray(v, dtype=np.float64) for v in (A, B, Q, R))
        self.goal = None
        self.f, self.costfun, self.df = (_DeviceCallback(self, r) for r in ("f", "costfun", "df"))


class PendcartModel:
    """Pendulum on a cart, Euler step (system_pendcart.jl:51-54,83-106)."""

    kind = 2
    terminal_cost = 1

    def __init__(self, g=9.82, l=0.35, h=0.01, d=0.99, Q=None, R=1.0, goal=None):
        self.p = (g, l, h, d)
        self.Q = np.diag([10.0, 1.0, 2.0, 1.0]) if Q is None else np.asarray(Q, dtype=np.float64)
        self.R = np.atleast_2d(np.asarray(R, dtype=np.float64))
        self.goal = np.array([math.pi, 0.0, 0.0, 0.0]) if goal is None else np.asarray(goal, dtype=np.float64)
        self.f, self.costfun, self.df = (_DeviceCallback(self, r) for r in ("f", "costfun", "df"))


def _model_of(f, costfun=None):
    model = getattr(f, "model", None)
    if not isinstance(f, _DeviceCallback) or (costfun is not None and getattr(costfun, "model", None) is not model):
        raise TypeError("f/costfun must be the callbacks of a device model descriptor (LinearModel, PendcartModel): "
                        "arbitrary host closures cannot run on the GPU and there is no CPU fallback")
    return model


def _pack_model(eng: Engine, model, B, N, n, m):
    keep = []
    M = L.Model()
    M.kind = model.kind
    M.terminal_cost = model.terminal_cost
    Qh = np.asarray(model.Q, dtype=np.float64)
    M.flags = 1 if (Qh.ndim == 2 and np.count_nonzero(Qh - np.diag(np.diagonal(Qh))) == 0) else 0   # isdiag(Q)
    if model.kind == 1:
        dev, M.A = _pack_mat(eng, model.A, B, N, n, n, "A"); keep.append(dev)
        dev, M.Bm = _pack_mat(eng, model.B, B, N, n, m, "B"); keep.append(dev)
    else:
        for i, v in enumerate(model.p):
            M.p[i] = v
    dev, M.Q = _pack_mat(eng, model.Q, B, 1, n, n, "Q"); keep.append(dev)
    dev, M.R = _pack_mat(eng, model.R, B, 1, m, m, "R"); keep.append(dev)
    if model.goal is not None:
        gd = eng.upload(model.goal); keep.append(gd)
        M.goal = gd.ptr
    return M, keep


# ---------------------------------------------------------------------------------------------
# forward_pass
# ---------------------------------------------------------------------------------------------


def forward_pass(traj_new: GaussianPolicy, x0, u, x, alpha, f, costfun, lims, diff=None, *, u_scale=1.0,
                 per_step_cost=False, want_derivs=False, force_generic=False, engine: Engine = None):
    """``forward_pass(traj_new,x0,u,x,α,f,costfun,lims,diff)`` -- forward_pass.jl:9-33.

    ``f``/``costfun`` are the callbacks of a device model descriptor; ``diff`` must be the default
    ``-``.  Returns ``(xnew, unew, cnew)`` (+ ``(cx, cu)`` with ``want_derivs``); ``cnew`` is the
    total cost (per-step vector with ``per_step_cost``).
    """
    if diff is not None:
        raise NotImplementedError("only the default diff_fun (-) is supported on the device")
    model = _model_of(f, costfun)
    u = np.asarray(u, dtype=np.float64)
    batched = u.ndim == 3
    B = u.shape[0] if batched else 1
    N, m = u.shape[-2:]
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.shape[-1]
    eng = engine or Engine(n, m, N, B, force_generic=force_generic)
    M, keep = _pack_model(eng, model, B, N, n, m)
    a = L.ForwardPassArgs()
    x0b = np.broadcast_to(x0.reshape(-1, n), (B, n)) if x0.ndim <= 1 or x0.shape[0] != B else x0
    dx0 = eng.upload(x0b); keep.append(dx0)
    a.x0 = _tensor(dx0.ptr, n, 0)
    du, a.u = _pack_vec(eng, u, B, N, m, "u"); keep.append(du)
    if traj_new is not None and not traj_new.isempty():
        K = np.asarray(traj_new.K, dtype=np.float64)
        dK = eng.upload(np.swapaxes(K.reshape(B, N, m, n), -1, -2)); keep.append(dK)
        dk = eng.upload(np.asarray(traj_new.k, dtype=np.float64).reshape(B, N, m)); keep.append(dk)
        a.K, a.k = dK.ptr, dk.ptr
        dx, a.x = _pack_vec(eng, x, B, N, n, "x"); keep.append(dx)
    alpha = np.asarray(alpha, dtype=np.float64)
    if alpha.ndim == 0:
        a.alpha_scalar = float(alpha)
    else:
        dal = eng.upload(np.broadcast_to(alpha, (B,))); keep.append(dal)
        a.alpha = dal.ptr
    a.u_scale = float(u_scale)
    ld, lst = _lims_dev(eng, lims, m)
    if ld is not None:
        keep.append(ld)
        a.lims = ld.ptr
        a.lims_stride_t = lst
    xnew, unew, cost = eng.empty((B, N, n)), eng.empty((B, N, m)), eng.empty((B,))
    a.xnew, a.unew, a.cost = xnew.ptr, unew.ptr, cost.ptr
    Tc = N + model.terminal_cost
    cost_t = eng.empty((B, Tc)) if per_step_cost else None
    if cost_t is not None:
        a.cost_t = cost_t.ptr
    cxo = cuo = None
    if want_derivs:
        cxo, cuo = eng.empty((B, N, n)), eng.empty((B, N, m))
        a.cx, a.cu = cxo.ptr, cuo.ptr
    eng._ck(eng.lib.ddp_forward_pass_f64(eng.h, C.byref(M), C.byref(a)))
    eng.synchronize()
    xn, un = xnew.numpy(), unew.numpy()
    cn = cost_t.numpy() if per_step_cost else cost.numpy()
    if not batched:
        xn, un, cn = xn[0], un[0], (cn[0] if per_step_cost else float(cn[0]))
    if want_derivs:
        cxn, cun = cxo.numpy(), cuo.numpy()
        return xn, un, cn, ((cxn, cun) if batched else (cxn[0], cun[0]))
    return xn, un, cn


def forward_costs(traj_new: GaussianPolicy, x0, u, x, alphas, f, costfun, lims, *, force_generic=False, engine: Engine = None):
    """Total cost of ``forward_pass(traj_new,x0,u,x,α,…)`` for every ``α`` in ``alphas`` in one call
    (``ddp_forward_costs_multi_f64``): the serial backtracking of iLQG.jl:267-281 with the policy gains read once.
    Returns an array ``(len(alphas),)`` / ``(len(alphas), B)``."""
    model = _model_of(f, costfun)
    u = np.asarray(u, dtype=np.float64)
    batched = u.ndim == 3
    B = u.shape[0] if batched else 1
    N, m = u.shape[-2:]
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.shape[-1]
    eng = engine or Engine(n, m, N, B, force_generic=force_generic)
    M, keep = _pack_model(eng, model, B, N, n, m)
    a = L.ForwardPassArgs()
    x0b = np.broadcast_to(x0.reshape(-1, n), (B, n)) if x0.ndim <= 1 or x0.shape[0] != B else x0
    dx0 = eng.upload(x0b); keep.append(dx0)
    a.x0 = _tensor(dx0.ptr, n, 0)
    du, a.u = _pack_vec(eng, u, B, N, m, "u"); keep.append(du)
    dK = eng.upload(np.swapaxes(np.asarray(traj_new.K, dtype=np.float64).reshape(B, N, m, n), -1, -2)); keep.append(dK)
    dk = eng.upload(np.asarray(traj_new.k, dtype=np.float64).reshape(B, N, m)); keep.append(dk)
    a.K, a.k = dK.ptr, dk.ptr
    dx, a.x = _pack_vec(eng, x, B, N, n, "x"); keep.append(dx)
    a.u_scale = 1.0
    ld, lst = _lims_dev(eng, lims, m)
    if ld is not None:
        keep.append(ld)
        a.lims = ld.ptr
        a.lims_stride_t = lst
    xnew, unew, cost = eng.empty((B, N, n)), eng.empty((B, N, m)), eng.empty((B,))
    a.xnew, a.unew, a.cost = xnew.ptr, unew.ptr, cost.ptr
    al = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64).reshape(-1))
    out = eng.empty((len(al), B))
    eng._ck(eng.lib.ddp_forward_costs_multi_f64(eng.h, C.byref(M), C.byref(a), len(al), al.ctypes.data_as(L.c_double_p), out.ptr))
    eng.synchronize()
    res = out.numpy()
    del keep
    return res if batched else res[:, 0]


# ---------------------------------------------------------------------------------------------
# KL divergence (forward_covariance + kl_div_wiki)
# ---------------------------------------------------------------------------------------------


def kl_div_wiki(xnew, xold, fx, R1, traj_new: GaussianPolicy, traj_prev: GaussianPolicy, *, engine: Engine = None,
                Sx_cache=None, Sx_mode: int = 0, Sx_count: int = 0):
    """Per-step KL divergence between the new and previous policy (klutils.jl:70-100) with the state
    covariance of ``forward_covariance`` (forward_pass.jl:37-56) propagated on the device.

    The reference takes ``Σ_new`` from ``forward_covariance(model, x, u, traj_new)``, whose ``fx`` and
    ``R1`` come from the un-vendored LinearTimeVaryingModelsBase; here they are arguments.
    Returns ``(kl_t, kl_mean)``.

    ``Sx_cache`` (a device buffer of ``engine.empty((B, N, 528))``) with ``Sx_mode`` 1 stores the state covariances of this call
    (they depend on ``fx`` and ``R1`` only), ``Sx_mode`` 2 reads them back instead of propagating (``ddp_kl_args.Sx_tri``);
    ``Sx_count`` > 0: the buffer holds the first ``Sx_count`` trajectories only, the others are propagated in every call.
    """
    xnew = np.asarray(xnew, dtype=np.float64)
    batched = xnew.ndim == 3
    B = xnew.shape[0] if batched else 1
    N, n = xnew.shape[-2:]
    m = traj_new.m
    eng = engine or Engine(n, m, N, B)
    keep = []
    a = L.KlArgs()
    dev, a.fx = _pack_mat(eng, fx, B, N, n, n, "fx"); keep.append(dev)
    dev, a.R1 = _pack_mat(eng, R1, B, 1, n, n, "R1"); keep.append(dev)
    up = lambda v, shp: eng.upload(np.asarray(v, dtype=np.float64).reshape(shp))
    dxn, dxo = up(xnew, (B, N, n)), up(xold, (B, N, n))
    dKn = eng.upload(np.swapaxes(np.asarray(traj_new.K).reshape(B, N, m, n), -1, -2))
    dkn = up(traj_new.k, (B, N, m))
    dSn = eng.upload(np.swapaxes(np.asarray(traj_new.Sigma).reshape(B, N, m, m), -1, -2))
    a.xnew, a.xold, a.K_new, a.k_new, a.Sig_new = dxn.ptr, dxo.ptr, dKn.ptr, dkn.ptr, dSn.ptr
    dev, a.K_prev = _pack_mat(eng, np.asarray(traj_prev.K).reshape(B, N, m, n), B, N, m, n, "K_prev"); keep.append(dev)
    dev, a.Sig_prev = _pack_mat(eng, np.asarray(traj_prev.Sigma).reshape(B, N, m, m), B, N, m, m, "Sig_prev"); keep.append(dev)
    dev, a.Sigi_prev = _pack_mat(eng, np.asarray(traj_prev.Sigmai).reshape(B, N, m, m), B, N, m, m, "Sigi_prev"); keep.append(dev)
    dev, a.k_prev = _pack_vec(eng, np.asarray(traj_prev.k).reshape(B, N, m), B, N, m, "k_prev"); keep.append(dev)
    klt, klm = eng.empty((B, N)), eng.empty((B,))
    a.kl_t, a.kl_mean = klt.ptr, klm.ptr
    if Sx_cache is not None:
        a.Sx_tri, a.Sx_mode, a.Sx_count = Sx_cache.ptr, int(Sx_mode), int(Sx_count)
    eng._ck(eng.lib.ddp_kl_div_f64(eng.h, C.byref(a)))
    eng.synchronize()
    t, mn = klt.numpy(), klm.numpy()
    return (t, mn) if batched else (t[0], float(mn[0]))


# ---------------------------------------------------------------------------------------------
# iLQG  (iLQG.jl:143-341), whole outer loop device resident
# ---------------------------------------------------------------------------------------------

DEFAULT_ALPHA = 10.0 ** np.linspace(0, -3, 11)        # iLQG.jl:145
STATUS = {-1: "running", 0: "SUCCESS: gradient norm < tol_grad", 1: "SUCCESS: cost change < tol_fun",
          2: "EXIT: lambda > lambda_max", 3: "EXIT: Maximum iterations reached",
          4: "EXIT: Initial control sequence caused divergence", 6: "outer-loop safety cap reached"}
TRACE_DTYPE = np.dtype([("lam", "f8"), ("dlam", "f8"), ("cost", "f8"), ("alpha", "f8"), ("grad_norm", "f8"), ("improvement", "f8"),
                        ("reduce_ratio", "f8"), ("accepted", "i4"), ("bp_retries", "i4")])


def iLQG(f, costfun, df, x0, u0, *, lims=None, alpha=None, tol_fun=1e-7, tol_grad=1e-4, max_iter=500, lam=1.0, dlam=1.0,
         lamfactor=1.6, lammax=1e10, lammin=1e-6, regType=1, reduce_ratio_min=0.0, diff_fun=None, cost=None, trace_iters=0,
         force_generic=False, engine: Engine = None):
    """``iLQG(f,costfun,df,x0,u0;kw...)`` -- iLQG.jl:143-341, for one trajectory or a batch.

    ``f, costfun, df`` must be the callbacks of one device model descriptor.  ``x0`` is ``(n,)`` /
    ``(B,n)``, ``u0`` is ``(N,m)`` / ``(B,N,m)``; a pre-rolled initial trajectory ``x0`` of shape ``(N,n)`` / ``(B,N,n)``
    together with its ``cost`` (total, or per step: it is summed) starts from that trajectory instead of rolling out
    (iLQG.jl:193-197).  Returns ``(x, u, traj_new, Vx, Vxx, cost, trace)`` like the
    reference; ``Vxx`` is the value Hessian at the first timestep only (the history is optional on
    the device), ``cost`` the total cost, ``trace`` a dict with the per-trajectory final
    ``status, iter, accepted_iter, lam, dlam, g_norm`` and ``n_outer``; with ``trace_iters = k > 0`` also
    ``trace["iterations"]``: a structured array ``(k,)`` / ``(k,B)`` with the reference's per-iteration trace keys
    (``lam, dlam, cost, alpha, grad_norm, improvement, reduce_ratio`` -- iLQG.jl:257, 325-330; NaN where the reference
    records nothing).  Unbatched calls return ``None`` when the initial controls diverge (iLQG.jl:209) and raise
    ``RuntimeError`` when no iteration completed (iLQG.jl:335), as the reference does.
    """
    if diff_fun is not None:
        raise NotImplementedError("only the default diff_fun (-) is supported on the device")
    model = _model_of(f, costfun)
    if getattr(df, "model", None) is not model:
        raise TypeError("df must belong to the same device model descriptor as f and costfun")
    u0 = np.asarray(u0, dtype=np.float64)
    batched = u0.ndim == 3
    B = u0.shape[0] if batched else 1
    N, m = u0.shape[-2:]
    x0 = np.asarray(x0, dtype=np.float64)
    prerolled = x0.ndim == u0.ndim and x0.ndim >= 2          # size(x0,2) == N branch (iLQG.jl:193)
    if prerolled:
        if x0.shape[-2] != N:
            raise RuntimeError("pre-rolled initial trajectory must be of correct length (size(x0,2) == N)")      # iLQG.jl:199
        if cost is None:
            raise RuntimeError("Initial trajectory supplied, initial cost must also be supplied")
    n = x0.shape[-1]
    alpha = DEFAULT_ALPHA if alpha is None else np.asarray(alpha, dtype=np.float64)
    eng = engine or Engine(n, m, N, B, force_generic=force_generic)
    M, keep = _pack_model(eng, model, B, N, n, m)
    o = L.IlqgOpts()
    o.n_alpha = len(alpha)
    for i, v in enumerate(alpha):
        o.alpha[i] = float(v)
    o.tol_fun, o.tol_grad, o.max_iter = tol_fun, tol_grad, max_iter
    o.lam, o.dlam, o.lam_factor, o.lam_max, o.lam_min = lam, dlam, lamfactor, lammax, lammin
    o.reg_type, o.reduce_ratio_min = int(regType), float(reduce_ratio_min)
    ld, lst = _lims_dev(eng, lims, m)
    if lst:
        raise NotImplementedError("time-varying lims are supported by back_pass / forward_pass, not by the iLQG driver")
    if ld is not None:
        keep.append(ld)
        o.lims = ld.ptr
    if prerolled:
        xi = eng.upload(x0.reshape(B, N, n)); keep.append(xi)
        c0 = np.asarray(cost, dtype=np.float64)
        c0 = c0.reshape(B, -1).sum(axis=1) if c0.size != B else c0.reshape(B)
        ci = eng.upload(c0); keep.append(ci)
        o.x_init, o.cost_init = xi.ptr, ci.ptr
        dx0 = eng.upload(np.ascontiguousarray(x0.reshape(B, N, n)[:, 0]))
    else:
        dx0 = eng.upload(np.broadcast_to(x0.reshape(-1, n), (B, n)))
    tr = None
    if trace_iters > 0:
        init = np.zeros((trace_iters, B), dtype=TRACE_DTYPE)
        for key in ("lam", "dlam", "cost", "alpha", "grad_norm", "improvement", "reduce_ratio"):
            init[key] = np.nan
        init["accepted"] = -1
        tr = eng.upload(init.view(np.uint8).reshape(trace_iters, B, TRACE_DTYPE.itemsize), np.uint8)
        o.trace, o.trace_cap = tr.ptr, trace_iters
    du0 = eng.upload(u0.reshape(B, N, m))
    # zero-initialised: a trajectory whose initial rollout diverges (status 4) never gets an x, u
    x, u = eng.empty((B, N, n)).zero(), eng.empty((B, N, m)).zero()
    K, k, Vx, Vxx1 = eng.empty((B, N, n, m)).zero(), eng.empty((B, N, m)).zero(), eng.empty((B, N, n)).zero(), eng.empty((B, n, n)).zero()
    st = eng.empty((B, C.sizeof(L.IlqgState)), np.uint8)
    n_outer = C.c_int32(0)
    rc = eng.lib.ddp_ilqg_solve_f64(eng.h, C.byref(M), C.byref(o), dx0.ptr, du0.ptr, x.ptr, u.ptr, K.ptr, k.ptr, Vx.ptr,
                                    Vxx1.ptr, st.ptr, C.byref(n_outer))
    if rc != -5:                                           # DDP_ERR_INCOMPLETE still delivers every output (status 6 marks the stragglers)
        eng._ck(rc)
    states = np.frombuffer(st.numpy().tobytes(), dtype=np.dtype(
        [("lam", "f8"), ("dlam", "f8"), ("cost", "f8"), ("g_norm", "f8"), ("last_dcost", "f8"), ("last_alpha", "f8"),
         ("iter", "i4"), ("accepted_iter", "i4"), ("status", "i4"), ("pad", "i4")]))
    trace = {key: states[key].copy() for key in ("status", "iter", "accepted_iter", "lam", "dlam", "g_norm", "last_dcost", "last_alpha")}
    trace["n_outer"] = int(n_outer.value)
    if tr is not None:
        trace["iterations"] = np.frombuffer(tr.numpy().tobytes(), dtype=TRACE_DTYPE).reshape(trace_iters, B).copy()
    xs, us = x.numpy(), u.numpy()
    Ks, ks = np.swapaxes(K.numpy(), -1, -2), k.numpy()
    Vxs, Vxxs = Vx.numpy(), np.swapaxes(Vxx1.numpy(), -1, -2)
    cost = states["cost"].copy()
    if batched:
        return xs, us, GaussianPolicy(N, n, m, Ks, ks), Vxs, Vxxs, cost, trace
    if trace["status"][0] == 4:
        return None
    if trace["iter"][0] == 1:
        raise RuntimeError("Failure: no iterations completed, something is wrong.")
    trace = {key: ((v[:, 0] if key == "iterations" else v[0]) if isinstance(v, np.ndarray) else v) for key, v in trace.items()}
    return xs[0], us[0], GaussianPolicy(N, n, m, Ks[0], ks[0]), Vxs[0], Vxxs[0], float(cost[0]), trace


# ---------------------------------------------------------------------------------------------
# iLQGkl  (iLQGkl.jl:25-183, 238-252), single-KL-constraint branch
# ---------------------------------------------------------------------------------------------


@dataclass
class SimpleLTVModel:
    """What ``iLQGkl`` needs of the reference's ``model`` argument (LinearTimeVaryingModelsBase.SimpleLTVModel,
    demo_linear.jl:118 -- third party, not vendored): ``fx`` = what ``df(model,x,u)`` returns as the state Jacobian and
    ``R1`` = what ``covariance(model,x,u)`` returns (forward_pass.jl:38,42)."""
    fx: np.ndarray
    R1: np.ndarray


def _split_model(model, R1):
    if R1 is not None:                      # (…, traj_prev, fx_model, R1) spelling
        return model, R1
    return model.fx, model.R1


def iLQGkl(dynamics, costfun, derivs, x0, traj_prev: GaussianPolicy, model, R1=None, *, kl_step=1.0, lims=None, max_iter=50,
           etabracket=(1e-8, 1.0, 1e16), del0=1e-4, cost=None, max_eta_retries=200, force_generic=False):
    """``iLQGkl(dynamics,costfun,derivs,x0,traj_prev,model;kw...)`` -- iLQGkl.jl:25-183 + 238-252.

    Every sweep runs on the device (``ddp_back_pass_gps_f64``, ``ddp_forward_pass_f64``, ``ddp_kl_div_f64``,
    ``ddp_model_derivs_f64``); only the scalar η-bracket update of ``calc_η`` (klutils.jl:110-130) and the
    loop control are host code.  ``model`` is a :class:`SimpleLTVModel` (``fx``, ``R1``:
    what ``df(model,x,u)`` and ``covariance(model,x,u)`` return, forward_pass.jl:38,42, of the un-vendored third-party type);
    the older spelling ``(…, traj_prev, fx_model, R1)`` works too.  ``x0`` is the
    pre-rolled trajectory (N,n) and ``cost`` its cost, as the reference requires (iLQGkl.jl:63-70).
    Unbatched only (one trajectory), like the reference.
    """
    fx_model, R1 = _split_model(model, R1)
    model = _model_of(dynamics, costfun)
    u = np.array(traj_prev.k, dtype=np.float64)                     # :47
    x = np.asarray(x0, dtype=np.float64)
    N, m = u.shape
    n = x.shape[1]
    if x.ndim != 2 or x.shape[0] != N:
        raise RuntimeError("pre-rolled initial trajectory must be of correct length (size(x0,2) == N)")
    if cost is None:
        raise RuntimeError("Initial trajectory supplied, initial cost must also be supplied")
    eta = np.array(etabracket, dtype=np.float64)
    prev0 = GaussianPolicy(N, n, m, traj_prev.K, np.zeros_like(traj_prev.k), traj_prev.Sigma, traj_prev.Sigmai)   # :52 k := 0
    # derivatives once, outside the loop (:88, quirk Q9)
    eng = Engine(n, m, N, 1, force_generic=force_generic)
    M, keep = _pack_model(eng, model, 1, N, n, m)
    dx, du = eng.upload(x[None]), eng.upload(u[None])
    cxd, cud = eng.empty((1, N, n)), eng.empty((1, N, m))
    fxd = eng.empty((1, N, n, n)) if model.kind == 2 else None
    fud = eng.empty((1, N, n, m)) if model.kind == 2 else None
    eng._ck(eng.lib.ddp_model_derivs_f64(eng.h, C.byref(M), dx.ptr, du.ptr, fxd.ptr if fxd else None, fud.ptr if fud else None,
                                         cxd.ptr, cud.ptr))
    eng.synchronize()
    cx, cu = cxd.numpy()[0], cud.numpy()[0]
    if model.kind == 2:
        fx, fu = np.swapaxes(fxd.numpy()[0], -1, -2), np.swapaxes(fud.numpy()[0], -1, -2)
    else:
        fx, fu = np.asarray(model.A), np.asarray(model.B)
    cxx, cuu, cxu = model.Q, model.R, np.zeros((n, m))
    trace = {key: [] for key in ("cost", "improvement", "reduce_ratio", "divergence", "eta")}
    trace["cost"].append((0, float(np.sum(cost))))
    satisfied = False
    traj_new = xnew = unew = costnew = Vx = Vxx = None
    it = 0
    for it in range(1, max_iter + 1):                               # :93
        diverge, retries = 1, 0
        while diverge > 0:                                          # :97
            diverge, traj_new, Vx, Vxx, dV = back_pass_gps(cx, cu, cxx, cxu, cuu, fx, fu, lims, x, u, (prev0, eta),
                                                           force_generic=force_generic)
            if diverge > 0:
                eta[1] += del0                                      # :104
                del0 *= 2
                retries += 1
                if retries > max_eta_retries:
                    raise RuntimeError("eta retry loop did not terminate")
        xnew, unew, costnew = forward_pass(traj_new, x[0], u, x, 1.0, dynamics, costfun, lims, force_generic=force_generic)   # :134
        kl_t, divergence = kl_div_wiki(xnew, x, fx_model, R1, traj_new, prev0)                                               # :135, :143
        dcost = float(np.sum(cost) - costnew)
        expected = -(dV[0] + dV[1])                                 # :138
        # calc_η (klutils.jl:110-130)
        if not (kl_step > 0):
            satisfied, divergence = True, 0.0
        else:
            violation = divergence - kl_step
            satisfied = abs(violation) < 0.1 * kl_step
            if not satisfied:
                if violation < 0:
                    eta[2] = eta[1]
                    eta[1] = max(math.sqrt(eta[0] * eta[2]), 0.1 * eta[2])
                else:
                    eta[0] = eta[1]
                    eta[1] = min(math.sqrt(eta[0] * eta[2]), 10.0 * eta[0])
        trace["improvement"].append((it, dcost))
        trace["cost"].append((it, float(costnew)))
        trace["reduce_ratio"].append((it, dcost / expected))
        trace["divergence"].append((it, float(divergence)))
        trace["eta"].append((it, float(eta[1])))
        if satisfied or eta[1] > 0.999 * eta[2]:                    # :173-181
            break
    traj_new.k = unew.copy()                                        # :240-241 (quirk Q11)
    trace.update(iters=it, satisfied=satisfied, etabracket=eta)
    return xnew, unew, traj_new, Vx, Vxx, costnew, trace


KL_STATUS = {-1: "running", 0: "KL constraint satisfied", 1: "eta > 0.999 eta_max", 3: "EXIT: Maximum iterations reached",
             5: "eta-retry limit reached"}


def iLQGkl_device(dynamics, costfun, derivs, x0, traj_prev: GaussianPolicy, model, R1=None, *, kl_step=1.0, lims=None, max_iter=50,
                  etabracket=(1e-8, 1.0, 1e16), del0=1e-4, cost=None, max_eta_retries=200, force_generic=False,
                  engine: Engine = None, covariance_cache: bool = True):
    """Whole ``iLQGkl`` outer loop on the device for one trajectory or a batch (``ddp_ilqgkl_solve_f64``):
    iLQGkl.jl:93-183 with ``calc_eta`` (klutils.jl:110-130) as per-trajectory state machines.

    ``x0`` is the pre-rolled trajectory ``(N,n)`` / ``(B,N,n)``, ``traj_prev`` holds ``K (…,N,m,n)``, ``k (…,N,m)``,
    ``Sigma``/``Sigmai (…,N,m,m)``, ``cost`` the total (or per-step) cost of ``x0``.  Returns
    ``(xnew, unew, traj_new, Vx, Vxx1, costnew, trace)``; ``trace`` holds the per-trajectory final
    ``status, iter, eta bracket, divergence, dcost, expected, retries`` (no per-iteration history: nothing is
    read back inside the loop except two counters).  ``covariance_cache``: keep the state covariances of
    ``forward_covariance`` from the first η iteration for the later ones (``ddp_ilqgkl_opts.no_covariance_cache = 0``).
    """
    fx_model, R1 = _split_model(model, R1)
    model = _model_of(dynamics, costfun)
    x = np.asarray(x0, dtype=np.float64)
    batched = x.ndim == 3
    if not batched:
        x = x[None]
    B, N, n = x.shape
    u = np.array(traj_prev.k, dtype=np.float64).reshape(B, N, -1)               # iLQGkl.jl:47
    m = u.shape[-1]
    if cost is None:
        raise RuntimeError("Initial trajectory supplied, initial cost must also be supplied")
    cost = np.asarray(cost, dtype=np.float64)
    cost = cost.reshape(B, -1).sum(axis=1) if cost.size != B else cost.reshape(B)
    eng = engine or Engine(n, m, N, B, force_generic=force_generic)
    M, keep = _pack_model(eng, model, B, N, n, m)
    o = L.IlqgklOpts()
    o.kl_step, o.max_iter, o.del0, o.max_eta_retries = float(kl_step), int(max_iter), float(del0), int(max_eta_retries)
    o.no_covariance_cache = 0 if covariance_cache else 1
    for i in range(3):
        o.eta_bracket[i] = float(etabracket[i])
    ld, _ = _lims_dev(eng, lims, m)
    if ld is not None:
        keep.append(ld)
        o.lims = ld.ptr
    a = L.IlqgklArgs()
    dx, du, dc = eng.upload(x), eng.upload(u), eng.upload(cost)
    a.x, a.u, a.cost = dx.ptr, du.ptr, dc.ptr
    dev, a.K_prev = _pack_mat(eng, np.asarray(traj_prev.K).reshape(B, N, m, n), B, N, m, n, "K_prev"); keep.append(dev)
    dev, a.Sig_prev = _pack_mat(eng, np.asarray(traj_prev.Sigma).reshape(B, N, m, m), B, N, m, m, "Sig_prev"); keep.append(dev)
    dev, a.Sigi_prev = _pack_mat(eng, np.asarray(traj_prev.Sigmai).reshape(B, N, m, m), B, N, m, m, "Sigi_prev"); keep.append(dev)
    dev, a.fx_model = _pack_mat(eng, fx_model, B, N, n, n, "fx_model"); keep.append(dev)
    dev, a.R1 = _pack_mat(eng, R1, B, 1, n, n, "R1"); keep.append(dev)
    xnew, unew, K, k = eng.empty((B, N, n)).zero(), eng.empty((B, N, m)).zero(), eng.empty((B, N, n, m)).zero(), eng.empty((B, N, m)).zero()
    Sig, Sigi, Vx, Vxx1 = eng.empty((B, N, m, m)).zero(), eng.empty((B, N, m, m)).zero(), eng.empty((B, N, n)).zero(), eng.empty((B, n, n)).zero()
    cnew = eng.empty((B,)).zero()
    st = eng.empty((B, C.sizeof(L.IlqgklState)), np.uint8)
    a.xnew, a.unew, a.K, a.k, a.Sig, a.Sigi, a.Vx, a.Vxx1, a.costnew, a.state = (xnew.ptr, unew.ptr, K.ptr, k.ptr, Sig.ptr, Sigi.ptr,
                                                                              Vx.ptr, Vxx1.ptr, cnew.ptr, st.ptr)
    n_outer = C.c_int32(0)
    eng._ck(eng.lib.ddp_ilqgkl_solve_f64(eng.h, C.byref(M), C.byref(o), C.byref(a), C.byref(n_outer)))
    states = np.frombuffer(st.numpy().tobytes(), dtype=np.dtype(
        [("eta_min", "f8"), ("eta", "f8"), ("eta_max", "f8"), ("del0", "f8"), ("divergence", "f8"), ("dcost", "f8"),
         ("expected", "f8"), ("cost", "f8"), ("iter", "i4"), ("status", "i4"), ("retries", "i4"), ("pad", "i4")]))
    trace = {key: states[key].copy() for key in ("status", "iter", "eta_min", "eta", "eta_max", "del0", "divergence", "dcost",
                                                 "expected", "retries")}
    trace["satisfied"] = trace["status"] == 0
    trace["n_outer"] = int(n_outer.value)
    pol = GaussianPolicy(N, n, m, np.swapaxes(K.numpy(), -1, -2), k.numpy(), np.swapaxes(Sig.numpy(), -1, -2),
                         np.swapaxes(Sigi.numpy(), -1, -2))
    out = (xnew.numpy(), unew.numpy(), pol, Vx.numpy(), np.swapaxes(Vxx1.numpy(), -1, -2), cnew.numpy(), trace)
    del keep
    if batched:
        return out
    pol1 = GaussianPolicy(N, n, m, pol.K[0], pol.k[0], pol.Sigma[0], pol.Sigmai[0])
    tr1 = {key: (v[0] if isinstance(v, np.ndarray) else v) for key, v in trace.items()}
    return out[0][0], out[1][0], pol1, out[3][0], out[4][0], float(out[5][0]), tr1


# ---------------------------------------------------------------------------------------------
# one end-to-end iteration on host buffers (the bench's e2e path)
# ---------------------------------------------------------------------------------------------


class HostIteration:
    """Pinned host buffers + ``ddp_ilqg_iter_host_f64``: one backward sweep and one forward rollout
    for a batch of per-trajectory LTI linear problems, host to host (device layout arrays)."""

    FIELDS_IN = ("fx", "fu", "cx", "cu", "x", "u", "lam")
    FIELDS_OUT = ("xnew", "unew", "cost", "dV")

    def __init__(self, eng: Engine, Q, R, cxu=None, reg_type=1, alpha=1.0, chunk=0, device_derivs=False, keep_policy=True):
        self.eng = eng
        n, m, T, B = eng.n, eng.m, eng.T, eng.B
        shapes = dict(fx=(B, n, n), fu=(B, m, n), cx=(B, T, n), cu=(B, T, m), x=(B, T, n), u=(B, T, m), lam=(B,),
                      xnew=(B, T, n), unew=(B, T, m), cost=(B,), dV=(B, 2))
        self._ptrs = []
        self.bufs = {}
        for name, shp in shapes.items():
            self.bufs[name] = self._pinned(shp, np.float64)
        self.bufs["diverge"] = self._pinned((B,), np.int32)
        self.Q = np.ascontiguousarray(np.asarray(Q, dtype=np.float64).T)
        self.R = np.ascontiguousarray(np.asarray(R, dtype=np.float64).T)
        self.cxu = np.zeros((m, n)) if cxu is None else np.ascontiguousarray(np.asarray(cxu, dtype=np.float64).T)
        self.args = L.IterHostArgs()
        for name in self.FIELDS_IN + self.FIELDS_OUT + ("diverge",):
            setattr(self.args, name, self.bufs[name].ctypes.data)
        if device_derivs:            # cx = Qx, cu = Ru are formed on the device from x, u (the reference's df step)
            self.args.cx = self.args.cu = None
        self.args.Q, self.args.R, self.args.cxu = self.Q.ctypes.data, self.R.ctypes.data, self.cxu.ctypes.data
        self.args.reg_type, self.args.alpha, self.args.chunk = reg_type, alpha, chunk
        self.args.keep_policy = 1 if keep_policy else 0
        self.args.q_diagonal = 1 if np.count_nonzero(self.Q - np.diag(np.diagonal(self.Q))) == 0 else 0

    def _pinned(self, shape, dtype):
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = self.eng.lib.ddp_host_alloc(C.byref(p), max(nbytes, 8))
        if rc != 0:
            raise MemoryError("ddp_host_alloc failed")
        self._ptrs.append(p.value)
        buf = (C.c_char * max(nbytes, 8)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)

    def run(self, inputs_resident=False, commit_accepted=False, cost_prev=None):
        """One iteration.  ``inputs_resident``: use the device copies of ``fx,fu,x,u,lam`` left by the previous call instead of
        uploading them; ``commit_accepted`` (+ ``cost_prev``, the cost of the current trajectories): afterwards the device copies
        of ``x,u`` hold ``xnew,unew`` where the step was accepted, so the next call can run with ``inputs_resident``."""
        self.args.inputs_resident = 1 if inputs_resident else 0
        self.args.commit_accepted = 1 if commit_accepted else 0
        if commit_accepted:
            self._cprev = np.ascontiguousarray(cost_prev, dtype=np.float64)
            self.args.cost_prev = self._cprev.ctypes.data
        self.eng._ck(self.eng.lib.ddp_ilqg_iter_host_f64(self.eng.h, C.byref(self.args)))
        return int(self.args.h2d_bytes), int(self.args.d2h_bytes)

    def policy_ptrs(self):
        """Device pointers ``(K, k, Vx, xnew, unew)`` of what the last ``run`` left on the device (ddp_iter_host_policy);
        ``K, k, Vx`` are ``None`` unless the iteration was created with ``keep_policy``."""
        ps = [C.c_void_p() for _ in range(5)]
        self.eng._ck(self.eng.lib.ddp_iter_host_policy(self.eng.h, *[C.byref(p) for p in ps]))
        return tuple(p.value for p in ps)

    def policy(self, b0=0, b1=None):
        """Download the policy of trajectories ``b0:b1`` kept on the device: ``(K (nb,T,m,n), k (nb,T,m), Vx (nb,T,n))``."""
        n, m, T, B = self.eng.n, self.eng.m, self.eng.T, self.eng.B
        b1 = B if b1 is None else b1
        Kp, kp, Vxp, _, _ = self.policy_ptrs()
        if Kp is None:
            raise RuntimeError("the policy was not kept on the device (keep_policy=False)")
        nb = b1 - b0
        K, k, Vx = np.empty((nb, T, n, m)), np.empty((nb, T, m)), np.empty((nb, T, n))
        for arr, p, per in ((K, Kp, T * n * m), (k, kp, T * m), (Vx, Vxp, T * n)):
            self.eng._ck(self.eng.lib.ddp_download(self.eng.h, arr.ctypes.data, p + 8 * per * b0, arr.nbytes))
        return np.swapaxes(K, -1, -2), k, Vx

    def close(self):
        for p in self._ptrs:
            self.eng.lib.ddp_host_free(p)
        self._ptrs = []


# ---------------------------------------------------------------------------------------------
# one device-resident iteration, chunked (ddp_ilqg_iter_f64): batches larger than the policy's HBM residency
# ---------------------------------------------------------------------------------------------


def iterate_chunked(model, x, u, lam, alpha=1.0, *, regType=1, lims=None, chunk=0, keep_policy=False, engine: Engine = None):
    """One ``back_pass`` + ``forward_pass(α)`` over a batch ``x (B,N,n)``, ``u (B,N,m)`` of a device model with the policy
    living in a chunk-sized scratch (``ddp_ilqg_iter_f64``; BASELINE config 5).  The derivative step of the model
    (``cx = Q(x-goal)``, ``cu = R u``; pendcart also ``fx, fu``) runs on the device per chunk.
    Returns ``dict(xnew, unew, cost, dV, diverge, n_chunks[, K, k, Vx])``."""
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    B, N, n = x.shape
    m = u.shape[-1]
    eng = engine or Engine(n, m, N, B)
    M, keep = _pack_model(eng, model, B, N, n, m)
    a = L.IterArgs()
    dx, du = eng.upload(x), eng.upload(u)
    dl = eng.upload(np.broadcast_to(np.asarray(lam, dtype=np.float64), (B,)))
    a.x, a.u, a.lam = dx.ptr, du.ptr, dl.ptr
    alpha = np.asarray(alpha, dtype=np.float64)
    if alpha.ndim == 0:
        a.alpha_scalar = float(alpha)
    else:
        dal = eng.upload(np.broadcast_to(alpha, (B,))); keep.append(dal)
        a.alpha = dal.ptr
    a.reg_type = int(regType)
    ld, lst = _lims_dev(eng, lims, m)
    if lst:
        raise NotImplementedError("time-varying lims: use back_pass / forward_pass")
    if ld is not None:
        keep.append(ld)
        a.lims = ld.ptr
    xnew, unew, cost, dV, dv = eng.empty((B, N, n)), eng.empty((B, N, m)), eng.empty((B,)), eng.empty((B, 2)), eng.empty((B,), np.int32)
    a.xnew, a.unew, a.cost, a.dV, a.diverge = xnew.ptr, unew.ptr, cost.ptr, dV.ptr, dv.ptr
    pol = None
    if keep_policy:
        pol = (eng.empty((B, N, n, m)), eng.empty((B, N, m)), eng.empty((B, N, n)))
        a.K, a.k, a.Vx = (p.ptr for p in pol)
    a.chunk = int(chunk)
    eng._ck(eng.lib.ddp_ilqg_iter_f64(eng.h, C.byref(M), C.byref(a)))
    eng.synchronize()
    out = dict(xnew=xnew.numpy(), unew=unew.numpy(), cost=cost.numpy(), dV=dV.numpy(), diverge=dv.numpy(), n_chunks=int(a.n_chunks))
    if pol is not None:
        out.update(K=np.swapaxes(pol[0].numpy(), -1, -2), k=pol[1].numpy(), Vx=pol[2].numpy())
    del keep
    return out
