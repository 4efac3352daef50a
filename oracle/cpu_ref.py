"""ctypes wrapper of oracle/_build/libcpu_ref.so (the C++/OpenMP restated reference).
Test / measurement infrastructure only -- see oracle/cpu_ref.cpp."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcpu_ref.so")
_lib = None


def _host_stamp():
    """Identifies the host CPU the library was compiled for (-march=native): model name + ISA flags."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        keep = [ln for ln in txt.splitlines() if ln.startswith(("model name", "flags"))][:2]
    except OSError:
        keep = ["unknown"]
    return hashlib.sha256("\n".join(keep).encode()).hexdigest()


def build(force=False):
    stamp = os.path.join(HERE, "_build", "host.stamp")
    cur = _host_stamp()
    stale = force or not os.path.exists(LIB) or not os.path.exists(stamp) or open(stamp).read() != cur
    if stale and os.path.exists(LIB):
        os.remove(LIB)                     # another host's -march=native binary: rebuild for this one
    try:
        subprocess.check_call(["make", "-s", "-C", HERE])
    except subprocess.CalledProcessError:  # a compiler without -march=native support for this CPU
        subprocess.check_call(["make", "-s", "-C", HERE, "MARCH=x86-64-v3"])
    with open(stamp, "w") as f:
        f.write(cur)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.cpu_ref_num_threads.restype = C.c_int
    return _lib


def host_cores():
    """CPUs this process may run on (what OpenMP should use; torchrun exports OMP_NUM_THREADS=1, which is not it)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads():
    return int(load().cpu_ref_num_threads())


def back_pass(cx, cu, cxx, cxu, cuu, fx, fu, lam, regType, lims, u, *, want_Vxx=True, nthreads=0):
    """Device-layout arrays: cx (B,T,n), cu (B,T,m), fx (B,[T,]n,n) column-major per step (i.e. the
    math-layout matrix transposed), cxx/cxu/cuu shared (n,n)/(m,n)/(m,m) column-major.
    Returns diverge, K (B,T,n,m), k, Vx, Vxx|None, Vxx1, Quu, dV  (device layout)."""
    lib = load()
    B, T, n = cx.shape
    m = cu.shape[2]
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    cx, cu, cxx, cxu, cuu, fx, fu = map(f8, (cx, cu, cxx, cxu, cuu, fx, fu))
    lam = f8(np.broadcast_to(lam, (B,)))
    tv = fx.ndim == 4
    K = np.empty((B, T, n, m)); k = np.empty((B, T, m)); Vx = np.empty((B, T, n))
    Vxx = np.empty((B, T, n, n)) if want_Vxx else None
    Vxx1 = np.empty((B, n, n)); Quu = np.empty((B, T, m, m)); dV = np.empty((B, 2))
    dv = np.empty((B,), dtype=np.int32)
    limsd = None if lims is None else f8(np.asarray(lims).reshape(m, 2).T)
    ud = None if u is None else f8(u)
    L = C.c_long
    lib.cpu_back_pass(C.c_int(n), C.c_int(m), C.c_int(T), L(B), _p(cx), _p(cu), _p(cxx), L(0), L(0), _p(cxu), L(0), L(0),
                      _p(cuu), L(0), L(0), _p(fx), L((T if tv else 1) * n * n), L(n * n if tv else 0),
                      _p(fu), L((T if tv else 1) * n * m), L(n * m if tv else 0), _p(lam), C.c_int(regType), _p(limsd), _p(ud),
                      _p(dv), _p(K), _p(k), _p(Vx), _p(Vxx), _p(Vxx1), _p(Quu), _p(dV), C.c_int(nthreads))
    return dv, K, k, Vx, Vxx, Vxx1, Quu, dV


def forward_pass_linear(K, k, x0, x, u, alpha, lims, A, Bm, Q, R, *, nthreads=0):
    lib = load()
    B, T, m = u.shape
    n = x0.shape[1]
    f8 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    K, k, x0, x, u, A, Bm, Q, R = map(f8, (K, k, x0, x, u, A, Bm, Q, R))
    alpha = f8(np.broadcast_to(alpha, (B,)))
    limsd = None if lims is None else f8(np.asarray(lims).reshape(m, 2).T)
    xnew = np.empty((B, T, n)); unew = np.empty((B, T, m)); cost = np.empty((B,))
    L = C.c_long
    lib.cpu_forward_pass_linear(C.c_int(n), C.c_int(m), C.c_int(T), L(B), _p(K), _p(k), _p(x0), _p(x), _p(u), _p(alpha),
                                _p(limsd), _p(A), L(n * n), _p(Bm), L(n * m), _p(Q), _p(R), _p(xnew), _p(unew), _p(cost),
                                C.c_int(nthreads))
    return xnew, unew, cost


def boxqp(H, g, lower, upper, x0):
    lib = load()
    B, m, _ = H.shape
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    Hc = f8(np.swapaxes(H, -1, -2))
    g, lower, upper, x0 = map(f8, (g, lower, upper, x0))
    x = np.empty((B, m)); res = np.empty((B,), np.int32); Hf = np.empty((B, m, m))
    fm = np.empty((B,), np.uint32); nf = np.empty((B,), np.int32)
    lib.cpu_boxqp(C.c_int(m), C.c_long(B), _p(Hc), _p(g), _p(lower), _p(upper), _p(x0), _p(x), _p(res), _p(Hf), _p(fm), _p(nf))
    free = ((fm[:, None] >> np.arange(m)[None, :]) & 1).astype(bool)
    return x, res, np.swapaxes(Hf, -1, -2), free, nf
