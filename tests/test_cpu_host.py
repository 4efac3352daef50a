"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/ddp.h,
fails loudly without a GPU, the C++/OpenMP baseline agrees with the NumPy oracle, and the N > 1
sharding logic works under gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import make_batch_lq, relerr
from oracle import cpu_ref as CR
from oracle import ddp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(ddp):
    hdr = open(os.path.join(ROOT, "include", "ddp.h")).read()
    declared = set(re.findall(r"DDP_API\s+[\w\s\*]+?\b(ddp_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(ddp.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    bound = {s[0] for s in ddp._lib.SYMBOLS}
    assert declared == bound                       # the ctypes mirror binds exactly the header's surface
    assert ddp.load().ddp_version() == 201


def struct_layout_from_header(ddp):
    """{struct: (sizeof, {field: offsetof})} compiled from include/ddp.h with gcc -- the C side of the layout tests."""
    L = ddp._lib
    lines = []
    for cname, cls in L.STRUCTS.items():
        lines.append(f'printf("S {cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            cf = L.FIELD_ALIASES.get(fname, fname)
            lines.append(f'printf("F {cname} {fname} %zu\\n", offsetof({cname}, {cf}));')
    src = "#include <stdio.h>\n#include <stddef.h>\n#include \"ddp.h\"\nint main(void) {\n" + "\n".join(lines) + "\nreturn 0; }\n"
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c"); exe = os.path.join(td, "s")
        open(c, "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    layout = {}
    for ln in out.splitlines():
        f = ln.split()
        if f[0] == "S":
            layout[f[1]] = (int(f[2]), {})
        else:
            layout[f[1]][1][f[2]] = int(f[3])
    return layout


def test_struct_layouts_match_the_header(ddp):
    """sizeof of every ABI struct AND offsetof of every field, compiled from include/ddp.h with gcc, equal the ctypes
    mirror: a field-order slip with an unchanged size cannot pass."""
    L = ddp._lib
    hdr = open(os.path.join(ROOT, "include", "ddp.h")).read()
    declared = set(re.findall(r"typedef struct (ddp_\w+) \{", hdr))
    assert declared == set(L.STRUCTS)                      # every struct of the header has a mirror
    layout = struct_layout_from_header(ddp)
    for cname, cls in L.STRUCTS.items():
        size, offs = layout[cname]
        assert ctypes.sizeof(cls) == size, cname
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == offs[fname], (cname, fname)


def test_julia_structs_match_the_header(ddp):
    """julia/DifferentialDynamicProgramming.jl cannot be executed here (no Julia in the container); its ccall structs are
    generated from the same table as the ctypes mirror (scripts/gen_julia_structs.py) and must be up to date, and every
    generated struct lists the C fields in the C order with the Julia type of the C type's size."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import gen_julia_structs as G
    generated = G.generate()
    committed = open(os.path.join(ROOT, "julia", "ddp_structs.jl")).read()
    assert generated == committed, "run python scripts/gen_julia_structs.py"
    layout = struct_layout_from_header(ddp)
    for cname, (jname, fields) in G.parse(committed).items():
        size, offs = layout[cname]
        off = 0
        for fname, jtype in fields:
            sz, al = G.JULIA_SIZES[jtype] if jtype in G.JULIA_SIZES else G.struct_size_align(jtype, committed, layout)
            off = (off + al - 1) // al * al
            cf = fname
            assert offs[G.c_field(cname, fname)] == off, (cname, fname, off)
            off += sz
        assert (off + 7) // 8 * 8 == size, cname


def test_no_cpu_fallback(ddp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ddp.DDPError, match="no CPU fallback"):
        ddp.Engine(4, 1, 10, 1)
    with pytest.raises(ddp.DDPError):
        ddp.back_pass(np.zeros((5, 3)), np.zeros((5, 1)), np.eye(3), np.zeros((3, 1)), np.eye(1), np.eye(3), np.ones((3, 1)), 1.0, 1,
                      None, np.zeros((5, 3)), np.zeros((5, 1)))
    pm = ddp.PendcartModel()
    with pytest.raises(RuntimeError, match="device model descriptor"):
        pm.f(np.zeros(4), np.zeros(1), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "differentialdynamicprogramming.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                # comments may mention the oracle; code must never import, link or load it
                for banned in ("import oracle", "from oracle", "cpu_ref", "ddp_oracle import", "oracle."):
                    assert banned not in txt, (f, banned)


@pytest.mark.parametrize("n,m,N,reg,lims", [(10, 2, 40, 1, None), (32, 8, 16, 2, None), (4, 1, 30, 2, 0.05), (8, 3, 30, 1, 0.05)])
def test_cpp_baseline_matches_oracle(n, m, N, reg, lims):
    B = 3
    A, Bm, Q, R, x, u = make_batch_lq(1, B, n, m, N)
    cx, cu = x @ Q.T, u @ R.T
    lim = None if lims is None else np.tile(np.array([[-lims, lims]]), (m, 1))
    lam = 1e-3 if lims else 1.0
    cxu = 0.01 * np.random.default_rng(2).standard_normal((n, m))
    dv, K, k, Vx, Vxx, Vxx1, Quu, dV = CR.back_pass(cx, cu, Q.T, cxu.T, R.T, np.swapaxes(A, -1, -2), np.swapaxes(Bm, -1, -2), lam, reg, lim, u)
    xn, un, cn = CR.forward_pass_linear(K, k, x[:, 0].copy(), x, u, 0.5, lim, np.swapaxes(A, -1, -2), np.swapaxes(Bm, -1, -2), Q.T, R.T)
    for b in range(B):
        d0, p0, Vx0, Vxx0, dV0 = O.back_pass(cx[b], cu[b], Q, cxu, R, A[b], Bm[b], lam, reg, lim, x[b], u[b])
        assert d0 == dv[b]                       # (the regType-2 case with random cxu diverges: compared too)
        assert max(relerr(np.swapaxes(K[b], -1, -2), p0.K), relerr(k[b], p0.k), relerr(Vx[b], Vx0),
                   relerr(np.swapaxes(Vxx[b], -1, -2), Vxx0), relerr(dV[b], dV0)) < 1e-9
        om = O.LinearModel(A[b], Bm[b], Q, R)
        x0_, u0_, c0_ = O.forward_pass(p0, x[b, 0], u[b], x[b], 0.5, om.f, om.costfun, lim)
        assert relerr(xn[b], x0_) < 1e-9 and relerr(un[b], u0_) < 1e-9 and abs(cn[b] - c0_) < 1e-9 * abs(c0_)


def test_cpp_boxqp_bit_identical_to_oracle():
    rng = np.random.default_rng(7)
    m, B = 6, 40
    Hs = []
    for _ in range(B):
        G = rng.standard_normal((m, m)); H = G @ G.T + 0.05 * np.eye(m); Hs.append((H + H.T) / 2)
    Hs = np.array(Hs); g = 3 * rng.standard_normal((B, m)); lo = -rng.random((B, m)); up = rng.random((B, m)); x0 = rng.standard_normal((B, m))
    x, res, Hf, free, nf = CR.boxqp(Hs, g, lo, up, x0)
    for b in range(B):
        x0_, r0, Hf0, free0, nf0 = O.boxQP(Hs[b], g[b], lo[b], up[b], x0[b])
        assert r0 == res[b] and nf0 == nf[b] and np.array_equal(free0, free[b]) and np.array_equal(x0_, x[b])


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import make_batch_lq
from oracle import cpu_ref as CR
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
B, n, m, N = 8, 6, 2, 20
A, Bm, Q, R, x, u = make_batch_lq(3, B, n, m, N)           # every rank builds the same global batch ...
lo, hi = rank * B // world, (rank + 1) * B // world         # ... and owns a contiguous shard (SURVEY 8e)
sl = slice(lo, hi)
cx, cu = x @ Q.T, u @ R.T
dv, K, k, Vx, _, _, _, dV = CR.back_pass(cx[sl], cu[sl], Q.T, np.zeros((m, n)), R.T, np.swapaxes(A[sl], -1, -2), np.swapaxes(Bm[sl], -1, -2), 1.0, 1, None, u[sl], want_Vxx=False)
xn, un, cn = CR.forward_pass_linear(K, k, x[sl, 0].copy(), x[sl], u[sl], 1.0, None, np.swapaxes(A[sl], -1, -2), np.swapaxes(Bm[sl], -1, -2), Q.T, R.T)
cost_old = 0.5 * np.sum(x[sl] * (x[sl] @ Q.T), axis=(1, 2)) + 0.5 * np.sum(u[sl] * (u[sl] @ R.T), axis=(1, 2))
ex = -(dV[:, 0] + dV[:, 1])
stats = torch.tensor([cn.sum(), (cost_old - cn).sum(), ex.sum(), float(((cost_old - cn) / ex > 0).sum()), float((dv > 0).sum()), float(hi - lo), 0, 0], dtype=torch.float64)
dist.all_reduce(stats)                                      # the one collective of the path: a 64-byte SUM
if rank == 0:
    dv, K, k, Vx, _, _, _, dV = CR.back_pass(cx, cu, Q.T, np.zeros((m, n)), R.T, np.swapaxes(A, -1, -2), np.swapaxes(Bm, -1, -2), 1.0, 1, None, u, want_Vxx=False)
    xn, un, cn = CR.forward_pass_linear(K, k, x[:, 0].copy(), x, u, 1.0, None, np.swapaxes(A, -1, -2), np.swapaxes(Bm, -1, -2), Q.T, R.T)
    assert abs(stats[0].item() - cn.sum()) < 1e-12 * abs(cn.sum()), (stats[0].item(), cn.sum())
    assert stats[5].item() == B and stats[3].item() == B and stats[4].item() == 0
    print("SHARD_OK")
dist.destroy_process_group()
'''


def test_batch_sharding_world_size_2_gloo(tmp_path):
    """trajectories shard by contiguous ranges with no data-path collective; the only exchange is the
    statistics vector.  Runs the sharded host logic under gloo, world_size 2."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29613", str(script), ROOT], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARD_OK" in out.stdout


def _build_c_host(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_host")
    libdir = os.path.join(root, "differentialdynamicprogramming.jl_b200")
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "c_host.c"),
                        "-L", libdir, "-lddp", "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe, libdir


def test_plain_c_host_compiles_and_links(tmp_path):
    """examples/c_host.c: a C (not C++) program that includes include/ddp.h and links libddp.so -- the boundary has no C++
    or torch types in it."""
    _build_c_host(tmp_path)


@pytest.mark.gpu
def test_plain_c_host_runs(tmp_path):
    import subprocess
    exe, libdir = _build_c_host(tmp_path)
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 diverged back passes" in r.stdout and "kernel variant: tile32x8" in r.stdout


def test_julia_shim_surface_is_consistent_with_the_abi(ddp):
    """The Julia shim cannot run here; what can be checked statically is: it exports the reference's list
    (DifferentialDynamicProgramming.jl:6) and defines what it exports, every `ccall` names a symbol of include/ddp.h, and every
    field it sets through `mk(DdpXxx; field = ...)` exists in the generated struct of that name."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import gen_julia_structs as G
    src = open(os.path.join(ROOT, "julia", "DifferentialDynamicProgramming.jl")).read()
    exported = set(re.findall(r"\b\w+\b", " ".join(re.findall(r"^export (.*?)(?:#.*)?$", src, re.M))))
    reference_exports = {"QPTrace", "boxQP", "demoQP", "iLQG", "iLQGkl", "demo_linear", "demo_linear_kl", "demo_pendcart", "GaussianPolicy"}
    assert reference_exports <= exported
    for name in exported:
        assert re.search(rf"(function {name}\b|^{name}\(|struct {name}\b)", src, re.M), f"exported but not defined: {name}"
    assert re.search(r"function iLQGkl\(dynamics, costfun, derivs, x0, traj_prev::GaussianPolicy, model::SimpleLTVModel;", src)
    hdr = open(os.path.join(ROOT, "include", "ddp.h")).read()
    declared = set(re.findall(r"DDP_API\s+[\w\s\*]+?\b(ddp_\w+)\s*\(", hdr))
    called = set(re.findall(r"ccall\(\(:(ddp_\w+), libddp\)", src))
    assert called and called <= declared, called - declared
    structs = {jname: [f for f, _ in fields] for jname, fields in G.parse(G.generate()).values()}
    for m in re.finditer(r"mk\((Ddp\w+);(.*?)\)\n", src, re.S):
        jname, body = m.group(1), m.group(2)
        assert jname in structs, jname
        # top-level `name = value` pairs of the keyword list
        depth, tok, keys = 0, "", []
        for ch in body:
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            if ch == "," and depth == 0:
                keys.append(tok); tok = ""
            else:
                tok += ch
        keys.append(tok)
        for kv in keys:
            if "=" in kv:
                key = kv.split("=", 1)[0].strip().split()[-1]
                if re.fullmatch(r"\w+", key):
                    assert key in structs[jname], (jname, key)
