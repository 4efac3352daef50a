#!/usr/bin/env python
"""BASELINE configs[0] as the reference runs it: ONE trajectory, demo_linear n=10 m=2 T=1000, whole iLQG solve (iLQG.jl:143-341)
through the Python mirror of the reference call (upload + device-resident solve + download), wall clock."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ddp_b200 as ddp
from helpers import make_lq

n, m, T = 10, 2, 1000
rng = np.random.default_rng(0)
A, Bm, Q, R = make_lq(rng, n, m, h=0.01)
x0 = np.ones(n)
u0 = rng.standard_normal((T, m))
model = ddp.LinearModel(A, Bm, Q, R)
out = []
for rep in range(4):
    t0 = time.perf_counter()
    x, u, traj, Vx, Vxx, cost, trace = ddp.iLQG(model.f, model.costfun, model.df, x0, u0)
    dt = time.perf_counter() - t0
    out.append(dt)
print(json.dumps(dict(n=n, m=m, T=T, wall_s=out, best_s=min(out), cost=float(np.sum(cost)), iters=int(np.atleast_1d(trace["iter"])[0]),
                      status=int(np.atleast_1d(trace["status"])[0]))))
