import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import oracle as O
from flecsolve_b200 import _lib as F, host as H
sys.path.insert(0, 'tests')
from test_solvers_gpu import _system, _pair
ctx = F.Context(0)
n, rp, col, val = _system(7, (12, 11, 10), scale_rows=True)
A, S, M = _pair(ctx, n, rp, col, val)
b, x0 = M.set_random(0), M.set_random(1)
for precond in (None, "dinv"):
    dinv = M.dinv() if precond else None
    _, oinfo, ohist = M.gmres(b, x0=x0, dinv=dinv, history_cap=100, rtol=1e-4, maxiter=100)
    ctx.set_option("trace", 0)
    x, info, hist = S.solve(b, x0, solver="gmres", precond=precond, history_cap=100, rtol=1e-4, maxiter=100)
    print(precond, "oracle", oinfo.reason, oinfo.iters, "device", info.reason, info.iters, info.callbacks)
    print(" ohist", ohist[:6], ohist[-3:])
    print(" dhist", hist[:6], hist[-3:])
ctx.set_option("trace", 1)
x, info, hist = S.solve(b, x0, solver="gmres", precond="dinv", history_cap=100, rtol=1e-4, maxiter=3)
