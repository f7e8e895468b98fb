import sys, time, numpy as np
sys.path.insert(0, ".")
from flecsolve_b200 import _lib as F, host as H
ctx = F.Context(0)
nn = 256; n = nn ** 3
A0 = F.ParCSR.stencil(ctx, 7, nn, nn, nn)
rng = np.random.default_rng(3)
def log(*a): print(*a, file=sys.stderr, flush=True)
# A: primitive loop
x0, y0, r0 = A0.vector(rng.random(n)), A0.vector(), A0.vector(rng.random(n))
for chunk in range(5):
    ctx.sync(); t0 = time.perf_counter()
    for i in range(300):
        A0.spmv(x0, y0)
        t = r0.dot_token(y0)
        ctx.get(t)
    log("A spmv+dot 300x", time.perf_counter() - t0)
# B: single-vector solvers, device resident
S = H.Session(ctx, A0)
b = rng.random(n)
for solver in ("bicgstab", "cg"):
    for it in (100, 200, 400, 800):
        S.b.upload(b); S.x.set_scalar(0.0); ctx.sync()
        t0 = time.perf_counter()
        _, info, _ = S.solve(solver=solver, precond="identity", rtol=1e-30, maxiter=it)
        ctx.sync()
        log("B", solver, it, time.perf_counter() - t0, info.iters, info.res_norm_final)
