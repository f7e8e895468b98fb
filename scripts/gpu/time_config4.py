import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flecsolve_b200 import _lib as F, host as H
ctx = F.Context(0)
nn = 256; n = nn ** 3
A0 = F.ParCSR.stencil(ctx, 7, nn, nn, nn)
A1 = F.ParCSR.stencil(ctx, 7, nn, nn, nn, 1e-3, 1.0)
rng = np.random.default_rng(3)
b = np.concatenate([rng.random(n) if "--rand0" in sys.argv else np.zeros(n), rng.random(n)])
x0 = np.full(2 * n, 2.0)
it = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H.solve_multi2(ctx, A0, A1, b, x0, solver="bicgstab", rtol=1e-6, maxiter=3)
ctx.sync()
if "--trace" in sys.argv:
    ctx.set_option("trace", 1)
ctx.reset_stats()
t0 = time.perf_counter()
x, info, hist = H.solve_multi2(ctx, A0, A1, b, x0, solver="cg" if "--cg" in sys.argv else "bicgstab", rtol=1e-30 if "--cg" in sys.argv else 1e-6, maxiter=it)
print("seconds", time.perf_counter() - t0, info.iters, file=sys.stderr)
print({k: ctx.stat(k) for k in ("launches", "host_syncs", "wait_ns", "flush_ns")}, file=sys.stderr)
