"""Scratch: host-side cost of a CG iteration (flush = statement queue -> launches, wait = polling for reduction results)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from bench import WORKLOADS, x_true
from flecsolve_b200 import _lib as F, host as H, dist as D
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 256
world = D.init(D.world_from_env())  # one rank, or the ranks torchrun started
ctx = D.make_context(world)
A = F.ParCSR.stencil(ctx, 7, nn, nn, nn)
S = H.Session(ctx, A)
xt = A.vector(x_true(A.row_begin, A.local_rows)); A.spmv(xt, S.b); ctx.sync()
NAMES = {9: 'bind_ns', 10: 'ew_launch_ns', 11: 'spmv_launch_ns', 13: 'armed_hits', 14: 'armed_misses'}
def raw(k):
    return ctx.stat(NAMES[k])
for solver in ("cg", "cg_device"):
    S.x.zero(); S.solve(solver=solver, precond="dinv", maxiter=50, rtol=0.0, lag=2)
    S.x.zero(); ctx.sync(); ctx.reset_stats()
    t0 = time.perf_counter()
    S.solve(solver=solver, precond="dinv", maxiter=400, rtol=0.0, lag=2)
    ctx.sync(); dt = time.perf_counter() - t0
    it = 400
    if world.is_root:
      print(f"{nn}^3 on {world.size} rank(s) {solver}: {dt/it*1e6:.1f} us/iteration wall; flush {ctx.stat('flush_ns')/it/1e3:.2f} us, wait {ctx.stat('wait_ns')/it/1e3:.2f} us, "
          f"host syncs {ctx.stat('host_syncs')/it:.2f}, launches {ctx.stat('launches')/it:.2f} per iteration; "
          f"of the flush: bind {raw(9)/it/1e3:.2f}, ew launch call {raw(10)/it/1e3:.2f}, spmv prepare+launch {raw(11)/it/1e3:.2f} us; "
          f"armed hits {raw(13)/it:.2f}, misses {raw(14)/it:.3f} per iteration")
