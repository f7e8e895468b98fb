#!/usr/bin/env python
"""BASELINE.json configs 1, 3, 4 on one GPU through the C++ host layer: timings + size-independent
checks at full size.  One JSON line per config on stdout (not part of the bench.py contract)."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flecsolve_b200 import _lib as F  # noqa: E402
from flecsolve_b200 import host as H  # noqa: E402


def config1_poisson2d(ctx, n=256):
    """examples/poisson: 2-D 5-point Poisson, unpreconditioned CG, rtol 1e-9, x0 = mt19937(7);
    f = 8 pi^2 sin(2 pi x) sin(2 pi y) h^2, u = sin(2 pi x) sin(2 pi y) (poisson.cc:27-84,170-177)."""
    out = {}
    errs = []
    for m in (n // 4, n // 2, n):
        A = F.ParCSR.stencil(ctx, 5, m, m, 1)
        S = H.Session(ctx, A)
        h = 1.0 / (m + 1)
        xs = (np.arange(m) + 1) * h
        X, Y = np.meshgrid(xs, xs, indexing="xy")
        u = (np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y)).ravel()
        b = 8 * np.pi ** 2 * u * h * h
        S.x.set_random(7)
        x0 = S.x.download()
        t0 = time.perf_counter()
        x, info, _ = S.solve(b, x0, solver="cg", rtol=1e-9, maxiter=1000 * (m // 64 + 1))
        dt = time.perf_counter() - t0
        errs.append(float(np.abs(x - u).max()))
        out = {"n": m, "iters": info.iters, "status": info.reason, "seconds": dt, "it_per_s": info.iters / dt}
        S.close(); A.destroy()
    rate = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    out.update(config="C1 poisson 2-D 5-pt CG", max_err=errs, convergence_order=rate, ok=bool(min(rate) > 1.8))
    return out


def config3_heat(ctx, nn=256, steps=10):
    """3-D heat equation, BDF2 (CN start) + restarted GMRES(50) rtol 1e-6 on (I - gamma L)."""
    length = 10.0
    h = length / (nn + 1)
    A = F.ParCSR.stencil(ctx, 7, nn, nn, nn, 0.0, -1.0 / (h * h))
    S = H.Session(ctx, A)
    g = np.arange(nn ** 3)
    i, j, k = g % nn, (g // nn) % nn, g // (nn * nn)
    mid = lambda a: (5 * a >= 2 * nn) & (5 * a < 3 * nn)
    u0 = np.where(mid(i) & mid(j) & mid(k), 50.0, 0.0)
    opts = H.make_bdf_options(method="BDF2", time_rtol=1e-2, time_atol=1e-4, initial_dt=1e-2, max_dt=1e-2, min_dt=1e-6,
                              final_time=0.1, error_scaling="fixed-resolution", norm="inf", max_attempts=steps)
    S.bdf_heat(u0, H.make_bdf_options(method="BDF2", time_rtol=1e-2, time_atol=1e-4, initial_dt=1e-2, max_dt=1e-2,
                                      min_dt=1e-6, final_time=0.1, max_attempts=1), solver="gmres", rtol=1e-6,
               maxiter=10000, max_krylov_dim=50, restart=True)  # warm-up: allocations
    ctx.sync(); ctx.reset_stats()
    t0 = time.perf_counter()
    u, res, dts, good, iters = S.bdf_heat(u0, opts, solver="gmres", rtol=1e-6, maxiter=10000, max_krylov_dim=50, restart=True)
    dt = time.perf_counter() - t0
    out = {"config": "C3 heat 256^3 BDF2 + GMRES(50)", "attempts": res.attempts, "accepted": res.steps, "rejects": res.rejects,
           "inner_iterations": res.inner_iterations, "seconds_with_host_copies": dt,
           "solve_seconds": res.solve_ms * 1e-3, "attempts_per_s": res.attempts / (res.solve_ms * 1e-3),
           "inner_it_per_s": res.inner_iterations / (res.solve_ms * 1e-3), "launches": ctx.stat("launches"),
           "generic_groups": ctx.stat("unmatched_groups"), "u_max": float(u.max()), "u_min": float(u.min()),
           "heat": float(u.sum() * h ** 3), "heat0": float(u0.sum() * h ** 3),
           "ok": bool(u.min() >= -1e-9 and u.max() <= 50.0 * (1 + 1e-12) and u.sum() <= u0.sum() * (1 + 1e-12) and res.steps > 0)}
    S.close(); A.destroy()
    return out


def config4_multi(ctx, nn=256):
    """two-component vec::multi, block-diagonal operator {7-pt Dirichlet, 7-pt + 1e-3 I}, BiCGStab rtol 1e-6."""
    n = nn ** 3
    A0 = F.ParCSR.stencil(ctx, 7, nn, nn, nn)
    A1 = F.ParCSR.stencil(ctx, 7, nn, nn, nn, 1e-3, 1.0)
    rng = np.random.default_rng(3)
    b = np.concatenate([np.zeros(n), rng.random(n)])
    x0 = np.full(2 * n, 2.0)
    H.solve_multi2(ctx, A0, A1, b, x0, solver="bicgstab", rtol=1e-6, maxiter=3)  # warm-up
    ctx.sync(); ctx.reset_stats()
    t0 = time.perf_counter()
    x, info, hist = H.solve_multi2(ctx, A0, A1, b, x0, solver="bicgstab", rtol=1e-6, maxiter=500, history_cap=600)
    dt = time.perf_counter() - t0
    # true residual through the device operators
    r = []
    for A, sl in ((A0, slice(0, n)), (A1, slice(n, 2 * n))):
        xv, yv = A.vector(x[sl]), A.vector()
        A.spmv(xv, yv)
        r.append(b[sl] - yv.download())
        xv.destroy(); yv.destroy()
    rel = float(np.linalg.norm(np.concatenate(r)) / np.linalg.norm(b))
    out = {"config": "C4 2-component vec::multi 256^3 BiCGStab", "iters": info.iters, "status": info.reason, "seconds_with_host_copies": dt,
           "solve_seconds": info.solve_ms * 1e-3, "it_per_s": info.iters / (info.solve_ms * 1e-3), "true_rel_residual": rel, "launches": ctx.stat("launches"),
           "generic_groups": ctx.stat("unmatched_groups"), "ok": bool(info.reason == "converged_rtol" and rel < 5e-6)}
    A0.destroy(); A1.destroy()
    return out


def main():
    ctx = F.Context(0)
    small = "--small" in sys.argv
    for fn, kw in ((config1_poisson2d, dict(n=64 if small else 256)), (config3_heat, dict(nn=48 if small else 256)),
                   (config4_multi, dict(nn=48 if small else 256))):
        print(json.dumps(fn(ctx, **kw)), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
