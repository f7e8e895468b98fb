#!/usr/bin/env python
"""BASELINE.json configs 1, 3, 4 through the C++ host layer on 1 GPU or row-sharded over N (torchrun): timings,
size-independent checks at full size, and the fraction of the measured HBM roofline each solve sustains
(algorithmic bytes of SURVEY.md 8(d) / device time).  One JSON line per config on stdout (rank 0).

    python scripts/run_configs.py [--small]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/run_configs.py [--small]
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flecsolve_b200 import _lib as F  # noqa: E402
from flecsolve_b200 import dist as D  # noqa: E402
from flecsolve_b200 import host as H  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def config1_poisson2d(world, ctx, n=256):
    """examples/poisson: 2-D 5-point Poisson, unpreconditioned CG, rtol 1e-9, x0 = mt19937(7);
    f = 8 pi^2 sin(2 pi x) sin(2 pi y) h^2, u = sin(2 pi x) sin(2 pi y) (poisson.cc:27-84,170-177)."""
    out = {}
    errs = []
    for m in (n // 4, n // 2, n):
        A = F.ParCSR.stencil(ctx, 5, m, m, 1)
        S = H.Session(ctx, A)
        lo, hi = A.row_begin, A.row_begin + A.local_rows
        h = 1.0 / (m + 1)
        xs = (np.arange(m) + 1) * h
        X, Y = np.meshgrid(xs, xs, indexing="xy")
        u = (np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y)).ravel()
        b = 8 * np.pi ** 2 * u * h * h
        # the reference draws x0 per colour with the same seed (topo_tasks.hh:305-314)
        S.x.set_random(7)
        x0 = S.x.download()
        t0 = time.perf_counter()
        x, info, _ = S.solve(b[lo:hi], x0, solver="cg", rtol=1e-9, maxiter=1000 * (m // 64 + 1))
        dt = D.max_over_ranks(world, time.perf_counter() - t0)
        errs.append(D.max_over_ranks(world, float(np.abs(x - u[lo:hi]).max())))
        nnz, N = D.sum_over_ranks(world, A.nnz(0) + A.nnz(1)), m * m
        out = {"n": m, "iters": info.iters, "status": info.reason, "seconds": dt, "it_per_s": info.iters / dt,
               "ms_per_iteration_device": info.solve_ms / max(info.iters, 1),
               "roofline_frac": (12 * nnz + 92 * N) / world.size / (info.solve_ms / max(info.iters, 1)) / 1e6 / peak_gbs()}
        S.close(); A.destroy()
    rate = [float(np.log2(errs[i] / errs[i + 1])) for i in range(2)]
    out.update(config="C1 poisson 2-D 5-pt CG", n_gpus=world.size, max_err=errs, convergence_order=rate, ok=bool(min(rate) > 1.8),
               roofline_note="12 nnz + 92 N bytes per iteration (SURVEY 8d, unpreconditioned CG); a 256^2 system is 0.6 MB: "
                             "launch- and host-read-bound, not HBM-bound")
    return out


def config3_heat(world, ctx, nn=256, steps=10):
    """3-D heat equation, BDF2 (CN start) + restarted GMRES(50) rtol 1e-6 on (I - gamma L)."""
    length = 10.0
    h = length / (nn + 1)
    A = F.ParCSR.stencil(ctx, 7, nn, nn, nn, 0.0, -1.0 / (h * h))
    S = H.Session(ctx, A)
    lo, hi = A.row_begin, A.row_begin + A.local_rows
    g = np.arange(lo, hi)
    i, j, k = g % nn, (g // nn) % nn, g // (nn * nn)
    mid = lambda a: (5 * a >= 2 * nn) & (5 * a < 3 * nn)
    u0 = np.where(mid(i) & mid(j) & mid(k), 50.0, 0.0)
    kw = dict(method="BDF2", time_rtol=1e-2, time_atol=1e-4, initial_dt=1e-2, max_dt=1e-2, min_dt=1e-6, final_time=0.1)
    opts = H.make_bdf_options(error_scaling="fixed-resolution", norm="inf", max_attempts=steps, **kw)
    # warm-up: allocations, and the one-off run-time compilation of the integrator's statement groups that have no
    # ahead-of-time kernel (cached per process)
    S.bdf_heat(u0, opts, solver="gmres", rtol=1e-6, maxiter=10000, max_krylov_dim=50, restart=True)
    ctx.sync(); ctx.reset_stats()
    D.barrier(world)
    t0 = time.perf_counter()
    u, res, dts, good, iters = S.bdf_heat(u0, opts, solver="gmres", rtol=1e-6, maxiter=10000, max_krylov_dim=50, restart=True)
    dt = D.max_over_ranks(world, time.perf_counter() - t0)
    solve_s = D.max_over_ranks(world, res.solve_ms * 1e-3)
    # bytes: GMRES inner step with k prior basis vectors = SpMV + (32 k + 24) N; the adapter's axpy (+24 N) rides on
    # every product; a solve of m iterations walks k = 0 .. m-1 (restart at 50)
    N = nn ** 3
    nnz = D.sum_over_ranks(world, A.nnz(0) + A.nnz(1))
    spmv = 12 * nnz + 4 * N + 16 * N
    gm = 0
    for m in iters:
        for kk in range(int(m)):
            gm += spmv + 24 * N + (32 * (kk % 50) + 24) * N
        gm += spmv + 48 * N  # initial residual of the solve and the correction z = sum y_i q_i (first restart block), roughly
    hsum = lambda v: D.sum_over_ranks(world, float(v))
    heat, heat0 = hsum(u.sum() * h ** 3), hsum(u0.sum() * h ** 3)
    umax, umin = D.max_over_ranks(world, float(u.max())), -D.max_over_ranks(world, float(-u.min()))
    out = {"config": f"C3 heat {nn}^3 BDF2 + GMRES(50)", "n_gpus": world.size, "attempts": res.attempts, "accepted": res.steps,
           "rejects": res.rejects, "inner_iterations": res.inner_iterations, "seconds_with_host_copies": dt,
           "solve_seconds": solve_s, "attempts_per_s": res.attempts / solve_s, "inner_it_per_s": res.inner_iterations / solve_s,
           "launches": ctx.stat("launches"), "generic_groups": ctx.stat("unmatched_groups"), "jit_groups": ctx.stat("jit_groups"),
           "u_max": umax, "u_min": umin, "heat": heat, "heat0": heat0,
           "algorithmic_bytes": gm, "roofline_frac": gm / world.size / solve_s / 1e9 / peak_gbs(),
           "roofline_note": "sum over inner iterations of SpMV + adapter axpy + (32 k + 24) N (modified Gram-Schmidt with k "
                            "prior vectors), per SURVEY 8d; integrator overhead (source terms, error norms) is in the time, "
                            "not in the bytes",
           "ok": bool(umin >= -1e-9 and umax <= 50.0 * (1 + 1e-12) and heat <= heat0 * (1 + 1e-12) and res.steps > 0)}
    S.close(); A.destroy()
    return out


def config4_multi(world, ctx, nn=256):
    """examples/equilibrium_diffusion's shape: two-component vec::multi, block-diagonal operator (the two variables couple
    only through the shared Krylov scalars): {7-pt with Dirichlet closure, 7-pt with Neumann closure + 1e-3 I},
    RHS 0 / random(3), x0 = 2, BiCGStab rtol 1e-6, maxiter 500."""
    N = nn ** 3
    A0 = F.ParCSR.stencil(ctx, 7, nn, nn, nn)
    A1 = F.ParCSR.stencil(ctx, 107, nn, nn, nn, 1e-3, 1.0)
    lo, n = A0.row_begin, A0.local_rows
    rng = np.random.default_rng(3)
    b1 = rng.random(N)[lo:lo + n]
    b = np.concatenate([np.zeros(n), b1])
    x0 = np.full(2 * n, 2.0)
    H.solve_multi2(ctx, A0, A1, b, x0, solver="bicgstab", rtol=1e-6, maxiter=8)  # warm-up (allocations, run-time kernels)
    ctx.sync(); ctx.reset_stats()
    D.barrier(world)
    t0 = time.perf_counter()
    x, info, hist = H.solve_multi2(ctx, A0, A1, b, x0, solver="bicgstab", rtol=1e-6, maxiter=500, history_cap=600)
    dt = D.max_over_ranks(world, time.perf_counter() - t0)
    solve_s = D.max_over_ranks(world, info.solve_ms * 1e-3)
    # true residual through the device operators
    r2 = 0.0
    for A, sl in ((A0, slice(0, n)), (A1, slice(n, 2 * n))):
        xv, yv = A.vector(x[sl]), A.vector()
        A.spmv(xv, yv)
        r2 += float(np.sum((b[sl] - yv.download()) ** 2))
        xv.destroy(); yv.destroy()
    rel = float(np.sqrt(D.sum_over_ranks(world, r2)) / np.sqrt(D.sum_over_ranks(world, float(b @ b))))
    nnz = D.sum_over_ranks(world, A0.nnz(0) + A0.nnz(1))
    per_it = 2 * (2 * (12 * nnz + 4 * N) + 200 * N)  # 2 (12 nnz + 4 N) + 200 N per component (identity preconditioner: -16 N)
    out = {"config": f"C4 2-component vec::multi {nn}^3 BiCGStab (Dirichlet + Neumann/shift blocks)", "n_gpus": world.size,
           "iters": info.iters, "status": info.reason, "seconds_with_host_copies": dt, "solve_seconds": solve_s,
           "it_per_s": info.iters / solve_s, "true_rel_residual": rel, "launches": ctx.stat("launches"),
           "generic_groups": ctx.stat("unmatched_groups"), "jit_groups": ctx.stat("jit_groups"),
           "roofline_frac": per_it * max(info.iters, 1) / world.size / solve_s / 1e9 / peak_gbs(),
           "roofline_note": "2 components x (2 (12 nnz + 4 N) + 200 N) bytes per BiCGStab iteration (SURVEY 8d)",
           "ok": bool(info.reason == "converged_rtol" and rel < 5e-6)}
    A0.destroy(); A1.destroy()
    return out


def main():
    world = D.init(D.world_from_env())
    ctx = D.make_context(world)
    small = "--small" in sys.argv
    for fn, kw in ((config1_poisson2d, dict(n=64 if small else 256)), (config3_heat, dict(nn=48 if small else 256)),
                   (config4_multi, dict(nn=48 if small else 256))):
        r = fn(world, ctx, **kw)
        if world.is_root:
            print(json.dumps(r), flush=True)
    ctx.close()
    D.finalize(world)


if __name__ == "__main__":
    main()
