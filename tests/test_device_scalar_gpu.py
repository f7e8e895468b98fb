"""Device-resident scalars, the halt flag and the device-scalar CG variant.

The variant must walk the same iterates as op::cg (reference flecsolve/solvers/cg.hh:44-139): same
iteration count and stop reason, residual history equal to rounding of the reduction order, and the
returned x is the iterate of the converging iteration although the host looks `lag` iterations late."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H

pytestmark = pytest.mark.gpu


def test_scalar_slots_and_device_coefficients(ctx):
    n = 10001
    rng = np.random.default_rng(0)
    xh, yh = rng.standard_normal(n), rng.standard_normal(n)
    x, y, z = ctx.vector(n), ctx.vector(n), ctx.vector(n)
    x.upload(xh); y.upload(yh)
    s, t = ctx.scalar(3.0), ctx.scalar(7.0)
    assert ctx.scalar_get(s) == 3.0 and ctx.scalar_get(t) == 7.0
    # z = (-1 * s/t) x + 2 y : one division, then the scaling, products and sum rounded separately
    z.linear_sum_c((-1.0, s, t), x, 2.0, y)
    a = -1.0 * (3.0 / 7.0)
    assert np.array_equal(z.download(), a * xh + 2.0 * yh)
    # aliased destination in the first operand (axpby form) keeps its coefficient
    z.upload(xh)
    z.linear_sum_c((1.0, t, s), z, 1.0, y)
    assert np.array_equal(z.download(), (7.0 / 3.0) * xh + yh)
    # a reduction deposits its value in a slot; the next statement reads it in a later launch
    ctx.reset_stats()
    tok = x.dot_opts_token(y, store=s)
    z.linear_sum_c((1.0, s, 0), x, 0.0, y)  # needs <x,y>: may not share the reduction's launch
    got = z.download()
    d = ctx.get(tok)
    assert ctx.scalar_get(s) == d
    assert ctx.stat("launches") == 2
    assert np.array_equal(got, d * xh + 0.0 * yh)
    assert abs(d - xh @ yh) <= 1e-12 * (np.abs(xh) @ np.abs(yh))
    with pytest.raises(F.FsbError):
        z.linear_sum_c((1.0, 60, 0), x, 0.0, y)  # never created
    ctx.scalar_destroy(s); ctx.scalar_destroy(t)
    for v in (x, y, z):
        v.destroy()


def test_halt_flag_freezes_vectors_but_tokens_still_arrive(ctx):
    n = 4097
    rng = np.random.default_rng(1)
    xh = rng.standard_normal(n)
    x, y = ctx.vector(n), ctx.vector(n)
    x.upload(xh); y.set_scalar(1.0)
    with pytest.raises(F.FsbError):
        x.dot_opts_token(x, halt_mode=F.HALT_IF_SQRT_LT, halt_threshold=1.0)  # not armed
    ctx.halt_arm()
    norm = float(np.sqrt(xh @ xh))
    t0 = x.dot_opts_token(x, halt_mode=F.HALT_IF_SQRT_LT, halt_threshold=0.5 * norm)  # does not pass
    y.axpy(2.0, x, y)
    t1 = x.dot_opts_token(x, halt_mode=F.HALT_IF_SQRT_LT, halt_threshold=2.0 * norm)  # passes: halt
    ctx.flush()  # the flag is looked at when a launch begins: end the launch that computes it
    y.axpy(5.0, x, y)  # skipped
    y.scale(3.0)  # skipped
    t2 = y.dot_token(y)  # token still delivered (value of a skipped pass: the fold identity)
    assert abs(ctx.get(t0) - xh @ xh) <= 1e-12 * (xh @ xh)
    assert abs(ctx.get(t1) - xh @ xh) <= 1e-12 * (xh @ xh)
    assert ctx.get(t2) == 0.0
    assert ctx.halt_disarm() is True
    assert np.array_equal(y.download(), 2.0 * xh + 1.0)
    y.scale(3.0)  # works again
    assert np.array_equal(y.download(), 3.0 * (2.0 * xh + 1.0))
    ctx.halt_arm()
    assert ctx.halt_disarm() is False
    x.destroy(); y.destroy()


def _system(kind, dims, scale_rows, seed=0):
    rp, col, val = O.stencil_csr(kind, *dims)
    n = len(rp) - 1
    if scale_rows:
        rng = np.random.default_rng(seed)
        d = 10 ** rng.uniform(-1.5, 1.5, n)
        A = (sp.diags(d) @ sp.csr_matrix((val, col, rp)) @ sp.diags(d)).tocsr()
        A.sort_indices()
        rp, col, val = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data
    return n, rp, col, val


@pytest.mark.parametrize("lag", [0, 1, 2, 5])
@pytest.mark.parametrize("precond", [None, "dinv"])
@pytest.mark.parametrize("kind,dims", [(5, (48, 48, 1)), (7, (20, 18, 16)), (27, (12, 12, 12))])
def test_cg_device_walks_the_same_iterates_as_cg(ctx, kind, dims, precond, lag):
    n, rp, col, val = _system(kind, dims, precond == "dinv")
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    S = H.Session(ctx, A)
    As = sp.csr_matrix((val, col, rp))
    b = As @ np.linspace(1, 2, n)
    x0 = np.random.default_rng(7).random(n)
    x, info, hist = S.solve(b, x0, solver="cg", precond=precond, rtol=1e-9, maxiter=2000, history_cap=2000)
    ctx.reset_stats()
    xd, dinfo, dhist = S.solve(b, x0, solver="cg_device", precond=precond, rtol=1e-9, maxiter=2000, history_cap=2000,
                               lag=lag)
    launches = ctx.stat("launches")
    assert dinfo.reason == info.reason == "converged_rtol"
    assert abs(dinfo.iters - info.iters) <= 1, (dinfo.iters, info.iters)
    assert dinfo.callbacks == dinfo.iters
    m = min(len(hist), len(dhist), 60)
    assert np.allclose(dhist[:m], hist[:m], rtol=1e-9)
    # x is the iterate of the converging iteration: its true residual matches the reported norm, and
    # running further ahead (lag) changed nothing
    res = np.linalg.norm(b - As @ xd)
    assert res <= 1.5 * np.float32(1e-9) * np.linalg.norm(b)
    assert abs(res - dinfo.res_norm_final) <= 1e-3 * res + 1e-9 * np.linalg.norm(b)
    if dinfo.iters == info.iters:
        assert np.allclose(xd, x, rtol=0, atol=1e-9 * np.abs(x).max())
    # 3 kernels per iteration (+ the prologue and the iterations issued past convergence)
    assert launches <= 3 * (dinfo.iters + lag) + 12
    assert abs(dinfo.sol_norm_final - np.linalg.norm(xd)) <= 1e-6 * np.linalg.norm(xd)
    S.close(); A.destroy()


@pytest.mark.parametrize("lag", [0, 2])
@pytest.mark.parametrize("precond", [None, "dinv"])
@pytest.mark.parametrize("kind,dims", [(5, (48, 48, 1)), (7, (20, 18, 16)), (27, (12, 12, 12))])
def test_single_reduction_cg(ctx, kind, dims, precond, lag):
    """solvers/cg_sr.hh: two launches per iteration (one fused vector pass, one SpMV that also carries <w,u> and the
    scalar recurrences), CG's iterates to rounding, same stop reason, iteration count +- 2"""
    n, rp, col, val = _system(kind, dims, precond == "dinv")
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    S = H.Session(ctx, A)
    As = sp.csr_matrix((val, col, rp))
    b = As @ np.linspace(1, 2, n)
    x0 = np.random.default_rng(7).random(n)
    x, info, hist = S.solve(b, x0, solver="cg", precond=precond, rtol=1e-9, maxiter=2000, history_cap=2000)
    ctx.reset_stats()
    xs, sinfo, shist = S.solve(b, x0, solver="cg_sr", precond=precond, rtol=1e-9, maxiter=2000, history_cap=2000, lag=lag)
    launches, unmatched = ctx.stat("launches"), ctx.stat("unmatched_groups")
    assert sinfo.reason == info.reason == "converged_rtol"
    assert abs(sinfo.iters - info.iters) <= 2, (sinfo.iters, info.iters)
    m = min(len(hist), len(shist)) * 2 // 3
    assert np.allclose(shist[:m], hist[:m], rtol=1e-6)
    res = np.linalg.norm(b - As @ xs)
    assert res <= 2 * np.float32(1e-9) * np.linalg.norm(b)
    assert launches <= 2 * (sinfo.iters + lag) + 14, launches
    assert unmatched <= 2  # the iteration's groups run through ahead-of-time kernels
    S.close(); A.destroy()


def test_scalar_statements_ride_on_a_reduction(ctx):
    """fsb_red_opts::post: slot arithmetic evaluated on the device once the reduction's value is stored"""
    n = 1000
    x = ctx.vector(n)
    x.set_scalar(2.0)
    a, b, c = ctx.scalar(3.0), ctx.scalar(0.0), ctx.scalar(0.0)
    launches = ctx.stat("launches")
    # b = <x,x> = 4000 ; then c = b / a ; a = a * a ; b = c - a ; c = copy of a
    t = x.dot_opts_token(x, store=b, post=[("div", c, b, a), ("mul", a, a, a), ("sub", b, c, a), ("copy", c, a, a)])
    assert ctx.get(t) == 4.0 * n
    assert ctx.stat("launches") - launches == 1
    assert ctx.scalar_get(a) == 9.0 and ctx.scalar_get(b) == 4000.0 / 3.0 - 9.0 and ctx.scalar_get(c) == 9.0
    with pytest.raises(F.FsbError):
        x.dot_opts_token(x, post=[("add", 0, a, a)])  # slot 0 is the constant 1
    for s in (a, b, c):
        ctx.scalar_destroy(s)
    x.destroy()


def test_cg_device_maxiter_user_stop_and_early_exit(ctx):
    n, rp, col, val = _system(7, (10, 10, 10), False)
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    S = H.Session(ctx, A)
    b = np.random.default_rng(3).random(n)
    _, info, hist = S.solve(b, None, solver="cg_device", rtol=1e-12, maxiter=3, history_cap=10)
    assert info.reason == "diverged_iters" and info.iters == 0 and len(hist) == 3
    _, ref, rhist = S.solve(b, None, solver="cg", rtol=1e-12, maxiter=3, history_cap=10)
    assert np.allclose(hist, rhist, rtol=1e-12)
    xs, _, _ = S.solve(b, None, solver="cg", rtol=1e-10, maxiter=1000)
    _, info, _ = S.solve(b, xs, solver="cg_device", rtol=1e-6, maxiter=1000)
    assert info.reason == "converged_rtol" and info.iters == 0 and info.callbacks == 0
    # vectors are usable afterwards (halt flag cleared)
    v = ctx.vector(16)
    v.set_scalar(2.0); v.scale(2.0)
    assert np.array_equal(v.download(), np.full(16, 4.0))
    v.destroy(); S.close(); A.destroy()
