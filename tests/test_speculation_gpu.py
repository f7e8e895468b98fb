"""Armed launches (FSB_OPT_SPECULATE, csrc/fuser.cu): a statement group launched ahead of the reduction its coefficients come
from runs the same kernel on the same coefficients -- results are bit-identical to ordinary launches; a request that does
not match tells the armed kernel to leave; nothing stays armed behind a call that touches the stream."""
import os

import numpy as np
import pytest

import oracle as O
from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("solver", ["cg", "bicgstab", "gmres"])
def test_solver_results_do_not_depend_on_it(ctx, solver):
    dims = (24, 22, 26)
    rp, col, val = O.stencil_csr(7, *dims)
    n = len(rp) - 1
    b = O.csr_spmv(rp, col, val, np.linspace(1.0, 2.0, n))
    A = F.ParCSR.stencil(ctx, 7, *dims)
    S = H.Session(ctx, A)
    runs = {}
    for mode in ("1", "0"):
        os.environ["FSB_SPECULATE"] = mode
        try:
            ctx.reset_stats()
            kw = dict(max_krylov_dim=30, restart=True) if solver == "gmres" else {}
            x, info, hist = S.solve(b, np.zeros(n), solver=solver, precond="dinv", rtol=1e-10, maxiter=400, history_cap=400, **kw)
            runs[mode] = (x.copy(), info.iters, info.reason, np.array(hist), ctx.stat("armed_hits"), ctx.stat("armed_misses"))
        finally:
            os.environ.pop("FSB_SPECULATE", None)
    on, off = runs["1"], runs["0"]
    assert on[2] == off[2] == "converged_rtol" and on[1] == off[1]
    assert np.array_equal(on[0], off[0]) and np.array_equal(on[3], off[3])
    assert off[4] == 0 and off[5] == 0
    if solver == "cg":  # three of the four launches of an iteration
        assert on[4] >= 2.5 * (on[1] - 3) and on[5] <= 6, (on[4], on[5], on[1])
    S.close(); A.destroy()


def test_hit_miss_and_release_at_the_c_abi(ctx):
    """a hand-written loop on the vector calls: the third time round the update is launched ahead; a different statement
    makes it leave; downloads, syncs and switching the option off never leave an armed kernel behind"""
    n = 50001
    rng = np.random.default_rng(3)
    xh, ph, rh, wh = (rng.standard_normal(n) for _ in range(4))
    mk = lambda a: F.Vector(ctx, n, 0).upload(a)
    x, p, r, w = mk(xh), mk(ph), mk(rh), mk(wh)
    ctx.set_option("speculate", 1)
    try:
        _hit_miss_and_release(ctx, n, xh, ph, rh, wh, x, p, r, w)
    finally:
        ctx.set_option("speculate", 0)  # the shared context must not keep launching ahead behind a failed assertion
    for v in (x, p, r, w):
        v.destroy()


def _hit_miss_and_release(ctx, n, xh, ph, rh, wh, x, p, r, w):
    ctx.reset_stats()
    ref_x, ref_r = xh.copy(), rh.copy()
    rho = ctx.get(r.dot_token(r))
    for it in range(8):
        alpha = 1.0 / (1.0 + rho)  # a coefficient the host computes from the reduction it waited for
        x.linear_sum(alpha, p, 1.0, x)
        r.linear_sum(-alpha, w, 1.0, r)
        ref_x = alpha * ph + 1.0 * ref_x
        ref_r = -alpha * wh + 1.0 * ref_r
        rho = ctx.get(r.dot_token(r))
    hits = ctx.stat("armed_hits")
    assert hits >= 4, hits  # learnt, confirmed, launched ahead from the fourth round on
    assert np.array_equal(x.download(), ref_x) and np.array_equal(r.download(), ref_r)  # download released what was armed
    misses = ctx.stat("armed_misses")
    assert misses >= 1
    # the loop once more, then a different statement where the update was expected
    for it in range(3):
        alpha = 0.5
        x.linear_sum(alpha, p, 1.0, x)
        r.linear_sum(-alpha, w, 1.0, r)
        ref_x = alpha * ph + 1.0 * ref_x
        ref_r = -alpha * wh + 1.0 * ref_r
        rho = ctx.get(r.dot_token(r))
    w.scale(2.0)  # not what follows that wait
    assert ctx.get(w.dot_token(w)) == pytest.approx(4.0 * float(wh @ wh), rel=1e-13)
    assert ctx.stat("armed_misses") > misses
    assert np.array_equal(x.download(), ref_x) and np.array_equal(r.download(), ref_r)
    rho = ctx.get(r.dot_token(r))  # arms again ...
    ctx.set_option("speculate", 0)  # ... and this lets it go
    ctx.sync()
    assert np.array_equal(w.download(), 2.0 * wh)
