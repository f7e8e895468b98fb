"""Worker for tests/test_multi_gpu.py: run under torchrun with one rank per GPU.  Every rank
checks its slice of the results against the serial oracle and exits non-zero on mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from flecsolve_b200 import _lib as F  # noqa: E402
from flecsolve_b200 import dist as D  # noqa: E402
from flecsolve_b200 import host as H  # noqa: E402


def main():
    world = D.init(D.world_from_env())
    ctx = D.make_context(world)
    P, me = world.size, world.rank
    rng = np.random.default_rng(0)

    # ---- general (host) path: 27-point operator with perturbed values, rows not divisible by P
    dims = (7, 6, 5 if P == 1 else 11)
    rp, col, val = O.stencil_csr(27, *dims)
    val = val * rng.uniform(0.5, 1.5, val.size)
    n = len(rp) - 1
    part = D.equal_map(n, P)
    lo, hi = part[me], part[me + 1]
    A = F.ParCSR.from_csr(ctx, n, part, rp[lo:hi + 1] - rp[lo], col[rp[lo]:rp[hi]], val[rp[lo]:rp[hi]])
    M = O.ParCSR(rp, col, val, colours=P)
    # split representation equals the oracle's colour `me` (topo/csr.hh color()/init_mats())
    for which in (0, 1):
        orp, ocol, oval, ocm = M.block(me, which)
        drp, dcol, dval = A.download(which)
        assert np.array_equal(drp, orp) and np.array_equal(dcol, ocol) and np.array_equal(dval, oval), ("split", which)
    assert np.array_equal(A.colmap(), M.block(me, 1)[3])
    x = rng.standard_normal(n)
    xv, yv = A.vector(x[lo:hi]), A.vector()
    A.spmv(xv, yv)
    ref = M.spmv(x)  # oracle does (diag.x) + (offd.x) per colour like parcsr.hh:61-68
    got = yv.download()
    assert np.array_equal(got, ref[lo:hi]), ("spmv", np.abs(got - ref[lo:hi]).max())
    # reductions across ranks
    d = xv.dot(yv)
    assert abs(d - x @ ref) <= 1e-12 * np.abs(x * ref).sum()
    assert xv.global_size() == n
    assert xv.max() == x.max() and xv.min() == x.min() and xv.inf_norm() == np.abs(x).max()
    # fused spmv + dot on several ranks
    A.spmv(xv, yv)
    t = yv.dot_token(xv)
    assert abs(ctx.get(t) - x @ ref) <= 1e-12 * np.abs(x * ref).sum()
    # solver parity on the distributed matrix
    S = H.Session(ctx, A)
    b = M.spmv(np.linspace(1, 2, n))
    for solver, fn in (("cg", M.cg), ("bicgstab", M.bicgstab)):
        xo, oinfo, _ = fn(b, dinv=M.dinv(), rtol=1e-9, maxiter=500)
        xs, info, _ = S.solve(b[lo:hi], np.zeros(hi - lo), solver=solver, precond="dinv", rtol=1e-9, maxiter=500)
        assert info.reason == oinfo.reason == "converged_rtol", (solver, info.reason, oinfo.reason)
        assert abs(info.iters - oinfo.iters) <= max(1, 0.05 * oinfo.iters), (solver, info.iters, oinfo.iters)
        assert np.abs(xs - xo[lo:hi]).max() <= 1e-6
    # weighted Jacobi sweeps need ghost values of x
    x0 = rng.standard_normal(n)
    w = float(np.float32(2 / 3))
    bv, xj, tv = A.vector(b[lo:hi]), A.vector(x0[lo:hi]), A.vector()
    A.jacobi_relax(w, 3, bv, xj, tv)
    refj = M.jacobi_relax(w, 3, b, x0)
    # owned-column entries (without the diagonal), then off-process entries, one running sum: the reference's order
    assert np.array_equal(xj.download(), refj[lo:hi]), np.abs(xj.download() - refj[lo:hi]).max()
    if P > 1 and os.environ.get("FSB_P2P_REDUCE") != "0" and os.environ.get("FSB_FUSED_HALO") != "0":
        assert A.info("fused_halo") == 1  # y = A x was ONE launch per rank, ghost exchange included
    S.close()
    A.destroy()

    # ---- device generator: row blocks that are not whole planes (equal_map gives the first N % P ranks one more row),
    # 27-point ghosts that are not a contiguous range, Neumann closure
    for gkind, gdims in ((7, (12, 11, 5 * P + 1)), (27, (9, 8, 4 * P + 1)), (107, (7, 9, 3 * P + 2)), (5, (13, 7 * P + 3, 1))):
        grp, gcol, gval = O.stencil_csr(gkind, *gdims)
        gn = len(grp) - 1
        GM = O.ParCSR(grp, gcol, gval, colours=P)
        GA = F.ParCSR.stencil(ctx, gkind, *gdims)
        glo, ghi = GA.row_begin, GA.row_begin + GA.local_rows
        assert [glo, ghi] == list(GM.partition()[me:me + 2]), (gkind, glo, ghi)
        for which in (0, 1):
            orp, ocol, oval, ocm = GM.block(me, which)
            drp, dcol, dval = GA.download(which)
            assert np.array_equal(drp, orp) and np.array_equal(dcol, ocol) and np.array_equal(dval, oval), ("gen", gkind, which)
        assert np.array_equal(GA.colmap(), GM.block(me, 1)[3]), ("colmap", gkind)
        gx = rng.standard_normal(gn)
        gxv, gyv = GA.vector(gx[glo:ghi]), GA.vector()
        GA.spmv(gxv, gyv)
        assert np.array_equal(gyv.download(), GM.spmv(gx)[glo:ghi]), ("gen spmv", gkind)
        gxv.destroy(); gyv.destroy(); GA.destroy()

    # ---- device generator path: plane-aligned z-slabs
    nn = 16
    rp, col, val = O.stencil_csr(7, nn, nn, nn * P if P > 1 else nn)
    n = len(rp) - 1
    M = O.ParCSR(rp, col, val, colours=P)
    B = F.ParCSR.stencil(ctx, 7, nn, nn, nn * P if P > 1 else nn)
    lo, hi = B.row_begin, B.row_begin + B.local_rows
    assert [lo, hi] == list(M.partition()[me:me + 2])
    for which in (0, 1):
        orp, ocol, oval, ocm = M.block(me, which)
        drp, dcol, dval = B.download(which)
        assert np.array_equal(drp, orp) and np.array_equal(dcol, ocol) and np.array_equal(dval, oval), ("gen", which)
    assert np.array_equal(B.colmap(), M.block(me, 1)[3])
    x = rng.standard_normal(n)
    xv, yv = B.vector(x[lo:hi]), B.vector()
    B.spmv(xv, yv)
    assert np.array_equal(yv.download(), M.spmv(x)[lo:hi])
    S = H.Session(ctx, B)
    b = M.spmv(np.linspace(1, 2, n))
    xo, oinfo, ohist = M.cg(b, dinv=M.dinv(), rtol=1e-9, maxiter=1000, history_cap=1000)
    xs, info, hist = S.solve(b[lo:hi], np.zeros(hi - lo), solver="cg", precond="dinv", rtol=1e-9, maxiter=1000,
                             history_cap=1000)
    assert info.reason == "converged_rtol" and abs(info.iters - oinfo.iters) <= max(1, 0.02 * oinfo.iters)
    m = min(len(hist), len(ohist), 30)
    assert np.allclose(hist[:m], ohist[:m], rtol=1e-8)
    assert ctx.stat("halo_exchanges") > 0 or P == 1
    # device-scalar CG: every rank raises its halt flag from the same all-rank value, in the same iteration
    xd, dinfo, dhist = S.solve(b[lo:hi], np.zeros(hi - lo), solver="cg_device", precond="dinv", rtol=1e-9, maxiter=1000,
                               history_cap=1000, lag=2)
    assert dinfo.reason == "converged_rtol" and abs(dinfo.iters - info.iters) <= 1, (dinfo.iters, info.iters)
    m = min(len(hist), len(dhist), 30)
    assert np.allclose(dhist[:m], hist[:m], rtol=1e-8)
    assert np.abs(xd - xo[lo:hi]).max() <= 1e-6
    # single-reduction CG: the scalar recurrences ride on the all-reduced <w,u>, identically on every rank
    xr, rinfo, rhist = S.solve(b[lo:hi], np.zeros(hi - lo), solver="cg_sr", precond="dinv", rtol=1e-9, maxiter=1000,
                               history_cap=1000, lag=2)
    assert rinfo.reason == "converged_rtol" and abs(rinfo.iters - info.iters) <= 2, (rinfo.iters, info.iters)
    assert np.abs(xr - xo[lo:hi]).max() <= 1e-6
    if P > 1 and os.environ.get("FSB_P2P_REDUCE") == "0":
        assert ctx.stat("allreduces") > 0  # the NCCL path really ran
    elif P > 1:
        assert ctx.stat("allreduces") == 0  # reductions went through peer memory inside the kernels
    S.close()
    B.destroy()

    # ---- BASELINE configs 3 and 4 sharded: implicit heat step (BDF2 + GMRES through operator_adapter) and BiCGStab on a
    # two-component vec::multi with the Dirichlet / Neumann+shift blocks -- against the same drivers on ONE rank
    # (a private single-rank context on this rank's GPU)
    solo = F.Context(world.local_rank)
    nn = 12
    dims = (nn, nn, nn * P if P > 1 else nn)
    nloc, N = None, dims[0] * dims[1] * dims[2]
    hgrid = 10.0 / (nn + 1)
    rng3 = np.random.default_rng(3)
    bfull = rng3.random(N)
    g = np.arange(N)
    gi, gj, gk = g % dims[0], (g // dims[0]) % dims[1], g // (dims[0] * dims[1])
    mid = lambda a, m: (5 * a >= 2 * m) & (5 * a < 3 * m)
    u0 = np.where(mid(gi, dims[0]) & mid(gj, dims[1]) & mid(gk, dims[2]), 50.0, 0.0)
    bdf = dict(method="BDF2", time_rtol=1e-2, time_atol=1e-4, initial_dt=1e-2, max_dt=1e-2, min_dt=1e-6, final_time=0.1,
               error_scaling="fixed-resolution", norm="inf", max_attempts=6)
    res = {}
    for tag, cx in (("sharded", ctx), ("solo", solo)):
        Ah = F.ParCSR.stencil(cx, 7, *dims, 0.0, -1.0 / (hgrid * hgrid))
        lo, hi = Ah.row_begin, Ah.row_begin + Ah.local_rows
        Sh = H.Session(cx, Ah)
        u, r, dts, good, iters = Sh.bdf_heat(u0[lo:hi], H.make_bdf_options(**bdf), solver="gmres", rtol=1e-6, maxiter=1000,
                                             max_krylov_dim=30, restart=True)
        A0 = F.ParCSR.stencil(cx, 7, *dims)
        A1 = F.ParCSR.stencil(cx, 107, *dims, 1e-3, 1.0)
        n0 = A0.local_rows
        bb = np.concatenate([np.zeros(n0), bfull[lo:hi]])
        xm, minfo, mhist = H.solve_multi2(cx, A0, A1, bb, np.full(2 * n0, 2.0), solver="bicgstab", rtol=1e-8, maxiter=500,
                                          history_cap=600)
        xc, cinfo, chist = H.solve_multi2(cx, A0, A1, bb, np.full(2 * n0, 2.0), solver="cg", rtol=1e-8, maxiter=500,
                                          history_cap=600)
        res[tag] = dict(lo=lo, hi=hi, u=u, steps=(r.steps, r.rejects, r.attempts), dts=dts, iters=iters, xm=xm, minfo=minfo,
                        xc=xc, cinfo=cinfo, chist=chist)
        Sh.close(); Ah.destroy(); A0.destroy(); A1.destroy()
    a, b1 = res["sharded"], res["solo"]
    lo, hi = a["lo"], a["hi"]
    assert a["steps"] == b1["steps"], (a["steps"], b1["steps"])
    assert np.allclose(a["dts"], b1["dts"], rtol=1e-6) and np.all(np.abs(a["iters"] - b1["iters"]) <= 1)
    assert np.abs(a["u"] - b1["u"][lo:hi]).max() <= 1e-6 * max(1.0, np.abs(b1["u"]).max())
    assert a["cinfo"].reason == b1["cinfo"].reason == "converged_rtol" and abs(a["cinfo"].iters - b1["cinfo"].iters) <= 1
    m = min(len(a["chist"]), len(b1["chist"]), 30)
    assert np.allclose(a["chist"][:m], b1["chist"][:m], rtol=1e-8)
    n1 = hi - lo
    for comp in range(2):
        full = b1["xc"][comp * N:(comp + 1) * N][lo:hi]
        assert np.abs(a["xc"][comp * n1:(comp + 1) * n1] - full).max() <= 1e-6
    assert a["minfo"].reason == b1["minfo"].reason == "converged_rtol"
    assert abs(a["minfo"].iters - b1["minfo"].iters) <= max(2, 0.1 * b1["minfo"].iters)
    for comp in range(2):
        full = b1["xm"][comp * N:(comp + 1) * N][lo:hi]
        assert np.abs(a["xm"][comp * n1:(comp + 1) * n1] - full).max() <= 1e-5
    solo.close()
    ctx.close()
    D.finalize(world)
    print(f"rank {me}/{P}: multi-gpu checks passed", flush=True)


if __name__ == "__main__":
    main()
