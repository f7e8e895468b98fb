"""The sharded path on ONE GPU: the ranks of a group are threads of this process sharing the device
(fsb_ctx_create_group), so the colour split (topo/csr.hh:482-618), the ghost exchange inside the SpMV kernel, the
cross-rank all-reduce inside the reduction kernels and the solvers on top of them run -- and are held against the
oracle -- on a single-GPU box as well (tests/test_multi_gpu.py needs one GPU per rank).  Same checks as
tests/multi_gpu_worker.py, sizes small enough for the ranks' kernels to be resident together."""
import threading

import numpy as np
import pytest

import oracle as O
from flecsolve_b200 import _lib as F
from flecsolve_b200 import dist as D
from flecsolve_b200 import host as H

pytestmark = pytest.mark.gpu


def run_ranks(nranks, body):
    """body(ctx, rank) -> objects to destroy, on one thread per rank; re-raises the first failure.  Matrices are
    destroyed only after every rank has drained its stream: releasing device memory waits for the whole device, and a
    peer's kernel may be waiting for this rank's next launch."""
    if F.device_count() == 0:
        pytest.skip("no CUDA device")
    ctxs = F.Context.group(0, nranks)
    errors = [None] * nranks
    gate = threading.Barrier(nranks)

    def work(r):
        garbage = []
        try:
            garbage = body(ctxs[r], r) or []
        except BaseException as e:  # noqa: BLE001 - reported below
            errors[r] = e
            gate.abort()
            return
        ctxs[r].sync()
        try:
            gate.wait(timeout=60)
        except threading.BrokenBarrierError:
            return
        for obj in garbage:
            obj.close() if hasattr(obj, "close") else obj.destroy()

    import faulthandler
    import sys
    faulthandler.dump_traceback_later(25, file=sys.stderr)  # where every rank is, should one of them never return
    threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(nranks)]  # a stuck rank must not keep the process alive
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=30)
    faulthandler.cancel_dump_traceback_later()
    for e in errors:  # a rank that failed leaves its peers waiting in the next collective: report the cause first
        if e is not None:
            raise e
    hung = [t for t in threads if t.is_alive()]
    assert not hung, "a rank did not finish"
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("nranks", [2, 3])
def test_split_spmv_reductions_and_jacobi(nranks):
    P = nranks
    rng0 = np.random.default_rng(0)
    dims = (7, 6, 11)
    rp, col, val = O.stencil_csr(27, *dims)
    val = val * rng0.uniform(0.5, 1.5, val.size)
    n = len(rp) - 1
    part = D.equal_map(n, P)
    M = O.ParCSR(rp, col, val, colours=P)
    x = rng0.standard_normal(n)
    ref = M.spmv(x)
    b = M.spmv(np.linspace(1, 2, n))
    x0 = rng0.standard_normal(n)
    w = float(np.float32(2 / 3))
    refj = M.jacobi_relax(w, 3, b, x0)

    def body(ctx, me):
        lo, hi = part[me], part[me + 1]
        A = F.ParCSR.from_csr(ctx, n, part, rp[lo:hi + 1] - rp[lo], col[rp[lo]:rp[hi]], val[rp[lo]:rp[hi]])
        for which in (0, 1):  # the oracle's colour `me` (topo/csr.hh color() / init_mats())
            orp, ocol, oval, ocm = M.block(me, which)
            drp, dcol, dval = A.download(which)
            assert np.array_equal(drp, orp) and np.array_equal(dcol, ocol) and np.array_equal(dval, oval), ("split", which)
        assert np.array_equal(A.colmap(), M.block(me, 1)[3])
        assert A.info("fused_halo") == 1
        xv, yv = A.vector(x[lo:hi]), A.vector()
        launches = ctx.stat("launches")
        A.spmv(xv, yv)
        got = yv.download()
        assert ctx.stat("launches") - launches == 1  # ghost exchange included
        assert np.array_equal(got, ref[lo:hi]), np.abs(got - ref[lo:hi]).max()
        d = xv.dot(yv)
        assert abs(d - x @ ref) <= 1e-12 * np.abs(x * ref).sum()
        assert xv.global_size() == n
        assert xv.max() == x.max() and xv.min() == x.min() and xv.inf_norm() == np.abs(x).max()
        A.spmv(xv, yv)
        t = yv.dot_token(xv)
        assert abs(ctx.get(t) - x @ ref) <= 1e-12 * np.abs(x * ref).sum()
        assert ctx.stat("allreduces") == 0  # reductions crossed the ranks inside the kernels
        bv, xj, tv = A.vector(b[lo:hi]), A.vector(x0[lo:hi]), A.vector()
        A.jacobi_relax(w, 3, bv, xj, tv)
        assert np.array_equal(xj.download(), refj[lo:hi])
        # explicit ghost update (the copy plan of topo/csr.hh:237-245)
        A.halo_exchange(xv)
        ghosts = xv.download(A.local_rows + A.num_ghosts)[A.local_rows:]
        assert np.array_equal(ghosts, x[A.colmap()])
        return [xv, yv, bv, xj, tv, A]

    run_ranks(nranks, body)


@pytest.mark.parametrize("kind,dims,P", [(7, (12, 11, 13), 2), (27, (9, 8, 11), 2), (107, (7, 9, 8), 2),
                                         (7, (7, 5, 11), 3), (27, (5, 7, 10), 3)])  # 385 and 350 rows on 3 ranks: uneven blocks
def test_device_generator_and_solvers(kind, dims, P):
    rp, col, val = O.stencil_csr(kind, *dims, 1e-3 if kind == 107 else 0.0, 1.0)
    n = len(rp) - 1
    M = O.ParCSR(rp, col, val, colours=P)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(n)
    b = M.spmv(np.linspace(1, 2, n))
    xo, oinfo, ohist = M.cg(b, dinv=M.dinv(), rtol=1e-9, maxiter=1000, history_cap=1000)

    def body(ctx, me):
        B = F.ParCSR.stencil(ctx, kind, *dims, 1e-3 if kind == 107 else 0.0, 1.0)
        lo, hi = B.row_begin, B.row_begin + B.local_rows
        assert [lo, hi] == list(M.partition()[me:me + 2])
        for which in (0, 1):
            orp, ocol, oval, ocm = M.block(me, which)
            drp, dcol, dval = B.download(which)
            assert np.array_equal(drp, orp) and np.array_equal(dcol, ocol) and np.array_equal(dval, oval), ("gen", which)
        assert np.array_equal(B.colmap(), M.block(me, 1)[3])
        xv, yv = B.vector(x[lo:hi]), B.vector()
        B.spmv(xv, yv)
        assert np.array_equal(yv.download(), M.spmv(x)[lo:hi])
        S = H.Session(ctx, B)
        for solver, lag in (("cg", 0), ("cg_device", 2), ("cg_sr", 2)):
            xs, info, hist = S.solve(b[lo:hi], np.zeros(hi - lo), solver=solver, precond="dinv", rtol=1e-9, maxiter=1000,
                                     history_cap=1000, lag=lag)
            assert info.reason == "converged_rtol" and abs(info.iters - oinfo.iters) <= (2 if solver == "cg_sr" else 1), solver
            m = min(len(hist), len(ohist), 25)
            assert np.allclose(hist[:m], ohist[:m], rtol=1e-6 if solver == "cg_sr" else 1e-8), solver
            assert np.abs(xs - xo[lo:hi]).max() <= 1e-6
        xb, binfo, _ = S.solve(b[lo:hi], np.zeros(hi - lo), solver="bicgstab", precond="dinv", rtol=1e-9, maxiter=1000)
        assert binfo.reason == "converged_rtol" and np.abs(xb - xo[lo:hi]).max() <= 1e-6
        assert ctx.stat("halo_exchanges") > 0
        return [S, xv, yv, B]

    run_ranks(P, body)


@pytest.mark.parametrize("nranks,dims", [(2, (9, 7, 10)), (3, (6, 5, 11)), (2, (11, 8))])
def test_structured_grid_operators_over_slabs(nranks, dims):
    """narray fields cut into slabs along the last axis, one padded array per rank: the pad plane facing a
    neighbour is a ghost plane refreshed by the operator (FleCSI's ghost copy) -- the stencil and the assembled finite-volume
    operator equal the one-array results (stencil: bit for bit), reductions span the ranks, and a CG written on the C ABI's
    vector calls converges to the same iterate as on one rank"""
    P = nranks
    dim = len(dims)
    rng = np.random.default_rng(dim * 10 + P)
    ext = tuple(n + 2 for n in dims)  # global padded array, x first
    gshape = ext[::-1]
    lo, hi = (1,) * dim, tuple(n + 1 for n in dims)
    ug = rng.standard_normal(gshape)  # boundary layers hold data too
    a = rng.uniform(0.5, 2.0, gshape)
    bface = [rng.uniform(0.5, 2.0, gshape) for _ in range(dim)]
    kface = list(rng.uniform(0.5, 2.0, dim))
    beta, alpha, vol = 1.3, 0.7, 0.11
    off = [-1.0, -1.25, -0.5][:dim]
    center = 2.0 * sum(-o for o in off) + 0.3
    box = tuple(slice(1, n + 1) for n in dims[::-1])
    # one-array references
    sten = np.zeros(gshape)
    terms = [(off[ax], np.roll(ug, 1, axis=dim - 1 - ax)) for ax in reversed(range(dim))] + [(center, ug)]
    terms += [(off[ax], np.roll(ug, -1, axis=dim - 1 - ax)) for ax in range(dim)]
    for c, arr in terms:
        sten = sten + c * arr
    fvm = O.fvm_diffusion_apply(ug, a, bface, beta, alpha, vol, kface, lo, hi)
    # slabs of the last mesh axis = numpy axis 0
    cuts = D.equal_map(dims[-1], P)
    bvec = rng.standard_normal(gshape)

    def cg(ctx, A, mk, b, iters):  # unpreconditioned CG on box vectors through the C ABI
        x, r, p, w = mk(), mk(), mk(), mk()
        x.set_scalar(0.0); r.copy(b); p.copy(b)
        rho = ctx.get(r.dot_token(r))
        hist = [rho]
        for _ in range(iters):
            A.spmv(p, w)
            alpha_ = rho / ctx.get(w.dot_token(p))
            x.linear_sum(alpha_, p, 1.0, x)
            r.linear_sum(-alpha_, w, 1.0, r)
            rho_new = ctx.get(r.dot_token(r))
            hist.append(rho_new)
            p.linear_sum(rho_new / rho, p, 1.0, r)
            rho = rho_new
        return x, hist, [r, p, w]

    # the same CG on ONE rank (whole array) as the reference for the solver check
    solo = F.Context(0)
    As = solo.box_stencil(ext, lo, hi, center, off)
    mk_s = lambda: solo.box_vector(ext, lo, hi).upload_all(np.zeros(ug.size))
    bs = mk_s().upload_all(bvec.ravel())
    xs, hist_solo, junk = cg(solo, As, mk_s, bs, 12)
    x_solo = xs.download_all().reshape(gshape)
    for v in [xs, bs] + junk:
        v.destroy()
    As.destroy(); solo.close()

    def body(ctx, me):
        z0, z1 = cuts[me], cuts[me + 1]  # my dof planes (0-based among the dofs of the last axis)
        lext = ext[:-1] + (z1 - z0 + 2,)
        llo, lhi = (1,) * dim, tuple(n + 1 for n in dims[:-1]) + (z1 - z0 + 1,)
        mine = slice(z0, z1 + 2)  # padded planes of the global array that are my padded array

        def local(arr, poison=False):
            out = arr[mine].copy()
            if poison:  # ghost planes start out wrong: the operator must refresh them
                if me > 0:
                    out[0] = 7.0
                if me < P - 1:
                    out[-1] = -7.0
            return out

        x = ctx.box_vector(lext, llo, lhi).upload_all(local(ug, poison=True).ravel())
        y = ctx.box_vector(lext, llo, lhi).upload_all(np.zeros(x.n + x.n_ghost))
        lbox = (slice(1, z1 - z0 + 1),) + box[1:]
        gbox = (slice(z0 + 1, z1 + 1),) + box[1:]
        A = ctx.box_stencil(lext, llo, lhi, center, off)
        assert A.global_rows == int(np.prod(dims)) and A.local_rows == x.n
        A.spmv(x, y)
        got = y.download_all().reshape(lext[::-1])
        assert np.array_equal(got[lbox], sten[gbox])
        # the ghost planes of x now hold the neighbours' planes, boundary pads of the other axes included
        xa = x.download_all().reshape(lext[::-1])
        assert np.array_equal(xa, ug[mine])
        d = ctx.get(y.dot_token(x))
        want = float(np.sum(sten[box] * ug[box]))
        assert abs(d - want) <= 1e-12 * float(np.sum(np.abs(sten[box] * ug[box])))
        B = ctx.box_fvm(lext, llo, lhi, beta, alpha, vol, kface, local(a), [local(b) for b in bface])
        x.upload_all(local(ug, poison=True).ravel())  # a write invalidates the ghost planes
        B.spmv(x, y)
        got = y.download_all().reshape(lext[::-1])
        assert np.abs(got[lbox] - fvm[gbox]).max() <= 1e-12 * np.abs(fvm[box]).max()
        B.halo_exchange(x)
        mk = lambda: ctx.box_vector(lext, llo, lhi).upload_all(np.zeros(x.n + x.n_ghost))
        b = mk().upload_all(local(bvec).ravel())
        xs, hist, junk = cg(ctx, A, mk, b, 12)
        assert np.allclose(hist, hist_solo, rtol=1e-10)
        assert np.abs(xs.download_all().reshape(lext[::-1])[lbox] - x_solo[gbox]).max() <= 1e-10 * np.abs(x_solo).max()
        return [x, y, b, xs] + junk + [A, B]

    run_ranks(P, body)

