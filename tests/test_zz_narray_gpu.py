"""Structured-grid ("narray") vectors and operators: fields are padded N-d arrays, vector
operations and reductions act on the interior sub-box only (reference examples/poisson/mesh.hh:157-161,
index_util.hh:33-71), operators read the boundary layers (poisson.cc:44-82).  Checked against numpy on the
padded arrays and against the flat CSR path of the same problem."""
import numpy as np
import pytest

from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H

pytestmark = pytest.mark.gpu

BOXES = [((37,), (1,), (36,)), ((19, 14), (1, 1), (18, 13)), ((12, 9, 8), (1, 1, 1), (11, 8, 7)),
         ((16, 11, 9), (2, 0, 3), (13, 11, 7))]


def _interior(ext, lo, hi):
    """boolean mask over the padded array in storage order (x fastest) and the dof order"""
    shape = tuple(reversed(ext))  # numpy C order: last axis fastest = x
    mask = np.zeros(shape, dtype=bool)
    sl = tuple(slice(l, h) for l, h in reversed(list(zip(lo, hi))))
    mask[sl] = True
    return mask


@pytest.mark.parametrize("ext,lo,hi", BOXES)
def test_vector_operations_touch_the_interior_only(ctx, ext, lo, hi):
    rng = np.random.default_rng(len(ext))
    mask = _interior(ext, lo, hi)
    n = int(mask.sum())
    full = {k: rng.standard_normal(mask.shape) for k in "xyz"}
    v = {k: ctx.box_vector(ext, lo, hi).upload_all(full[k].ravel()) for k in "xyz"}
    x, y, z = v["x"], v["y"], v["z"]
    assert x.n == n and x.n + x.n_ghost == mask.size
    assert np.array_equal(x.download(), full["x"][mask])  # dof order = storage order restricted to the box
    # a fused group: z = 2x + 3y ; z = z * y ; <z, x> ; max z   -- one launch, interior only
    ctx.reset_stats()
    z.linear_sum(2.0, x, 3.0, y)
    z.multiply(z, y)
    t = z.dot_token(x)
    m = z.max()
    zi = (2.0 * full["x"][mask] + 3.0 * full["y"][mask]) * full["y"][mask]
    got = z.download_all().reshape(mask.shape)
    assert np.array_equal(got[mask], zi)
    assert np.array_equal(got[~mask], full["z"][~mask])  # boundary layers untouched
    assert abs(ctx.get(t) - zi @ full["x"][mask]) <= 1e-12 * (np.abs(zi) @ np.abs(full["x"][mask]))
    assert m == zi.max()
    # every reduction kind, aliasing forms, set/scale/abs/recip/add_scalar
    assert x.min() == full["x"][mask].min() and x.inf_norm() == np.abs(full["x"][mask]).max()
    assert abs(x.l1norm() - np.abs(full["x"][mask]).sum()) <= 1e-12 * np.abs(full["x"][mask]).sum()
    assert abs(x.l2norm() - np.linalg.norm(full["x"][mask])) <= 1e-12 * np.linalg.norm(full["x"][mask])
    y.set_scalar(1.5); y.scale(2.0); y.add_scalar(y, -4.0); y.abs(y); y.reciprocal(y); y.axpy(0.5, x, y)
    gy = y.download_all().reshape(mask.shape)
    assert np.array_equal(gy[mask], 0.5 * full["x"][mask] + 1.0 / abs(1.5 * 2.0 - 4.0))
    assert np.array_equal(gy[~mask], full["y"][~mask])
    # partial host copies address dofs
    part = np.arange(5.0) if n >= 8 else np.arange(1.0)
    x.upload(part, 3 if n >= 8 else 0)
    assert np.array_equal(x.download(part.size, 3 if n >= 8 else 0), part)
    assert x.global_size() == n
    for k in v.values():
        k.destroy()


def test_layouts_do_not_mix(ctx):
    a = ctx.box_vector((10, 10), (1, 1), (9, 9))
    b = ctx.box_vector((10, 10), (1, 1), (9, 8))
    c = ctx.box_vector((12, 8), (2, 0), (10, 8))  # 64 dofs like a, different box
    flat = ctx.vector(64)
    for other in (b, c, flat):
        with pytest.raises(F.FsbError):
            a.add(a, other)
        with pytest.raises(F.FsbError):
            a.dot(other)
    A = F.ParCSR.stencil(ctx, 5, 8, 8, 1)
    with pytest.raises(F.FsbError):
        A.spmv(a, a if False else ctx.box_vector((10, 10), (1, 1), (9, 9)))
    B = ctx.box_stencil((10, 10), (1, 1), (9, 9), 4.0, (-1.0, -1.0))
    with pytest.raises(F.FsbError):
        B.spmv(flat, ctx.vector(64))
    with pytest.raises(F.FsbError):
        B.spmv(c, ctx.box_vector((12, 8), (2, 0), (10, 8)))
    with pytest.raises(F.FsbError):
        ctx.box_stencil((10, 10), (0, 1), (9, 9), 4.0, (-1.0, -1.0))  # no boundary layer on the low x side
    A.destroy(); B.destroy()


@pytest.mark.parametrize("ext,lo,hi", BOXES[:3] + [((70, 66, 40), (1, 1, 1), (69, 65, 39))])
def test_box_stencil_reads_boundary_layers(ctx, ext, lo, hi):
    dim = len(ext)
    rng = np.random.default_rng(10 + dim)
    mask = _interior(ext, lo, hi)
    xf = rng.standard_normal(mask.shape)  # boundary layers hold data too (inhomogeneous Dirichlet)
    off = [-1.0, -1.25, -0.5][:dim]
    center = 4.5
    A = ctx.box_stencil(ext, lo, hi, center, off)
    assert A.local_rows == int(mask.sum()) and A.nnz(0) == (2 * dim + 1) * int(mask.sum())
    x = ctx.box_vector(ext, lo, hi).upload_all(xf.ravel())
    y = ctx.box_vector(ext, lo, hi).upload_all(np.full(mask.size, 7.0))
    u = ctx.box_vector(ext, lo, hi).upload_all(rng.standard_normal(mask.size))
    ctx.reset_stats()
    A.spmv(x, y)
    t = u.dot_token(y)
    d = ctx.get(t)
    assert ctx.stat("launches") == 1  # the dot rides in the SpMV here too
    # reference: same accumulation order as the CSR row (z-1, y-1, x-1, centre, x+1, y+1, z+1), products and sums
    # rounded separately
    ref = np.zeros(mask.shape)
    terms = []
    for a in reversed(range(dim)):
        terms.append((off[a], np.roll(xf, 1, axis=dim - 1 - a)))
    terms.append((center, xf))
    for a in range(dim):
        terms.append((off[a], np.roll(xf, -1, axis=dim - 1 - a)))
    for c, arr in terms:
        ref = ref + c * arr
    got = y.download_all().reshape(mask.shape)
    assert np.array_equal(got[mask], ref[mask])
    assert np.all(got[~mask] == 7.0)  # rows exist for dofs only
    ui = u.download()
    assert abs(d - ui @ ref[mask]) <= 1e-12 * (np.abs(ui) @ np.abs(ref[mask]))
    for v in (x, y, u):
        v.destroy()
    A.destroy()


@pytest.mark.parametrize("solver", ["cg", "cg_device"])
def test_poisson_example_on_its_mesh_matches_the_flat_csr_path(ctx, solver):
    """config 1 (examples/poisson) with u, f as narray fields vs the same system as a flat parallel CSR matrix"""
    m = 64
    u, info, hist = H.poisson_narray(ctx, m, seed=7, solver=solver, rtol=1e-9, maxiter=1000, history_cap=1000)
    A = F.ParCSR.stencil(ctx, 5, m, m, 1)
    S = H.Session(ctx, A)
    h = 1.0 / (m + 1)
    xs = (np.arange(m) + 1) * h
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    exact = (np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y)).ravel()
    b = 8 * np.pi ** 2 * exact * h * h
    S.x.set_random(7)
    xf, finfo, fhist = S.solve(b, S.x.download(), solver=solver, rtol=1e-9, maxiter=1000, history_cap=1000)
    assert info.reason == finfo.reason == "converged_rtol"
    assert abs(info.iters - finfo.iters) <= 1, (info.iters, finfo.iters)
    k = min(len(hist), len(fhist), 60)
    assert np.allclose(hist[:k], fhist[:k], rtol=1e-9)
    assert np.abs(u - xf).max() <= (1e-8 if info.iters == finfo.iters else 1e-5)
    assert np.abs(u - exact).max() < 1.2e-3  # discretisation error of the 64^2 grid (second order)
    S.close(); A.destroy()


# ---------------------------------------------------------------- finite-volume coefficient assembler

def _unit_cube(n):
    dx = 1.0 / n
    return (n + 2,) * 3, (1, 1, 1), (n + 1,) * 3, dx ** 3, [dx * dx * (1.0 / dx)] * 3


def _apply(ctx, A, ext, lo, hi, u):
    x = ctx.box_vector(ext, lo, hi).upload_all(np.asarray(u).ravel())
    y = ctx.box_vector(ext, lo, hi).upload_all(np.zeros(u.size))
    A.spmv(x, y)
    out = y.download_all().reshape(u.shape)
    x.destroy(); y.destroy()
    return out


def test_box_fvm_reference_known_answers(ctx):
    """the reference's own checks of its diffusion operator (physics/test/fvm_diffusion.cc:169-210: zero flux, source only,
    boundary sink on the 8^3 unit cube) on the assembled operator"""
    from tests.test_fvm_oracle import dirichlet, neumann
    ext, lo, hi, vol, k = _unit_cube(8)
    ones = np.ones(ext)
    inner = (slice(1, -1),) * 3
    A = ctx.box_fvm(ext, lo, hi, 1.0, 0.0, vol, k, ones, [ones] * 3)
    assert A.local_rows == 512 and A.nnz(0) == 7 * 512
    assert np.abs(_apply(ctx, A, ext, lo, hi, neumann(ones))[inner]).max() < 1e-12  # zero flux
    v = _apply(ctx, A, ext, lo, hi, dirichlet(ones, 1e-9))[inner]  # boundary sink
    edge = np.ones_like(v, dtype=bool)
    edge[1:-1, 1:-1, 1:-1] = False
    assert np.abs(v[~edge]).max() < 1e-12 and np.all(v[edge] < 1.0) and np.all(v[edge] > 0.0)
    A.destroy()
    A = ctx.box_fvm(ext, lo, hi, 0.0, 1.0, vol, k, ones, [ones] * 3)  # source only
    assert np.abs(_apply(ctx, A, ext, lo, hi, neumann(ones))[inner] - vol).max() < 1e-12
    A.destroy()


@pytest.mark.parametrize("shape", [(33,), (21, 17), (14, 11, 9), (40, 38, 36)])
def test_box_fvm_matches_the_matrix_free_operator(ctx, shape):
    """variable cell and face coefficients: the assembled operator equals the oracle's restatement of the reference's
    matrix-free flux form to 1e-12 of the row's absolute sum, and its entries are the documented expressions bit for bit"""
    import oracle as O
    dim = len(shape)
    rng = np.random.default_rng(100 + dim)
    ext = tuple(n + 2 for n in shape)
    npshape = ext[::-1]
    lo, hi = (1,) * dim, tuple(n + 1 for n in shape)
    a, u = rng.uniform(0.5, 2.0, npshape), rng.standard_normal(npshape)
    bface = [rng.uniform(0.5, 2.0, npshape) for _ in range(dim)]
    kface = list(rng.uniform(0.5, 2.0, dim))
    beta, alpha, vol = 1.3, 0.7, 0.11
    A = ctx.box_fvm(ext, lo, hi, beta, alpha, vol, kface, a, bface)
    n = int(np.prod(shape))
    assert A.local_rows == n and A.nnz(0) == (2 * dim + 1) * n
    got = _apply(ctx, A, ext, lo, hi, u)
    ref = O.fvm_diffusion_apply(u, a, bface, beta, alpha, vol, kface, lo, hi)
    box = tuple(slice(lo[dim - 1 - d], hi[dim - 1 - d]) for d in range(dim))
    rowabs = np.zeros(npshape)[box]
    for ax in range(dim):
        npax = dim - 1 - ax
        lower = tuple(slice(s.start - 1, s.stop - 1) if d == npax else s for d, s in enumerate(box))
        upper = tuple(slice(s.start + 1, s.stop + 1) if d == npax else s for d, s in enumerate(box))
        fu, fl = bface[ax][box] * kface[ax], bface[ax][lower] * kface[ax]
        rowabs += beta * (fu * (np.abs(u[upper]) + np.abs(u[box])) + fl * (np.abs(u[lower]) + np.abs(u[box])))
    rowabs += alpha * vol * a[box] * np.abs(u[box])
    assert np.all(np.abs(got[box] - ref[box]) <= 1e-12 * rowabs), (np.abs(got[box] - ref[box]) / rowabs).max()
    # entries: towards c +- e  -(beta (b kface)), centre  beta sum (b kface) + (alpha vol) a
    rp, col, val = A.download(0)
    w = 2 * dim + 1
    val = val.reshape(n, w)
    for ax in range(dim):
        npax = dim - 1 - ax
        lower = tuple(slice(s.start - 1, s.stop - 1) if d == npax else s for d, s in enumerate(box))
        assert np.array_equal(val[:, dim + 1 + ax], (-(beta * (bface[ax][box] * kface[ax]))).ravel())
        assert np.array_equal(val[:, dim - 1 - ax], (-(beta * (bface[ax][lower] * kface[ax]))).ravel())
    diag = np.zeros(npshape)[box]
    for ax in range(dim):
        npax = dim - 1 - ax
        lower = tuple(slice(s.start - 1, s.stop - 1) if d == npax else s for d, s in enumerate(box))
        diag = diag + (bface[ax][box] * kface[ax] + bface[ax][lower] * kface[ax])
    assert np.array_equal(val[:, dim], (beta * diag + (alpha * vol) * a[box]).ravel())
    A.destroy()
