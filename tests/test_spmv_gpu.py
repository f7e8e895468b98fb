"""Device SpMV through the C ABI vs the oracle's restatement of matrices/seq.hh:178-194."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from flecsolve_b200 import _lib as F

pytestmark = pytest.mark.gpu


def _device_spmv(ctx, n, rp, col, val, x):
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    xv, yv = A.vector(x), A.vector()
    A.spmv(xv, yv)
    y = yv.download()
    xv.destroy(); yv.destroy()
    return A, y


@pytest.mark.parametrize("kind,dims", [(5, (64, 64, 1)), (5, (33, 21, 1)), (7, (17, 13, 11)), (7, (32, 32, 32)),
                                       (27, (9, 10, 11)), (27, (24, 24, 24))])
def test_stencil_bit_exact(ctx, kind, dims):
    rp, col, val = O.stencil_csr(kind, *dims)
    n = len(rp) - 1
    rng = np.random.default_rng(kind)
    x = rng.standard_normal(n)
    ref = O.csr_spmv(rp, col, val, x)
    A, y = _device_spmv(ctx, n, rp, col, val, x)
    assert np.array_equal(y, ref)  # same accumulation order, no FMA contraction
    assert np.abs(y - sp.csr_matrix((val, col, rp)) @ x).max() <= 1e-12 * np.abs(ref).max()
    # the device generator builds the identical matrix
    B = F.ParCSR.stencil(ctx, kind, *dims)
    rp2, col2, val2 = B.download(0)
    assert np.array_equal(rp2, rp) and np.array_equal(col2, col) and np.array_equal(val2, val)
    xv, yv = B.vector(x), B.vector()
    B.spmv(xv, yv)
    assert np.array_equal(yv.download(), ref)
    xv.destroy(); yv.destroy(); A.destroy(); B.destroy()


def _random_csr(n, rng, density_rows):
    """rows with 0..k entries incl. empty rows, unsorted columns, repeated structure"""
    counts = density_rows(rng, n)
    rp = np.zeros(n + 1, dtype=np.int64)
    rp[1:] = np.cumsum(counts)
    col = rng.integers(0, n, rp[-1])
    val = rng.standard_normal(rp[-1])
    return rp, col, val


@pytest.mark.parametrize("case", ["ragged", "empty_rows", "long_rows", "one_huge_row", "all_empty", "single"])
def test_irregular_matrices(ctx, case):
    rng = np.random.default_rng(11)
    if case == "ragged":
        n = 5000
        rp, col, val = _random_csr(n, rng, lambda r, n: r.integers(0, 40, n))
    elif case == "empty_rows":
        n = 3000
        rp, col, val = _random_csr(n, rng, lambda r, n: r.integers(0, 3, n) * (r.random(n) < 0.3))
    elif case == "long_rows":  # few rows, hundreds of entries: warp-per-row path
        n = 700
        rp, col, val = _random_csr(n, rng, lambda r, n: r.integers(300, 600, n))
    elif case == "one_huge_row":  # a row longer than a shared-memory stage
        n = 2000
        rp, col, val = _random_csr(n, rng, lambda r, n: np.where(np.arange(n) == 777, 20000, r.integers(0, 9, n)))
    elif case == "all_empty":
        n = 100
        rp, col, val = np.zeros(n + 1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0)
    else:
        n = 1
        rp, col, val = np.array([0, 1]), np.array([0]), np.array([2.5])
    x = rng.standard_normal(n)
    ref = O.csr_spmv(rp, col, val, x)
    A, y = _device_spmv(ctx, n, rp, col, val, x)
    scale = np.abs(ref).max() if ref.size and np.abs(ref).max() > 0 else 1.0
    rowabs = np.zeros(n)
    if val.size:
        np.add.at(rowabs, np.repeat(np.arange(n), np.diff(rp)), np.abs(val * x[col]))
    assert np.all(np.abs(y - ref) <= 1e-12 * np.maximum(rowabs, scale * 1e-3)), np.abs(y - ref).max()
    if case in ("ragged", "empty_rows", "all_empty", "single"):
        assert np.array_equal(y, ref)  # thread-per-row path keeps the reference order
    A.destroy()


def test_fused_dot_matches_separate(ctx):
    rp, col, val = O.stencil_csr(7, 20, 19, 18)
    n = len(rp) - 1
    rng = np.random.default_rng(2)
    x = rng.standard_normal(n)
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    xv, yv, uv = A.vector(x), A.vector(), A.vector(rng.standard_normal(n))
    ref = O.csr_spmv(rp, col, val, x)
    for other, expect in ((xv, ref @ x), (uv, ref @ uv.download()), (yv, ref @ ref)):
        ctx.reset_stats()
        A.spmv(xv, yv)
        t = yv.dot_token(other)
        got = ctx.get(t)
        assert abs(got - expect) <= 1e-13 * np.abs(ref).sum() * max(np.abs(x).max(), np.abs(ref).max())
        assert ctx.stat("launches") == 1  # the dot (and its fold) ride in the SpMV kernel
        assert np.array_equal(yv.download(), ref)
    for v in (xv, yv, uv):
        v.destroy()
    A.destroy()


def test_fused_dot_reaches_across_a_burst(ctx):
    """vec::multi issues A0 x0, A1 x1, <r0,y0>, <r1,y1> (vectors/operations/multi.hh): each dot still rides in its SpMV."""
    rp, col, val = O.stencil_csr(7, 12, 11, 10)
    n = len(rp) - 1
    rng = np.random.default_rng(9)
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    x0, x1, r0, r1 = (A.vector(rng.standard_normal(n)) for _ in range(4))
    y0, y1 = A.vector(), A.vector()
    e0 = O.csr_spmv(rp, col, val, x0.download())
    e1 = O.csr_spmv(rp, col, val, x1.download())
    ctx.reset_stats()
    A.spmv(x0, y0)
    A.spmv(x1, y1)
    t0 = r0.dot_token(y0)
    t1 = r1.dot_token(y1)
    g0, g1 = ctx.get(t0), ctx.get(t1)
    assert ctx.stat("launches") == 2
    assert np.array_equal(y0.download(), e0) and np.array_equal(y1.download(), e1)
    for got, e, r in ((g0, e0, r0), (g1, e1, r1)):
        rr = r.download()
        assert abs(got - rr @ e) <= 1e-13 * (np.abs(rr) @ np.abs(e))
    # a dot whose other operand is overwritten by an intervening SpMV must not move ahead of it
    ctx.reset_stats()
    A.spmv(x0, y0)
    A.spmv(x1, y1)
    t = y0.dot_token(y1)
    assert abs(ctx.get(t) - e0 @ e1) <= 1e-13 * (np.abs(e0) @ np.abs(e1))
    for v in (x0, x1, r0, r1, y0, y1):
        v.destroy()
    A.destroy()


def test_fused_dot_is_bitwise_reproducible(ctx):
    """One rank: row blocks are assigned to CTAs statically (FSB_OPT_REPRODUCIBLE), so the dot fused into the
    SpMV gives the same bits on every run; the dynamic schedule is allowed to differ in the last bits only."""
    A = F.ParCSR.stencil(ctx, 27, 96, 96, 96)
    n = A.local_rows
    rng = np.random.default_rng(5)
    x, u, y = A.vector(rng.standard_normal(n)), A.vector(rng.standard_normal(n)), A.vector()
    vals = []
    for _ in range(6):
        A.spmv(x, y)
        vals.append(ctx.get(u.dot_token(y)))
    assert len(set(vals)) == 1
    ctx.set_option("reproducible", 0)
    A.spmv(x, y)
    dyn = ctx.get(u.dot_token(y))
    ctx.set_option("reproducible", 1)
    yy, uu = y.download(), u.download()
    assert abs(dyn - vals[0]) <= 1e-13 * (np.abs(yy) @ np.abs(uu))
    for v in (x, u, y):
        v.destroy()
    A.destroy()


def test_spmv_rejects_aliasing_and_size_mismatch(ctx):
    rp, col, val = O.stencil_csr(5, 8, 8)
    A = F.ParCSR.from_csr(ctx, 64, [0, 64], rp, col, val)
    x = A.vector()
    with pytest.raises(F.FsbError):
        A.spmv(x, x)
    bad = ctx.vector(63)
    with pytest.raises(F.FsbError):
        A.spmv(x, bad)
    x.destroy(); bad.destroy(); A.destroy()


def test_dinv_and_jacobi_relax(ctx):
    rp, col, val = O.stencil_csr(27, 7, 8, 9)
    n = len(rp) - 1
    rng = np.random.default_rng(4)
    val = val * rng.uniform(0.5, 1.5, val.size)
    M = O.ParCSR(rp, col, val, colours=1)
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    d = A.vector()
    A.extract_dinv(d)
    assert np.array_equal(d.download(), M.dinv())  # 1/a_rr exactly
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    w = float(np.float32(2 / 3))
    bv, xv, tv = A.vector(b), A.vector(x0), A.vector()
    launches = ctx.stat("launches")
    A.jacobi_relax(w, 4, bv, xv, tv)
    assert ctx.stat("launches") - launches == 4  # one pass over the matrix per sweep, nothing else
    ref = M.jacobi_relax(w, 4, b, x0)
    # the kernel skips the diagonal while summing, in the reference's order (mg/jacobi.hh:73-89): same bits
    assert np.array_equal(xv.download(), ref)
    assert np.array_equal(tv.download(), M.jacobi_relax(w, 3, b, x0))  # tmp = the previous iterate, like std::copy leaves it
    for v in (d, bv, xv, tv):
        v.destroy()
    A.destroy()


@pytest.mark.parametrize("kind,dims", [(7, (12, 11, 10)), (27, (9, 8, 7)), (5, (17, 13, 1)), (27, (33, 5, 3))])
@pytest.mark.parametrize("nrelax", [1, 2, 5])
def test_jacobi_relax_bit_exact(ctx, kind, dims, nrelax):
    """weighted Jacobi (solvers/mg/jacobi.hh:44-93) in the gather and the window format, owned and wrapped vectors,
    rows whose diagonal is huge (a cancellation-prone formulation would lose them) or stored twice"""
    rp, col, val = O.stencil_csr(kind, *dims)
    n = len(rp) - 1
    rng = np.random.default_rng(kind + nrelax)
    val = val * rng.uniform(0.5, 1.5, val.size)
    diag = np.flatnonzero(col == np.repeat(np.arange(n), np.diff(rp)))
    val[diag[::7]] = 1e30  # penalty rows
    M = O.ParCSR(rp, col, val, colours=1)
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    w = float(np.float32(2 / 3))
    bv, xv, tv = A.vector(b), A.vector(x0), A.vector()
    A.jacobi_relax(w, nrelax, bv, xv, tv)
    ref = M.jacobi_relax(w, nrelax, b, x0)
    assert np.array_equal(xv.download(), ref)
    for v in (bv, xv, tv):
        v.destroy()
    A.destroy()


@pytest.mark.parametrize("dims", [(8, 8, 8), (9, 7, 5), (33, 17, 3), (64, 64, 4), (5, 5, 5), (130, 3, 3)])
def test_window_format_27pt_bit_exact(ctx, dims):
    """the 27-point operators take the window format (x segments staged in shared memory, 16-bit local columns):
    same bits as the reference's row loop, also with the fused dot, for odd sizes and rows cut by the domain"""
    rp, col, val = O.stencil_csr(27, *dims)
    n = len(rp) - 1
    rng = np.random.default_rng(n)
    val = val * rng.uniform(0.5, 1.5, val.size)
    for A in (F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val),):
        assert A.info("window_format") == 1 and A.info("window_x") > 0
        x = rng.standard_normal(n)
        xv, yv, uv = A.vector(x), A.vector(), A.vector(rng.standard_normal(n))
        A.spmv(xv, yv)
        ref = O.csr_spmv(rp, col, val, x)
        assert np.array_equal(yv.download(), ref)
        A.spmv(xv, yv)
        t = yv.dot_token(uv)
        assert abs(ctx.get(t) - ref @ uv.download()) <= 1e-13 * (np.abs(ref) @ np.abs(uv.download()))
        assert np.array_equal(yv.download(), ref)
        A.spmv(xv, yv)
        t = yv.sumsq_token()
        assert abs(ctx.get(t) - ref @ ref) <= 1e-13 * (ref @ ref)
        for v in (xv, yv, uv):
            v.destroy()
        A.destroy()
    # the device generator builds the same thing
    B = F.ParCSR.stencil(ctx, 27, *dims)
    assert B.info("window_format") == 1
    rp, col, val = O.stencil_csr(27, *dims)
    xv, yv = B.vector(x), B.vector()
    B.spmv(xv, yv)
    assert np.array_equal(yv.download(), O.csr_spmv(rp, col, val, x))
    xv.destroy(); yv.destroy(); B.destroy()


def test_window_format_general_matrices(ctx):
    """banded matrices with ragged rows: a few column clusters per row block -> window format; scattered columns ->
    the builder declines and the gather kernel runs.  Both bit-exact."""
    rng = np.random.default_rng(11)
    n = 5000
    rows, cols = [], []
    for offs in (-700, -699, -3, -1, 0, 1, 2, 650, 651, 652):
        r = np.arange(n)
        keep = (r + offs >= 0) & (r + offs < n) & (rng.random(n) < 0.8)
        rows.append(r[keep]); cols.append((r + offs)[keep])
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    import scipy.sparse as sp
    S = sp.csr_matrix((rng.standard_normal(rows.size), (rows, cols)), shape=(n, n))
    S.sum_duplicates(); S.sort_indices()
    rp, col, val = S.indptr.astype(np.int64), S.indices.astype(np.int64), S.data
    import os
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    x = rng.standard_normal(n)
    xv, yv = A.vector(x), A.vector()
    A.spmv(xv, yv)
    assert np.array_equal(yv.download(), O.csr_spmv(rp, col, val, x))
    xv.destroy(); yv.destroy(); A.destroy()
    # scattered: 12 random columns per row
    cols = rng.integers(0, n, size=(n, 12))
    S = sp.csr_matrix((rng.standard_normal(n * 12), (np.repeat(np.arange(n), 12), cols.ravel())), shape=(n, n))
    S.sum_duplicates(); S.sort_indices()
    rp, col, val = S.indptr.astype(np.int64), S.indices.astype(np.int64), S.data
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    assert A.info("window_format") == 0
    xv, yv = A.vector(x), A.vector()
    A.spmv(xv, yv)
    assert np.array_equal(yv.download(), O.csr_spmv(rp, col, val, x))
    xv.destroy(); yv.destroy(); A.destroy()


@pytest.mark.parametrize("kind,dims", [(7, (19, 17, 23)), (27, (12, 9, 10)), (5, (40, 37, 1))])
@pytest.mark.parametrize("distinct", [1, 2, 3, 255, 256, 257, 5000])
def test_value_dictionary_bit_exact(ctx, kind, dims, distinct):
    """matrices with at most 256 distinct values stream one byte per nonzero (spmv_window_kernel<..., VD>): the same bits
    as the reference's row loop -- plain, with the fused dots and in a Jacobi sweep -- with -0.0, denormals and huge values
    in the table; 257 distinct values and more keep the fp64 stream; switching the option off changes nothing"""
    rp, col, val = O.stencil_csr(kind, *dims)
    n = len(rp) - 1
    rng = np.random.default_rng(distinct + kind)
    table = rng.standard_normal(distinct) * 10.0 ** rng.integers(-3, 4, distinct)
    if distinct >= 3:
        table[:3] = [-0.0, 5e-324, 1e300]
    pick = rng.integers(0, distinct, val.size)
    pick[:distinct] = np.arange(distinct)  # every entry of the table occurs
    val = table[pick]
    diag = np.flatnonzero(col == np.repeat(np.arange(n), np.diff(rp)))
    val[diag] = np.where(np.abs(val[diag]) < 1e-200, 3.0, val[diag])  # Jacobi divides by the diagonal
    expected = len(np.unique(val.view(np.uint64)))
    M = O.ParCSR(rp, col, val, colours=1)
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    assert A.info("window_format") == 1
    assert A.info("value_dictionary") == (expected if expected <= 256 else 0)
    x = rng.standard_normal(n)
    ref = O.csr_spmv(rp, col, val, x)
    xv, yv, uv = A.vector(x), A.vector(), A.vector(rng.standard_normal(n))
    w = float(np.float32(2 / 3))
    b = rng.standard_normal(n)
    refj = M.jacobi_relax(w, 2, b, x)
    for dictionary in (1, 0):
        ctx.set_option("spmv_dictionary", dictionary)
        A.spmv(xv, yv)
        assert np.array_equal(yv.download().view(np.uint64), ref.view(np.uint64))
        A.spmv(xv, yv)
        t = yv.dot_token(xv)  # CG's <Ap, p>: the operand comes out of the staged x window
        assert abs(ctx.get(t) - ref @ x) <= 1e-12 * (np.abs(ref) @ np.abs(x))
        A.spmv(xv, yv)
        t = yv.dot_token(uv)
        assert abs(ctx.get(t) - ref @ uv.download()) <= 1e-12 * (np.abs(ref) @ np.abs(uv.download()))
        bv, xj, tv = A.vector(b), A.vector(x), A.vector()
        A.jacobi_relax(w, 2, bv, xj, tv)
        assert np.array_equal(xj.download(), refj, equal_nan=True)
        for v in (bv, xj, tv):
            v.destroy()
    ctx.set_option("spmv_dictionary", 1)
    for v in (xv, yv, uv):
        v.destroy()
    A.destroy()


def test_value_dictionary_of_the_generated_operators(ctx):
    """the device generator's constant-coefficient stencils have 2 distinct values (3 with a diagonal shift)"""
    for kind, shift, want in ((7, 0.0, 2), (27, 0.0, 2), (7, 1e-3, 2), (107, 1e-3, None)):
        A = F.ParCSR.stencil(ctx, kind, 20, 18, 22, shift, 1.0)
        got = A.info("value_dictionary")
        rp, col, val = A.download(0)
        assert got == len(np.unique(val.view(np.uint64))) and (want is None or got == want)
        x = np.random.default_rng(kind).standard_normal(A.local_rows)
        xv, yv = A.vector(x), A.vector()
        A.spmv(xv, yv)
        assert np.array_equal(yv.download(), O.csr_spmv(rp, col, val, x))
        xv.destroy(); yv.destroy(); A.destroy()


def test_large_stencil_properties(ctx):
    """256^3 7-point (BASELINE config 2): size-independent checks."""
    nn = 256
    A = F.ParCSR.stencil(ctx, 7, nn, nn, nn)
    n = A.local_rows
    assert n == nn ** 3 and A.nnz(0) == 7 * n - 6 * nn * nn
    ones, y = A.vector(), A.vector()
    ones.set_scalar(1.0)
    A.spmv(ones, y)
    got = y.download().reshape(nn, nn, nn)
    # row sum = 6 - number of neighbours = number of faces on the boundary
    idx = np.arange(nn)
    edge = ((idx == 0) | (idx == nn - 1)).astype(np.float64)
    expect = edge[:, None, None] + edge[None, :, None] + edge[None, None, :]
    assert np.array_equal(got, expect)
    # linearity on random vectors: A(a u + b v) == a A u + b A v to rounding
    rng = np.random.default_rng(0)
    u, v = A.vector(rng.standard_normal(n)), A.vector(rng.standard_normal(n))
    w, Au, Av, Aw = A.vector(), A.vector(), A.vector(), A.vector()
    w.linear_sum(0.3, u, -1.7, v)
    A.spmv(u, Au); A.spmv(v, Av); A.spmv(w, Aw)
    Au.linear_sum(0.3, Au, -1.7, Av)
    Aw.subtract(Aw, Au)
    assert Aw.inf_norm() <= 1e-13 * 12 * 5
    # symmetry: u.Av == v.Au
    A.spmv(u, Au); A.spmv(v, Av)
    assert abs(u.dot(Av) - v.dot(Au)) <= 1e-12 * n
    for t in (ones, y, u, v, w, Au, Av, Aw):
        t.destroy()
    A.destroy()
