"""Device vector operations through the C ABI vs the oracle (vectors/operations/topo_tasks.hh).
Element-wise results must be BIT-IDENTICAL to the oracle (products and sums are rounded
separately on both sides); reductions within 1e-13 relative."""
import numpy as np
import pytest

import oracle as O
from flecsolve_b200 import _lib as F

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 31, 32, 1000, 100003]


def _rng_vectors(n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n), rng.standard_normal(n), rng.uniform(0.5, 2.0, n)


# (name, device call, oracle call) for every alias pattern the reference dispatches on
def _cases(a=1.7, b=-0.3):
    return [
        ("copy", lambda z, x, y: z.copy(x), lambda Z, X, Y: O.vec_op("copy", Z, X)),
        ("set", lambda z, x, y: z.set_scalar(a), lambda Z, X, Y: O.vec_op("set", Z, a=a)),
        ("scale", lambda z, x, y: z.scale(a, x), lambda Z, X, Y: O.vec_op("scale", Z, X, a=a)),
        ("add", lambda z, x, y: z.add(x, y), lambda Z, X, Y: O.vec_op("add", Z, X, Y)),
        ("subtract", lambda z, x, y: z.subtract(x, y), lambda Z, X, Y: O.vec_op("subtract", Z, X, Y)),
        ("multiply", lambda z, x, y: z.multiply(x, y), lambda Z, X, Y: O.vec_op("multiply", Z, X, Y)),
        ("divide", lambda z, x, y: z.divide(x, y), lambda Z, X, Y: O.vec_op("divide", Z, X, Y)),
        ("reciprocal", lambda z, x, y: z.reciprocal(y), lambda Z, X, Y: O.vec_op("reciprocal", Z, Y)),
        ("linear_sum", lambda z, x, y: z.linear_sum(a, x, b, y), lambda Z, X, Y: O.vec_op("linear_sum", Z, X, Y, a, b)),
        ("axpy", lambda z, x, y: z.axpy(a, x, y), lambda Z, X, Y: O.vec_op("axpy", Z, X, Y, a)),
        ("axpby", lambda z, x, y: z.axpby(a, b, x), lambda Z, X, Y: O.vec_op("axpby", Z, X, a=a, b=b)),
        ("abs", lambda z, x, y: z.abs(x), lambda Z, X, Y: O.vec_op("abs", Z, X)),
        ("add_scalar", lambda z, x, y: z.add_scalar(x, a), lambda Z, X, Y: O.vec_op("add_scalar", Z, X, a=a)),
    ]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("alias", ["none", "z=x", "z=y", "x=y", "all"])
def test_elementwise_bit_exact(ctx, n, alias):
    X, Y, Zinit = _rng_vectors(n)
    for name, dev, ora in _cases():
        hx, hy, hz = X.copy(), Y.copy(), Zinit.copy()
        x, y, z = ctx.vector(n, data=hx), ctx.vector(n, data=hy), ctx.vector(n, data=hz)
        if alias == "z=x":
            dx, dy, dz, ox, oy, oz = z, y, z, hz, hy, hz
        elif alias == "z=y":
            dx, dy, dz, ox, oy, oz = x, z, z, hx, hz, hz
        elif alias == "x=y":
            dx, dy, dz, ox, oy, oz = x, x, z, hx, hx, hz
        elif alias == "all":
            dx, dy, dz, ox, oy, oz = z, z, z, hz, hz, hz
        else:
            dx, dy, dz, ox, oy, oz = x, y, z, hx, hy, hz
        if name == "copy" and dz is dx:
            for v in (x, y, z):
                v.destroy()
            continue
        dev(dz, dx, dy)
        ora(oz, ox, oy)
        got = dz.download()
        assert np.array_equal(got, oz, equal_nan=True), (name, alias, n, np.abs(got - oz).max())
        for v in (x, y, z):
            v.destroy()


@pytest.mark.parametrize("n", SIZES)
def test_reductions(ctx, n):
    X, Y, _ = _rng_vectors(n, seed=3)
    x, y = ctx.vector(n, data=X), ctx.vector(n, data=Y)
    tol = 1e-13

    def close(got, ref, scale):
        return abs(got - ref) <= tol * max(scale, 1e-300)

    assert close(x.dot(y), O.vec_reduce("dot", X, Y), np.abs(X * Y).sum())
    assert close(x.l2norm(), np.sqrt(O.vec_reduce("dot", X, X)), np.linalg.norm(X))
    assert close(x.l1norm(), O.vec_reduce("l1", X), np.abs(X).sum())
    assert x.inf_norm() == O.vec_reduce("inf", X)  # exact
    assert x.min() == O.vec_reduce("min", X)
    assert x.max() == O.vec_reduce("max", X)
    P = np.abs(X) + 0.5
    p = ctx.vector(n, data=P)
    ref = O.vec_reduce("powsum", P, a=3)
    assert abs(p.lp_norm(3) ** 3 - ref) <= 1e-12 * ref
    assert x.global_size() == n
    for v in (x, y, p):
        v.destroy()


def test_closed_forms_n32(ctx):
    """vectors/test/flecsi_vector.cc:338-392 on the device (n = 32, x = gid, y = 2 gid, z = 3 gid)."""
    n = 32
    gid = np.arange(n, dtype=np.float64)
    x, y, z, tmp = (ctx.vector(n, data=d) for d in (gid, 2 * gid, 3 * gid, np.zeros(n)))
    tol = n * 1e-8
    tmp.add(x, z); assert np.abs(tmp.download() - 4 * gid).sum() < tol
    tmp.subtract(x, z); assert np.abs(tmp.download() + 2 * gid).sum() < tol
    tmp.multiply(x, z); assert np.abs(tmp.download() - 3 * gid * gid).sum() < tol
    x.add_scalar(x, 1); assert np.abs(x.download() - (gid + 1)).sum() < tol
    tmp.divide(y, x); assert np.abs(tmp.download() - 2 * gid / (gid + 1)).sum() < tol
    x.add_scalar(x, -1)
    tmp.scale(2, x); assert np.abs(tmp.download() - 2 * gid).sum() < tol
    y.add_scalar(y, 1)
    tmp.reciprocal(y); assert np.abs(tmp.download() - 1 / (2 * gid + 1)).sum() < tol
    y.add_scalar(y, -1)
    tmp.linear_sum(8, y, 9, z); assert np.abs(tmp.download() - (16 * gid + 27 * gid)).sum() < tol
    tmp.axpy(7, x, y); assert np.abs(tmp.download() - 9 * gid).sum() < tol
    tmp.copy(y); tmp.axpby(4, 11, z); assert np.abs(tmp.download() - (12 * gid + 22 * gid)).sum() < tol
    tmp.add_scalar(y, -4); tmp.abs(tmp); assert np.abs(tmp.download() - np.abs(2 * gid - 4)).sum() < tol
    tmp.add_scalar(y, -7)
    assert tmp.min() == -7
    assert z.max() == 93
    x.set_scalar(1.5); y.set_scalar(3.8)
    assert abs(x.dot(y) - 32 * 1.5 * 3.8) < 1e-8
    x.set_scalar(-3.141719)
    assert abs(x.l1norm() - 32 * 3.141719) < 1e-8
    assert abs(x.l2norm() - np.sqrt(32 * 3.141719 ** 2)) < 1e-8
    for v in (x, y, z, tmp):
        v.destroy()


def test_fusion_is_value_preserving(ctx):
    """The deferred queue must not change any value: same statements with fusion on and off."""
    n = 50001
    rng = np.random.default_rng(5)
    data = [rng.standard_normal(n) for _ in range(6)]

    def run(fuse):
        ctx.set_option("fusion", int(fuse))
        ctx.reset_stats()
        p, x, w, r, dinv, z = (ctx.vector(n, data=d) for d in data)
        out = []
        for alpha in (0.3, -1.1):
            x.axpy(alpha, p, x)
            r.axpy(-alpha, w, r)
            t1 = r.sumsq_token()
            z.multiply(dinv, r)
            t2 = r.dot_token(z)
            p.axpy(0.7, p, z)
            out += [ctx.get(t1), ctx.get(t2)]
        vecs = [v.download() for v in (p, x, r, z)]
        launches = ctx.stat("launches")
        for v in (p, x, w, r, dinv, z):
            v.destroy()
        return out, vecs, launches

    try:
        red_f, vec_f, launches_f = run(True)
        red_u, vec_u, launches_u = run(False)
    finally:
        ctx.set_option("fusion", 1)
    for a, b in zip(vec_f, vec_u):
        assert np.array_equal(a, b)
    assert red_f == red_u  # same kernels' grid => same summation order
    assert launches_f < launches_u
    assert launches_f == 2 and launches_u == 2 * 6  # fused: one (generic) launch per flush


def test_set_random_matches_oracle(ctx):
    rp, col, val = O.stencil_csr(5, 16, 16)
    M = O.ParCSR(rp, col, val, colours=1)
    v = ctx.vector(M.n)
    v.set_random(7)
    assert np.array_equal(v.download(), M.set_random(7))
    v.destroy()


def test_wrap_external_memory(ctx):
    import torch
    from flecsolve_b200 import _lib as F
    import ctypes as C
    t = torch.arange(1000, dtype=torch.float64, device="cuda")
    h = C.c_void_p()
    F.check(F.lib().fsb_vec_wrap(ctx.h, C.c_void_p(t.data_ptr()), 1000, 0, C.byref(h)))
    v = F.Vector(ctx, 1000, 0, handle=h)
    v.scale(2.0)
    ctx.sync()
    assert torch.equal(t.cpu(), torch.arange(1000, dtype=torch.float64) * 2)
    v.destroy()


def test_argument_errors(ctx):
    from flecsolve_b200 import _lib as F
    a, b = ctx.vector(10), ctx.vector(11)
    with pytest.raises(F.FsbError):
        a.add(a, b)  # size mismatch
    with pytest.raises(F.FsbError):
        ctx.get(10 ** 9)  # unknown token
    a.destroy(); b.destroy()


def test_generic_program_kernel(ctx):
    """Statement groups without a compile-time instantiation run through the generic program kernel:
    one launch, same bits as statement-by-statement evaluation (here: a BDF3-like source-term chain,
    time-integrators/bdf.hh:319-345, followed by unrelated statements and two reductions)."""
    n = 77777
    rng = np.random.default_rng(9)
    H = [rng.standard_normal(n) for _ in range(6)]
    u0, u1, u2, f, rhs, w = (ctx.vector(n, data=h) for h in H)
    a1, a2, a3 = -1.7, 0.9, -0.2
    ctx.sync()
    ctx.reset_stats()
    f.scale(a1, u0)            # f = a1 u_n
    f.axpy(a2, u1, f)          # f += a2 u_{n-1}
    f.axpy(a3, u2, f)          # f += a3 u_{n-2}
    rhs.scale(-1.0, f)         # rhs = -f
    w.divide(w, u0)            # w = w / u0   (aliased)
    w.abs(w)
    t1 = w.dot_token(rhs)
    t2 = f.sumsq_token()
    d1, d2 = ctx.get(t1), ctx.get(t2)
    assert ctx.stat("launches") == 1 and ctx.stat("unmatched_groups") == 1
    F_ = O.vec_op("scale", H[3].copy(), H[0], a=a1)
    O.vec_op("axpy", F_, H[1], F_, a=a2)
    O.vec_op("axpy", F_, H[2], F_, a=a3)
    R_ = O.vec_op("scale", H[4].copy(), F_, a=-1.0)
    W_ = O.vec_op("divide", H[5].copy(), H[5].copy(), H[0])
    O.vec_op("abs", W_, W_)
    assert np.array_equal(f.download(), F_) and np.array_equal(rhs.download(), R_) and np.array_equal(w.download(), W_)
    assert abs(d1 - W_ @ R_) <= 1e-13 * np.abs(W_ * R_).sum()
    assert abs(d2 - F_ @ F_) <= 1e-13 * (F_ @ F_)
    # min / max / inf-norm folds through the generic epilogue
    ctx.reset_stats()
    w.add_scalar(w, -0.5)
    t3, t4 = ctx.vector, None
    import ctypes as C
    from flecsolve_b200 import _lib as F
    toks = [w._tok(F.lib().fsb_vec_min, w.h), w._tok(F.lib().fsb_vec_max, w.h), w._tok(F.lib().fsb_vec_amax, w.h)]
    got = [ctx.get(t) for t in toks]
    W2 = W_ - 0.5
    assert got == [W2.min(), W2.max(), np.abs(W2).max()]
    assert ctx.stat("launches") == 1
    for v in (u0, u1, u2, f, rhs, w):
        v.destroy()


def test_multivector_closed_forms_of_the_reference_test(ctx):
    """vectors/test/flecsi_multivector.cc:84-241 on a four-component vec::multi (n = 32 as there, and a size that
    spans many CTAs): element-wise results to the reference's 1e-8, min == -7, combined reductions exactly the
    combination of the component reductions, subset() by variable / multivariable"""
    from flecsolve_b200 import host as H
    for n in (32, 70001):
        rp = np.arange(n + 1, dtype=np.int64)
        A = F.ParCSR.from_csr(ctx, n, [0, n], rp, np.arange(n, dtype=np.int64), np.ones(n))
        S = H.Session(ctx, A)
        out = S.multivector_selftest()
        scale = float(n) ** 2  # values grow like gid^2 in the multiply check
        assert np.all(out[:11] <= 1e-8 * max(1.0, scale * 1e-6)), out[:11]
        assert out[11] == -7
        assert np.all(out[12:18] == 1.0), out[12:18]
        S.close(); A.destroy()


def test_handles_over_the_same_memory_are_one_vector(ctx):
    """fsb_vec_wrap: two handles over the same storage behave as one vector inside a fused group (a statement reading
    through one sees what an earlier statement wrote through the other); partially overlapping ranges never share a
    launch, and are rejected inside one call"""
    n = 4096
    rng = np.random.default_rng(5)
    base, y = ctx.vector(n + 2), ctx.vector(n)
    yh = rng.standard_normal(n)
    y.upload(yh)
    ptr = base.device_ptr
    a, b = ctx.wrap(ptr, n), ctx.wrap(ptr, n)
    launches = ctx.stat("launches")
    a.set_scalar(2.0)          # writes through a
    y.axpy(3.0, b, y)          # reads through b: must see 2.0 although both statements share a launch
    t = b.sumsq_token()
    got = y.download()
    assert ctx.stat("launches") - launches == 1
    assert np.array_equal(got, 3.0 * 2.0 + yh) and ctx.get(t) == 4.0 * n
    # shifted by two entries: element i of c is element i + 2 of a
    c = ctx.wrap(ptr + 16, n)
    launches = ctx.stat("launches")
    a.set_scalar(1.0)
    c.scale(5.0, c)            # its own launch, after the first one completed
    ctx.sync()
    assert ctx.stat("launches") - launches == 2
    full = np.zeros(n + 2); base_h = base.download(); full[:] = base_h
    assert np.array_equal(full[:2], [1.0, 1.0]) and np.all(full[2:n] == 5.0)
    with pytest.raises(F.FsbError):
        a.add(a, c)
        ctx.sync()
    for v in (a, b, c, base, y):
        v.destroy()
