"""Scratch: first GPU shake-out of libfsb (vec ops, spmv, python-level CG timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import faulthandler; faulthandler.dump_traceback_later(240, exit=True)
from flecsolve_b200 import _lib as F
import oracle

ctx = F.Context(0)
n = 1000003
rng = np.random.default_rng(0)
X, Y = rng.random(n), rng.random(n)
x, y, z = ctx.vector(n, data=X), ctx.vector(n, data=Y), ctx.vector(n)
z.axpy(2.5, x, y)
print("axpy err", np.abs(z.download() - (2.5 * X + Y)).max())
print("dot", x.dot(y), X @ Y, "l2", x.l2norm(), np.linalg.norm(X), "min", x.min(), X.min(), "max", y.max(), Y.max())
z.axpy(-1.0, x, z); 
print("fused axpy+norm", z.l2norm(), np.linalg.norm(Y + 1.5 * X))
print("launches", ctx.stat("launches"), "fused", ctx.stat("fused_statements"))

# spmv small vs oracle
for kind, dims in ((7, (17, 13, 11)), (27, (9, 10, 11)), (5, (33, 21, 1))):
    rp, c, v = oracle.stencil_csr(kind, *dims)
    N = len(rp) - 1
    A = F.ParCSR.from_csr(ctx, N, [0, N], rp, c, v)
    xv = rng.random(N)
    xx, yy = A.vector(xv), A.vector()
    A.spmv(xx, yy)
    ref = oracle.csr_spmv(rp, c, v, xv)
    got = yy.download()
    print(kind, dims, "host-csr spmv maxdiff", np.abs(got - ref).max(), "bitexact", np.array_equal(got, ref))
    B = F.ParCSR.stencil(ctx, kind, *dims)
    rp2, c2, v2 = B.download(0)
    print("   generator match", np.array_equal(rp2, rp), np.array_equal(c2, c), np.array_equal(v2, v))
    B.spmv(xx, yy)
    print("   stencil spmv bitexact", np.array_equal(yy.download(), ref))
    w = xx.dot(yy); 
    A.spmv(xx, yy); t = yy.dot_token(xx); print("   fused dot", ctx.get(t), w, ref @ xv)

# big: 256^3 7pt
import ctypes
for kind, nn in ((7, 256), (27, 160)):
    t0 = time.time()
    A = F.ParCSR.stencil(ctx, kind, nn, nn, nn)
    ctx.sync(); print(f"{kind}-pt {nn}^3 build {time.time()-t0:.2f}s nnz {A.nnz(0)}")
    N = A.local_rows
    p, w, xv, r, zv, dinv = (A.vector() for _ in range(6))
    p.set_scalar(1.0); A.extract_dinv(dinv)
    def timeit(fn, reps=20):
        fn(); ctx.sync()
        ctx.event_record(0)
        for _ in range(reps): fn()
        ctx.event_record(1)
        return ctx.event_elapsed_ms(0, 1) / reps
    nnz = A.nnz(0)
    ms = timeit(lambda: (A.spmv(p, w), ctx.flush()))
    byt = 12 * nnz + 4 * (N + 1) + 16 * N
    print(f"  spmv {ms:.4f} ms  {byt/ms/1e6:.1f} GB/s")
    ms = timeit(lambda: (A.spmv(p, w), w.dot(p)))
    print(f"  spmv+dot(get) {ms:.4f} ms  {byt/ms/1e6:.1f} GB/s")
    ms = timeit(lambda: (xv.axpy(0.5, p, xv), r.axpy(-0.5, w, r), r.l2norm()))
    print(f"  cg_update {ms:.4f} ms  {48*N/ms/1e6:.1f} GB/s")
    ms = timeit(lambda: (zv.multiply(dinv, r), r.dot(zv)))
    print(f"  jacobi+dot {ms:.4f} ms  {24*N/ms/1e6:.1f} GB/s")
    ms = timeit(lambda: (p.axpy(0.9, p, zv), ctx.flush()))
    print(f"  p update {ms:.4f} ms  {24*N/ms/1e6:.1f} GB/s")
    ms = timeit(lambda: (zv.copy(r), ctx.flush()))
    print(f"  copy {ms:.4f} ms  {16*N/ms/1e6:.1f} GB/s")
    ms = timeit(lambda: (r.dot(zv)))
    print(f"  dot {ms:.4f} ms  {16*N/ms/1e6:.1f} GB/s")
    # python-level PCG
    b = A.vector(); xs = A.vector(); ones = A.vector(); ones.set_scalar(1.0)
    A.spmv(ones, b); xs.zero(); r.copy(b)
    bn = b.l2norm(); zv.multiply(dinv, r); rho = zv.dot(r); p.copy(zv)
    ctx.sync(); t0 = time.time(); its = 0
    for it in range(200):
        A.spmv(p, w); alpha = rho / w.dot(p)
        xs.axpy(alpha, p, xs); r.axpy(-alpha, w, r); res = r.l2norm()
        zv.multiply(dinv, r); rho0, rho = rho, r.dot(zv)
        p.axpy(rho / rho0, p, zv); its += 1
    ctx.sync(); dt = time.time() - t0
    cgb = 12 * nnz + 108 * N
    print(f"  python PCG: {its/dt:.1f} it/s, {dt/its*1e3:.4f} ms/it, res {res/bn:.3e}, roofline frac {cgb/(dt/its)/6546.9e9:.3f}, launches/it {ctx.stat('launches')}")
    for v in (p, w, xv, r, zv, dinv, b, xs, ones): v.destroy()
    A.destroy()
print("DONE")
