"""Pin the CPU oracle: closed forms of the reference's own vector tests, scipy as an independent
SpMV, structural properties of the colour split, and solver invariants the reference's tests check."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O

N = 32
GID = np.arange(N, dtype=np.float64)


def xyz():
    return GID.copy(), 2 * GID, 3 * GID  # vectors/test/flecsi_vector.cc:25-34 (rconv<0,1,2>)


def test_vector_closed_forms():
    """The sequence of vectors/test/flecsi_vector.cc:338-380 with its expected values (tol 1e-8 * n)."""
    x, y, z = xyz()
    tmp = np.zeros(N)
    tol = N * 1e-8
    assert np.abs(O.vec_op("add", tmp, x, z) - (x + z)).sum() < tol
    assert np.abs(O.vec_op("subtract", tmp, x, z) - (GID - 3 * GID)).sum() < tol
    assert np.abs(O.vec_op("multiply", tmp, x, z) - GID * 3 * GID).sum() < tol
    O.vec_op("add_scalar", x, x, a=1)  # aliased: x.add_scalar(x, 1)
    assert np.abs(x - (GID + 1)).sum() < tol
    assert np.abs(O.vec_op("divide", tmp, y, x) - 2 * GID / (GID + 1)).sum() < tol
    O.vec_op("add_scalar", x, x, a=-1)
    assert np.abs(O.vec_op("scale", tmp, x, a=2) - 2 * GID).sum() < tol
    O.vec_op("add_scalar", y, y, a=1)
    assert np.abs(O.vec_op("reciprocal", tmp, y) - 1 / (2 * GID + 1)).sum() < tol
    O.vec_op("add_scalar", y, y, a=-1)
    assert np.abs(O.vec_op("linear_sum", tmp, y, z, a=8, b=9) - (8 * 2 * GID + 9 * 3 * GID)).sum() < tol
    assert np.abs(O.vec_op("axpy", tmp, x, y, a=7) - (7 * GID + 2 * GID)).sum() < tol
    O.vec_op("copy", tmp, y)
    assert np.abs(O.vec_op("axpby", tmp, z, a=4, b=11) - (4 * 3 * GID + 11 * 2 * GID)).sum() < tol
    O.vec_op("add_scalar", tmp, y, a=-4)
    O.vec_op("abs", tmp, tmp)  # aliased abs
    assert np.abs(tmp - np.abs(2 * GID - 4)).sum() < tol
    O.vec_op("add_scalar", tmp, y, a=-7)
    assert O.vec_reduce("min", tmp) == -7  # flecsi_vector.cc:376-377
    assert O.vec_reduce("max", z) == 93  # :379


def test_reduction_closed_forms():
    """flecsi_vector.cc:269-307, 382-392: constant vectors."""
    a, b = np.full(N, 1.5), np.full(N, 3.8)
    assert abs(O.vec_reduce("dot", a, b) - 32 * 1.5 * 3.8) < 1e-8
    c = np.full(N, -3.141719)
    assert abs(O.vec_reduce("l1", c) - 32 * 3.141719) < 1e-8
    assert abs(np.sqrt(O.vec_reduce("dot", c, c)) - np.sqrt(32 * 3.141719 ** 2)) < 1e-8
    assert O.vec_reduce("inf", c) == 3.141719
    assert abs(O.vec_reduce("powsum", np.full(N, 2.0), a=3) - 32 * 8) < 1e-12


@pytest.mark.parametrize("kind,dims", [(5, (13, 9, 1)), (7, (7, 6, 5)), (27, (5, 6, 4))])
def test_stencil_generators(kind, dims):
    rp, col, val = O.stencil_csr(kind, *dims)
    n = dims[0] * dims[1] * dims[2]
    A = sp.csr_matrix((val, col, rp), shape=(n, n))
    assert (A != A.T).nnz == 0  # symmetric
    diag = {5: 4.0, 7: 6.0, 27: 26.0}[kind]
    assert np.all(A.diagonal() == diag)
    assert np.all(np.diff(col[rp[1]:rp[2]]) > 0)  # ascending columns
    full = {5: 5, 7: 7, 27: 27}[kind]
    assert np.diff(rp).max() == full
    # interior row sums vanish (Dirichlet truncation only removes entries at the boundary)
    sums = np.asarray(A.sum(axis=1)).ravel()
    assert sums.min() >= 0 and (sums == 0).sum() == max(dims[0] - 2, 0) * max(dims[1] - 2, 0) * (
        max(dims[2] - 2, 0) if kind != 5 else 1)


@pytest.mark.parametrize("colours", [1, 2, 3, 4, 7])
def test_parallel_split_matches_serial(colours):
    """matrices/parcsr.hh:61-68 over topo/csr.hh:482-618 == serial seq.hh:178-194 up to the
    (diag.x) + (offd.x) association."""
    rp, col, val = O.stencil_csr(27, 6, 5, 7)
    rng = np.random.default_rng(1)
    val = val * rng.uniform(0.5, 1.5, val.size)
    x = rng.standard_normal(len(rp) - 1)
    M = O.ParCSR(rp, col, val, colours=colours)
    y_par = M.spmv(x)
    y_ser = O.csr_spmv(rp, col, val, x)
    y_sp = sp.csr_matrix((val, col, rp)) @ x
    assert np.abs(y_ser - y_sp).max() <= 1e-13 * np.abs(y_sp).max()
    assert np.abs(y_par - y_ser).max() <= 1e-14 * np.abs(y_ser).max()
    if colours == 1:
        assert np.array_equal(y_par, y_ser)
    part = M.partition()
    assert part[0] == 0 and part[-1] == len(rp) - 1 and np.all(np.diff(part) > 0)
    n = len(rp) - 1
    assert np.diff(part).max() - np.diff(part).min() <= 1  # equal_map
    for p in range(colours):
        drp, dcol, dval, cm = M.block(p, 0)
        orp, ocol, oval, _ = M.block(p, 1)
        n_owned = part[p + 1] - part[p]
        assert np.all(np.diff(cm) > 0)  # sorted unique ghosts (force_unique, topo/csr.hh:524)
        assert np.all((cm < part[p]) | (cm >= part[p + 1]))
        assert dcol.size == 0 or (dcol.min() >= 0 and dcol.max() < n_owned)
        assert ocol.size == 0 or (ocol.min() >= n_owned and ocol.max() < n_owned + cm.size)
        assert drp[-1] + orp[-1] == rp[part[p + 1]] - rp[part[p]]
        if ocol.size:
            assert set(np.unique(ocol - n_owned)) == set(range(cm.size))  # every ghost is referenced


def test_cg_invariants():
    """solvers/test/cg.cc:16-54: A-norm of the error decreases monotonically and obeys the
    condition-number bound; iteration count equals the number of diagnostic calls."""
    rp, col, val = O.stencil_csr(7, 10, 9, 8)
    A = sp.csr_matrix((val, col, rp))
    M = O.ParCSR(rp, col, val, colours=3)
    x0 = M.set_random(7)
    b = np.zeros(A.shape[0])
    x, info, hist = M.cg(b, x0=x0, rtol=1e-9, maxiter=1000, history_cap=1000)
    assert info.reason == "diverged_iters" or info.reason == "converged_rtol"
    # b = 0 -> b_norm := 1 (cg.hh:57-58), so the solve converges to |r| < 1e-9
    assert info.reason == "converged_rtol" and info.iters == len(hist)
    assert hist[-1] < np.float32(1e-9)
    assert np.linalg.norm(A @ x) < 1e-8


def test_solver_parity_between_colour_counts():
    """Partitioning only changes summation order: iteration counts agree (north_star: within 2 %)."""
    rp, col, val = O.stencil_csr(7, 12, 12, 12)
    A = sp.csr_matrix((val, col, rp))
    b = A @ np.linspace(1, 2, A.shape[0])
    its = {}
    for colours in (1, 4):
        M = O.ParCSR(rp, col, val, colours=colours)
        for name, fn, kw in (("cg", M.cg, {}), ("bicgstab", M.bicgstab, {}),
                             ("gmres", M.gmres, dict(max_krylov_dim=30, restart=True))):
            x, info, _ = fn(b, dinv=M.dinv(), rtol=1e-8, maxiter=500, **kw)
            assert info.reason == "converged_rtol", (name, colours)
            assert np.linalg.norm(b - A @ x) <= 2e-8 * np.linalg.norm(b) * (10 if name == "gmres" else 1)
            its.setdefault(name, []).append(info.iters)
    for name, (a, c) in its.items():
        assert abs(a - c) <= max(1, 0.02 * a), (name, a, c)


def test_gmres_jacobi_helps_and_restart_consistency():
    """solvers/test/gmres.cc:74-110: a diagonal preconditioner reduces the iteration count on a
    badly scaled system; residual estimates decrease monotonically."""
    rp, col, val = O.stencil_csr(7, 8, 8, 8)
    n = len(rp) - 1
    rng = np.random.default_rng(0)
    d = 10 ** rng.uniform(-2, 2, n)
    A = sp.diags(d) @ sp.csr_matrix((val, col, rp)) @ sp.diags(d)  # symmetric diagonal scaling
    A = A.tocsr()
    A.sort_indices()
    M = O.ParCSR(A.indptr, A.indices, A.data, colours=2)
    b = M.set_random(0)
    x0 = M.set_random(1)
    _, plain, h0 = M.gmres(b, x0=x0, rtol=1e-4, maxiter=100, history_cap=100)
    _, prec, h1 = M.gmres(b, x0=x0, dinv=M.dinv(), rtol=1e-4, maxiter=100, history_cap=100)
    assert prec.reason == "converged_rtol"
    assert prec.iters < (plain.iters if plain.iters else 100)
    assert np.all(np.diff(h1) <= 1e-12)


def test_bicgstab_converges_and_quirk():
    rp, col, val = O.stencil_csr(5, 24, 24)
    A = sp.csr_matrix((val, col, rp))
    M = O.ParCSR(rp, col, val, colours=2)
    b = A @ np.ones(A.shape[0])
    x, info, hist = M.bicgstab(b, rtol=1e-9, maxiter=400, history_cap=400)
    assert info.reason == "converged_rtol" and np.abs(x - 1).max() < 1e-6
    # tolerance is relative to |b| only and float-typed (solver_settings.hh:31-36)
    assert hist[-1] < np.float32(1e-9) * np.linalg.norm(b)


def test_jacobi_relax_matches_dense_formula():
    """mg/jacobi.hh:58-93 against x <- w D^-1 (b - (A - D) x) + (1 - w) x."""
    rp, col, val = O.stencil_csr(27, 5, 4, 6)
    A = sp.csr_matrix((val, col, rp))
    M = O.ParCSR(rp, col, val, colours=3)
    rng = np.random.default_rng(3)
    b, x = rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])
    w = float(np.float32(2 / 3))
    D = A.diagonal()
    ref = x.copy()
    for _ in range(3):
        ref = w / D * (b - (A @ ref - D * ref)) + (1 - w) * ref
    got = M.jacobi_relax(w, 3, b, x)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()


def test_set_random_is_mt19937_uniform():
    """topo_tasks.hh:305-314: same seed on every colour, sequential draws."""
    rp, col, val = O.stencil_csr(5, 8, 8)
    M = O.ParCSR(rp, col, val, colours=2)
    x = M.set_random(7)
    half = len(x) // 2
    assert np.array_equal(x[:half], x[half:])  # both colours restart the generator
    # libstdc++ mt19937(7) + uniform_real_distribution<double>(0,1): first value is a 53-bit draw
    assert 0 <= x.min() and x.max() < 1 and len(np.unique(x[:half])) == half
