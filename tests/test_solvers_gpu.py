"""Krylov solvers of the C++ host layer (flecsolve-shaped templates over device vectors) vs the
oracle's restatement of solvers/cg.hh, gmres.hh, bicgstab.hh.  Parity bar (BASELINE.json):
iteration counts within 2 %, same stop reason, final residual below the same tolerance."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from flecsolve_b200 import _lib as F
from flecsolve_b200 import host as H

pytestmark = pytest.mark.gpu


def _system(kind, dims, seed=0, scale_rows=False):
    rp, col, val = O.stencil_csr(kind, *dims)
    n = len(rp) - 1
    if scale_rows:  # badly scaled SPD system D A D: Jacobi matters
        rng = np.random.default_rng(seed)
        d = 10 ** rng.uniform(-1.5, 1.5, n)
        A = (sp.diags(d) @ sp.csr_matrix((val, col, rp)) @ sp.diags(d)).tocsr()
        A.sort_indices()
        rp, col, val = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data
    return n, rp, col, val


def _pair(ctx, n, rp, col, val):
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    return A, H.Session(ctx, A), O.ParCSR(rp, col, val, colours=1)


def _same(info, oinfo, tol_frac=0.02):
    assert info.reason == oinfo.reason, (info.reason, oinfo.reason)
    assert abs(info.iters - oinfo.iters) <= max(1, tol_frac * oinfo.iters), (info.iters, oinfo.iters)


@pytest.mark.parametrize("precond", [None, "dinv"])
@pytest.mark.parametrize("kind,dims", [(5, (48, 48, 1)), (7, (20, 18, 16)), (27, (12, 12, 12))])
def test_cg_parity(ctx, kind, dims, precond):
    n, rp, col, val = _system(kind, dims, scale_rows=(precond == "dinv"))
    A, S, M = _pair(ctx, n, rp, col, val)
    As = sp.csr_matrix((val, col, rp))
    b = As @ np.linspace(1, 2, n)
    x0 = M.set_random(7)
    dinv = M.dinv() if precond else None
    xo, oinfo, ohist = M.cg(b, x0=x0, dinv=dinv, rtol=1e-9, maxiter=2000, history_cap=2000)
    x, info, hist = S.solve(b, x0, solver="cg", precond=precond, rtol=1e-9, maxiter=2000, history_cap=2000)
    _same(info, oinfo)
    assert info.callbacks == info.iters  # diagnostic runs once per iteration (solvers/test/cg.cc:91)
    m = min(len(hist), len(ohist), 50)
    assert np.allclose(hist[:m], ohist[:m], rtol=1e-8)  # same residual history while rounding has not diverged
    assert np.linalg.norm(b - As @ x) <= 1.5 * np.float32(1e-9) * np.linalg.norm(b)
    assert abs(info.rhs_norm - oinfo.rhs_norm) <= 1e-6 * oinfo.rhs_norm
    assert abs(info.sol_norm_initial - oinfo.sol_norm_initial) <= 1e-6 * oinfo.sol_norm_initial
    S.close(); A.destroy()


def test_cg_zero_guess_zero_rhs_and_maxiter(ctx):
    n, rp, col, val = _system(7, (10, 10, 10))
    A, S, M = _pair(ctx, n, rp, col, val)
    # zero rhs with random x0: |b| == 0 -> 1 (cg.hh:57-58)
    x0 = M.set_random(7)
    _, oinfo, _ = M.cg(np.zeros(n), x0=x0, rtol=1e-9, maxiter=1000)
    x, info, _ = S.solve(np.zeros(n), x0, solver="cg", rtol=1e-9, maxiter=1000)
    _same(info, oinfo)
    assert np.abs(x).max() < 1e-8
    # use_zero_guess ignores x0
    b = M.set_random(3)
    _, oinfo, _ = M.cg(b, x0=x0, rtol=1e-6, maxiter=1000, use_zero_guess=True)
    _, info, _ = S.solve(b, x0, solver="cg", rtol=1e-6, maxiter=1000, use_zero_guess=True)
    _same(info, oinfo)
    assert info.sol_norm_initial == 0
    # hitting maxiter reports diverged_iters with iters == 0 (cg.hh:135-136)
    _, info, hist = S.solve(b, x0, solver="cg", rtol=1e-12, maxiter=3, history_cap=10)
    assert info.reason == "diverged_iters" and info.iters == 0 and len(hist) == 3
    # already converged: early return with converged_rtol and no iterations
    xs, _, _ = S.solve(b, None, solver="cg", rtol=1e-10, maxiter=1000, use_zero_guess=True)
    _, info, _ = S.solve(b, xs, solver="cg", rtol=1e-6, maxiter=1000)
    assert info.reason == "converged_rtol" and info.iters == 0 and info.callbacks == 0
    S.close(); A.destroy()


@pytest.mark.parametrize("precond", [None, "dinv"])
def test_gmres_parity(ctx, precond):
    n, rp, col, val = _system(7, (12, 11, 10), scale_rows=True)
    A, S, M = _pair(ctx, n, rp, col, val)
    As = sp.csr_matrix((val, col, rp))
    b, x0 = M.set_random(0), M.set_random(1)
    dinv = M.dinv() if precond else None
    kw = dict(rtol=1e-4, maxiter=100)
    _, oinfo, ohist = M.gmres(b, x0=x0, dinv=dinv, history_cap=100, **kw)
    x, info, hist = S.solve(b, x0, solver="gmres", precond=precond, history_cap=100, **kw)
    _same(info, oinfo)
    m = min(len(hist), len(ohist))
    assert np.allclose(hist[:m], ohist[:m], rtol=1e-6)
    assert np.all(np.diff(hist) <= 1e-12)  # residual estimate is monotone (solvers/test/gmres.cc:33-34)
    if info.reason == "converged_rtol":
        assert np.linalg.norm(b - As @ x) <= 2e-4 * np.linalg.norm(b)
    S.close(); A.destroy()


def test_gmres_restart_left_and_jacobi_gain(ctx):
    n, rp, col, val = _system(7, (10, 10, 10), scale_rows=True)
    A, S, M = _pair(ctx, n, rp, col, val)
    b, x0 = M.set_random(0), M.set_random(1)
    its = {}
    for precond in (None, "dinv"):
        for side in ("right", "left"):
            kw = dict(rtol=1e-5, maxiter=400, max_krylov_dim=20, restart=True)
            _, oinfo, _ = M.gmres(b, x0=x0, dinv=M.dinv() if precond else None, right_precond=(side == "right"), **kw)
            _, info, _ = S.solve(b, x0, solver="gmres", precond=precond, pre_side=side, **kw)
            _same(info, oinfo, tol_frac=0.05)
            assert info.restarts == oinfo.restarts or abs(info.iters - oinfo.iters) > 0
            its[(precond, side)] = info.iters if info.iters else 400
    assert its[("dinv", "right")] < its[(None, "right")]  # cf. 73 -> 18 in solvers/test/gmres.cc:79,86
    S.close(); A.destroy()


@pytest.mark.parametrize("precond", [None, "dinv"])
def test_bicgstab_parity(ctx, precond):
    n, rp, col, val = _system(27, (10, 9, 8), scale_rows=(precond == "dinv"))
    A, S, M = _pair(ctx, n, rp, col, val)
    As = sp.csr_matrix((val, col, rp))
    b = As @ np.ones(n)
    x0 = M.set_random(2)
    _, oinfo, ohist = M.bicgstab(b, x0=x0, dinv=M.dinv() if precond else None, rtol=1e-9, maxiter=500, history_cap=500)
    x, info, hist = S.solve(b, x0, solver="bicgstab", precond=precond, rtol=1e-9, maxiter=500, history_cap=500)
    _same(info, oinfo, tol_frac=0.1)  # BiCGStab amplifies rounding differences; the reference bounds it with <=
    assert info.reason == "converged_rtol"
    assert np.linalg.norm(b - As @ x) <= 1e-7 * np.linalg.norm(b)
    m = min(len(hist), len(ohist), 10)
    assert np.allclose(hist[:m], ohist[:m], rtol=1e-6)
    S.close(); A.destroy()


def test_relaxation_preconditioner_runs(ctx):
    """mg::bound_jacobi as a preconditioner handle (the reference's jacobi test only smoke-runs it)."""
    n, rp, col, val = _system(7, (12, 12, 12))
    A, S, M = _pair(ctx, n, rp, col, val)
    b = M.set_random(5)
    _, plain, _ = S.solve(b, None, solver="cg", rtol=1e-8, maxiter=500, use_zero_guess=True)
    x, info, _ = S.solve(b, None, solver="fcg", precond="relax", nrelax=2, rtol=1e-8, maxiter=500, use_zero_guess=True)
    assert info.reason == "converged_rtol" and info.iters <= plain.iters
    As = sp.csr_matrix((val, col, rp))
    assert np.linalg.norm(b - As @ x) <= 2e-8 * np.linalg.norm(b)
    S.close(); A.destroy()


def test_cg_on_multivector_subset_equals_single(ctx):
    """solvers/test/cgmulti.cc:63-76."""
    n, rp, col, val = _system(7, (9, 9, 9))
    A, S, M = _pair(ctx, n, rp, col, val)
    b = np.concatenate([np.zeros(n), M.set_random(3)])
    x0 = np.concatenate([M.set_random(7), M.set_random(4)])
    for which in (0, 1):
        sl = slice(which * n, (which + 1) * n)
        xs, single, _ = S.solve(b[sl], x0[sl], solver="cg", rtol=1e-9, maxiter=1000)
        xm, multi = H.solve_subset(ctx, A, which, b, x0, solver="cg", rtol=1e-9, maxiter=1000)
        assert multi.iters == single.iters and multi.reason == single.reason
        assert np.array_equal(xm[sl], xs)
        other = slice((1 - which) * n, (2 - which) * n)
        assert np.array_equal(xm[other], x0[other])  # the other variable is untouched
    S.close(); A.destroy()


@pytest.mark.parametrize("solver", ["cg", "bicgstab"])
def test_two_component_multivector_system(ctx, solver):
    """Block-diagonal operator on a vec::multi (the shape of examples/equilibrium_diffusion): the
    components are coupled only through the Krylov scalars; result solves both blocks."""
    n, rp, col, val = _system(7, (8, 8, 8))
    rp2, col2, val2 = O.stencil_csr(7, 8, 8, 8, diag_shift=1e-1)
    A0 = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    A1 = F.ParCSR.from_csr(ctx, n, [0, n], rp2, col2, val2)
    rng = np.random.default_rng(3)
    b = np.concatenate([np.zeros(n), rng.random(n)])
    x0 = np.full(2 * n, 2.0)
    # oracle: the same system as one block-diagonal CSR
    Ablk = sp.block_diag([sp.csr_matrix((val, col, rp)), sp.csr_matrix((val2, col2, rp2))]).tocsr()
    Ablk.sort_indices()
    M = O.ParCSR(Ablk.indptr, Ablk.indices, Ablk.data, colours=1)
    fn = M.cg if solver == "cg" else M.bicgstab
    _, oinfo, _ = fn(b, x0=x0, rtol=1e-6, maxiter=500)
    x, info, hist = H.solve_multi2(ctx, A0, A1, b, x0, solver=solver, rtol=1e-6, maxiter=500, history_cap=500)
    _same(info, oinfo, tol_frac=0.1)
    assert np.linalg.norm(b - Ablk @ x) <= 2e-6 * np.linalg.norm(b)
    A0.destroy(); A1.destroy()


def test_host_layer_vector_selftest(ctx):
    """vectors/test/flecsi_vector.cc:338-392 executed by the C++ header layer itself."""
    rp, col, val = O.stencil_csr(5, 8, 4)  # 32 rows: global ids 0..31
    A = F.ParCSR.from_csr(ctx, 32, [0, 32], rp, col, val)
    S = H.Session(ctx, A)
    out = S.vector_selftest()
    assert np.all(out[:11] < 32 * 1e-8), out[:11]
    assert out[11] == -7 and out[12] == 93
    assert np.all(out[13:16] < 1e-8)
    S.close(); A.destroy()


def test_operator_adapter(ctx):
    """y = x - gamma A x (time-integrators/operator_adapter.hh:29-35)."""
    n, rp, col, val = _system(7, (9, 8, 7))
    A, S, M = _pair(ctx, n, rp, col, val)
    x = M.set_random(9)
    y = S.adapter_apply(0.37, x)
    Ax = O.csr_spmv(rp, col, val, x)
    assert np.array_equal(y, O.vec_op("axpy", Ax.copy(), Ax, x, a=-0.37))
    S.close(); A.destroy()


def test_config2_small_scale_iteration_parity(ctx):
    """BASELINE config 2 at 64^3: Jacobi-CG, b = A x_true, x0 = 0, rtol 1e-9f."""
    nn = 64
    rp, col, val = O.stencil_csr(7, nn, nn, nn)
    n = nn ** 3
    g = np.arange(n, dtype=np.uint64)
    h = (g * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    x_true = 1.0 + (h % np.uint64(1000)).astype(np.float64) / 1000.0
    M = O.ParCSR(rp, col, val, colours=1)
    b = M.spmv(x_true)
    xo, oinfo, ohist = M.cg(b, dinv=M.dinv(), rtol=1e-9, maxiter=5000, history_cap=5000)
    A = F.ParCSR.stencil(ctx, 7, nn, nn, nn)
    S = H.Session(ctx, A)
    x, info, hist = S.solve(b, np.zeros(n), solver="cg", precond="dinv", rtol=1e-9, maxiter=5000, history_cap=5000)
    _same(info, oinfo)
    assert np.abs(x - x_true).max() < 1e-6
    m = min(len(hist), len(ohist), 100)
    assert np.allclose(hist[:m], ohist[:m], rtol=1e-7)
    S.close(); A.destroy()


# ---- against golden vectors generated by the reference's own solver templates (tests/golden/)
from tests import golden_util as G  # noqa: E402


@pytest.mark.parametrize("entry", G.load(), ids=lambda e: "-".join(map(str, e["case"][:4])).replace(" ", ""))
def test_device_matches_reference_golden(ctx, entry):
    (rp, col, val), M, b, x0, solver, precond, kw = G.problem(entry["case"])
    n = len(rp) - 1
    A = F.ParCSR.from_csr(ctx, n, [0, n], rp, col, val)
    info = entry["info"]
    if solver == "spmv":
        xv, yv = A.vector(x0), A.vector()
        A.spmv(xv, yv)
        assert G.sha(yv.download()) == entry["x_sha256"]  # bit-identical to the reference's serial SpMV
        xv.destroy(); yv.destroy(); A.destroy()
        return
    S = H.Session(ctx, A)
    x, dinfo, hist = S.solve(b, x0, solver=solver, precond="dinv" if precond else None, history_cap=2000, **kw)
    assert dinfo.status == info["status"]
    tol = 0.02 if solver != "bicgstab" else 0.1
    assert abs(dinfo.iters - info["iters"]) <= max(1, tol * info["iters"]), (dinfo.iters, info["iters"])
    ref_hist = G.history(entry)
    m = min(len(hist), len(ref_hist), 25)
    assert np.allclose(hist[:m], ref_hist[:m], rtol=1e-7)
    assert abs(np.linalg.norm(x) - float.fromhex(entry["x_norm2"])) <= 1e-6 * float.fromhex(entry["x_norm2"])
    S.close(); A.destroy()


def test_config5_residual_history_parity(ctx):
    """BASELINE config 5 (27-point, diag 26 / off -1, b = A 1, x0 = 0, Jacobi-CG) at 64^3: the first 50
    residual norms agree with the oracle to 1e-8, plus a sampled-row SpMV check."""
    nn = 64
    rp, col, val = O.stencil_csr(27, nn, nn, nn)
    n = nn ** 3
    M = O.ParCSR(rp, col, val, colours=1)
    b = M.spmv(np.ones(n))
    _, oinfo, ohist = M.cg(b, dinv=M.dinv(), rtol=0.0, maxiter=50, history_cap=50)
    A = F.ParCSR.stencil(ctx, 27, nn, nn, nn)
    S = H.Session(ctx, A)
    _, info, hist = S.solve(b, np.zeros(n), solver="cg", precond="dinv", rtol=0.0, maxiter=50, history_cap=50)
    assert len(hist) == len(ohist) == 50
    assert np.max(np.abs(hist - ohist) / ohist) <= 1e-8
    rng = np.random.default_rng(1)
    x = rng.standard_normal(n)
    xv, yv = A.vector(x), A.vector()
    A.spmv(xv, yv)
    y = yv.download()
    rows = rng.integers(0, n, 2000)
    ref = np.array([val[rp[r]:rp[r + 1]] @ x[col[rp[r]:rp[r + 1]]] for r in rows])
    assert np.max(np.abs(y[rows] - ref)) <= 1e-12 * np.abs(ref).max()
    xv.destroy(); yv.destroy(); S.close(); A.destroy()
